"""Float64 Q-networks with gradients for tiny boards (the autograd path of train_step,
src/dqn_lib.py:146-161).

`accelerate(net)` picks the evaluator: the reference's conv Q-network (configs/double_dqn_conv.py:19-28)
gets `qfused.TrainableConvQ` (K6 forward with saved activations, hand-built backward); any other small
conv net gets `FastQNet` below.  cuDNN has no fast float64 path for 2x2 convolutions on a 4x4 board
(12.9 ms per update at batch 5000 on B200, 0.3 TFLOP/s); written as patch gather + one float64 GEMM per
layer (cuBLAS DGEMM) on row-matrix activations, with the weight / bias gradients from the K7 kernels,
the same arithmetic runs an order of magnitude faster.  The weights stay in the caller's nn.Sequential
— the same parameter tensors are used, so training the wrapper trains the original module (checkpoints,
target sync and `Experiment.save` keep working) — and only the summation order changes (differences
~1e-15 relative, inside the 1e-9 parity gate).
"""
from __future__ import annotations

import torch
from torch import nn


class _Patches(torch.autograd.Function):
    """x [n*h*w, c] (activation as a row matrix) -> cols [n*oh*ow, c*kh*kw] with one gather kernel each
    way (conv_patches_f64 / conv_patches_grad_f64).  torch's x.unfold(...).unfold(...) backward costs two
    scatter kernels plus fills (~90 us per update for the second conv), F.unfold is ~100x slower in
    float64."""

    @staticmethod
    def forward(ctx, x, n, c, h, w, kh, kw):
        from . import _lib
        from .env import _dev, _ptr, _stream
        x = x.contiguous()
        _lib.init(_dev(x))
        ctx.geom = (n, c, h, w, kh, kw)
        cols = torch.empty((n * (h - kh + 1) * (w - kw + 1), c * kh * kw), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().conv_patches_f64(_ptr(x), _ptr(cols), n, c, h, w, kh, kw, _stream(x)),
                       "conv_patches_f64")
        return cols

    @staticmethod
    def backward(ctx, dcols):
        from . import _lib
        from .env import _ptr, _stream
        n, c, h, w, kh, kw = ctx.geom
        dcols = dcols.contiguous()
        dx = torch.empty((n * h * w, c), dtype=dcols.dtype, device=dcols.device)
        with torch.cuda.device(dcols.device):
            _lib.check(_lib.lib().conv_patches_grad_f64(_ptr(dcols), _ptr(dx), n, c, h, w, kh, kw, _stream(dcols)),
                       "conv_patches_grad_f64")
        return dx, None, None, None, None, None, None


def _wgrad_kind(c: int, k: int):
    """Which hand-written weight-gradient kernel covers a [c, k] weight matrix, if any."""
    if c <= 64 and k <= 64 and c * k <= 1024:
        return "small"                       # layer_wgrad_small_f64 (conv1 64x4, output layer 4x64)
    if c == 64 and k % 32 == 0 and k <= 256:
        return "dmma64"                      # layer_wgrad64_f64 (conv2 and fc1 of the conv net: 64x256)
    return None


class _AddmmOwnWgrad(torch.autograd.Function):
    """y = x @ W^T + b with the forward on cuBLAS (addmm) and dW, db from the hand-written kernels in
    csrc/wgrad_kernels.cu.  These gradients are reductions over the 5 000 .. 45 000 rows of the batch
    into a tiny matrix; cuBLAS runs them as a tall-skinny DGEMM on a handful of CTAs and ATen adds a
    generic column reduction for the bias (conv net, inside the update graph: 57+21, 50+14, 22+10 and
    8+7 us for the four layers).  The kernels split the rows over all SMs and add the per-CTA partial
    results in a fixed order, so the update stays bit-reproducible."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, gy):
        from . import _lib
        from .env import _dev, _ptr, _stream
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        rows, c = gy.shape
        k = x.shape[1]
        _lib.init(_dev(gy))
        L = _lib.lib()
        gw = torch.empty((c, k), dtype=gy.dtype, device=gy.device)
        gb = torch.empty(c, dtype=gy.dtype, device=gy.device)
        with torch.cuda.device(gy.device):
            if _wgrad_kind(c, k) == "small":
                scratch = torch.empty(L.layer_wgrad_small_scratch_elems(rows, c, k), dtype=gy.dtype, device=gy.device)
                _lib.check(L.layer_wgrad_small_f64(_ptr(gy), _ptr(x), _ptr(gw), _ptr(gb), _ptr(scratch), rows, c, k,
                                                   _stream(gy)), "layer_wgrad_small_f64")
            else:
                scratch = torch.empty(L.layer_wgrad64_scratch_elems(rows, k), dtype=gy.dtype, device=gy.device)
                _lib.check(L.layer_wgrad64_f64(_ptr(gy), _ptr(x), _ptr(gw), _ptr(gb), _ptr(scratch), rows, k,
                                               _stream(gy)), "layer_wgrad64_f64")
        gx = torch.mm(gy, weight) if ctx.needs_input_grad[0] else None
        return gx, gw, gb


def _affine(x: torch.Tensor, weight2d: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    own = (x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and torch.is_grad_enabled()
           and weight2d.requires_grad and _wgrad_kind(*weight2d.shape) is not None)
    if own:
        return _AddmmOwnWgrad.apply(x.contiguous(), weight2d, bias)
    return torch.addmm(bias, x, weight2d.t())


def _plain_conv(m: nn.Conv2d) -> bool:
    return (m.stride == (1, 1) and m.padding == (0, 0) and m.dilation == (1, 1) and m.groups == 1
            and m.bias is not None and m.padding_mode == "zeros")


class FastQNet(nn.Module):
    """Wraps the reference's conv Q-network (an nn.Sequential of Conv2d / ReLU / Flatten / Linear on a
    float64 CUDA device).  Activations live as row matrices [n*h*w, c]: every Conv2d is a patch gather
    + `_affine` (cuBLAS addmm forward, own weight-gradient kernels backward), nn.Flatten restores the
    (c, h, w) feature order the Linear weights expect, ReLU is torch's.  The parameters are the wrapped
    module's own tensors."""

    def __init__(self, net: nn.Sequential):
        super().__init__()
        self.net = net

    @staticmethod
    def supports(net: nn.Module) -> bool:
        if not isinstance(net, nn.Sequential) or not any(isinstance(m, nn.Conv2d) for m in net):
            return False
        for m in net:
            if isinstance(m, nn.Conv2d):
                if not _plain_conv(m):
                    return False
            elif isinstance(m, nn.Linear):
                if m.bias is None:
                    return False
            elif not isinstance(m, (nn.ReLU, nn.Flatten)):
                return False
        return all(p.is_cuda and p.dtype == torch.float64 for p in net.parameters())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, c, h, w = x.shape
        rows = (x if c == 1 else x.permute(0, 2, 3, 1)).reshape(n * h * w, c)
        for m in self.net:
            if isinstance(m, nn.Conv2d):
                kh, kw = m.kernel_size
                rows = _affine(_Patches.apply(rows, n, c, h, w, kh, kw), m.weight.reshape(m.out_channels, -1), m.bias)
                c, h, w = m.out_channels, h - kh + 1, w - kw + 1
            elif isinstance(m, nn.Linear):
                rows = _affine(rows.reshape(n, -1), m.weight, m.bias)
                c, h, w = m.out_features, 1, 1
            elif isinstance(m, nn.Flatten):                # nn.Flatten on NCHW: feature index (c, h, w)
                if h * w > 1:
                    rows = rows.reshape(n, h * w, c).transpose(1, 2).reshape(n, c * h * w)
                c, h, w = c * h * w, 1, 1
            else:
                rows = torch.relu(rows)
        return rows.reshape(n, -1) if h * w == 1 else rows.reshape(n, h, w, c).permute(0, 3, 1, 2)


def accelerate(net: nn.Module) -> nn.Module:
    """The fastest evaluator with gradients for `net`: `qfused.TrainableConvQ` (K6 forward, K7 backward)
    for the reference's conv Q-network, `qdense.TrainableDenseQ` (K8) for its dense one and other
    Linear / ReLU stacks with four outputs, FastQNet for other float64 CUDA Sequentials of plain small
    convolutions + ReLU / Flatten / Linear, otherwise the module itself."""
    from . import qdense, qfused
    if qfused.matches(net):
        return qfused.TrainableConvQ(net)
    if qdense.matches(net):
        return qdense.TrainableDenseQ(net)
    return FastQNet(net) if FastQNet.supports(net) else net
