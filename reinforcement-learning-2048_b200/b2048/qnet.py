"""Float64 Q-networks with gradients (the autograd path of train_step, src/dqn_lib.py:146-163).

The reference's conv net (configs/double_dqn_conv.py:19-28) is two 2x2 convolutions on a 4x4 board.
cuDNN has no fast float64 path for it (12.9 ms per update at batch 5000 on B200, 0.3 TFLOP/s);
written as patch-gather + one float64 GEMM per layer (cuBLAS DGEMM) the same arithmetic runs an
order of magnitude faster.  Both reference networks are then chains of "GEMM + bias + ReLU": the
GEMMs go to cuBLAS, everything around them is two fused kernels per layer (csrc/layer_kernels.cu)
instead of ATen's bias broadcast, clamp, threshold_backward and generic column reductions.  The
weights stay in the caller's nn.Sequential — the same parameter tensors are used, so training the
wrapper trains the original module (checkpoints, target sync and `Experiment.save` keep working) —
and only the summation order inside a convolution changes (differences ~1e-15 relative, inside the
1e-9 parity gate).
"""
from __future__ import annotations

import torch
from torch import nn


def _cuda_call(name, t, *args):
    from . import _lib
    from .env import _stream
    with torch.cuda.device(t.device):
        _lib.check(getattr(_lib.lib(), name)(*args, _stream(t)), name)


class _PatchesRows(torch.autograd.Function):
    """im2col of an activation kept as a row matrix [n*h*w, c] (what a GEMM over patches produces), so
    that consecutive convolutions never go back to NCHW (conv_patches_rows_f64 / _grad_f64)."""

    @staticmethod
    def forward(ctx, x, n, c, h, w, kh, kw):
        from .env import _ptr
        x = x.contiguous()
        ctx.geom = (n, c, h, w, kh, kw)
        cols = torch.empty((n * (h - kh + 1) * (w - kw + 1), c * kh * kw), dtype=x.dtype, device=x.device)
        _cuda_call("conv_patches_rows_f64", x, _ptr(x), _ptr(cols), n, c, h, w, kh, kw)
        return cols

    @staticmethod
    def backward(ctx, dcols):
        from .env import _ptr
        n, c, h, w, kh, kw = ctx.geom
        dcols = dcols.contiguous()
        dx = torch.empty((n * h * w, c), dtype=dcols.dtype, device=dcols.device)
        _cuda_call("conv_patches_rows_grad_f64", dcols, _ptr(dcols), _ptr(dx), n, c, h, w, kh, kw)
        return dx, None, None, None, None, None, None


class _LinearAct(torch.autograd.Function):
    """y = act(x @ W^T + b) for row-major float64 x [R,K], W [C,K]: cuBLAS DGEMM + one in-place
    bias/ReLU kernel forward; backward = one kernel for ReLU mask and bias gradient (fixed summation
    order) + two DGEMMs.  Replaces addmm's bias broadcast, clamp, threshold_backward and the generic
    reduce kernel ATen runs for every bias gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        from .env import _ptr
        y = torch.mm(x, weight.t())
        _cuda_call("layer_bias_act_f64", y, _ptr(y), _ptr(bias), y.shape[0], y.shape[1], int(relu))
        ctx.relu = relu
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import _lib
        from .env import _ptr
        x, weight, y = ctx.saved_tensors
        gy = gy.contiguous()
        rows, cols = gy.shape
        g = torch.empty_like(gy)
        db = torch.empty(cols, dtype=gy.dtype, device=gy.device)
        scratch = torch.empty(_lib.lib().layer_act_grad_scratch_elems(rows, cols), dtype=gy.dtype, device=gy.device)
        _cuda_call("layer_act_grad_bias_f64", gy, _ptr(gy), _ptr(y) if ctx.relu else None, _ptr(g), _ptr(db),
                   _ptr(scratch), rows, cols, int(ctx.relu))
        gx = torch.mm(g, weight) if ctx.needs_input_grad[0] else None
        gw = torch.mm(g.t(), x) if ctx.needs_input_grad[1] else None
        return gx, gw, (db if ctx.needs_input_grad[2] else None), None


def _fusable_param(p: torch.Tensor) -> bool:
    return p.is_cuda and p.dtype == torch.float64 and p.is_contiguous() and p.data_ptr() % 16 == 0


def _plain_conv(m: nn.Conv2d) -> bool:
    return (m.stride == (1, 1) and m.padding == (0, 0) and m.dilation == (1, 1) and m.groups == 1
            and m.bias is not None and m.padding_mode == "zeros")


class FastQNet(nn.Module):
    """Wraps an nn.Sequential Q-network of Conv2d / Linear / ReLU / Flatten layers (the reference's
    conv and dense configs).  Activations live as float64 row matrices [n*h*w, c]; every Conv2d /
    Linear (+ following ReLU) is `_LinearAct`, convolutions gather their patches with
    `_PatchesRows`, and nn.Flatten restores the (c, h, w) feature order the Linear weights expect.
    The parameters are the wrapped module's own tensors."""

    def __init__(self, net: nn.Sequential):
        super().__init__()
        self.net = net

    @staticmethod
    def supports(net: nn.Module) -> bool:
        if not isinstance(net, nn.Sequential) or len(net) == 0:
            return False
        for m in net:
            if isinstance(m, nn.Conv2d):
                if not _plain_conv(m) or m.out_channels % 2:
                    return False
            elif isinstance(m, nn.Linear):
                if m.bias is None or m.out_features % 2:
                    return False
            elif not isinstance(m, (nn.ReLU, nn.Flatten)):
                return False
        return isinstance(net[0], (nn.Conv2d, nn.Linear)) and all(_fusable_param(p) for p in net.parameters())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mods = list(self.net)
        n = x.shape[0]
        if x.dim() == 4:                                   # NCHW input -> rows (n, h, w) x c
            c, h, w = x.shape[1:]
            rows = x.reshape(n, h * w) if c == 1 else x.permute(0, 2, 3, 1).reshape(n * h * w, c)
            rows = rows.reshape(n * h * w, c).contiguous()
        else:
            c, h, w = x.shape[1], 1, 1
            rows = x.contiguous()
        i = 0
        while i < len(mods):
            m = mods[i]
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            if isinstance(m, nn.Conv2d):
                kh, kw = m.kernel_size
                cols = _PatchesRows.apply(rows, n, c, h, w, kh, kw)
                rows = _LinearAct.apply(cols, m.weight.reshape(m.out_channels, -1), m.bias, relu)
                c, h, w = m.out_channels, h - kh + 1, w - kw + 1
                i += 2 if relu else 1
            elif isinstance(m, nn.Linear):
                rows = _LinearAct.apply(rows.reshape(n, -1), m.weight, m.bias, relu)
                c, h, w = m.out_features, 1, 1
                i += 2 if relu else 1
            elif isinstance(m, nn.Flatten):                # NCHW flatten order: feature = (c, h, w)
                if h * w > 1:
                    rows = rows.reshape(n, h * w, c).transpose(1, 2).reshape(n, c * h * w)
                c, h, w = c * h * w, 1, 1
                i += 1
            else:                                          # a ReLU that follows nothing fusable
                rows = torch.relu(rows)
                i += 1
        return rows.reshape(n, -1) if h * w == 1 else rows.reshape(n, h, w, c).permute(0, 3, 1, 2)


def accelerate(net: nn.Module) -> nn.Module:
    """FastQNet for float64 CUDA Sequentials of Conv2d / Linear / ReLU / Flatten (both reference
    configs), otherwise the module itself."""
    return FastQNet(net) if FastQNet.supports(net) else net
