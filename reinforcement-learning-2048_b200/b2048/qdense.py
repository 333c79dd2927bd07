"""K8: the reference's dense Q-network (src/configs/double_dqn_dense.py:7-15: Linear 16 -> 512 -> 512 -> 256 -> 4 with
ReLU between, float64) on the FP64 tensor cores, without cuBLAS and without autograd.

`DenseQ(net)` evaluates any float64 CUDA `nn.Sequential` of the form Linear (ReLU Linear)* whose last layer has four
outputs; it reads the module's own parameter tensors (nothing is copied), so optimizer steps and
`load_state_dict` are seen immediately.  `TrainableDenseQ` adds the backward pass the Double-DQN update needs
(`forward_saving` / `backward_into`, the same contract as `qfused.TrainableConvQ`): d loss / d Q(s,.) from K3 goes
through `dense_linear_dgrad_f64` (input gradient fused with the ReLU mask) and `dense_linear_wgrad_f64` (weight and
bias gradients written straight into the flat gradient buffer)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .env import _dev, _ptr, _stream


def matches(net: nn.Module) -> bool:
    """Linear (ReLU Linear)*, float64 parameters on one CUDA device, even widths, four outputs."""
    if not isinstance(net, nn.Sequential) or len(net) < 1 or len(net) % 2 == 0:
        return False
    mods = list(net)
    for i, m in enumerate(mods):
        if i % 2 == 0:
            if not isinstance(m, nn.Linear) or m.bias is None:
                return False
            if m.weight.dtype != torch.float64 or not m.weight.is_cuda or not m.weight.is_contiguous():
                return False
            if m.in_features % 2 or (m.out_features % 2):
                return False
        elif not isinstance(m, nn.ReLU):
            return False
    return mods[-1].out_features == 4


class DenseQ:
    """No-gradient evaluator: `q = DenseQ(net)(states)`, states float64 [n, n_in] (or anything reshapeable to it)."""

    def __init__(self, net: nn.Sequential):
        if not matches(net):
            raise ValueError("DenseQ needs a float64 CUDA nn.Sequential of Linear / ReLU layers ending in 4 outputs")
        self.net = net
        self.linears = [m for m in net if isinstance(m, nn.Linear)]
        self.device = self.linears[0].weight.device
        _lib.init(_dev(self.linears[0].weight))

    def _forward(self, x: torch.Tensor, keep: bool):
        n = x.shape[0]
        x = x.reshape(n, self.linears[0].in_features)
        if not (x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()):
            x = x.to(self.device, torch.float64).contiguous()
        acts = [x]
        dev = _dev(x)
        with torch.cuda.device(dev):
            for li, lin in enumerate(self.linears):
                last = li == len(self.linears) - 1
                out = torch.empty((n, lin.out_features), dtype=torch.float64, device=x.device)
                if n:
                    _lib.check(_lib.lib().dense_linear_forward_f64(_ptr(acts[-1]), _ptr(lin.weight), _ptr(lin.bias), _ptr(out),
                                                                   n, lin.in_features, lin.out_features, 0 if last else 1,
                                                                   _stream(x)), "dense_linear_forward_f64")
                acts.append(out)
        return (acts[-1], acts[:-1]) if keep else (acts[-1], None)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward(x, False)[0]


class _DenseQTrain(torch.autograd.Function):
    """Autograd entry point (dqn_lib.train_step with the caller's own loss / optimizer): forward = K8 with saved
    activations, backward = K8 gradients; gradients flow to the parameters only."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        q, saved = mod._fused._forward(x, True)
        ctx.mod, ctx.saved = mod, saved
        return q

    @staticmethod
    def backward(ctx, gq):
        grads = [torch.empty_like(p) for p in ctx.mod.params()]
        ctx.mod.backward_into(ctx.saved, gq.contiguous(), grads)
        return (None, None) + tuple(grads)


class TrainableDenseQ(nn.Module):
    def __init__(self, net: nn.Sequential):
        super().__init__()
        if not matches(net):
            raise ValueError("TrainableDenseQ needs a float64 CUDA nn.Sequential of Linear / ReLU layers ending in 4 outputs")
        self.net = net
        self._fused = DenseQ(net)
        self._scratch = {}

    def params(self):
        """weight, bias of every Linear in nn.Sequential order (= the order of net.parameters())."""
        out = []
        for lin in self._fused.linears:
            out += [lin.weight, lin.bias]
        return tuple(out)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.requires_grad or not x.is_cuda:
            return self.net(x)                        # a gradient w.r.t. the input is asked for: plain torch
        if x.shape[0] == 0 or not (torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters())):
            return self._fused(x)
        return _DenseQTrain.apply(self, x, *self.params())

    @torch.no_grad()
    def forward_saving(self, x: torch.Tensor):
        """Q(s) plus the layer inputs `backward_into` needs, outside autograd (DDQNUpdater's direct path)."""
        return self._fused._forward(x, True)

    def _scratch_for(self, rows: int, lin: nn.Linear) -> torch.Tensor:
        key = (rows, lin.in_features, lin.out_features)
        buf = self._scratch.get(key)
        if buf is None:
            n = int(_lib.lib().dense_linear_wgrad_scratch_elems(rows, lin.in_features, lin.out_features))
            buf = torch.empty(n, dtype=torch.float64, device=lin.weight.device)
            self._scratch[key] = buf
        return buf

    @torch.no_grad()
    def backward_into(self, saved, gq: torch.Tensor, grads) -> None:
        """Writes d loss / d parameter into `grads` (contiguous tensors in `params()` order), overwriting them.
        `saved` = the layer inputs from forward_saving (x, h1, h2, ...), `gq` = d loss / d Q [n, 4]."""
        lins = self._fused.linears
        n = gq.shape[0]
        dev = _dev(gq)
        g = gq
        st = _stream(gq)
        with torch.cuda.device(dev):
            for li in range(len(lins) - 1, -1, -1):
                lin, x_in = lins[li], saved[li]
                if li > 0:          # gradient w.r.t. this layer's input, masked by the ReLU that produced it
                    dz = torch.empty((n, lin.in_features), dtype=torch.float64, device=gq.device)
                    _lib.check(_lib.lib().dense_linear_dgrad_f64(_ptr(g), _ptr(lin.weight), _ptr(x_in), _ptr(dz), n,
                                                                 lin.in_features, lin.out_features, st), "dense_linear_dgrad_f64")
                _lib.check(_lib.lib().dense_linear_wgrad_f64(_ptr(g), _ptr(x_in), _ptr(grads[2 * li]), _ptr(grads[2 * li + 1]),
                                                             _ptr(self._scratch_for(n, lin)), n, lin.in_features,
                                                             lin.out_features, st), "dense_linear_wgrad_f64")
                if li > 0:
                    g = dz
