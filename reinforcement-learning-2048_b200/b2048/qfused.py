"""K6: the reference's convolutional Q-network evaluated by one fused kernel (no gradient).

`FusedConvQ` wraps the caller's nn.Sequential of the conv config (src/configs/double_dqn_conv.py:19-28)
without copying anything: the kernel reads the module's own parameter tensors at every call, so
optimizer steps and target syncs are seen immediately.  It serves the places where the reference runs
the network without needing gradients — action selection (src/dqn_lib.py:24-25), greedy play
(src/player.py:47) and Q(s') of the Double-DQN target (src/dqn_lib.py:126-128).  The autograd forward
of Q(s) stays on the GEMM path (`qnet.FastQNet`).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .env import _chk, _dev, _ptr, _stream

SCALINGS = {"log2": 0, "normalized": 1}


def matches(net: nn.Module) -> bool:
    """True for Conv2d(1,64,2) ReLU Conv2d(64,64,2) ReLU Flatten Linear(256,64) ReLU Linear(64,4) in
    float64 on a CUDA device, all layers with bias and default stride / padding."""
    if not isinstance(net, nn.Sequential) or len(net) != 8:
        return False
    c1, r1, c2, r2, fl, l1, r3, l2 = net
    if not (isinstance(c1, nn.Conv2d) and isinstance(c2, nn.Conv2d) and isinstance(fl, nn.Flatten)
            and isinstance(l1, nn.Linear) and isinstance(l2, nn.Linear)
            and all(isinstance(r, nn.ReLU) for r in (r1, r2, r3))):
        return False
    for c, (cin, cout) in ((c1, (1, 64)), (c2, (64, 64))):
        if (c.in_channels, c.out_channels) != (cin, cout) or c.kernel_size != (2, 2) or c.stride != (1, 1) \
                or c.padding != (0, 0) or c.dilation != (1, 1) or c.groups != 1 or c.bias is None \
                or c.padding_mode != "zeros":
            return False
    if (l1.in_features, l1.out_features) != (256, 64) or (l2.in_features, l2.out_features) != (64, 4):
        return False
    if l1.bias is None or l2.bias is None or (fl.start_dim, fl.end_dim) != (1, -1):
        return False
    return all(p.dtype == torch.float64 and p.is_cuda and p.is_contiguous() for p in net.parameters())


class FusedConvQ:
    def __init__(self, net: nn.Sequential):
        if not matches(net):
            raise ValueError("FusedConvQ needs the float64 conv Q-network of configs/double_dqn_conv.py on a CUDA device")
        self.net = net
        c1, _, c2, _, _, l1, _, l2 = net
        self._params = (c1.weight, c1.bias, c2.weight, c2.bias, l1.weight, l1.bias, l2.weight, l2.bias)
        self.device = c1.weight.device

    def _run(self, boards, states, scaling, n, out):
        dev = _dev(self._params[0])
        _lib.init(dev)
        if out is None:
            out = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        _chk(out, torch.float64, 4 * n, "out")
        ref = boards if boards is not None else states
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().qnet_conv_forward_f64(
                _ptr(boards), _ptr(states), SCALINGS[scaling], *[_ptr(p) for p in self._params], _ptr(out), n,
                _stream(ref)), "qnet_conv_forward_f64")
        return out

    @torch.no_grad()
    def forward_boards(self, boards: torch.Tensor, scaling: str = "log2", out=None) -> torch.Tensor:
        """Q[n,4] for packed boards; `scaling` picks the input the network sees: "log2" = exponents
        (board.log_scale(), what training uses) or "normalized" = tile / max tile (player.py)."""
        if scaling not in SCALINGS:
            raise ValueError(f"unknown input scaling {scaling!r}")
        _chk(boards, torch.int64, name="boards")
        return self._run(boards, None, scaling, boards.numel(), out)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, out=None) -> torch.Tensor:
        """Q[n,4] for network inputs x: float64 [n,1,4,4] or [n,16] (contiguous)."""
        n = x.shape[0]
        _chk(x, torch.float64, 16 * n, "x")
        return self._run(None, x, "log2", n, out)


def accelerate_inference(net: nn.Module):
    """The fastest no-gradient evaluator for `net`: the fused kernel for the reference's conv
    Q-network, otherwise `qnet.accelerate(net)` (float64 GEMM path)."""
    from .qnet import accelerate
    return FusedConvQ(net) if matches(net) else accelerate(net)
