"""K3 fused Double-DQN target + summed-MSE (autograd-aware) and K0 batched epsilon-greedy."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .env import _chk, _dev, _ptr, _stream, _U64


def gamma_f32(discount_factor: float) -> float:
    """The reference multiplies an int64 tensor by the python float, which torch evaluates in
    float32 (src/dqn_lib.py:131).  Returns that float32 value as a python float."""
    return float(np.float32(discount_factor))


def ddqn_target_loss(q_next_online, q_next_target, q_cur, actions, rewards, dones, discount_factor,
                     use_double=True, want_grad=True):
    """Raw kernel call -> (loss f64[1], target f64[B], q_sa f64[B], grad f64[B,4] or None)."""
    B = q_cur.shape[0]
    dev = _dev(q_cur)
    _lib.init(dev)
    for t, nm in ((q_next_target, "q_next_target"), (q_cur, "q_cur")):
        _chk(t, torch.float64, 4 * B, nm)
    if use_double:
        _chk(q_next_online, torch.float64, 4 * B, "q_next_online")
    for t, nm in ((actions, "actions"), (rewards, "rewards"), (dones, "dones")):
        _chk(t, torch.int64, B, nm)
    kw = dict(dtype=torch.float64, device=q_cur.device)
    target, q_sa, loss = torch.empty(B, **kw), torch.empty(B, **kw), torch.empty(1, **kw)
    grad = torch.empty((B, 4), **kw) if want_grad else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ddqn_target_loss(_ptr(q_next_online) if use_double else None, _ptr(q_next_target),
                                               _ptr(q_cur), _ptr(actions), _ptr(rewards), _ptr(dones),
                                               gamma_f32(discount_factor), 1 if use_double else 0,
                                               _ptr(target), _ptr(q_sa), _ptr(loss), _ptr(grad), B,
                                               _stream(q_cur)), "ddqn_target_loss")
    return loss, target, q_sa, grad


class _DDQNLoss(torch.autograd.Function):
    """loss = sum_i (Q(s_i,a_i) - target_i)^2 with the target treated as a constant.

    The reference does not detach the target (src/dqn_lib.py:126-132), but the only extra effect
    is junk gradient accumulating on target_model, which nothing reads; the online
    network's gradient and the loss value are identical."""

    @staticmethod
    def forward(ctx, q_cur, q_next_online, q_next_target, actions, rewards, dones, discount_factor, use_double):
        loss, target, q_sa, grad = ddqn_target_loss(q_next_online, q_next_target, q_cur.contiguous(), actions,
                                                    rewards, dones, discount_factor, use_double, True)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(target, q_sa)
        return loss.reshape(()), target, q_sa

    @staticmethod
    def backward(ctx, g_loss, _g_target, _g_qsa):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None, None, None, None


def ddqn_loss(q_cur, q_next_online, q_next_target, actions, rewards, dones, discount_factor, use_double=True):
    """Autograd entry point: returns (loss 0-d tensor, target[B], q_sa[B]); backward flows into q_cur."""
    qno = q_next_online.detach().contiguous() if q_next_online is not None else None
    return _DDQNLoss.apply(q_cur, qno, q_next_target.detach().contiguous(), actions, rewards, dones,
                           float(discount_factor), bool(use_double))


def egreedy_select(q, flags, epsilon, seed=2052, ctr=0, index_base=0, override=None, out=None):
    """Batched epsilon_greedy_policy (src/dqn_lib.py:16-30) -> (actions uint8[n], max_q f64[n]).
    `flags`: legal-mask bytes; `override`: uint8[n], 0xFF = Philox, 0x80 = force greedy,
    0..3 = force that random action."""
    n = flags.numel()
    dev = _dev(flags)
    _lib.init(dev)
    _chk(q, torch.float64, 4 * n, "q")
    _chk(flags, torch.uint8, n, "flags")
    if override is not None:
        _chk(override, torch.uint8, n, "override")
    if out is None:
        out = (torch.empty(n, dtype=torch.uint8, device=flags.device),
               torch.empty(n, dtype=torch.float64, device=flags.device))
    actions, max_q = out
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().egreedy_select(_ptr(q), _ptr(flags), float(epsilon), seed & _U64, ctr & _U64,
                                             index_base & _U64, _ptr(override), _ptr(actions), _ptr(max_q), n,
                                             _stream(flags)), "egreedy_select")
    return actions, max_q


class FusedAdam:
    """torch.optim.Adam (no weight decay, no amsgrad) as ONE kernel over flat float64 buffers
    (`ddqn_adam_step`); the step counter lives on the device, so the update is graph-capturable.
    The stock capturable Adam launches ~35 kernels per step for the conv Q-network."""

    def __init__(self, flat_params: torch.Tensor, flat_grads: torch.Tensor, lr=1e-2, betas=(0.9, 0.999), eps=1e-8):
        assert flat_params.dtype == torch.float64 and flat_params.is_contiguous() and flat_grads.is_contiguous()
        self.p, self.g = flat_params, flat_grads
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=flat_params.device)

    def step(self) -> None:
        dev = _dev(self.p)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ddqn_adam_step(_ptr(self.p), _ptr(self.g), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                                 _ptr(self.step_count), self.p.numel(), self.lr, self.betas[0],
                                                 self.betas[1], self.eps, _stream(self.p)), "ddqn_adam_step")

    def state_dict(self) -> dict:
        return {"exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(), "step": self.step_count.cpu(),
                "lr": self.lr, "betas": list(self.betas), "eps": self.eps}

    def load_state_dict(self, st: dict) -> None:
        self.exp_avg.copy_(st["exp_avg"]); self.exp_avg_sq.copy_(st["exp_avg_sq"]); self.step_count.copy_(st["step"])
        self.lr, self.betas, self.eps = st["lr"], tuple(st["betas"]), st["eps"]
