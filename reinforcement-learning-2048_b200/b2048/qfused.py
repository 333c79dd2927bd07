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

    def forward_boards_raw(self, boards_ptr: int, out_ptr: int, n: int, stream: int) -> None:
        """forward_boards ("log2" scaling) on raw device addresses of buffers the caller has validated (int64 [n] packed
        boards, float64 [n, 4] output, on this network's device, which must be current): the launch-bound rollout loop
        (`VectorEnv.step`) calls this once per step."""
        _lib.check(_lib.lib().qnet_conv_forward_f64(boards_ptr, None, SCALINGS["log2"], *[p.data_ptr() for p in self._params],
                                                    out_ptr, n, stream), "qnet_conv_forward_f64")

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


_SIDE = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def conv_q_forward_saving(x, params):
    """K6 with saved activations on x [n,16] (float64, contiguous): -> (q [n,4], saved) where saved =
    (x, patches2 [4n,256], act2 [n,256], act3 [n,64]) is what `conv_q_backward` needs.  No autograd."""
    n = x.shape[0]
    dev = _dev(x)
    _lib.init(dev)
    kw = dict(dtype=torch.float64, device=x.device)
    q, p2 = torch.empty((n, 4), **kw), torch.empty((4 * n, 256), **kw)
    a2, a3 = torch.empty((n, 256), **kw), torch.empty((n, 64), **kw)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qnet_conv_forward_train_f64(_ptr(x), *[_ptr(p) for p in params], _ptr(q), _ptr(p2), _ptr(a2),
                                                          _ptr(a3), n, _stream(x)), "qnet_conv_forward_train_f64")
    return q, (x, p2, a2, a3)


def conv_q_forward_update(x, x_next, params, target_params, want_online_next=True):
    """The forwards of one Double-DQN update as ONE K6 launch (`qnet_conv_forward_update_f64`): Q(s) of the
    online network with saved activations, Q(s') of the online network (optional) and Q(s') of the target.
    -> (q, saved, q_next_online | None, q_next_target); bit-identical to the three separate launches."""
    import ctypes
    n = x.shape[0]
    dev = _dev(x)
    _lib.init(dev)
    kw = dict(dtype=torch.float64, device=x.device)
    q, p2 = torch.empty((n, 4), **kw), torch.empty((4 * n, 256), **kw)
    a2, a3 = torch.empty((n, 256), **kw), torch.empty((n, 64), **kw)
    qno = torch.empty((n, 4), **kw) if want_online_next else None
    qnt = torch.empty((n, 4), **kw)
    arr = ctypes.c_void_p * 8
    on, tg = arr(*[_ptr(p) for p in params]), arr(*[_ptr(p) for p in target_params])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().qnet_conv_forward_update_f64(_ptr(x), _ptr(x_next), on, tg, _ptr(q), _ptr(p2), _ptr(a2), _ptr(a3),
                                                           _ptr(qno) if qno is not None else None, _ptr(qnt), n, _stream(x)),
                   "qnet_conv_forward_update_f64")
    return q, (x, p2, a2, a3), qno, qnt


def conv_q_backward(saved, params, gq, out=None):
    """Gradients of the eight parameter tensors given gq = d loss / d q [n,4]: per layer the K7 weight /
    bias gradient kernels on the saved matrices (on a side stream: they are leaves of the dependency chain
    g4 -> g3 -> g2 -> g1), tensor-core kernels for the input gradients, fused with the ReLU masks (fc1's is stored
    regrouped as (board, position) x channel rows; conv2's never leaves the accumulators: the first convolution's whole
    backward is applied to it in the same kernel).  No library GEMM.  `out` (optional): eight contiguous tensors in
    parameter order that receive the gradients (overwritten, not accumulated) — e.g. the views of a flat gradient buffer."""
    x, p2, a2, a3 = saved
    w2, w3, w4 = params[2], params[4], params[6]
    n = x.shape[0]
    dev = _dev(x)
    L = _lib.lib()
    kw = dict(dtype=torch.float64, device=x.device)
    main = torch.cuda.current_stream(x.device)
    side = _side_stream(x.device)
    st = main.cuda_stream
    if out is None:
        out = [torch.empty(p.shape, **kw) for p in params]
    gw1, gb1, gw2, gb2, gw3, gb3, gw4, gb4 = out

    def wgrad(g, xin, gw, gb, c, k, stream):
        if c * k <= 1024:
            scratch = torch.empty(L.layer_wgrad_small_scratch_elems(g.shape[0], c, k), **kw)
            _lib.check(L.layer_wgrad_small_f64(_ptr(g), _ptr(xin), _ptr(gw), _ptr(gb), _ptr(scratch), g.shape[0], c, k,
                                               stream), "layer_wgrad_small_f64")
        else:
            scratch = torch.empty(L.layer_wgrad64_scratch_elems(g.shape[0], k), **kw)
            _lib.check(L.layer_wgrad64_f64(_ptr(g), _ptr(xin), _ptr(gw), _ptr(gb), _ptr(scratch), g.shape[0], k, stream),
                       "layer_wgrad64_f64")

    def ready():                                # marks "g is complete" on the main stream
        ev = torch.cuda.Event()
        ev.record(main)
        return ev

    def wgrad_aside(ev, g, xin, gw, gb, c, k):  # fork at `ev`, join at the end; capturable into a CUDA graph
        side.wait_event(ev)
        with torch.cuda.stream(side):
            wgrad(g, xin, gw, gb, c, k, side.cuda_stream)

    # Per layer: the input-gradient GEMM (critical path) is launched first, the weight-gradient kernel of
    # the same layer second, on the side stream, waiting only for g: it fills in behind the GEMM's CTAs
    # instead of taking the SMs' shared memory ahead of it.
    def dgrad(g, w2d, h, rows, n_in, n_out):    # (g w) * (h > 0) on the FP64 tensor cores (K8); h None: no mask
        dz = torch.empty((rows, n_in), **kw)
        _lib.check(L.dense_linear_dgrad_f64(_ptr(g), _ptr(w2d), _ptr(h), _ptr(dz), rows, n_in, n_out, st),
                   "dense_linear_dgrad_f64")
        return dz

    with torch.cuda.device(dev):
        g4 = gq.contiguous()                                              # [n, 4]
        ev = ready()
        g3 = dgrad(g4, w4, a3, n, 64, 4)                                  # [n, 64] incl. the ReLU mask of fc1
        wgrad_aside(ev, g4, a3, gw4, gb4, 4, 64)
        ev = ready()
        # fc1's input gradient, written directly as rows (board, position) x channel (feature = channel*4 + position)
        g2 = torch.empty((4 * n, 64), **kw)
        _lib.check(L.dense_linear_dgrad_regroup_f64(_ptr(g3), _ptr(w3), _ptr(a2), _ptr(g2), n, 256, 64, 4, st),
                   "dense_linear_dgrad_regroup_f64")
        wgrad_aside(ev, g3, a2, gw3, gb3, 64, 256)
        ev = ready()
        # conv2's input gradient (g2 W2, the gradient of the patch matrix) never leaves the tensor-core accumulators:
        # relu'(conv1), dW1 and db1 against the boards' cells are taken from them in the same kernel
        scratch = torch.empty(L.conv2_dgrad_conv1_wgrad_scratch_elems(n), **kw)
        _lib.check(L.conv2_dgrad_conv1_wgrad_f64(_ptr(g2), _ptr(w2), _ptr(p2), _ptr(x), _ptr(gw1), _ptr(gb1), _ptr(scratch),
                                                 n, st), "conv2_dgrad_conv1_wgrad_f64")
        wgrad_aside(ev, g2, p2, gw2, gb2, 64, 256)
        main.wait_stream(side)
    return out


class _ConvQTrain(torch.autograd.Function):
    """Q(s) of the conv Q-network with a hand-built backward (train_step, src/dqn_lib.py:146-161):
    `conv_q_forward_saving` / `conv_q_backward` behind autograd.  Same arithmetic as autograd on the
    nn.Sequential, different summation order (1e-10 relative on the gradients)."""

    @staticmethod
    def forward(ctx, x, *params):
        q, saved = conv_q_forward_saving(x, params)
        ctx.save_for_backward(*saved, *params)
        return q

    @staticmethod
    def backward(ctx, gq):
        t = ctx.saved_tensors
        return (None, *conv_q_backward(t[:4], t[4:], gq))


class TrainableConvQ(nn.Module):
    """The reference's conv Q-network as a module whose forward is K6 and whose backward is K7 + the fused
    input-gradient kernels (`_ConvQTrain`); without gradients it is the plain fused forward.  Wraps the caller's nn.Sequential:
    the parameters are its own tensors."""

    def __init__(self, net: nn.Sequential):
        super().__init__()
        if not matches(net):
            raise ValueError("TrainableConvQ needs the float64 conv Q-network of configs/double_dqn_conv.py on a CUDA device")
        self.net = net
        self._fused = FusedConvQ(net)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        x2 = x.reshape(n, 16)
        if not (x2.is_cuda and x2.dtype == torch.float64 and x2.is_contiguous()) or x.requires_grad:
            from .qnet import FastQNet               # e.g. a gradient w.r.t. the input is asked for: GEMM path
            return FastQNet(self.net)(x.reshape(n, 1, 4, 4)) if FastQNet.supports(self.net) else self.net(x)
        if n == 0 or not (torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters())):
            return self._fused(x2)
        return _ConvQTrain.apply(x2, *self.params())

    def params(self):
        """The eight parameter tensors in the order the kernels take them (= nn.Sequential order)."""
        c1, _, c2, _, _, l1, _, l2 = self.net
        return (c1.weight, c1.bias, c2.weight, c2.bias, l1.weight, l1.bias, l2.weight, l2.bias)

    @torch.no_grad()
    def forward_saving(self, x: torch.Tensor):
        """Q(s) plus the activations `backward_into` needs, outside autograd (DDQNUpdater's direct path)."""
        return conv_q_forward_saving(x.reshape(x.shape[0], 16), self.params())

    @torch.no_grad()
    def forward_update(self, x: torch.Tensor, x_next: torch.Tensor, target: "TrainableConvQ", use_double: bool = True):
        """Every forward of one update in one launch: -> (Q(s), saved, Q_online(s') | None, Q_target(s'))."""
        return conv_q_forward_update(x.reshape(x.shape[0], 16), x_next.reshape(x_next.shape[0], 16), self.params(),
                                     target.params(), want_online_next=use_double)

    @torch.no_grad()
    def backward_into(self, saved, gq: torch.Tensor, grads) -> None:
        """Writes d loss / d parameter into `grads` (eight contiguous tensors in `params()` order),
        overwriting them — no zeroing and no accumulate kernels."""
        conv_q_backward(saved, self.params(), gq, out=list(grads))


def accelerate_inference(net: nn.Module):
    """The fastest no-gradient evaluator for `net`: the fused kernel K6 for the reference's conv
    Q-network, the tensor-core layers K8 for its dense one, otherwise `qnet.accelerate(net)`."""
    from . import qdense
    from .qnet import accelerate
    if matches(net):
        return FusedConvQ(net)
    return qdense.DenseQ(net) if qdense.matches(net) else accelerate(net)
