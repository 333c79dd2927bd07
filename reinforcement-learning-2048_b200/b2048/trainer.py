"""Double-DQN updates at scale, captured in one CUDA graph: GPU replay sampling (K2), the float64
Q-network forwards (conv config: ONE launch of the fused kernel K6 for Q(s) with saved activations, Q_online(s')
and Q_target(s'); dense config: K8; otherwise torch), fused target / loss (K3), backward (conv: K7 weight gradients +
fused input-gradient kernels; dense: K8; no library GEMM in either), gradient exchange + Adam
(one GPU: fused Adam kernel; several: the NVLink peer-memory kernel K5, or NCCL).

`DDQNUpdater.update()` is the "real" update (zero_grad -> backward -> allreduce -> step); the
reference's train_step order, which never changes the weights, is what the drop-in
`dqn_lib.train_step` reproduces by default.
"""
from __future__ import annotations

import copy

import torch

from . import ddqn, dist as bdist
from .qdense import DenseQ
from .qfused import FusedConvQ, TrainableConvQ, accelerate_inference
from .qnet import accelerate
from .replay import ReplayRing


class DDQNUpdater:
    def __init__(self, model: torch.nn.Module, ring: ReplayRing, batch_size: int = 5000, gamma: float = 0.8,
                 lr: float = 1e-2, use_double: bool = True, conv: bool = True, target_model=None,
                 use_graph: bool = True, seed: int = 2051, exchange: str = "p2p"):
        self.model, self.ring = model, ring
        self.target = target_model if target_model is not None else copy.deepcopy(model)
        for p in self.target.parameters():
            p.requires_grad_(False)
        self.B, self.gamma, self.use_double, self.conv, self.seed = int(batch_size), float(gamma), use_double, conv, seed
        self.device = next(model.parameters()).device
        # same parameter tensors, convolutions evaluated as float64 GEMMs (see qnet.py)
        self.f_model, self.f_target = accelerate(self.model), accelerate(self.target)
        # no-gradient evaluators (Q(s') here, action selection in the rollout): the fused kernel K6 for
        # the reference's conv network, else the same GEMM path
        self.i_model, self.i_target = accelerate_inference(self.model), accelerate_inference(self.target)
        bdist.broadcast_module(self.model)
        bdist.broadcast_module(self.target)
        self.params = bdist.FlatParams(model)        # parameters and gradients as two flat buffers:
        world = bdist.world()[1]
        use_p2p = world > 1 and exchange == "p2p" and self.device.type == "cuda"
        self.grads = bdist.FlatGrads(model, tail=2 * world if use_p2p else 0)   # one allreduce, one fused Adam kernel
        self.opt = ddqn.FusedAdam(self.params.flat, self.grads.flat, lr=lr)
        self.side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.side2 = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        # gradient exchange: "p2p" = fused NVLink allreduce+Adam kernel, "nccl" = all_reduce then Adam
        self.exchange = None
        if use_p2p:
            from .p2p import PeerGradExchange
            self.exchange = PeerGradExchange(self.grads)
        kw = dict(device=self.device)
        B = self.B
        self.batch = (torch.empty((B, 16), dtype=torch.float64, **kw), torch.empty(B, dtype=torch.int64, **kw),
                      torch.empty(B, dtype=torch.int64, **kw), torch.empty((B, 16), dtype=torch.float64, **kw),
                      torch.empty(B, dtype=torch.int64, **kw))
        self.loss = torch.zeros((), dtype=torch.float64, **kw)
        self.updates = 0
        self.graph = None
        self.use_graph = use_graph and self.device.type == "cuda"

    def _shape(self, x):
        return x.view(self.B, 1, 4, 4) if self.conv else x

    def _infer(self, net, x):
        return net(x) if isinstance(net, (FusedConvQ, DenseQ)) else net(self._shape(x))

    def _update_eager(self):
        states, actions, rewards, next_states, dones = self.ring.sample(self.B, seed=self.seed, ctr=ReplayRing.CTR_AUTO,
                                                                        out=self.batch)
        main = torch.cuda.current_stream(self.device)
        direct = hasattr(self.f_model, "forward_saving")       # conv (K6/K7) and dense (K8) configs: no autograd graph at all
        if hasattr(self.f_model, "forward_update") and type(self.f_target) is type(self.f_model) \
                and states.is_contiguous() and next_states.is_contiguous():
            # conv config: Q(s) (+ saved activations), Q_online(s') and Q_target(s') are ONE K6 launch
            q_cur, saved, q_next_online, q_next_target = self.f_model.forward_update(states, next_states, self.f_target,
                                                                                     self.use_double)
        else:
            # the two no-grad forwards run on side streams next to the training forward (the fork/join is
            # captured into the CUDA graph)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side), torch.no_grad():
                q_next_target = self._infer(self.i_target, next_states)
            q_next_online = None
            if self.use_double:      # its own stream: a short launch leaves most SMs free after its first round
                self.side2.wait_stream(main)
                with torch.cuda.stream(self.side2), torch.no_grad():
                    q_next_online = self._infer(self.i_model, next_states)
            if direct:
                q_cur, saved = self.f_model.forward_saving(states)
            else:
                q_cur = self.f_model(self._shape(states))
            main.wait_stream(self.side)
            if self.use_double:
                main.wait_stream(self.side2)
            if not torch.cuda.is_current_stream_capturing():      # eager mode: tell the allocator about the hand-over
                q_next_target.record_stream(main)
                if q_next_online is not None:
                    q_next_online.record_stream(main)
        if direct:
            # K3 also emits d loss / d Q(s,.); K7 writes every parameter gradient straight into the flat
            # buffer (overwrite: no zeroing, no per-parameter accumulate kernels)
            loss, _, _, grad_q = ddqn.ddqn_target_loss(q_next_online, q_next_target, q_cur, actions, rewards, dones,
                                                       self.gamma, self.use_double, want_grad=True)
            # the loss leaves for its reporting slot next to the backward pass, not behind Adam
            self.side2.wait_stream(main)
            with torch.cuda.stream(self.side2):
                self.loss.copy_(loss.detach().reshape(()))
            self.f_model.backward_into(saved, grad_q, [p.grad for p in self.grads.params])
        else:
            loss, _, _ = ddqn.ddqn_loss(q_cur, q_next_online, q_next_target, actions, rewards, dones, self.gamma,
                                        self.use_double)
            self.grads.zero_()
            loss.backward()
        if self.exchange is not None:    # sum over ranks == gradient of the summed loss over the global batch
            self.exchange.allreduce_adam(self.opt)
        else:
            self.grads.allreduce_()
            self.opt.step()
        if direct:
            main.wait_stream(self.side2)
        else:
            self.loss.copy_(loss.detach().reshape(()))

    def update(self) -> torch.Tensor:
        """One update; returns the (device) loss tensor of this rank's batch."""
        if not self.use_graph:
            self._update_eager()
        else:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        self.updates += 1
        return self.loss

    def _trainable_state(self):
        """Everything one `_update_eager()` changes: weights, Adam moments and step, the ring's
        sample counter, the reported loss."""
        return (self.params.flat, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.step_count,
                self.ring.head_size, self.loss)

    def _capture(self):
        """Warm-up (allocator, lazy kernel configuration) + graph capture, WITHOUT side effects: the three warm-up
        updates are real ones, so the state they touch is saved before and put back after — the first
        `update()` in graph mode then applies exactly one optimizer step, and a resumed run continues
        bit-identically (with several ranks every rank runs the same warm-up exchanges, so the
        replicas stay in step)."""
        saved = [t.clone() for t in self._trainable_state()]
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._update_eager()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._update_eager()                 # capture only records; nothing executes
        with torch.no_grad():
            for t, v in zip(self._trainable_state(), saved):
                t.copy_(v)
        torch.cuda.synchronize(self.device)

    def check(self) -> None:
        """Surface a failed gradient exchange (synchronises; call where a scalar is read back anyway)."""
        if self.exchange is not None:
            self.exchange.check()

    def sync_target(self) -> None:
        """target <- online (the reference's load_state_dict(deepcopy(...)), src/dqn_lib.py:227-228);
        a local copy on every rank: the weights are already identical after the allreduce."""
        with torch.no_grad():
            for t, p in zip(self.target.parameters(), self.model.parameters()):
                t.copy_(p)
            for t, p in zip(self.target.buffers(), self.model.buffers()):
                t.copy_(p)
