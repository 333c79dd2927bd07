"""Drop-in replacement for the reference's ``dqn_lib`` module (``src/dqn_lib.py``) on the CUDA engine.

Same entry points, argument order and return types as the reference (``src/dqn_lib.py:8-244``), so
``double_dqn_conv.py``, ``double_dqn_dense.py`` and ``player.py`` run unchanged:

  epsilon_greedy_policy   legal mask (kernel K1) + Q forward (conv net: fused kernel K6) + fused
                          Q-normalise/mask/argmax (kernel K0)
  play_one_step           Board2048.peek_action on the GPU + replay append
  sample_experiences      GPU ring: fused sample + gather + unpack to float64 (kernel K2)
  train_step              Q forwards in torch float64 + fused Double-DQN target / summed MSE (K3)
  training_loop           the reference's episode loop on top of the above

The reference's observable quirks are kept on purpose and documented in SURVEY.md §0: the
optimizer is stepped after ``zero_grad`` so the weights never move (Q1; set
``FIX_UPDATE_ORDER = True`` or env ``B2048_FIX_UPDATE_ORDER=1`` for a real update), gamma is applied
in float32 (Q2), ``done`` belongs to the pre-action board (Q5), random actions ignore legality
(Q6), the greedy branch normalises with ``Q - min*max - min`` (Q7).  The host-side random draws
(``np.random.rand``, ``np.random.randint``) are the reference's, so a seeded run takes the same
explore/exploit decisions and samples the same replay indices.
"""
from __future__ import annotations

import copy
import os
import weakref
from collections import deque
from typing import Callable

import numpy as np
import torch

from board import Board2048
from b2048 import ddqn as _ddqn
from b2048 import env as _env
from b2048 import qfused as _qfused
from b2048.qnet import accelerate as _accelerate
from b2048.replay import ReplayDeque

FIX_UPDATE_ORDER = os.environ.get("B2048_FIX_UPDATE_ORDER", "0") == "1"


def _cuda_device(device) -> torch.device:
    d = torch.device(device)
    if d.type == "cuda":
        return d
    if not torch.cuda.is_available():
        raise RuntimeError("dqn_lib on the CUDA engine needs a GPU: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _exponent_tensor(boards, shape, device) -> torch.Tensor:
    """Tile values of one or many boards -> float64 exponents via the GPU pack/unpack kernels."""
    dev = _cuda_device(device)
    tiles = np.stack([np.ascontiguousarray(b.state, dtype=np.int64).reshape(16) for b in boards])
    packed = _env.pack(torch.from_numpy(tiles).to(dev))
    out = _env.unpack_f64(packed).reshape(shape)
    return out if torch.device(device) == dev else out.to(device)


def board_as_4d_tensor(board: Board2048, device: str) -> torch.Tensor:
    """float64 [1,1,4,4] of tile exponents on `device` (src/dqn_lib.py:8-9)."""
    return _exponent_tensor([board], (1, 1, 4, 4), device)


def board_as_flattened_tensor(board: Board2048, device: str) -> torch.Tensor:
    """float64 [16] of tile exponents on `device` (src/dqn_lib.py:12-13)."""
    return _exponent_tensor([board], (16,), device)


_FUSED = weakref.WeakKeyDictionary()        # model -> FusedConvQ (or None when the model is something else)


def _fused_for(model):
    """FusedConvQ for the reference's conv Q-network on a CUDA device (cached per module), else None."""
    try:
        return _FUSED[model]
    except KeyError:
        fq = _FUSED[model] = _qfused.FusedConvQ(model) if _qfused.matches(model) else None
        return fq
    except TypeError:
        return None


def _fusable(x) -> bool:
    return x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == 16 * x.shape[0]


def _q_values(model, state, fallback=None):
    """model(state) where the reference never takes a gradient through the call (action selection,
    src/dqn_lib.py:24-25; Q(s') of the target, :126-128): the conv Q-network runs as the single fused
    kernel K6, any other model through `fallback` (default: the module itself)."""
    fq = _fused_for(model)
    if fq is not None and _fusable(state):
        return fq(state)
    return (fallback or model)(state)


def epsilon_greedy_policy(board, epsilon, model, device, board_to_tensor_function: Callable = board_as_4d_tensor):
    """-> (action, done, max_q) exactly as src/dqn_lib.py:16-30."""
    available_moves = board.available_moves_as_torch_unit_vector(device=device)
    done = torch.max(available_moves) == 0
    if np.random.rand() < epsilon:
        return np.random.randint(4), done, torch.zeros(size=(1,), device=device)
    state = board_to_tensor_function(board, device)
    q_values = _q_values(model, state)
    dev = _cuda_device(device)
    q = q_values.detach().reshape(1, 4).to(dev, torch.float64).contiguous()
    legal = (available_moves.to(dev) != 0).to(torch.uint8)
    flags = (legal * torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=dev)).sum().to(torch.uint8).reshape(1)
    action, _ = _ddqn.egreedy_select(q, flags, 0.0, override=torch.full((1,), 0x80, dtype=torch.uint8, device=dev))
    return int(action.item()), int(done), torch.max(q_values)


def _known_tensor_fn(fn) -> bool:
    return fn in (board_as_4d_tensor, board_as_flattened_tensor)


def _extract(batch_size, batch, device, actions, rewards, dones, board_to_tensor_function, conv: bool):
    for _, action, reward, _, done in batch:
        actions.append(action)
        rewards.append(reward)
        dones.append(int(done))
    shape = (len(batch), 1, 4, 4) if conv else (batch_size, 16)
    if _known_tensor_fn(board_to_tensor_function):
        states = _exponent_tensor([e[0] for e in batch], shape, device)
        next_states = _exponent_tensor([e[3] for e in batch], shape, device)
    else:  # a caller-supplied board->tensor function is honoured sample by sample, like the reference
        states = torch.stack([board_to_tensor_function(e[0], device).reshape(shape[1:]) for e in batch])
        next_states = torch.stack([board_to_tensor_function(e[3], device).reshape(shape[1:]) for e in batch])
    return states, actions, rewards, next_states, dones


def extract_samples_conv(batch_size, batch, device, actions, rewards, dones, board_to_tensor_function):
    """States as float64 [N,1,4,4]; appends to the caller's lists (src/dqn_lib.py:33-46)."""
    return _extract(batch_size, batch, device, actions, rewards, dones, board_to_tensor_function, True)


def extract_samples_dense(batch_size, batch, device, actions, rewards, dones, board_to_tensor_function):
    """States as float64 [batch_size,16] (src/dqn_lib.py:49-64)."""
    return _extract(batch_size, batch, device, actions, rewards, dones, board_to_tensor_function, False)


def sample_experiences(batch_size: int, replay_buffer, device: str, board_to_tensor_function: Callable,
                       extract_sample_function: Callable):
    """Uniform sampling with replacement (src/dqn_lib.py:67-84).  With the GPU-resident
    `ReplayDeque` that `training_loop` creates, the drawn indices go straight to the fused
    sample/gather/unpack kernel; a plain deque of 5-tuples is accepted too."""
    random_sample = np.random.randint(len(replay_buffer), size=batch_size)
    conv = extract_sample_function is not extract_samples_dense
    if (isinstance(replay_buffer, ReplayDeque) and _known_tensor_fn(board_to_tensor_function)
            and extract_sample_function in (extract_samples_conv, extract_samples_dense)):
        states, actions, rewards, next_states, dones = replay_buffer.sample(batch_size, indices=random_sample)
        if conv:
            states, next_states = states.view(batch_size, 1, 4, 4), next_states.view(batch_size, 1, 4, 4)
        if torch.device(device) != states.device:
            states, actions, rewards, next_states, dones = (t.to(device) for t in
                                                            (states, actions, rewards, next_states, dones))
        return states, actions, rewards, next_states, dones
    if isinstance(replay_buffer, ReplayDeque):
        raise TypeError("custom board_to_tensor/extract functions need a plain deque replay buffer")
    batch = [replay_buffer[index] for index in random_sample]
    actions, rewards, dones = [], [], []
    states, actions, rewards, next_states, dones = extract_sample_function(
        batch_size, batch, device, actions, rewards, dones, board_to_tensor_function)
    actions = torch.tensor(actions, device=device)
    rewards = torch.tensor(rewards, device=device)
    dones = torch.tensor(dones, device=device)
    return states, actions, rewards, next_states, dones


def reward_func_merge_score(board: Board2048, next_board: Board2048, action: int, done: int) -> int:
    return next_board.merge_score() - board.merge_score()


def play_one_step(board: Board2048, epsilon: float, model, replay_buffer, device: str,
                  reward_function: Callable = reward_func_merge_score,
                  board_to_tensor_function: Callable = board_as_4d_tensor):
    """One environment step + replay append (src/dqn_lib.py:91-107)."""
    action, done, max_q_value = epsilon_greedy_policy(board, epsilon=epsilon, model=model, device=device,
                                                      board_to_tensor_function=board_to_tensor_function)
    next_board = board.peek_action(action)
    reward = reward_function(board, next_board, action, done)
    replay_buffer.append((board, action, reward, next_board, done))
    return next_board, action, reward, done, max_q_value


def one_hot(tensor: torch.Tensor, no_outputs: int, device: str):
    """float32 one-hot rows (src/dqn_lib.py:110-116)."""
    assert tensor.max().item() + 1 <= no_outputs, \
        "One hot encoded array size has to be bigger or equal than max scalar value"
    assert len(tensor.shape) == 1, "should be 1D"
    return torch.zeros(tensor.shape[0], no_outputs, device=device).scatter_(1, tensor.reshape(-1, 1), 1.0)


def _is_sum_mse(loss_fn) -> bool:
    return isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "sum"


def train_step(batch_size: int, discount_factor, model, target_model, replay_buffer, loss_fn: Callable,
               optimizer, device: str, use_double_dqn: bool = True,
               board_to_tensor_function: Callable = board_as_4d_tensor,
               extract_samples_function: Callable = extract_samples_conv):
    """One (Double-)DQN update on a sampled batch; returns the loss tensor (src/dqn_lib.py:119-164)."""
    states, actions, rewards, next_states, dones = sample_experiences(
        batch_size, replay_buffer, device, board_to_tensor_function, extract_samples_function)
    dev = _cuda_device(device)
    on_dev = states.device == dev
    # same parameters; 2x2 convolutions are evaluated as float64 GEMMs (b2048/qnet.py)
    f_model, f_target = _accelerate(model), _accelerate(target_model)
    # Q(s') enters the loss as a constant: no autograd graph needed for it
    q_next_target = _q_values(target_model, next_states, f_target)
    q_next_online = _q_values(model, next_states, f_model) if use_double_dqn else None
    q_cur = f_model(states)
    if on_dev and _is_sum_mse(loss_fn) and q_cur.dtype == torch.float64:
        loss, _, _ = _ddqn.ddqn_loss(q_cur, q_next_online, q_next_target, actions, rewards, dones,
                                     discount_factor, use_double_dqn)
    else:
        # any other loss / device: targets and Q(s,a) still come from the fused kernel
        k = lambda t: None if t is None else t.detach().to(dev, torch.float64).contiguous()  # noqa: E731
        _, target, _, _ = _ddqn.ddqn_target_loss(k(q_next_online), k(q_next_target), k(q_cur), actions.to(dev),
                                                 rewards.to(dev), dones.to(dev), discount_factor, use_double_dqn,
                                                 want_grad=False)
        q_values = q_cur.gather(1, actions.reshape(-1, 1)).reshape(-1)
        loss = loss_fn(q_values, target.to(q_values.device))
    if FIX_UPDATE_ORDER:
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    else:          # the reference's order: the step sees zeroed gradients
        loss.backward()
        optimizer.zero_grad()
        optimizer.step()
    return loss


def _make_replay_buffer(length, override, gpu_ok: bool, device):
    """deque(maxlen) of the reference (src/dqn_lib.py:169-172), GPU-resident whenever the stock
    board->tensor / extract functions are in use; an A*-seeded override deque is ingested."""
    if override:
        if gpu_ok and not isinstance(override, ReplayDeque):
            return ReplayDeque(override, maxlen=override.maxlen or length, device=_cuda_device(device))
        return override
    if gpu_ok:
        return ReplayDeque(maxlen=length, device=_cuda_device(device))
    return deque(maxlen=length)


def _epsilon_schedule(ep, ramp_episodes, min_epsilon, warm_episodes):
    """Linear decay max((E - ep) / E, min_eps); 0 while filling an existing model's buffer
    (src/dqn_lib.py:184-188)."""
    if ep < warm_episodes:
        return 0
    return max((ramp_episodes - ep) / ramp_episodes, min_epsilon)


def _play_episode(epsilon, model, replay_buffer, reward_function, board_to_tensor_function, device):
    """One game from a fresh board until the dead-board transition (src/dqn_lib.py:176-205)."""
    board, done = Board2048(), False
    history, rewards, q_values = [], [], []
    while not done:
        nxt, action, reward, done, max_q = play_one_step(
            board, epsilon, model, replay_buffer, reward_function=reward_function,
            board_to_tensor_function=board_to_tensor_function, device=device)
        history.append((board.state, "udlr"[int(action)], reward))
        rewards.append(reward)
        q_values.append(float(max_q))
        board = nxt
    return board, history, rewards, q_values


def training_loop(replay_buffer_length, no_episodes, no_episodes_to_reach_epsilon,
                  no_episodes_to_fill_up_existing_model_replay_buffer, min_epsilon, model, reward_function,
                  board_to_tensor_function, device, experiment, snapshot_game_every_n_episodes,
                  no_episodes_before_training, batch_size, discount_factor, target_model, loss_fn, optimizer,
                  use_double_dqn, no_episodes_before_updating_target, extract_samples_function,
                  replay_buffer_override=None):
    """The reference's episode loop (src/dqn_lib.py:167-244) with its exact positional signature:
    one train_step per finished episode, a target sync every `no_episodes_before_updating_target`
    episodes, the same console lines and the same Experiment calls; an exception saves the
    experiment and is re-raised, Ctrl-C saves and returns."""
    warm = no_episodes_to_fill_up_existing_model_replay_buffer
    stock = (_known_tensor_fn(board_to_tensor_function)
             and extract_samples_function in (extract_samples_conv, extract_samples_dense))
    try:
        replay_buffer = _make_replay_buffer(replay_buffer_length, replay_buffer_override, stock, device)
        for ep in range(no_episodes):
            print(ep)
            epsilon = _epsilon_schedule(ep, no_episodes_to_reach_epsilon, min_epsilon, warm)
            board, history, rewards, q_values = _play_episode(
                epsilon, model, replay_buffer, reward_function, board_to_tensor_function, device)
            experiment.add_episode(board, epsilon, ep, np.mean(np.array(rewards)), np.mean(np.array(q_values)))
            if ep % snapshot_game_every_n_episodes == 0:
                experiment.snapshot_game(history, ep)
            if ep % 10 == 0:
                print(f"Episode: {ep}: {board.merge_score()}, {np.max(board.state.flatten())}, "
                      f"{len(board._action_history)}")
            if ep > no_episodes_before_training:
                train_step(batch_size, discount_factor, model, target_model, replay_buffer, loss_fn, optimizer,
                           device=device, use_double_dqn=use_double_dqn,
                           board_to_tensor_function=board_to_tensor_function,
                           extract_samples_function=extract_samples_function)
            if ep % no_episodes_before_updating_target == 0 and ep >= warm:
                target_model.load_state_dict(copy.deepcopy(model.state_dict()))
            if ep % 1000 == 0:
                experiment.save()
                print("Saved game")
        experiment.save()
    except KeyboardInterrupt as stop:
        print(stop)
        print(f"\nKeyboard interrut caught. Saving current experiment in {experiment.folder}")
        experiment.save()
    except Exception:
        experiment.save()
        print(f"\nSaving current experiment in {experiment.folder}\n")
        raise
