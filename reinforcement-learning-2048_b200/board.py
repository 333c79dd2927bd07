"""Drop-in replacement for the reference's ``board`` module (``src/board.py``) on the CUDA engine.

Put this directory first on ``sys.path`` and the reference's ``player.py``, ``double_dqn_conv.py``
and ``double_dqn_dense.py`` import it unchanged (``from board import Board2048``).  The public
surface — constructor, attributes (``state`` stays an assignable ``np.ndarray`` of tile values),
methods, return types, the ``ValueError`` for a bad action — is the reference's
(``src/board.py:8-240``); every slide/merge, spawn and legal-move test runs on the GPU through the
C-ABI library (``include/b2048.h``).  There is no CPU fallback: without a CUDA device the first
board operation raises.

Deliberate, documented differences:
  * only ``k == 4`` boards exist (packed 64-bit format); other ``k`` raise ``NotImplementedError``;
  * tiles must be 0 or powers of two up to 32768 (4-bit exponents);
  * ``clone()`` does not burn random numbers (the reference spawns two throw-away tiles, Q4);
  * spawns come from a counter-based Philox stream (``seed(n)`` to fix it), "4" with probability
    0.5 like the reference (``set_spawn_four_probability`` to change it, e.g. 0.1).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from b2048 import single as _single
from b2048.env import FLAG_CHANGED

_MOVES = ("up", "down", "left", "right")          # action index order, src/board.py:129,191


def seed(value: int) -> None:
    """Fix the spawn stream of all boards created through this module."""
    _single.engine().reseed(value)


def set_spawn_four_probability(p: float) -> None:
    """Probability that a spawned tile is a 4 (reference: 0.5, src/board.py:12,49)."""
    _single.engine().set_p_four(p)


class Board2048:
    """4x4 game board; every move returns a new board (src/board.py:8)."""

    def __init__(self, k: int = 4, populate_empty_cells=True):
        if k != 4:
            raise NotImplementedError("Board2048 on the CUDA engine supports k=4 only (packed 64-bit boards)")
        self.state: np.ndarray = np.zeros(shape=(k, k), dtype=int)
        self._empty_spot_numbers: List[int] = [2, 4]
        self._mergescore = 0
        self._action_history: List[str] = []
        self.k = k
        self.populate_empty_cells = populate_empty_cells
        if populate_empty_cells:                       # two starting tiles (src/board.py:18-20)
            self.state = _single.engine().fresh().astype(int)

    # -- bookkeeping -----------------------------------------------------------------------------
    def clone(self) -> "Board2048":
        other = Board2048(k=self.k, populate_empty_cells=False)
        other.populate_empty_cells = self.populate_empty_cells
        other.state = np.copy(self.state)
        other._mergescore = self._mergescore
        other._action_history = list(self._action_history)
        return other

    def __repr__(self):
        return str(self.state)

    def __contains__(self, element) -> bool:
        return bool(np.isin(element, self.state).all())

    def __eq__(self, other):
        return (self.state == other.state).all()

    __hash__ = None

    # -- primitives --------------------------------------------------------------------------------
    def _populate_empty_cell(self) -> "Board2048":
        """One new 2 or 4 in a uniformly random empty cell (src/board.py:41-51), in place."""
        self.state = _single.engine().spawn(self.state).astype(int)
        return self

    def _reverse_vector(self, vector):
        return np.flip(vector)

    def _apply_action_to_vector(self, vector) -> np.ndarray:
        """Slide/merge one length-4 vector toward index 0 (src/board.py:92-126); the merged tile
        values are added to this board's merge score like the reference does (:114)."""
        v = np.asarray(vector)
        if v.shape != (4,):
            raise NotImplementedError("rows have 4 cells on the CUDA engine")
        scratch = np.zeros((4, 4), dtype=np.int64)
        scratch[0] = v
        nxt, reward, _ = _single.engine().move(scratch, 2, spawn=False)      # 2 = left
        self._mergescore += reward
        return nxt[0].astype(v.dtype if np.issubdtype(v.dtype, np.integer) else int)

    def _moved(self, action: int) -> "Board2048":
        board = self.clone()
        board._action_history.append(_MOVES[action])
        nxt, reward, flags = _single.engine().move(self.state, action, spawn=True)
        if flags & FLAG_CHANGED:                       # unchanged boards keep their state object's values
            board.state = nxt.astype(int)
        if reward:
            board._mergescore = board._mergescore + np.int64(reward)
        return board

    # -- legal moves -----------------------------------------------------------------------------
    def available_moves_as_torch_unit_vector(self, device=None):
        """float32[4] on `device`, 1 where [up, down, left, right] changes the board (:128-135)."""
        flags = _single.engine().legal(self.state)
        unit_vector = torch.zeros(4, device=device)
        for i in range(4):
            if flags >> i & 1:
                unit_vector[i] = 1
        return unit_vector

    def available_moves(self) -> Dict[str, "Board2048"]:
        """{move name: successor board} for the moves that change the board (:138-145)."""
        nxt4, rew4, flags = _single.engine().all4(self.state)
        mapping = dict()
        for i, move in enumerate(_MOVES):
            if flags >> i & 1:
                b = self.clone()
                b._action_history.append(move)
                b.state = nxt4[i].astype(int)
                if rew4[i]:
                    b._mergescore = b._mergescore + np.int64(rew4[i])
                mapping[move] = b
        return mapping

    # -- the four moves (src/board.py:147-183) ---------------------------------------------------
    def up(self) -> "Board2048":
        return self._moved(0)

    def down(self) -> "Board2048":
        return self._moved(1)

    def left(self) -> "Board2048":
        return self._moved(2)

    def right(self) -> "Board2048":
        return self._moved(3)

    def peek_action(self, action) -> "Board2048":
        """The board after `action` (a move name, its first letter, an int or a 0-d tensor index
        into [up, down, left, right]); this board is left untouched (src/board.py:185-202)."""
        if type(action) is not str:
            action = "udlr"[int(action)]
        key = action.lower()[0]
        if key in "udlr":
            return self._moved("udlr".index(key))
        raise ValueError(f"Action: {key} is invalid.")

    # -- scores and views --------------------------------------------------------------------------
    def simple_score(self):
        return self.state.flatten().sum(axis=0)

    def merge_score(self):
        return self._mergescore

    def show(self, ignore_zeros=False):
        print(f"Simple Score: {self.simple_score()}")
        print(f"Merge Score: {self.merge_score()}")
        print(self.__repr__().replace("0", "_") if ignore_zeros else self)

    def normalized(self) -> "Board2048":
        out = self.clone()
        out.state = out.state / np.max(out.state)
        return out

    def log_scale(self) -> "Board2048":
        """Clone whose state holds tile exponents (0 for empty): the network input (:224-231)."""
        out = self.clone()
        s = np.asarray(out.state)
        out.state = np.where(s > 0, np.log2(np.where(s > 0, s, 1)).astype(s.dtype if s.dtype.kind in "iu" else int), 0)
        return out

    def flattened_state_as_tensor(self):
        return torch.from_numpy(np.ascontiguousarray(self.state).flatten()).double()

    def state_as_4d_tensor(self):
        return torch.from_numpy(np.ascontiguousarray(self.state)[np.newaxis][np.newaxis]).double()

    def number_of_empty_cells(self) -> int:
        return int((self.state == 0).sum())


def basic_updown_algorithm(k=4):
    """Up/left until stuck, then down/right; the reference's demo loop (src/board.py:244-261)."""
    board = Board2048(k=k)
    score = board.simple_score()
    while True:
        for first in ("up", "left"):
            board = board.peek_action(first)
            board.show(ignore_zeros=True)
        if score == board.simple_score():
            board = board.peek_action("down")
            board.show(ignore_zeros=True)
            board = board.peek_action("right")
            if score == board.simple_score():
                break
        board.show(ignore_zeros=True)
        score = board.simple_score()
    board.show()
    return board
