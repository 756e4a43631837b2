"""The measurement seam on the GPU (reference: qtttgym/qeval.py:5-51, called from
qtttgym/board.py:51; MCTS._step's both-outcomes enumeration mcts.py:233-267).

``qeval_both`` is the batched form.  ``QEvalB200`` is the duck-typed plugin: an object with
``eval(entangled_moves) -> list[int]`` that can be handed to the reference's own
``qtttgym.Board(qevaluator)`` (board.py:2,7) in place of ``QEvalClassic``.
"""
from __future__ import annotations

import random as _random

import torch

from . import _lib
from .actions import move2ind
from .env import _stream_ptr, pack_states


def qeval_both(state, actions, *, want_states=True, want_boards=True, want_squares=False,
               want_probs=True, out=None):
    """Both collapse outcomes of ``actions`` (uint8[N], 0..35) applied to packed ``state``
    (int32[N,4]).  Returns a dict with

    next0/next1  int32[N,4] successor states for coin 0 / 1 (equal when no cycle closes)
    board0/board1 int64[N]  successor boards, 4 bits per square, value ``board[s] + 1``
    sq0/sq1      int8[N,9]  square each move index collapses into (-1: not in the measurement)
    closes       uint8[N]   1 when the action closes a cycle
    result_prob  f32[N,3]   P(X wins), P(O wins), P(neither) over the two equiprobable outcomes

    ``out``: a dict returned by an earlier call with the same shapes; its tensors are reused
    (no allocation, so the call can be captured in a CUDA graph).
    """
    lib = _lib.lib()
    dev, n = state.device, state.shape[0]
    actions = actions.to(torch.uint8).contiguous()
    assert actions.device == dev and actions.shape == (n,)
    if out is not None:
        g = lambda k: _lib.ptr(out.get(k))   # noqa: E731
        with torch.cuda.device(dev):
            _lib.check(lib.qttt_qeval_both(state.data_ptr(), actions.data_ptr(), g("next0"), g("next1"),
                                           g("board0"), g("board1"), g("sq0"), g("sq1"),
                                           out["closes"].data_ptr(), g("result_prob"), n,
                                           _stream_ptr(dev)))
        return out
    out = {"closes": torch.empty(n, dtype=torch.uint8, device=dev)}
    if want_states:
        out["next0"] = torch.empty_like(state)
        out["next1"] = torch.empty_like(state)
    if want_boards:
        out["board0"] = torch.empty(n, dtype=torch.int64, device=dev)
        out["board1"] = torch.empty(n, dtype=torch.int64, device=dev)
    if want_squares:
        out["sq0"] = torch.empty((n, 9), dtype=torch.int8, device=dev)
        out["sq1"] = torch.empty((n, 9), dtype=torch.int8, device=dev)
    if want_probs:
        out["result_prob"] = torch.empty((n, 3), dtype=torch.float32, device=dev)
    g = lambda k: _lib.ptr(out.get(k))   # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_qeval_both(state.data_ptr(), actions.data_ptr(), g("next0"), g("next1"),
                                       g("board0"), g("board1"), g("sq0"), g("sq1"),
                                       out["closes"].data_ptr(), g("result_prob"), n,
                                       _stream_ptr(dev)))
    return out


def square_probabilities(sq0, sq1):
    """P[n, k, s] = probability that move k of the measured component collapses into square s
    (values in {0, 1/2, 1}; SURVEY section 8(a) 'derived quantity for config 3')."""
    n = sq0.shape[0]
    p = torch.zeros((n, 9, 9), dtype=torch.float32, device=sq0.device)
    for sq in (sq0, sq1):
        valid = sq >= 0
        idx = sq.clamp(min=0).long().unsqueeze(-1)
        p.scatter_add_(2, idx, valid.unsqueeze(-1).float() * 0.5)
    return p


class QEvalB200:
    """Drop-in for ``qtttgym.QEvalClassic`` at the plugin seam (board.py:51).

    ``eval(entangled_moves)`` takes the component's ``(a, b, idx)`` moves in idx order, the last
    one having closed the cycle, and returns the square each collapses into.  The coin is one
    call to ``rng.choice((0, 1))`` -- the stdlib ``random`` module by default, which is what
    the reference consumes (qeval.py:35) -- or a forced bit via ``force``.

    One call is one kernel launch (``qttt_qeval1``): the component is packed on the host (a few
    shifts), travels as kernel arguments, and both outcomes come back through mapped pinned host
    memory; the host spins on the record's sequence word.  No tensors are created per call.
    """

    def __init__(self, device="cuda", rng=None):
        import ctypes as C
        self.lib = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("qtttgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rng = rng if rng is not None else _random
        self.forced: list[int] = []
        self._host = torch.zeros(32, dtype=torch.uint8).pin_memory()
        self._np = self._host.numpy()
        self._seq_view = self._np[28:32].view("uint32")
        self._mv = memoryview(self._np).cast("b")
        self._state = (C.c_uint32 * 4)()
        self._seq = 0

    def force(self, *bits):
        self.forced.extend(int(b) & 1 for b in bits)

    def eval(self, entangled_moves):
        k = len(entangled_moves)
        if not 2 <= k <= 9:
            raise ValueError("a measured component has 2..9 moves")
        # pack the k - 1 earlier moves as an all-quantum position (csrc/qttt_core.cuh: E fields of 9
        # bits, three per word, len(moves) in bits 27..30 of word 0); the last move is the action
        words = [0, 0, 0, 0]
        for r, m in enumerate(entangled_moves[:-1]):
            a, b = int(m[0]), int(m[1])
            if not (0 <= a < 9 and 0 <= b < 9 and a != b):
                raise ValueError("moves are pairs of distinct squares 0..8")
            words[r // 3] |= ((1 << a) | (1 << b)) << (9 * (r % 3))
        words[0] |= (k - 1) << 27
        st = self._state
        st[0], st[1], st[2], st[3] = words
        last = entangled_moves[-1]
        self._seq = seq = (self._seq % 0xFFFFFFFE) + 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_qeval1(st, move2ind(int(last[0]), int(last[1])), self._host.data_ptr(), seq,
                                            _stream_ptr(self.device)))
        view, spins = self._seq_view, 0
        while view[0] != seq:
            spins += 1
            if spins > 5_000_000:
                torch.cuda.synchronize(self.device)
                if view[0] != seq:
                    raise RuntimeError("qttt_qeval1: the record never arrived")
        mv = self._mv
        if mv[18] != 1:
            raise ValueError("the last move does not close a cycle in this component")
        coin = self.forced.pop(0) if self.forced else self.rng.choice((0, 1))
        off = 9 if coin else 0
        return list(mv[off:off + k])
