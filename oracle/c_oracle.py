"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy face of the C oracle (oracle/qttt_oracle.c)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

GAME_BYTES = 40
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        assert _lib.orc_sizeof_game() == GAME_BYTES
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Games:
    """n reference-shaped games held in one numpy byte array (orc_game structs)."""

    def __init__(self, n: int):
        self.n = int(n)
        self.raw = np.zeros((self.n, GAME_BYTES), dtype=np.uint8)
        lib().orc_reset_batch(_p(self.raw), C.c_int64(self.n))

    def copy(self) -> "Games":
        g = Games.__new__(Games)
        g.n, g.raw = self.n, self.raw.copy()
        return g

    def select(self, idx) -> "Games":
        g = Games.__new__(Games)
        g.raw = np.ascontiguousarray(self.raw[idx])
        g.n = g.raw.shape[0]
        return g

    @staticmethod
    def from_arrays(classical, moves, n_moves) -> "Games":
        classical = np.ascontiguousarray(classical, dtype=np.int8)
        moves = np.ascontiguousarray(moves, dtype=np.int8)
        n_moves = np.ascontiguousarray(n_moves, dtype=np.uint8)
        g = Games(classical.shape[0])
        lib().orc_from_arrays(_p(g.raw), C.c_int64(g.n), _p(classical), _p(moves), _p(n_moves))
        return g

    def step(self, action_pairs, coin=None):
        """env.py:34-53 batched.  action_pairs int8[n,2]; coin uint8[n] or None (=0)."""
        n = self.n
        ap = np.ascontiguousarray(action_pairs, dtype=np.int8).reshape(n, 2)
        co = None if coin is None else np.ascontiguousarray(coin, dtype=np.uint8)
        reward = np.empty(n, np.float32)
        done = np.empty(n, np.uint8)
        mask = np.empty(n, np.uint64)
        status = np.empty(n, np.uint8)
        collapsed = np.empty(n, np.uint8)
        lib().orc_step_batch(_p(self.raw), C.c_int64(n), _p(ap), _p(co), _p(reward), _p(done),
                             _p(mask), _p(status), _p(collapsed))
        return dict(reward=reward, done=done, mask=mask, status=status, collapsed=collapsed)

    def observe(self):
        n = self.n
        out = dict(classical=np.empty((n, 9), np.int8), moves=np.empty((n, 9, 2), np.int8),
                   n_moves=np.empty(n, np.uint8), q_p1=np.empty((n, 5, 2), np.int8),
                   q_p2=np.empty((n, 4, 2), np.int8), turn=np.empty(n, np.uint8),
                   rounds=np.empty((n, 2), np.int8), reward_p1=np.empty(n, np.float32),
                   winner=np.empty(n, np.uint8))
        lib().orc_observe_batch(_p(self.raw), C.c_int64(n), _p(out["classical"]), _p(out["moves"]),
                                _p(out["n_moves"]), _p(out["q_p1"]), _p(out["q_p2"]),
                                _p(out["turn"]), _p(out["rounds"]), _p(out["reward_p1"]),
                                _p(out["winner"]))
        return out

    def legal_mask(self):
        return self.step(np.full((self.n, 2), -1, np.int8))["mask"]   # illegal => no-op

    def qeval_both(self, action_idx):
        n = self.n
        act = np.ascontiguousarray(action_idx, dtype=np.uint8)
        out0 = np.empty(n, np.uint64); out1 = np.empty(n, np.uint64)
        closes = np.empty(n, np.uint8)
        sq0 = np.empty((n, 9), np.int8); sq1 = np.empty((n, 9), np.int8)
        lib().orc_qeval_both(_p(self.raw), C.c_int64(n), _p(act), _p(out0), _p(out1), _p(closes),
                             _p(sq0), _p(sq1))
        return dict(out0=out0, out1=out1, closes=closes, sq0=sq0, sq1=sq1)

    def rollout(self, n_rollouts: int, seed: int):
        tallies = np.empty((self.n, 3), np.int32)
        steps = C.c_int64(0)
        lib().orc_rollout(_p(self.raw), C.c_int64(self.n), C.c_int32(n_rollouts),
                          C.c_uint64(seed), _p(tallies), C.byref(steps))
        return tallies, steps.value


def selfplay(lo: int, hi: int, seed: int):
    stats = np.zeros(6, np.int64)
    hist = np.zeros(10, np.int64)
    lib().orc_selfplay(C.c_int64(lo), C.c_int64(hi), C.c_uint64(seed), _p(stats), _p(hist))
    return stats, hist


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(o))
    return tuple(int(x) for x in o)


def playout_trace(games: Games, i: int, seed: int, gid: int, domain: int):
    g = games.raw[i].copy()
    acts = np.zeros(9, np.uint8); coins = np.zeros(9, np.uint8); n = C.c_int(0)
    lib().orc_playout_trace.restype = C.c_int
    w = lib().orc_playout_trace(_p(g), C.c_uint64(seed), C.c_uint64(gid), C.c_uint32(domain),
                                _p(acts), _p(coins), C.byref(n))
    return w, acts[:n.value].copy(), coins[:n.value].copy(), g


def num_threads() -> int:
    return lib().orc_num_threads()
