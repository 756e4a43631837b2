"""TEST INFRASTRUCTURE ONLY -- trace generation and replay harness shared by the oracle tests
and ``oracle/make_golden.py``.

A *trace* is the input side of a game: ``[(a, b, coin), ...]`` -- the pair passed to
``Env.step`` and the coin bit that is consumed only if that step collapses.  A *record* is
everything observable after that step.  The same trace can be replayed through the live
reference (``replay_reference``; build container only) or the oracle (``replay_oracle``).
"""
from __future__ import annotations

import random
import struct

from . import qttt_oracle as O


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def random_trace(rng: random.Random, *, illegal_rate: float = 0.0, overrun: int = 0):
    """Plays one random game on the oracle to obtain a legal trace, optionally injecting
    illegal actions (Q2) and ``overrun`` extra steps after termination (Q3)."""
    g = O.Game()
    trace = []
    extra = overrun
    while True:
        term = g.terminal()
        if term:
            if extra == 0:
                break
            extra -= 1
        if rng.random() < illegal_rate or (term and not g.legal_actions()):
            kind = rng.randrange(4)
            if kind == 0:
                a = b = rng.randrange(9)                   # same square
            elif kind == 1:
                a, b = rng.randrange(9, 16), rng.randrange(0, 16)   # off board
            elif kind == 2:
                a, b = rng.randrange(0, 9), rng.randrange(9, 16)
            else:                                          # classical square if any
                cl = [s for s in range(9) if g.board[s] != -1]
                a = rng.choice(cl) if cl else rng.randrange(9)
                b = rng.randrange(9)
                if not cl:
                    b = a
                if rng.random() < 0.5:
                    a, b = b, a
            coin = rng.randrange(2)
            trace.append((a, b, coin))
            if g.is_legal(a, b):   # can happen for kind 3 with no classical squares? no: b=a
                g.place(a, b, lambda: coin)
            continue
        acts = g.legal_actions()
        if not acts:
            break
        a, b = O.PAIRS[rng.choice(acts)]
        if rng.random() < 0.5:
            a, b = b, a                                    # order must not matter
        coin = rng.randrange(2)
        trace.append((a, b, coin))
        g.place(a, b, lambda: coin)
    return trace


def _record(board, moves, comps, obs, r, term, mask, rounds, reward_p1, winner):
    return {
        "board": list(board),
        "moves": [list(m) for m in moves],
        "comps": [sorted(c) for c in comps],
        "q1": [list(p) for p in obs["q_states_p1"]],
        "q2": [list(p) for p in obs["q_states_p2"]],
        "turn": obs["turn"],
        "reward_bits": f32_bits(r),
        "terminated": bool(term),
        "mask": int(mask),
        "rounds": list(rounds),
        "reward_p1": float(reward_p1),
        "winner": int(winner),
    }


def replay_oracle(trace):
    env = O.Env()
    env.reset()
    out = []
    for a, b, coin in trace:
        obs, r, term, trunc, info = env.step((a, b), coin=lambda: coin)
        g = env.game
        assert trunc is False and info == {}
        out.append(_record(g.board, g.moves, g.comps, obs, r, term, g.legal_mask(),
                           g.win_rounds(), g.reward_p1(), g.winner()))
    return out


def replay_reference(trace, ns):
    """Same through the unmodified reference (``ns = refload.load_reference()``)."""
    env = ns.qtttgym.Env()
    env.reset()
    out = []
    for a, b, coin in trace:
        ns.coin.bits.clear()
        ns.coin.feed(coin)
        obs, r, term, trunc, info = env.step((a, b))
        assert trunc is False and info == {}
        bd = env._gameboard
        # legal mask per mcts.py:19-27 / 87-91 on the live board
        gs = ns.mcts.MCTS.GameState(bd.board, bd.moves, True, None, False)
        mask = 0
        for k, v in enumerate(gs.action_mask()):
            if v:
                mask |= 1 << k
        # winner per strat_eval.py:21-32 / mcts.py:52-65
        gs.update_winner()
        winner = 0 if gs.winner is None else (1 if gs.winner else 2)
        out.append(_record(bd.board, bd.moves, bd.qstructs, obs, r, term, mask,
                           bd.check_win(), env._reward(), winner))
    return out
