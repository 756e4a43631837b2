"""TEST INFRASTRUCTURE ONLY -- records tests/golden/simulate_freq_v1.json: win/draw counts of the
UNMODIFIED reference's MCTS._simulate (mcts.py:185-208) with its OWN randomness (numpy's global
RNG for the playout policy, stdlib MT19937 for the collapse coin) from a set of roots reached by
replay.  Config 4's statistical bridge: the Philox-driven rollout kernels must reproduce these
frequencies within binomial confidence.  Run in the build container:
    python -m oracle.make_golden_simulate"""
from __future__ import annotations

import json
import os
import random

from . import qttt_oracle as O
from .make_golden import OUT
from .refload import load_reference


def main():
    ns = load_reference()
    ref = ns.mcts
    import numpy as real_np
    ref.np = real_np                                  # undo any keyed stand-in
    ns.qeval_module.random = ns.real_random           # the reference's own coin
    try:
        real_np.random.seed(20261018)
        ns.real_random.seed(20261018)
        rng = random.Random(99)
        cases = []
        playouts = 600
        for case_id in range(24):
            board = ns.qtttgym.Board(ns.qtttgym.QEvalClassic())
            g = O.Game()
            prefix = []
            for _ in range(case_id % 7):
                act = rng.choice(g.legal_actions())
                a, b = O.PAIRS[act]
                c = rng.randrange(2)
                trial = g.clone()
                trial.place(a, b, lambda: c)
                if trial.terminal():
                    break
                g = trial
                # replay on the reference with the same coin: force it for this one move only
                ns.qeval_module.random = ns.coin
                ns.coin.bits.clear(); ns.coin.feed(c)
                board.make_move((a, b))
                ns.qeval_module.random = ns.real_random
                prefix.append([a, b, c])
            assert board.board == g.board and [tuple(m) for m in board.moves] == g.moves
            mc = ref.MCTS(rollouts=1, num_simulations=1)
            mc.reset(board)
            mc.root.qstructs = [set(c) for c in board.qstructs]
            counts = {1: 0, -1: 0, 0: 0}
            for _ in range(playouts):
                counts[mc._simulate(mc.root)] += 1
            cases.append({"prefix": prefix, "playouts": playouts, "x": counts[1], "o": counts[-1],
                          "draw": counts[0]})
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "simulate_freq_v1.json"), "w") as f:
            json.dump(cases, f, separators=(",", ":"))
        print(sum(c["x"] for c in cases), sum(c["o"] for c in cases), sum(c["draw"] for c in cases))
    finally:
        ns.qeval_module.random = ns.coin


if __name__ == "__main__":
    main()
