"""TEST INFRASTRUCTURE ONLY -- records tests/golden/mcts_search_v1.json.gz from the LIVE,
UNMODIFIED reference MCTS (mcts.py) driven by the keyed random stream of oracle/mcts_oracle.py.
Run in the build container:  python -m oracle.make_golden_mcts

Each case: a root reached by replaying `prefix` [(a, b, coin)...] from the empty board, then
`stages`: contemplate `rollouts` rollouts with `num_simulations` playouts each, record the
root's N / Q / Ntot and choose(); then play `move` = (action, coin) on the board, sync, and
continue with the next stage (mcts.py:294-337, strat_eval.py:34-63 usage pattern)."""
from __future__ import annotations

import random

from . import mcts_oracle as MO
from . import qttt_oracle as O
from .make_golden import _dump
from .refload import load_reference


def main():
    ns = load_reference()
    rng = random.Random(2026)
    cases = []
    for case_id in range(40):
        board = ns.qtttgym.Board(ns.qtttgym.QEvalClassic())
        g = O.Game()
        prefix = []
        for _ in range(rng.choice([0, 0, 1, 2, 3, 4, 5])):
            if g.terminal():
                break
            act = rng.choice(g.legal_actions())
            a, b = O.PAIRS[act]
            c = rng.randrange(2)
            trial = g.clone()
            trial.place(a, b, lambda: c)
            if trial.terminal():
                break
            g = trial
            ns.coin.bits.clear(); ns.coin.feed(c)
            board.make_move((a, b))
            prefix.append([a, b, c])
        seed, sims = 77000 + case_id, rng.choice([1, 4, 10, 32])
        rollouts = rng.choice([20, 50, 120])
        ref, fake = MO.shim_reference_mcts(ns, seed, case_id, rollouts, sims)
        ref.reset(board)
        ref.root.qstructs = [set(c) for c in board.qstructs]   # reset() forgets them
        stages = []
        for stage in range(3):
            for _ in range(rollouts):
                ref._rollout()
            root = ref.root
            rec = {"rollouts": rollouts,
                   "N": [int(root.N.get(a, 0)) for a in range(36)],
                   "Q": [float(root.Q.get(a, 0.0)) for a in range(36)],
                   "Ntot": int(root.Ntot), "choose": int(ref.choose()), "move": None}
            stages.append(rec)
            if root.terminal:
                break
            act = rng.choice(list(root.actions))
            a, b = O.PAIRS[act]
            c = rng.randrange(2)
            ns.coin.bits.clear(); ns.coin.feed(c)
            board.make_move((a, b))
            g.place(a, b, lambda: c)
            rec["move"] = [act, c]
            # sync needs self.game (set by Strategy.reset) to be the live board
            ref.sync(act)
            if ref.root.terminal:
                break
        cases.append({"prefix": prefix, "seed": seed, "root_index": case_id, "num_simulations": sims,
                      "stages": stages})
    _dump("mcts_search_v1.json.gz", cases)


if __name__ == "__main__":
    main()
