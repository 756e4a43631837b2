"""The MCTS restatement (oracle/mcts_oracle.py) against search statistics recorded from the
live reference MCTS (tests/golden/mcts_search_v1.json.gz), and live when the reference is here."""
import pytest

from oracle import mcts_oracle as MO
from oracle import qttt_oracle as O
from oracle.refload import reference_available

from helpers import load_golden


def replay_case(case, make_searcher):
    g = O.Game()
    for a, b, c in case["prefix"]:
        g.place(a, b, lambda: c)
    s = make_searcher(case, g)
    out = []
    for st in case["stages"]:
        s.contemplate(st["rollouts"])
        out.append(s.root_stats() + (s.choose(),))
        if st["move"] is None:
            break
        act, c = st["move"]
        a, b = O.PAIRS[act]
        g.place(a, b, lambda: c)
        s.sync(act, g)
    return out


def test_mcts_oracle_matches_reference_search_statistics():
    for case in load_golden("mcts_search_v1.json.gz"):
        def make(case, g):
            m = MO.MCTS(0, case["num_simulations"], case["seed"], case["root_index"])
            m.reset(g)
            return m
        for (n, q, ntot, choice), st in zip(replay_case(case, make), case["stages"]):
            assert n == st["N"] and q == st["Q"] and ntot == st["Ntot"] and choice == st["choose"]


@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_mcts_oracle_equals_live_reference():
    import random
    from oracle.refload import load_reference
    ns = load_reference()
    rng = random.Random(5)
    for trial in range(6):
        board = ns.qtttgym.Board(ns.qtttgym.QEvalClassic())
        g = O.Game()
        for _ in range(rng.randrange(0, 4)):
            act = rng.choice(g.legal_actions())
            a, b = O.PAIRS[act]
            c = rng.randrange(2)
            g.place(a, b, lambda: c)
            ns.coin.bits.clear(); ns.coin.feed(c)
            board.make_move((a, b))
        if g.terminal():
            continue
        ref, _ = MO.shim_reference_mcts(ns, 500 + trial, trial, 40, 5)
        ref.reset(board)
        ref.root.qstructs = [set(c) for c in board.qstructs]
        mine = MO.MCTS(40, 5, 500 + trial, trial)
        mine.reset(g)
        for _ in range(40):
            ref._rollout()
            mine._rollout()
        n, q, ntot = mine.root_stats()
        assert [ref.root.N.get(a, 0) for a in range(36)] == n
        assert [ref.root.Q.get(a, 0.0) for a in range(36)] == q
        assert ref.root.Ntot == ntot and ref.choose() == mine.choose()
