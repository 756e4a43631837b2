"""TEST INFRASTRUCTURE ONLY -- records golden fixtures from the LIVE, UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Writes tests/golden/*.json.gz.  The fixtures are data recorded from the reference's own
code paths (Env.step, Board, QEvalClassic.eval, MCTS._step / GameState), so they can pin
the oracle and the CUDA path on the GPU box where the reference itself is absent.

Fixtures
  kat_appendix_a.json.gz   SURVEY.md Appendix A known-answer games (actions, coins) with the
                           full per-step record re-captured from the live reference
  traces_v1.json.gz        600 random games: 200 plain, 200 with ~12 % illegal actions (Q2),
                           200 with illegal actions + 3 post-terminal steps (Q3)
  qeval_v1.json.gz         direct QEvalClassic.eval calls: (entangled_moves, coin) -> squares
  mcts_step_v1.json.gz     MCTS._step(node, action): children (board, moves, turn, winner,
                           terminal) incl. both collapse outcomes; action lists / masks
  population_v1.json       tallies of 20,000 reference random games (MT19937 seed 12345)
  features_v1.json.gz      GameState.to_vector() (mcts.py:67-85) and displayBoard() text
                           (qtttgym/display.py:4-32) along random MCTS._step walks
"""
from __future__ import annotations

import gzip
import json
import os
import random

from . import tracegen as T
from . import qttt_oracle as O
from .refload import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

APPENDIX_A = {  # name: (actions, coins-in-collapse-order)
    "kat1": ([(0, 1), (1, 2), (2, 0)], [0]),
    "kat2": ([(0, 1), (1, 2), (2, 0)], [1]),
    "kat3": ([(4, 5), (5, 4)], [1]),
    "kat4": ([(0, 1), (1, 2), (2, 3), (3, 4), (1, 3)], [0]),
    "kat5": ([(0, 1), (1, 2), (2, 3), (3, 4), (1, 3)], [1]),
    "kat6": ([(0, 1), (1, 0), (3, 3), (0, 3), (9, 1)], [0]),
    "kat7": ([(0, 1), (3, 4), (1, 2), (4, 5), (0, 2)], [0]),
    "kat8": ([(0, 1), (0, 1), (2, 3), (2, 3), (4, 5), (4, 5), (6, 7), (6, 7)], [0, 0, 0, 0]),
    "kat9": ([(0, 3), (1, 4), (3, 6), (4, 7), (0, 6), (1, 7)], [0, 0]),
    "kat10": ([(3, 7), (1, 7), (0, 8), (2, 6), (1, 6), (0, 6), (1, 7)], [1]),
}


def _dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    data = json.dumps(obj, separators=(",", ":"), sort_keys=True)
    if name.endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data.encode())
    else:
        with open(path, "w") as f:
            f.write(data)
    print(f"{path}: {os.path.getsize(path)} bytes")


def _coins_per_step(actions, coins):
    """Distribute collapse-ordered coins to steps by replaying on the oracle."""
    g = O.Game()
    it = iter(coins)
    out = []
    for a, b in actions:
        used = []

        def coin():
            c = next(it)
            used.append(c)
            return c

        if g.is_legal(a, b):
            g.place(a, b, coin)
        out.append((a, b, used[0] if used else 0))
    return out


def main():
    ns = load_reference()
    # ---- Appendix A
    kat = {}
    for name, (actions, coins) in APPENDIX_A.items():
        trace = _coins_per_step(actions, coins)
        kat[name] = {"trace": trace, "records": T.replay_reference(trace, ns)}
    _dump("kat_appendix_a.json.gz", kat)

    # ---- random traces
    rng = random.Random(20261018)
    games = []
    for g in range(600):
        kind = g // 200
        trace = T.random_trace(rng, illegal_rate=(0.0, 0.12, 0.12)[kind], overrun=(0, 0, 3)[kind])
        games.append({"trace": trace, "records": T.replay_reference(trace, ns)})
    _dump("traces_v1.json.gz", games)

    # ---- direct eval calls (the plugin seam, board.py:51 -> qeval.py:5)
    ev = ns.qtttgym.QEvalClassic()
    cases = []
    rng = random.Random(7)
    seen = set()
    while len(cases) < 400:
        g = O.Game()
        while not g.terminal():
            act = rng.choice(g.legal_actions())
            a, b = O.PAIRS[act]
            ia = next((i for i, c in enumerate(g.comps) if a in c), -1)
            ib = next((i for i, c in enumerate(g.comps) if b in c), -2)
            if ia == ib:
                ent = [m for m in g.moves if m[0] in g.comps[ia]] + [(a, b, len(g.moves))]
                key = tuple(ent)
                if key not in seen:
                    seen.add(key)
                    res = []
                    for coin in (0, 1):
                        ns.coin.bits.clear()
                        ns.coin.feed(coin)
                        res.append(ev.eval(list(ent)))
                    cases.append({"entangled": ent, "out0": res[0], "out1": res[1]})
            g.place(a, b, lambda: rng.randrange(2))
    _dump("qeval_v1.json.gz", cases)

    # ---- MCTS._step / GameState (mcts.py:19-27, 52-65, 87-91, 233-267)
    M = ns.mcts.MCTS
    steps = []
    rng = random.Random(11)
    for _ in range(150):
        mc = M()
        board = ns.qtttgym.Board(ns.qtttgym.QEvalClassic())
        mc.reset(board)
        node = mc.root
        while not node.terminal:
            act = int(rng.choice(node.actions))
            ns.coin.bits.clear()
            ns.coin.feed(0, 1)
            kids = mc._step(node, act)
            steps.append({
                "board": list(node.board), "moves": [list(m) for m in node.moves],
                "action": act, "actions": list(node.actions),
                "mask": [bool(x) for x in node.action_mask()],
                "children": [{"board": list(k.board), "moves": [list(m) for m in k.moves],
                              "turn": bool(k.turn),
                              "winner": None if k.winner is None else bool(k.winner),
                              "terminal": bool(k.terminal), "actions": list(k.actions)} for k in kids],
            })
            node = kids[rng.randrange(len(kids))]
    _dump("mcts_step_v1.json.gz", steps)

    # ---- to_vector / displayBoard along _step walks (children carry their qstructs)
    import contextlib
    import io
    feats = []
    rng = random.Random(13)
    for _ in range(60):
        mc = M()
        mc.reset(ns.qtttgym.Board(ns.qtttgym.QEvalClassic()))
        node = mc.root
        while True:
            vec = node.to_vector()
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ns.qtttgym.displayBoard(node)
            feats.append({"board": list(node.board), "moves": [list(m) for m in node.moves],
                          "nonzero": [[int(i), float(v)] for i, v in enumerate(vec.flatten()) if v != 0.0],
                          "display": buf.getvalue()})
            if node.terminal:
                break
            ns.coin.bits.clear()
            ns.coin.feed(0, 1)
            kids = mc._step(node, int(rng.choice(node.actions)))
            node = kids[rng.randrange(len(kids))]
    _dump("features_v1.json.gz", feats)

    # ---- population tallies with the reference's own MT19937 coin
    ns.qeval_module.random = ns.real_random
    try:
        ns.real_random.seed(12345)
        tally = {"x": 0, "o": 0, "draw": 0, "steps": 0, "collapses": 0, "games": 20000,
                 "steps_hist": [0] * 10, "autofill": 0}
        for _ in range(tally["games"]):
            bd = ns.qtttgym.Board(ns.qtttgym.QEvalClassic())
            n = 0
            while True:
                legal = [(i, j) for (i, j) in O.PAIRS if bd.board[i] == -1 and bd.board[j] == -1]
                before = list(bd.board)
                bd.make_move(ns.real_random.choice(legal))
                n += 1
                tally["collapses"] += int(before != bd.board)
                p1, p2 = bd.check_win()
                if p1 > 0 or p2 > 0 or len(bd.moves) > 8:
                    break
            w = None
            if p1 > 0 and p2 > 0:
                w = p1 < p2
            elif p1 > 0:
                w = True
            elif p2 > 0:
                w = False
            tally["x" if w is True else ("o" if w is False else "draw")] += 1
            tally["steps"] += n
            tally["steps_hist"][n] += 1
            tally["autofill"] += int(bd.moves[-1][0] == bd.moves[-1][1])
    finally:
        ns.qeval_module.random = ns.coin
    _dump("population_v1.json", tally)


if __name__ == "__main__":
    main()
