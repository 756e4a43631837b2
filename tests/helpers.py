"""Shared comparison helpers: golden records / oracle outputs / CUDA outputs in one shape."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return json.loads(f.read().decode())
    with open(path) as f:
        return json.load(f)


def traces_to_arrays(traces):
    """list of [(a, b, coin), ...] -> (pairs int8[T, G, 2] padded with (-1,-1), coins uint8[T, G],
    lengths)."""
    G = len(traces)
    T = max(len(t) for t in traces)
    pairs = np.full((T, G, 2), -1, np.int8)
    coins = np.zeros((T, G), np.uint8)
    for g, t in enumerate(traces):
        for s, (a, b, c) in enumerate(t):
            pairs[s, g] = (a, b)
            coins[s, g] = c
    return pairs, coins, np.array([len(t) for t in traces])


def check_against_record(rec, g, step_out, obs, where=""):
    """``rec``: a golden / python-oracle record; ``step_out`` / ``obs``: dicts of numpy arrays as
    produced by oracle.c_oracle.Games.step/observe or the CUDA path; ``g``: row."""
    tag = f"{where} game {g}"
    assert obs["classical"][g].tolist() == rec["board"], tag
    nm = int(obs["n_moves"][g])
    assert nm == len(rec["moves"]), tag
    mv = obs["moves"][g].tolist()
    assert [m + [i] for i, m in enumerate(mv[:nm])] == rec["moves"], tag
    assert all(m == [-1, -1] for m in mv[nm:]), tag
    assert int(np.asarray(step_out["reward"][g:g + 1]).view(np.uint32)[0]) == rec["reward_bits"], tag
    assert bool(step_out["done"][g]) == rec["terminated"], tag
    assert int(step_out["mask"][g]) == rec["mask"], tag
    assert obs["rounds"][g].tolist() == rec["rounds"], tag
    assert float(obs["reward_p1"][g]) == rec["reward_p1"], tag
    assert int(obs["winner"][g]) == rec["winner"], tag
    q1 = [p for p in obs["q_p1"][g].tolist() if p[0] >= 0]
    q2 = [p for p in obs["q_p2"][g].tolist() if p[0] >= 0]
    assert q1 == rec["q1"] and q2 == rec["q2"], tag
    assert int(obs["turn"][g]) == rec["turn"], tag


def run_traces_and_check(make_games, traces, records, where=""):
    """Drives a batched implementation (``make_games(n)`` -> object with .step(pairs, coins) and
    .observe()) through ``traces`` ply by ply and checks every record."""
    pairs, coins, lens = traces_to_arrays(traces)
    games = make_games(len(traces))
    for s in range(pairs.shape[0]):
        out = games.step(pairs[s], coins[s])
        obs = games.observe()
        for g in np.nonzero(lens > s)[0]:
            check_against_record(records[g][s], int(g), out, obs, where=f"{where} step {s}")
    return games
