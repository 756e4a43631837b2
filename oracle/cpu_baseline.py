"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU path timed on host cores.

The reference is pure Python and cannot travel to the GPU box, so the timed thing is the
Python oracle port (oracle/qttt_oracle.py: same language, same data structures -- lists of
tuples and sets -- same algorithm), driven by the config-1 loop of BASELINE.md section 4:
one Env per process, reset(), uniform-random legal (a, b) via random.choice until terminated.
In the build container the port runs at ~1.1-1.3x the speed of the unmodified reference
(DESIGN.md records the measurement), i.e. it is a slightly *stronger* baseline.
"""
from __future__ import annotations

import os
import random
import time


def play_games(n_games: int, seed: int) -> tuple[int, float]:
    """config 1: plays n_games random-vs-random games; returns (env_steps, seconds)."""
    from . import qttt_oracle as O

    rng = random.Random(seed)
    pairs = O.PAIRS
    env = O.Env(coin=lambda: rng.getrandbits(1))
    steps = 0
    t0 = time.perf_counter()
    for _ in range(n_games):
        obs, _ = env.reset()
        term = False
        while not term:
            board = obs["classical"]
            legal = [p for p in pairs if board[p[0]] == -1 and board[p[1]] == -1]
            obs, _, term, _, _ = env.step(rng.choice(legal))
            steps += 1
    return steps, time.perf_counter() - t0


def play_for(seconds: float, seed: int) -> tuple[int, float]:
    steps, t0 = 0, time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        s, _ = play_games(200, seed + 7919 * k)
        steps += s
        k += 1
    return steps, time.perf_counter() - t0


def _worker(args):
    kind, amount, seed = args
    return play_for(amount, seed) if kind == "time" else play_games(int(amount), seed)


def run_pool(kind: str, amount: float, cores: int | None = None, seed: int = 1):
    """All-cores aggregate: ``cores`` processes each run the config-1 loop.  Returns
    (total env_steps, wall seconds of the slowest worker, per-process steps/s list)."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")   # callers create pools BEFORE initialising CUDA
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_worker, [(kind, amount, seed + 104729 * i) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(s for s, _ in res)
    slowest = max(t for _, t in res)
    return total, slowest, wall, [s / t for s, t in res]


class PersistentPool:
    """Keeps worker processes alive across bench steps (spawn cost stays out of the timing)."""

    def __init__(self, cores: int | None = None):
        import multiprocessing as mp

        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("fork").Pool(self.cores)   # create before CUDA init

    def step(self, games_per_proc: int, seed: int):
        t0 = time.perf_counter()
        res = self.pool.map(_worker, [("games", games_per_proc, seed + 104729 * i)
                                      for i in range(self.cores)])
        wall = time.perf_counter() - t0
        return sum(s for s, _ in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()
