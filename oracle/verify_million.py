"""TEST INFRASTRUCTURE ONLY -- one-off check at the north-star scale: 1,000,000 random games
(random legal actions in random order, random forced coins) through the LIVE unmodified
reference (Env.step) and the Python oracle, every field compared after every step, on all
cores.  Build container only.  `python -m oracle.verify_million [n_games]`
(The GPU suite then diffs the CUDA path against the oracle on 1e6 games as well, which closes
the chain reference == oracle == CUDA at that scale.)"""
from __future__ import annotations

import multiprocessing as mp
import os
import random
import sys
import time


def _worker(args):
    seed, n_games = args
    from . import tracegen as T
    from .refload import load_reference
    ns = load_reference()
    rng = random.Random(seed)
    steps = 0
    for _ in range(n_games):
        trace = T.random_trace(rng)
        a, b = T.replay_oracle(trace), T.replay_reference(trace, ns)
        if a != b:
            return ("MISMATCH", trace)
        steps += len(trace)
    return ("ok", steps)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    cores = os.cpu_count() or 1
    chunks = 64
    per = -(-n // chunks)
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_worker, [(1000 + i, per) for i in range(chunks)])
    bad = [r for r in res if r[0] != "ok"]
    steps = sum(r[1] for r in res if r[0] == "ok")
    print(f"games {per * chunks} steps {steps} mismatches {len(bad)} in {time.time() - t0:.0f} s on {cores} cores")
    if bad:
        print(bad[0])
        raise SystemExit(1)


if __name__ == "__main__":
    main()
