"""Times one lock-step pass (9 plies, 2^24 envs) through BatchedEnv.step_obs (qttt_step_obs) and
through step followed by observation (qttt_step_ex + qttt_observe), per ply with CUDA events.

    python profiles/stepobs_bench.py [--envs 16777216]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import torch
    import qtttgym_b200 as Q
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    E, seed, reps = args.envs, 20261018, args.reps
    gen = Q.BatchedEnv(E, device=dev, seed=seed)
    actions = torch.empty((9, E), dtype=torch.uint8, device=dev)
    coins = torch.empty((9, E), dtype=torch.uint8, device=dev)
    for ply in range(9):
        info = gen.step_random(record=True)[4]
        actions[ply].copy_(info["action"])
        coins[ply].copy_(info["coin"])
    del gen
    env = Q.BatchedEnv(E, device=dev, seed=seed)
    buf = env.observation()

    def fused(ply):
        env.step_obs(actions[ply], coins[ply], out=buf, fresh=(ply == 0))

    def split(ply):
        (env.reset_step if ply == 0 else env.step)(actions[ply], coins[ply])
        env.observation(out=buf)

    out = {}
    for name, fn in (("fused", fused), ("step_then_observe", split)):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(18)] for _ in range(reps)]
        for it in range(2 + reps):
            for ply in range(9):
                if it >= 2:
                    ev[it - 2][2 * ply].record()
                fn(ply)
                if it >= 2:
                    ev[it - 2][2 * ply + 1].record()
        torch.cuda.synchronize()
        per = [sum(ev[k][2 * p].elapsed_time(ev[k][2 * p + 1]) for k in range(reps)) / reps for p in range(9)]
        out[name] = {"pass_ms": round(sum(per), 4), "ply_us": [round(1e3 * x, 1) for x in per]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
