"""Times the step kernel (K1) ply by ply for a lock-step pass and for a desynchronised batch,
under the tuning knob of csrc/qttt_kernels.cu (QTTT_STEP_ITERS: chunks of 256 games per block), one
subprocess per setting (the knob is read once per process).

    python profiles/tune_step.py [--envs 16777216] [--reps 5] [--grid "4,8,16"]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_STEP = 47


def child(args):
    import torch
    import qtttgym_b200 as Q
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    E, seed, reps = args.envs, 20261018, args.reps
    peak = 6461.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    env = Q.BatchedEnv(E, device=dev, seed=seed)
    actions = torch.empty((9, E), dtype=torch.uint8, device=dev)
    coins = torch.empty((9, E), dtype=torch.uint8, device=dev)
    accepted = []
    for ply in range(9):
        _, _, _, _, info = env.step_random(record=True)
        actions[ply].copy_(info["action"])
        coins[ply].copy_(info["coin"])
        accepted.append(int((info["status"] == 0).sum().item()))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(18)] for _ in range(reps)]
    for it in range(2 + reps):
        for ply in range(9):
            if it >= 2:
                ev[it - 2][2 * ply].record()
            (env.reset_step if ply == 0 else env.step)(actions[ply], coins[ply])
            if it >= 2:
                ev[it - 2][2 * ply + 1].record()
    torch.cuda.synchronize()
    per_ply = [sum(ev[k][2 * p].elapsed_time(ev[k][2 * p + 1]) for k in range(reps)) / reps for p in range(9)]
    alg = BYTES_PER_STEP * sum(accepted) - 16 * accepted[0]
    frac = alg / (sum(per_ply) * 1e-3) / 1e9 / peak
    # desynchronised batch: mix with autoreset self-play, then replay 9 recorded steps
    mix = Q.BatchedEnv(E, device=dev, seed=seed + 1)
    for _ in range(31):
        mix.step_random(autoreset=True)
    start = mix.state.clone()
    start_epoch = mix.epoch
    for t in range(9):
        mix.step_random(autoreset=True, out=(actions[t], coins[t]))
    final = mix.state.clone()
    plies = torch.bincount(((start[:, 0] >> 27) & 15).long(), minlength=10).tolist()
    dv = [[torch.cuda.Event(enable_timing=True) for _ in range(18)] for _ in range(reps)]
    for it in range(2 + reps):
        mix.state.copy_(start)
        mix.epoch = start_epoch
        for t in range(9):
            if it >= 2:
                dv[it - 2][2 * t].record()
            mix.step(actions[t], coins[t], autoreset=True)
            if it >= 2:
                dv[it - 2][2 * t + 1].record()
    torch.cuda.synchronize()
    assert torch.equal(mix.state, final), "desync replay diverged"
    d_ms = [sum(dv[k][2 * p].elapsed_time(dv[k][2 * p + 1]) for k in range(reps)) / reps for p in range(9)]
    d_frac = BYTES_PER_STEP * E * 9 / (sum(d_ms) * 1e-3) / 1e9 / peak
    print(json.dumps({"iters": os.environ.get("QTTT_STEP_ITERS"),
                      "pass_ms": round(sum(per_ply), 4), "frac": round(frac, 4),
                      "ply_us": [round(1e3 * x, 1) for x in per_ply],
                      "desync_ms_per_launch": round(sum(d_ms) / 9, 4), "desync_frac": round(d_frac, 4),
                      "desync_start_plies": plies}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--grid", default="8")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for cfg in args.grid.split(","):
        env = dict(os.environ, QTTT_STEP_ITERS=cfg.split(":")[0])
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--envs", str(args.envs),
                        "--reps", str(args.reps)], env=env, check=False)


if __name__ == "__main__":
    main()
