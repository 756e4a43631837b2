"""Times qttt_qeval_both (boards-only shape, config 3) the way bench.py does: 2^24 ply-4 boards
(one ply per batch), 2^20 boards under a CUDA graph, and a batch that mixes every ply inside every
warp (positions of an auto-resetting self-play batch).

    python profiles/qeval_bench.py [--envs 16777216]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_BOARD = 33


def main():
    import torch
    import qtttgym_b200 as Q
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 24)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    E, seed = args.envs, 20261018
    peak = 6461.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def timed(fn, reps):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def frac(n, ms):
        return round(n / (ms * 1e-3) * BYTES_PER_BOARD / 1e9 / peak, 4)

    out = {}
    for ply in (4, 7):
        env = Q.BatchedEnv(E, device=dev, seed=seed)
        for _ in range(ply):
            env.step_random()
        qa = env.step_random(record=True)[4]["action"].clone()          # a legal action of the position before it
        env2 = Q.BatchedEnv(E, device=dev, seed=seed)
        for _ in range(ply):
            env2.step_random()
        qa = torch.where(qa < 36, qa, torch.zeros_like(qa))
        buf = Q.qeval_both(env2.state, qa, want_states=False, want_probs=False)
        ms = timed(lambda: Q.qeval_both(env2.state, qa, out=buf), 20)
        out[f"ply{ply}_16M"] = {"us": round(ms * 1e3, 1), "frac_at_33B": frac(E, ms),
                                "closes": float(buf["closes"].float().mean().item()) if isinstance(buf, dict) else None}
        if ply == 4:
            nb = 1 << 20
            st, a1 = env2.state[:nb].clone(), qa[:nb].clone()
            small = Q.qeval_both(st, a1, want_states=False, want_probs=False)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10):
                    Q.qeval_both(st, a1, out=small)
            ms = timed(g.replay, 20) / 10
            out["ply4_1M_graph"] = {"us": round(ms * 1e3, 2), "frac_at_33B": frac(nb, ms)}
        del env, env2, buf
    mix = Q.BatchedEnv(E, device=dev, seed=seed + 1)
    for _ in range(31):
        mix.step_random(autoreset=True)
    state = mix.state.clone()
    qa = mix.step_random(record=True, autoreset=True)[4]["action"].clone()
    qa = torch.where(qa < 36, qa, torch.zeros_like(qa))
    buf = Q.qeval_both(state, qa, want_states=False, want_probs=False)
    ms = timed(lambda: Q.qeval_both(state, qa, out=buf), 20)
    out["mixed_plies_16M"] = {"us": round(ms * 1e3, 1), "frac_at_33B": frac(E, ms)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
