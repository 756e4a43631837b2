"""Launches every kernel that bench.py times, once each at the bench's sizes, so that one
`ncu --set full -k regex:qttt` capture of this short program covers all of them (profiles/README.md).
Run plain first (prints CUDA-event times per launch), then under ncu:

    python profiles/drive_kernels.py [--envs 16777216] [--only step,qeval,...]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import qtttgym_b200 as Q  # noqa: E402


def timed(label, fn, reps=1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b) / reps * 1e3:.1f} us", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only   # noqa: E731
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    E, seed = args.envs, 20261018
    env = Q.BatchedEnv(E, device=dev, seed=seed)
    actions = torch.empty((9, E), dtype=torch.uint8, device=dev)
    coins = torch.empty((9, E), dtype=torch.uint8, device=dev)
    accepted = []
    for ply in range(9):                                  # trace generation: k_step<0,1,...>
        _, _, _, _, info = env.step_random(record=True)
        actions[ply].copy_(info["action"])
        coins[ply].copy_(info["coin"])
        accepted.append(int((info["status"] == 0).sum().item()))
    print("accepted by ply:", accepted, flush=True)

    if want("step"):
        for rep in range(2):                              # pass 0 warms up, pass 1 is the one to read
            for ply in range(9):
                f = env.reset_step if ply == 0 else env.step
                timed(f"k_step ply {ply} (pass {rep})", lambda: f(actions[ply], coins[ply]))
    # mid-game positions (ply 4) for the kernels that take states
    env.reset()
    for ply in range(4):
        env.step(actions[ply], coins[ply])
    mid = env.state.clone()
    qa = torch.where(actions[4] < 36, actions[4], torch.zeros_like(actions[4]))
    if want("desync"):
        mix = Q.BatchedEnv(E, device=dev, seed=seed + 1)
        for _ in range(31):
            mix.step_random(autoreset=True)
        da = torch.empty(E, dtype=torch.uint8, device=dev)
        dc = torch.empty(E, dtype=torch.uint8, device=dev)
        start = mix.state.clone()
        mix.step_random(autoreset=True, out=(da, dc))
        for rep in range(2):
            mix.state.copy_(start)
            timed(f"k_step desync autoreset (rep {rep})", lambda: mix.step(da, dc, autoreset=True))
        del mix, start
    if want("packed"):
        ac = Q.pack_actions(actions[4], coins[4])
        res = torch.empty(E, dtype=torch.int16, device=dev)
        st = mid.clone()

        def packed():
            Q._lib.check(env.lib.qttt_step_packed(st.data_ptr(), ac.data_ptr(), res.data_ptr(), E,
                                                  torch.cuda.current_stream().cuda_stream))
        timed("k_step_packed ply 4", packed)
        timed("k_step_packed ply 4 (again)", packed)
        h_ac = ac.cpu().pin_memory()
        h_res = torch.empty(E, dtype=torch.int16).pin_memory()

        def mapped():
            Q._lib.check(env.lib.qttt_step_packed_mapped(st.data_ptr(), h_ac.data_ptr(), h_res.data_ptr(), None, E,
                                                         torch.cuda.current_stream().cuda_stream))
        timed("k_step_packed_zc ply 4 (mapped pinned host buffers)", mapped)
        timed("k_step_packed_zc ply 4 (again)", mapped)
        h_res12 = torch.empty(3 * ((E + 3) // 4), dtype=torch.int16).pin_memory()

        def mapped12():
            Q._lib.check(env.lib.qttt_step_packed12_mapped(st.data_ptr(), h_ac.data_ptr(), h_res12.data_ptr(), E,
                                                           torch.cuda.current_stream().cuda_stream))
        timed("k_step_packed_zc<12-bit results> ply 4", mapped12)
        timed("k_step_packed_zc<12-bit results> ply 4 (again)", mapped12)
    if want("observe"):
        buf = Q.observe_states(mid, extras=True)
        timed("k_observe all outputs", lambda: Q.observe_states(mid, extras=True, out=buf))
        buf2 = Q.observe_states(mid)
        timed("k_observe env.py outputs", lambda: Q.observe_states(mid, out=buf2))
        del buf, buf2
    if want("stepobs"):
        oenv = Q.BatchedEnv(E, device=dev, seed=seed)
        obuf = oenv.observation()
        for rep in range(2):
            oenv.state.copy_(mid)
            timed(f"k_step_obs ply 4 (rep {rep})", lambda: oenv.step_obs(actions[4], coins[4], out=obuf))
        del oenv, obuf
    if want("features"):
        timed("k_features 2^20", lambda: Q.to_vector(mid[:1 << 20]))
        timed("k_features 2^20 (again)", lambda: Q.to_vector(mid[:1 << 20]))
        timed("k_get_mask 2^24", lambda: Q.get_mask(mid))
        nf = 1 << 20
        fenv = Q.BatchedEnv(nf, device=dev, seed=seed)
        fenv.state.copy_(mid[:nf])
        info = fenv.step_features(actions[4, :nf], coins[4, :nf], want_mask=True)[4]
        fenv.state.copy_(mid[:nf])
        timed("k_step_features 2^20 (step + to_vector + get_mask)",
              lambda: fenv.step_features(actions[4, :nf], coins[4, :nf], want_mask=True, out=info))
        del fenv, info
    if want("qeval"):
        out_big = Q.qeval_both(mid, qa, want_states=False, want_probs=False)
        timed("k_qeval_both 2^24 boards", lambda: Q.qeval_both(mid, qa, out=out_big))
        nb = 1 << 20
        out_small = Q.qeval_both(mid[:nb], qa[:nb], want_states=False, want_probs=False)
        timed("k_qeval_both 2^20 boards", lambda: Q.qeval_both(mid[:nb], qa[:nb], out=out_small))
        del out_big, out_small
    if want("rollout"):
        for nr in (1024, 65536):
            roots = mid[:nr].clone()
            o = Q.rollout_eval(roots, 256, seed)
            timed(f"k_rollout {nr}x256", lambda: Q.rollout_eval(roots, 256, seed, out=o))
    if want("sweep"):
        timed("k_sweep 1.25e8 games", lambda: Q.selfplay_sweep(0, 125_000_000, seed, device=dev))
    if want("mcts"):
        roots = mid[:1024].clone()
        mc = Q.BatchedMCTS(rollouts=500, num_simulations=10, seed=seed, device=dev)
        mc.reset(roots, total_rollouts=500)
        timed("k_mcts_run 1024 roots 500x10", lambda: mc.contemplate(500))
    if want("env1"):
        single = Q.Env(device=dev, seed=1)
        timed("k_env1 (single-env step, record to mapped host memory)", lambda: single.step((0, 1)), reps=20)
    torch.cuda.synchronize()
    print("drive_kernels done", flush=True)


if __name__ == "__main__":
    main()
