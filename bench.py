#!/usr/bin/env python
"""bench.py -- batched QTTT env-steps/sec on B200 (and % of the HBM roofline) next to the
reference's CPU path timed on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--envs E]

One *pass* of the hot path over one batch is ``reset`` + 9 ``step`` calls (kernel K1) that play
E games from the empty board to termination on pre-generated random-legal actions and forced
collapse coins (config 2 of BASELINE.json scaled to fill the GPU: E = 2^24 envs per GPU; the
literal 4096-env config is reported under ``extra``).  One bench *step* is ``--passes-per-step``
such passes (40: the timed region of the default run is ~1 s, long enough to show sustained
clocks).  The metric counts *accepted* moves only (an env-step = one accepted make_move;
finished games idle as illegal no-ops and are not counted).

``value``  : inputs already resident in HBM when the timed region starts.
``e2e``    : the same pass through the public API with pinned HOST buffers
             (BatchedEnv.step_host): actions + coins copied host->device and reward / done /
             legal mask copied device->host inside the timed region, every ply.
``roofline``: kernel k_step, 47 algorithmic bytes per env-step,
             duration from CUDA events around every launch in the timed region.
``cpu_baseline``: the Python oracle port on all host cores (N=1, rank 0 only).

Multi-GPU (torchrun, one rank per GPU): every rank plays its own E games (weak scaling, no
data-path collective); the only collective is one all_reduce(SUM) of the win/draw tallies of
the self-play sweep reported under ``extra`` (NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched QTTT env-steps/sec"
UNIT = "env-steps/s"
BYTES_PER_STEP = 47            # 16 state in + 16 out + 1 action + 1 coin + 4 reward + 1 done + 8 mask
BYTES_PER_BOARD_QEVAL = 33     # 16 state + 1 action + 2 x 8 boards
PLIES = 9
WORKLOAD = ("step API (K1): reset + 9 steps per pass (the reset is fused into the first step's launch), "
            "random legal actions with forced "
            "collapse coins, games played from the empty board to termination "
            "(config 2 of BASELINE.json scaled to fill the GPU)")


def bench_config(envs_per_gpu, world):
    """The workload description both arms print (identical for `--impl b200` and `--impl reference`)."""
    return {"workload": WORKLOAD, "envs_per_gpu": envs_per_gpu, "global_envs": envs_per_gpu * world,
            "plies_per_pass": PLIES,
            "l2": ("inputs exceed L2: 16 B x E state + 2 B x E actions/coins + 13 B x E outputs per launch "
                   f"= {31 * envs_per_gpu / 1e6:.0f} MB vs 126 MB L2" if 31 * envs_per_gpu > 126e6
                   else "inputs fit in L2 (small E)"),
            "parallelism": f"dp{world} (independent games per rank, no data-path collective)"}


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per k_step launch from the committed
    `ncu --set full` capture (profiles/k_step_traffic.json, written by profiles/summarize.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "k_step_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons through NVML from a background
    thread every ~2 ms DURING the timed region (the timed region of a short run is only tens
    of milliseconds, too short for an `nvidia-smi -lms` subprocess to see)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, torch_device):
        import threading
        self.samples, self.masks = [], []
        self.max_mhz = None
        self._stop = threading.Event()
        self.err = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            handle = None
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index or 0)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                        self.masks.append(int(get_reasons(handle)))
                    except Exception as e:   # pragma: no cover
                        self.err = repr(e)
                        return
                    time.sleep(0.002)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)
            self.thread = None

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [f"no samples ({self.err})"]}
        seen = 0
        for m in self.masks:
            seen |= m
        reasons = sorted(name for bit, name in self.REASONS.items() if seen & bit)
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples),
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": reasons,
                "how": "NVML SM clock + clock-event reasons sampled every ~2 ms inside the timed region"}


def bind_to_gpu_numa(torch_device):
    """Pins this rank's CPU affinity to the cores NVML reports as local to its GPU, so that the
    pinned host buffers of the e2e path are allocated on the GPU's own NUMA node (with 8 ranks
    the PCIe copies otherwise cross the socket interconnect)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:   # best effort
        return f"unavailable ({e!r})"


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.  The reference is
    pure Python and is not present on the GPU box, so this times the Python oracle port
    (oracle/qttt_oracle.py) with the config-1 loop on every core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as B
    cores = os.cpu_count() or 1
    games_per_proc = args.ref_games
    pool = B.PersistentPool(cores)
    for w in range(args.warmup):
        pool.step(max(50, games_per_proc // 10), 1000 + w)
    steps_total, t_total = 0, 0.0
    for k in range(args.steps):
        s, wall = pool.step(games_per_proc, 2000 + k)
        steps_total += s
        t_total += wall
    pool.close()
    value = steps_total / t_total
    sample = (f"config-1 loop (one Env per process, random.choice over legal pairs until terminated), "
              f"{games_per_proc} games per process per step on {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
        "config": bench_config(args.envs, args.gpus),
        "reference_sample": {"what": "the same workload (random-vs-random games from the empty board to "
                                     "termination through Env.reset/step) on the CPU: one env per process, "
                                     f"{cores} processes x {games_per_proc} games per step",
                             "envs_per_process": 1, "processes": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def pcie_ceiling(torch, dev, barrier, max_over_ranks, nbytes=256 << 20, reps=4):
    """Pinned-host copy ceiling of this rank's link, measured with plain cudaMemcpyAsync
    (torch .copy_ of pinned tensors): each direction alone and both at once on two streams.
    Under torchrun every rank measures at the same time (barrier first), so the numbers are
    what the box gives N ranks concurrently.  GB/s per rank."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1)
        e2.record(s2)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(e1)
        cur.wait_event(e2)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) * 1e-3
    run(True, True)
    t_h2d, t_d2h, t_both = run(True, False), run(False, True), run(True, True)
    gb = nbytes * reps / 1e9
    return {"h2d_alone_gbs": gb / t_h2d, "d2h_alone_gbs": gb / t_d2h,
            "h2d_concurrent_gbs": gb / t_both, "d2h_concurrent_gbs": gb / t_both,
            "how": f"{reps} x {nbytes >> 20} MiB pinned cudaMemcpyAsync per direction; 'concurrent' = both "
                   "directions at once on two streams; per rank, all ranks measuring at the same time"}


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, single-GPU runs only): forked workers, before CUDA exists.
    cpu_baseline = None
    cpu_c = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as B
        from oracle import c_oracle as CO
        cores = os.cpu_count() or 1
        total, slowest, wall, per = B.run_pool("time", args.cpu_seconds, cores)
        cpu_baseline = {
            "value": total / slowest, "unit": UNIT, "cores": cores, "kind": "port",
            "per_core": statistics.mean(per),
            "sample": f"Python oracle port, config-1 loop (Env.reset/step, random.choice over legal "
                      f"pairs), {cores} processes x {args.cpu_seconds:.0f} s = {total} env-steps",
        }
        t0 = time.perf_counter()
        st, _ = CO.selfplay(0, 4_000_000, 1)
        dt = time.perf_counter() - t0
        cpu_c = {"value": float(st[3]) / dt, "unit": UNIT, "cores": CO.num_threads(), "kind": "port",
                 "sample": "C oracle (oracle/qttt_oracle.c, OpenMP), 4,000,000 Philox self-play games"}

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import qtttgym_b200 as Q
    affinity = bind_to_gpu_numa(dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    E, K, W, P = args.envs, args.steps, args.warmup, args.passes_per_step
    seed = 20261018
    from qtttgym_b200 import _lib as qlib
    launches = 0          # our kernels launched inside the timed regions (value, value_eager, e2e)
    peak, peak_src = measured_hbm_peak()

    # ---- synthetic workload: random-policy traces generated on the device (untimed)
    env = Q.BatchedEnv(E, device=dev, seed=seed, game_base=rank * E)
    actions = torch.empty((PLIES, E), dtype=torch.uint8, device=dev)
    coins = torch.empty((PLIES, E), dtype=torch.uint8, device=dev)
    accepted = []
    for ply in range(PLIES):
        _, _, _, _, info = env.step_random(out=(actions[ply], coins[ply]))
        accepted.append(int((info["status"] == 0).sum().item()))
    steps_per_pass = sum(accepted)
    final_winner = torch.bincount(env.winner().long(), minlength=3)
    total_steps_pass = sum_over_ranks(float(steps_per_pass))

    # ---- value: inputs resident in HBM; one bench step = P passes, each pass one CUDA graph of the
    #      9 launches (reset fused into the first).  K steps are timed.
    graph_full = env.capture_episode(actions, coins)
    for _ in range(W * P):                     # W warm-up steps
        graph_full.replay()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler = ClockSampler(dev)
    g0.record()
    for _ in range(K * P):
        graph_full.replay()
    g1.record()
    barrier()
    clocks = sampler.stop()
    graph_ms = max_over_ranks(g0.elapsed_time(g1))
    launches += K * P * PLIES
    value_graph = total_steps_pass * K * P / (graph_ms * 1e-3)
    assert torch.equal(torch.bincount(env.winner().long(), minlength=3), final_winner), \
        "timed replay diverged from the generated trace"

    # ---- the same pass issued eagerly, with CUDA events around every launch (per-ply durations)
    Ke = max(3, min(K, 10))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * PLIES)] for _ in range(Ke)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3 + Ke):
        if it == 3:
            barrier()
            launches_before = qlib.LAUNCHES
            start.record()
        for ply in range(PLIES):
            if it >= 3:
                ev[it - 3][2 * ply].record()
            if ply == 0:
                env.reset_step(actions[0], coins[0])       # Env.reset fused into the first ply
            else:
                env.step(actions[ply], coins[ply])
            if it >= 3:
                ev[it - 3][2 * ply + 1].record()
    end.record()
    barrier()
    t_ms = max_over_ranks(start.elapsed_time(end))
    launches += qlib.LAUNCHES - launches_before
    value_eager = total_steps_pass * Ke / (t_ms * 1e-3)
    assert torch.equal(torch.bincount(env.winner().long(), minlength=3), final_winner)
    per_ply_ms = [sum(ev[k][2 * p].elapsed_time(ev[k][2 * p + 1]) for k in range(Ke)) / Ke for p in range(PLIES)]
    ev_launch_ms = sum(per_ply_ms) / PLIES

    # algorithmic bytes of one pass: 47 B per accepted step, except the first ply, whose launch has
    # the reset fused in and does not read the state (47 - 16 = 31 B per step)
    alg_bytes_pass = BYTES_PER_STEP * steps_per_pass - 16 * accepted[0]
    avg_launch_ms = graph_ms / (K * P * PLIES)            # the timed region itself: 9 launches per pass
    achieved = (alg_bytes_pass / PLIES) / (avg_launch_ms * 1e-3) / 1e9
    achieved_ev = (alg_bytes_pass / PLIES) / (ev_launch_ms * 1e-3) / 1e9
    alg_by_ply = [BYTES_PER_STEP * a - (16 * a if p == 0 else 0) for p, a in enumerate(accepted)]
    roofline = {"bound": "hbm", "kernel": "k_step<QTTT_ACT_INDEX, forced coins, all outputs>", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (args.traffic_bytes if args.traffic_bytes is not None else
                            (profiled_traffic() or {}).get("dram_bytes_per_launch")),
                "traffic_source": (profiled_traffic() or {}).get("source"),
                "algorithmic_bytes_per_launch": alg_bytes_pass / PLIES,
                "avg_launch_ms": avg_launch_ms,
                "avg_launch_source": "the timed region of `value` itself: CUDA events around K x P graph replays of the "
                                     "9-launch pass, divided by 9 K P (kernel time plus the gaps between graph nodes)",
                "per_launch_events": {"avg_launch_ms": ev_launch_ms, "frac": achieved_ev / peak,
                                      "launch_ms_by_ply": per_ply_ms,
                                      "frac_by_ply": [b / (m * 1e-3) / 1e9 / peak for b, m in zip(alg_by_ply, per_ply_ms)],
                                      "note": "the same pass issued eagerly with one CUDA event pair per launch "
                                              f"({Ke} passes); each pair adds ~2 us of event latency to its launch"},
                "bytes_per_env_step": BYTES_PER_STEP,
                "note": "mean over the 9 launches of a pass; the first ply's launch (reset fused in, state not "
                        "read) is charged 31 B per step; finished games idle as no-ops and earn no bytes although "
                        "their state is still read and their outputs written (ply 8: half of the lanes)"}

    extra = {}

    def timed(fn, reps):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    # ---- K1 on a DESYNCHRONISED batch: envs spread over all plies inside every warp (what a policy-
    #      driven vector env with autoreset looks like).  31 autoreset self-play steps mix the batch,
    #      9 more are recorded and replayed (forced actions + coins, autoreset on) as one CUDA graph.
    mix = Q.BatchedEnv(E, device=dev, seed=seed + 1, game_base=rank * E)
    for _ in range(31):
        mix.step_random(autoreset=True)
    d_start, d_epoch = mix.state.clone(), mix.epoch
    d_act = torch.empty((PLIES, E), dtype=torch.uint8, device=dev)
    d_coin = torch.empty((PLIES, E), dtype=torch.uint8, device=dev)
    d_accepted = 0
    for t in range(PLIES):
        _, _, _, _, info = mix.step_random(autoreset=True, out=(d_act[t], d_coin[t]))
        d_accepted += int(((info["status"] & 3) == 0).sum().item())
    d_final = mix.state.clone()
    ply_hist = torch.bincount(((d_start[:, 0] >> 27) & 15).long(), minlength=10).tolist()

    def desync_pass():
        mix.epoch = d_epoch
        for t in range(PLIES):
            mix.step(d_act[t], d_coin[t], autoreset=True)
    mix.state.copy_(d_start)
    desync_pass()
    torch.cuda.synchronize()
    assert torch.equal(mix.state, d_final), "desync replay diverged"
    dg = torch.cuda.CUDAGraph()
    mix.state.copy_(d_start)
    with torch.cuda.graph(dg):
        desync_pass()
    # (replaying from wherever the last replay ended keeps the batch just as mixed; the actions were
    #  legal for the recorded trajectory only, so every replay starts from the recorded start state)
    def desync_replay():
        mix.state.copy_(d_start)
        dg.replay()
    copy_ms = timed(lambda: mix.state.copy_(d_start), 20)
    ms = timed(desync_replay, 20) - copy_ms
    d_bytes = BYTES_PER_STEP * d_accepted
    extra["k1_desync"] = {
        "ms_per_launch": ms / PLIES, "env_steps_per_s": sum_over_ranks(float(d_accepted)) / (ms * 1e-3),
        "roofline_frac": d_bytes / (ms * 1e-3) / 1e9 / peak, "accepted_steps_per_launch": d_accepted / PLIES,
        "envs_by_len_moves_at_start": ply_hist,
        "note": "2^24 envs at mixed plies in every warp, step(..., autoreset=True): a game that is over restarts "
                "inside the launch, so every lane plays every step; 9 launches as one CUDA graph; 47 B per step"}
    del mix, dg, d_start, d_final, d_act, d_coin

    # ---- step + full observation in the timed region (classical, q lists, turn after every ply):
    #      the fused launch (qttt_step_obs) and, beside it, step followed by qttt_observe
    obs_env = Q.BatchedEnv(E, device=dev, seed=seed, game_base=rank * E)
    obs_buf = obs_env.observation()

    def obs_pass_two_launches():
        for ply in range(PLIES):
            (obs_env.reset_step if ply == 0 else obs_env.step)(actions[ply], coins[ply])
            obs_env.observation(out=obs_buf)

    def obs_pass():
        for ply in range(PLIES):
            obs_env.step_obs(actions[ply], coins[ply], out=obs_buf, fresh=(ply == 0))
    ms2 = timed(obs_pass_two_launches, 3)
    ms = timed(obs_pass, 3)
    extra["value_with_observation"] = {
        "env_steps_per_s": total_steps_pass / (ms * 1e-3), "ms_per_pass": ms,
        "env_steps_per_s_step_then_observe": total_steps_pass / (ms2 * 1e-3), "ms_per_pass_step_then_observe": ms2,
        "note": "the same pass through BatchedEnv.step_obs (qttt_step_obs: Env.step and the env.py:68-85 tensors "
                "classical int8[N,9], q_states_p1/p2, turn of the new state from one launch), i.e. what the "
                "reference's Env.step returns as obs, decoded on the device; beside it step + qttt_observe"}
    del obs_env, obs_buf

    # ---- e2e: host buffers through the public API, copies inside the timed region
    ceiling = pcie_ceiling(torch, dev, barrier, max_over_ranks)
    h_ac = Q.pack_actions(actions, coins).cpu().pin_memory()          # 1 byte per env per ply
    h_res = torch.empty(E, dtype=torch.int16).pin_memory()            # 2 bytes per env per ply
    h_obs = torch.empty((E, 4), dtype=torch.int32).pin_memory()       # 16 bytes per env per ply
    e2e_K = max(1, min(K, args.e2e_steps))

    def e2e_run(step_fn):
        nonlocal launches
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(2 + e2e_K):
            if it == 2:
                barrier()
                before = qlib.LAUNCHES
                s2.record()
            env.reset()
            for ply in range(PLIES):
                step_fn(ply)
        e2.record()
        barrier()
        launches += qlib.LAUNCHES - before
        ms = max_over_ranks(s2.elapsed_time(e2))
        return total_steps_pass * e2e_K / (ms * 1e-3), ms / e2e_K

    def link_floor_ms(b_in, b_out):
        # the traffic is lopsided (1 B in, 2 or 18 B out per env-step), so each direction is held against
        # what the link gave that direction ALONE: the most favourable ceiling, frac <= 1
        return 1e3 * max(b_in / (ceiling["h2d_alone_gbs"] * 1e9), b_out / (ceiling["d2h_alone_gbs"] * 1e9))

    variants = {}
    for name, kw, b_in, b_out in (
            ("packed_copy", dict(), 1, 2),
            ("packed_copy_obs", dict(obs_host=h_obs), 1, 18),
            ("packed_mapped", dict(mapped=True), 1, 2),
            ("packed_mapped_obs", dict(obs_host=h_obs, mapped=True), 1, 18)):
        v, ms = e2e_run(lambda ply, kw=kw: env.step_host_packed(h_ac[ply], h_res, **kw))
        _, term_chk, _, _ = Q.unpack_result(h_res)
        assert int(term_chk.sum()) == E, "e2e pass did not finish every game"
        if "obs_host" in kw:
            assert torch.equal(h_obs, env.state.cpu()), "e2e observation is not the final state"
        floor = link_floor_ms(b_in * E * PLIES, b_out * E * PLIES)
        variants[name] = {"value": v, "ms_per_pass": ms, "h2d_bytes_per_pass": b_in * E * PLIES,
                          "d2h_bytes_per_pass": b_out * E * PLIES, "link_floor_ms_per_pass": floor,
                          "frac_of_link_ceiling": floor / ms}
    # the mapped path with the result words bit-packed: 12 bits per env, four envs in three words
    h_res12 = torch.empty(3 * ((E + 3) // 4), dtype=torch.int16).pin_memory()      # 1.5 bytes per env per ply
    v, ms = e2e_run(lambda ply: env.step_host_packed12(h_ac[ply], h_res12))
    _, term_chk, _, _ = Q.unpack_result(Q.unpack_result12(h_res12, E))
    assert int(term_chk.sum()) == E, "e2e pass did not finish every game"
    floor = link_floor_ms(E * PLIES, 3 * ((E + 3) // 4) * 2 * PLIES)
    variants["packed12_mapped"] = {"value": v, "ms_per_pass": ms, "h2d_bytes_per_pass": E * PLIES,
                                   "d2h_bytes_per_pass": 3 * ((E + 3) // 4) * 2 * PLIES,
                                   "link_floor_ms_per_pass": floor, "frac_of_link_ceiling": floor / ms}
    v, ms = e2e_run(lambda ply: env.step_host_packed12(h_ac[ply], h_res12, mapped=False))
    _, term_chk, _, _ = Q.unpack_result(Q.unpack_result12(h_res12, E))
    assert int(term_chk.sum()) == E, "e2e pass did not finish every game"
    variants["packed12_copy"] = dict(variants["packed12_mapped"], value=v, ms_per_pass=ms, frac_of_link_ceiling=floor / ms)
    del h_res12
    best = max(("packed_copy", "packed_mapped", "packed12_mapped", "packed12_copy"), key=lambda k: variants[k]["value"])
    # the observation as 12-byte records (env.py observation + flags) instead of 16-byte state + result word
    h_rec = torch.empty((E, 3), dtype=torch.int32).pin_memory()
    v, ms = e2e_run(lambda ply: env.step_host_obs12(h_ac[ply], h_rec))
    obs_chk, _, term_chk, _, _ = Q.unpack_obs12(h_rec[:4096])
    want_chk = Q.observe_states(env.state[:4096])
    assert bool(term_chk.all()) and all(torch.equal(obs_chk[k], want_chk[k].cpu()) for k in want_chk), \
        "e2e obs12 records do not decode to the final observation"
    floor = link_floor_ms(E * PLIES, 12 * E * PLIES)
    variants["obs12_copy"] = {"value": v, "ms_per_pass": ms, "h2d_bytes_per_pass": E * PLIES,
                              "d2h_bytes_per_pass": 12 * E * PLIES, "link_floor_ms_per_pass": floor,
                              "frac_of_link_ceiling": floor / ms}
    del h_rec
    best_obs = max(("packed_copy_obs", "packed_mapped_obs", "obs12_copy"), key=lambda k: variants[k]["value"])
    e2e = {"value": variants[best]["value"], "unit": UNIT, "h2d_bytes_per_step": variants[best]["h2d_bytes_per_pass"] * P,
           "d2h_bytes_per_step": variants[best]["d2h_bytes_per_pass"] * P, "passes_timed": e2e_K,
           "ms_per_pass": variants[best]["ms_per_pass"],
           "variant": best,
           "fields_returned": "per env per ply one result word: free-square set (= the 36-bit legal mask, re-expanded by "
                              "unpack_result), terminated, line (= reward -1.0 / -0.0), status -- 16 bits per env, or "
                              "(packed12_mapped) 12 bits per env with four envs in three words.  The observation stays "
                              "in HBM (the packed state tensor); `with_obs` is the same path with the observation "
                              "(packed 16-B state per env) copied to the host as well",
           "api": "BatchedEnv.reset + 9 x BatchedEnv.step_host_packed (pinned host buffers: 1 B action|coin in, 2 B out "
                  "per env per ply).  packed_copy = qttt_step_packed_host_obs: 8 slices pipelined over 4 side "
                  "streams, one cudaMemcpyAsync per array per slice; packed_mapped = qttt_step_packed_mapped: one "
                  "launch whose threads read / write the pinned host buffers across PCIe themselves; packed12_mapped = "
                  "qttt_step_packed12_mapped (BatchedEnv.step_host_packed12): the same with 1.5 B out per env per ply; "
                  "packed12_copy = qttt_step_packed12_host: the 12-bit results through the cudaMemcpyAsync pipeline",
           "with_obs": {"value": variants[best_obs]["value"], "variant": best_obs,
                        "ms_per_pass": variants[best_obs]["ms_per_pass"],
                        "h2d_bytes_per_pass": variants[best_obs]["h2d_bytes_per_pass"],
                        "d2h_bytes_per_pass": variants[best_obs]["d2h_bytes_per_pass"],
                        "note": "obs12_copy = BatchedEnv.step_host_obs12 (qttt_step_packed_host_obs12): the env.py "
                                "observation and the step flags as one 12-byte record per env, decoded by "
                                "unpack_obs12; packed_*_obs = the 16-byte packed state + the result word"},
           "variants": variants, "pcie_ceiling": ceiling,
           "note": "frac_of_link_ceiling = (bytes that must cross the link / pinned cudaMemcpyAsync bandwidth measured "
                   "in this run for that direction alone, all ranks copying at once) / measured time.  With several "
                   "ranks the per-rank ceiling itself drops (one host, one NUMA node feeds every GPU): that, not "
                   "the kernels, is what bounds multi-GPU e2e"}
    del h_obs

    # ---- extras: the other configs of BASELINE.json
    # config 2 literal: 4096 envs (latency-bound: 9 launches of ~2 us of work each)
    small = Q.BatchedEnv(4096, device=dev, seed=seed, game_base=rank * E)
    sa, sc = actions[:, :4096].contiguous(), coins[:, :4096].contiguous()
    small_steps = int(sum(int((sa[p] < 36).sum().item()) for p in range(PLIES)))

    def small_pass():
        small.reset()
        for p in range(PLIES):
            small.step(sa[p], sc[p])
    ms = timed(small_pass, 50)
    graph = small.capture_episode(sa, sc)
    ms_graph = timed(graph.replay, 200)
    extra["config2_4096_envs"] = {"env_steps_per_s": small_steps / (ms_graph * 1e-3) * world,
                                  "ms_per_pass_cuda_graph": ms_graph, "ms_per_pass_eager": ms,
                                  "env_steps_per_s_eager": small_steps / (ms * 1e-3) * world,
                                  "note": "launch-latency bound (9-10 launches of 4096 threads per pass); the whole "
                                          "episode is captured in one CUDA graph (BatchedEnv.capture_episode)"}

    # config 5: fused self-play sweep (K5) + the one NCCL all_reduce of the tallies
    G = args.sweep_games
    stats_box = {}

    def sweep_pass():
        stats_box["s"] = Q.sharded_sweep(G * world, seed, dev)
    ms = timed(sweep_pass, 3)
    st = stats_box["s"].cpu().tolist()
    extra["config5_sweep"] = {
        "env_steps_per_s": st[3] / (ms * 1e-3), "games": st[5], "env_steps": st[3], "ms": ms,
        "x_wins": st[0], "o_wins": st[1], "draws": st[2], "collapses": st[4],
        "hbm_frac_at_47B": st[3] / (ms * 1e-3) * BYTES_PER_STEP / 1e9 / (peak * world),
        "collective": "all_reduce(SUM) int64[16]" if world > 1 else "none (1 rank)"}

    # config 3: qeval both outcomes over 2^20 mid-game boards
    nb = 1 << 20
    qa = actions[4, :nb].clone()
    qenv = Q.BatchedEnv(nb, device=dev, seed=seed, game_base=rank * E)
    for p in range(4):
        qenv.step(actions[p, :nb], coins[p, :nb])
    qa = torch.where(qa < 36, qa, torch.zeros_like(qa))
    qout = Q.qeval_both(qenv.state, qa, want_states=False, want_probs=False)
    ms_api = timed(lambda: Q.qeval_both(qenv.state, qa, want_states=False, want_probs=False), 20)
    qgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(qgraph):
        for _ in range(10):
            Q.qeval_both(qenv.state, qa, out=qout)
    ms = timed(qgraph.replay, 20) / 10
    extra["config3_qeval_1M_boards"] = {
        "boards_per_s": nb / (ms * 1e-3) * world, "ms": ms, "ms_per_python_call": ms_api,
        "hbm_frac_at_33B": nb / (ms * 1e-3) * BYTES_PER_BOARD_QEVAL / 1e9 / peak,
        "note": "device time per launch (10 launches per CUDA-graph replay); 34.6 MB per launch fit the 126 MB L2, "
                "so this size is launch-latency / L2 bound; a bare Python call costs ms_per_python_call"}

    # the same kernel at a size that fills the GPU (2^24 boards: the trace's ply-4 positions)
    qa_big = torch.where(actions[4] < 36, actions[4], torch.zeros_like(actions[4]))
    big = Q.BatchedEnv(E, device=dev, seed=seed, game_base=rank * E)
    for p in range(4):
        big.step(actions[p], coins[p])
    qout_big = Q.qeval_both(big.state, qa_big, want_states=False, want_probs=False)
    ms = timed(lambda: Q.qeval_both(big.state, qa_big, out=qout_big), 10)
    extra["qeval_16M_boards"] = {"boards_per_s": E / (ms * 1e-3) * world, "ms": ms,
                                 "hbm_frac_at_33B": E / (ms * 1e-3) * BYTES_PER_BOARD_QEVAL / 1e9 / peak}
    del qout_big

    # config 4: 1024 roots x 256 rollouts
    roots = qenv.state[:1024].clone()
    box = {}

    box["r"] = Q.rollout_eval(roots, 256, seed)

    def roll():
        Q.rollout_eval(roots, 256, seed, out=box["r"])
    ms_call = timed(roll, 50)
    rsteps = int(box["r"][2].item())
    rgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(rgraph):
        for _ in range(10):
            roll()
    ms = timed(rgraph.replay, 20) / 10
    extra["config4_rollout_1024x256"] = {"playouts_per_s": 1024 * 256 / (ms * 1e-3) * world,
                                         "env_steps_per_s": rsteps / (ms * 1e-3) * world, "ms": ms,
                                         "ms_per_python_call": ms_call,
                                         "env_steps_per_s_python_calls": rsteps / (ms_call * 1e-3) * world,
                                         "note": "device time per launch (10 launches + their counter resets per "
                                                 "CUDA-graph replay); one launch is 20 us of GPU work, so bare "
                                                 "Python calls (ms_per_python_call) are launch-latency bound"}
    # config 1 through the drop-in single-env adapter (qtttgym_b200.Env): the reference's own loop
    if rank == 0:
        import random as _random
        single = Q.Env(device=dev, seed=seed)
        prng = _random.Random(1)
        t0 = time.perf_counter()
        n_single = 0
        for _ in range(300):
            obs, _ = single.reset()
            term = False
            while not term:
                board = obs["classical"]
                legal = [p for p in Q.PAIRS if board[p[0]] == -1 and board[p[1]] == -1]
                obs, _, term, _, _ = single.step(prng.choice(legal))
                n_single += 1
        dt = time.perf_counter() - t0
        extra["config1_single_env_adapter"] = {
            "env_steps_per_s": n_single / dt, "us_per_step": 1e6 * dt / n_single,
            "note": "one env, one step per call through qtttgym_b200.Env: wall time of the reference's own loop "
                    "(legal-move list + random.choice + step) per step"}

    # MCTS search around the leaf evaluator (next row #1): 1024 mid-game roots, the reference's
    # default num_simulations=10, 500 rollouts per root; and the config-4 shape (256 playouts per leaf)
    # (a search is one warp per tree, so it is latency-bound until the machine is full of trees:
    # the third line is the same search over 32,768 roots)
    many_roots = qenv.state[:32768].clone()
    for name, n_roll, n_sim, rts in (("mcts_1024_roots_500x10", 500, 10, roots),
                                     ("mcts_1024_roots_100x256", 100, 256, roots),
                                     ("mcts_32768_roots_100x10", 100, 10, many_roots)):
        n_rts = rts.shape[0]
        mc = Q.BatchedMCTS(rollouts=n_roll, num_simulations=n_sim, seed=seed, root_base=rank * n_rts, device=dev)

        def search():
            mc.reset(rts, total_rollouts=n_roll)
            mc.contemplate(n_roll)
        ms = timed(search, 3)

        assert int(mc.errors().max().item()) == 0
        extra[name] = {"ms": ms, "rollouts_per_s": n_rts * n_roll / (ms * 1e-3) * world,
                       "playouts_per_s": n_rts * n_roll * n_sim / (ms * 1e-3) * world,
                       "nodes_per_tree_mean": float(mc.node_counts().float().mean().item())}
        del mc
    del many_roots

    # a9 observation decode (env.py:68-85 + extras) and the to_vector feature encoder
    obs_buf = env.observation(extras=True)
    ms = timed(lambda: env.observation(extras=True, out=obs_buf), 10)
    extra["observe_all_outputs"] = {"ms": ms, "envs": E, "bytes_per_env": 16 + 90,
                                    "gb_per_s": E * (16 + 90) / (ms * 1e-3) / 1e9,
                                    "note": "qttt_observe, every output (classical, moves, n_moves, q lists, turn, "
                                            "rounds, reward_p1, winner, bool mask) into preallocated tensors"}
    del obs_buf
    ms = timed(lambda: Q.to_vector(qenv.state), 20)
    extra["to_vector_1M_states"] = {"ms": ms, "states_per_s": nb / (ms * 1e-3) * world,
                                    "gb_per_s": nb * 736 / (ms * 1e-3) / 1e9}
    # the net-input path: step fused with to_vector + get_mask (one launch) next to step, then encoders
    nf = 1 << 20
    fenv = Q.BatchedEnv(nf, device=dev, seed=seed, game_base=rank * E)
    fa, fc = actions[4, :nf].contiguous(), coins[4, :nf].contiguous()
    fstart = qenv.state.clone()

    def fused():
        fenv.state.copy_(fstart)
        return fenv.step_features(fa, fc, want_mask=True, out=fbox.get("info"))
    fbox = {}
    fbox["info"] = fused()[4]
    ms_copy = timed(lambda: fenv.state.copy_(fstart), 20)
    ms_fused = timed(fused, 20) - ms_copy

    def separate():
        fenv.state.copy_(fstart)
        fenv.step(fa, fc)
        Q.to_vector(fenv.state)
        Q.get_mask(fenv.state)
    ms_sep = timed(separate, 20) - ms_copy
    extra["step_features_1M_envs"] = {
        "ms_fused": ms_fused, "ms_step_then_to_vector_then_get_mask": ms_sep,
        "gb_per_s_fused": nf * (47 + 720 + 36) / (ms_fused * 1e-3) / 1e9,
        "note": "qttt_step_features (Env.step + GameState.to_vector + nn.Model.get_mask of the new state in one "
                "launch, 803 B per env) against the three separate launches (which also allocate their outputs)"}
    del fenv, fbox
    ms = timed(lambda: Q.get_mask(qenv.state), 20)
    extra["get_mask_1M_states"] = {"ms": ms, "gb_per_s": nb * 52 / (ms * 1e-3) / 1e9}

    # MCTS node-pool occupancy with pruning (MCTS._prune): 256 games, 200 rollouts per move, both players
    # searching; peak live nodes per tree vs nodes that would have been needed without reclamation
    if rank == 0:
        ar_env = Q.BatchedEnv(256, device=dev, seed=seed)
        mx, mo = Q.MCTSStrategy(rollouts=200, num_simulations=10, seed=1), Q.MCTSStrategy(rollouts=200, num_simulations=10, seed=2)
        mx.reset(ar_env)
        mo.reset(ar_env)
        created = torch.zeros(256, dtype=torch.int64, device=dev)
        for ply in range(9):
            mover = mx if ply % 2 == 0 else mo
            before = mover.search.live_counts().clone()
            mover.contemplate()
            created += (mover.search.live_counts() - before) if mover is mx else 0
            a_ = mover.choose()
            ar_env.step(a_)
            mx.sync(a_)
            mo.sync(a_)
        extra["mcts_pool_with_prune"] = {
            "peak_live_nodes_per_tree_mean": float(mx.search.peak_counts().float().mean().item()),
            "peak_live_nodes_per_tree_max": int(mx.search.peak_counts().max().item()),
            "pool_high_water_mean": float(mx.search.node_counts().float().mean().item()),
            "nodes_created_mean": float(created.float().mean().item()) + 1.0,
            "capacity_per_tree": mx.search.capacity, "node_bytes": mx.search.node_bytes,
            "errors": int(mx.search.errors().max().item()),
            "note": "X's search over a 9-ply game against an MCTS opponent, 200 rollouts x 10 playouts per move, "
                    "sync() after every ply (own and opponent's): nodes of pruned subtrees return to the pool"}

    # the rollout kernel at a size that fills the GPU: 65,536 roots x 256 playouts
    big_roots = big.state[:65536].clone()
    box["rb"] = Q.rollout_eval(big_roots, 256, seed)
    ms = timed(lambda: Q.rollout_eval(big_roots, 256, seed, out=box["rb"]), 20)
    extra["rollout_65536x256"] = {"playouts_per_s": 65536 * 256 / (ms * 1e-3) * world,
                                  "env_steps_per_s": int(box["rb"][2].item()) / (ms * 1e-3) * world, "ms": ms}
    del big
    if cpu_c:
        extra["cpu_c_oracle"] = cpu_c
    if affinity is not None:
        extra["rank0_cpu_affinity"] = (f"{len(affinity)} cores local to the GPU (NVML)"
                                       if isinstance(affinity, list) else affinity)
    extra["population"] = {"x_wins": int(final_winner[1]), "o_wins": int(final_winner[2]),
                           "draws": int(final_winner[0]), "games": E}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value_graph, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": graph_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": bench_config(E, world),
            "run": {"passes_per_step": P, "timed_region_s": graph_ms * 1e-3,
                    "env_steps_per_pass_per_gpu": steps_per_pass, "accepted_steps_by_ply": accepted,
                    "step": "one bench step = passes_per_step passes; a pass = reset + 9 steps over all envs, "
                            "replayed as one CUDA graph of 9 kernel launches"},
            "value_eager": {"value": value_eager, "ms_per_pass": t_ms / Ke,
                            "note": "the same pass issued as 9 separate BatchedEnv.reset_step/step calls per pass "
                                    "(the loop the per-launch events are recorded in)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 24, help="envs per GPU")
    ap.add_argument("--sweep-games", type=int, default=125_000_000, help="config-5 games per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3, help="passes timed per e2e variant")
    ap.add_argument("--passes-per-step", type=int, default=40,
                    help="passes (CUDA-graph replays of reset + 9 steps) per bench step: 20 steps x 40 passes ~ 1 s")
    ap.add_argument("--cpu-seconds", type=float, default=3.0, help="CPU baseline wall seconds per core")
    ap.add_argument("--ref-games", type=int, default=4000, help="--impl reference: games per process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per k_step launch from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "b200":
        # launched directly with --gpus N: re-launch as one rank per GPU (what the driver does itself)
        import socket
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
