"""Turns the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under
profiles/.  Usage: python profiles/summarize.py r01   (reads gpurun_out/launches_<tag>.csv,
gpurun_out/prof_step_<tag>.ncu-rep, gpurun_out/prof_sweep_<tag>.ncu-rep)."""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(tag):
    path = os.path.join(GP, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            name = r["Kernel Name"].split("(")[0].replace(",", ";")
            f.write(f'{r["ID"]},{name},{r["Grid Size"].replace(",", " ")},{r["Block Size"].replace(",", " ")},{r["Metric Value"]}\n')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` "
                "over `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1`\n\n"
                "Per-launch times are cold-cache and serialised: compare shares, not absolutes.  The first 400\n"
                "launches cover trace generation (k_step<0,1>, 9 launches), torch bookkeeping of the untimed\n"
                "setup, and the warm-up + timed passes of the step API (k_reset + 9 x k_step<0,0,1> per pass).\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |\n")
        ours = sum(v[1] for k, v in agg.items() if "qttt::" in k)
        f.write(f"\nqttt:: kernels: {100 * ours / tot:.1f}% of the captured GPU time.\n")
        # the timed region is a sequence of passes: k_step<0,0,1,1> (reset fused in) then 8 x k_step<0,0,1,0>
        names = [r["Kernel Name"].split("(")[0] for r in rows]
        durs = [float(r["Metric Value"]) for r in rows]
        passes = []
        for i, nm in enumerate(names):
            if "k_step<0, 0, 1, 1>" in nm and i + 8 < len(names) and all("k_step<0, 0, 1, 0>" in x for x in names[i + 1:i + 9]):
                passes.append(durs[i:i + 9])
        if passes:
            steps = [sum(p[k] for p in passes) / len(passes) for k in range(9)]
            gap = [nm for nm in set(names) if "qttt::" not in nm]
            f.write(f"\n## One pass of the step API (the bench step), mean of {len(passes)} captured passes\n\n"
                    f"9 x k_step = {sum(steps) / 1e3:.1f} us (by ply: {', '.join(f'{x / 1e3:.0f}' for x in steps)} us); "
                    "the first launch has Env.reset fused in.  No other kernel runs inside a pass: "
                    "k_step is 100% of the GPU time of the bench step.\n")


def full(tag, which):
    rep = os.path.join(GP, f"prof_{which}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if which == "step" and "dram__bytes_read.sum" in hdr:
        import json
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data]
        with open(os.path.join(OUT, "k_step_traffic.json"), "w") as jf:
            json.dump({"dram_bytes_per_launch": sum(per) / len(per),
                       "source": f"profiles/{tag}_step_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum, "
                                 f"mean of {len(per)} launches of 2^24 envs)"}, jf)
    with open(os.path.join(OUT, f"{tag}_{which}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full, kernel k_{which} ({tag})\n\nOne column per captured launch.\n\n| metric | unit | "
                + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---:|" * len(data) + "\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
    # per-instruction hot spots from the source page (needs -lineinfo)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(OUT, f"{tag}_{which}_ncu_source_head.csv"), "w") as f:
        f.write("\n".join(src.splitlines()[:400]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    full(tag, "step")
    full(tag, "sweep")
    print("written to", OUT)
