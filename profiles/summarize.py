"""Turns the raw ncu exports brought back in gpurun_out/ by profiles/capture.sh into the tracked
summaries under profiles/.    python profiles/summarize.py r02

  <tag>_kernels.md        one table: every captured launch with duration, DRAM bytes next to the
                          algorithmic bytes, pipe / issue utilisation, lanes per instruction, the
                          shared-memory conflict ratio, occupancy and the top stall reasons
  <tag>_step_hot.md       k_step (ply 7 launch): executed-instruction count per SASS region
  k_step_traffic.json     dram bytes per k_step launch (bench.py copies it into roofline.traffic)
  sass/<kernel>.sass      full SASS of k_step (headline variant) and k_sweep (cuobjdump -sass)
"""
import csv
import gzip
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

M = {
    "dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "inst": "smsp__inst_executed.sum", "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size",
    "conf": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "wave": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
}
STALLS = ["long_scoreboard", "math_pipe_throttle", "not_selected", "wait", "short_scoreboard", "barrier",
          "branch_resolving", "mio_throttle", "lg_throttle", "dispatch_stall", "no_instruction"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def load_raw(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    return rows[0], rows[1], rows[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def labels_from_drive(tag):
    """algorithmic bytes per launch, by kernel and order of appearance (drive_kernels.py's sequence)"""
    E = 1 << 24
    acc = None
    try:
        log = open(os.path.join(GP, f"drive_{tag}.log")).read()
        m = re.search(r"accepted by ply: \[(.*?)\]", log)
        if m:
            acc = [int(x) for x in m.group(1).split(",")]
    except OSError:
        pass
    return E, acc


def kernels(tag):
    E, acc = labels_from_drive(tag)
    lines = []
    traffic = []
    for part in ("step", "desync", "packed", "io", "stepobs", "qeval", "play", "mcts"):
        path = os.path.join(GP, f"raw_{part}_{tag}.csv")
        if not os.path.exists(path):
            continue
        raw = load_raw(path)
        if raw is None:
            continue
        hdr, units, data = raw
        ki = hdr.index("Kernel Name")
        seen = {}
        for r in data:
            name = re.sub(r"^void ", "", r[ki]).split("(")[0]
            name = re.sub(r"\((int|bool)\)", "", name).replace("qttt::", "")
            k = seen[name] = seen.get(name, -1) + 1
            g = {a: (num(r[hdr.index(m)]) if m in hdr else float("nan")) for a, m in M.items()}
            for a in ("dur", "rd", "wr"):
                if M[a] in hdr:
                    g[a] *= SCALE.get(units[hdr.index(M[a])], 1)
            st = {}
            for s in STALLS:
                key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
                if key in hdr:
                    st[s] = num(r[hdr.index(key)])
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            what, alg = name, None
            if part == "desync":
                what, alg = "k_step DESYNC batch (autoreset, forced actions, plies 1-9 mixed in every warp)", 47 * E
            elif part == "packed":
                what, alg = (("k_step_packed ply 4", 35 * (acc[4] if acc else E)) if "zc" not in name else
                             (("k_step_packed_zc<12-bit results> ply 4 (1 B in, 1.5 B out per env over PCIe)"
                               if "<1>" in name.replace(" ", "") else
                               "k_step_packed_zc ply 4 (kernel reads/writes mapped pinned host memory)"), None))
            elif part == "step" and name.startswith("k_step<"):
                if "k_step<0, 0, 1, 1" in name:
                    what, alg = "k_step ply 0 (reset fused in)", (31 * acc[0] if acc else None)
                elif "k_step<0, 0, 1, 0" in name:
                    ply = k + 1
                    if ply <= 8:
                        what, alg = f"k_step ply {ply}", (47 * acc[ply] if acc else None)
                        traffic.append(g["rd"] + g["wr"])
                    else:
                        what = f"k_step (launch {k})"
                elif "k_step<0, 0, 1, 2" in name:
                    what, alg = "k_step DESYNC batch (autoreset, forced actions)", 47 * E
                elif "k_step<0, 1, 0, 2" in name:
                    what, alg = "k_step desync batch, random policy (Philox inside)", 47 * E
            elif name.startswith("k_step_obs"):
                what, alg = "k_step_obs ply 4 (step + env.py observation)", 47 * (acc[4] if acc else E) + 28 * E
            elif name.startswith("k_step_packed_zc"):
                what, alg = "k_step_packed_zc ply 4 (host-mapped I/O)", None
            elif name.startswith("k_step_packed"):
                what, alg = "k_step_packed ply 4", 35 * (acc[4] if acc else E)
            elif name.startswith("k_qeval"):
                n = E if k < 2 else (1 << 20)
                what, alg = f"{name.split(chr(60))[0]} {n} boards", 33 * n
            elif name.startswith("k_observe"):
                everything = "<2>" in name or (chr(60) not in name and k < 2)
                what, alg = ("k_observe all outputs" if everything else "k_observe env.py outputs"), E * (16 + (90 if everything else 28))
            elif name.startswith("k_features"):
                what, alg = "k_features 2^20", (1 << 20) * 736
            elif name.startswith("k_get_mask"):
                what, alg = "k_get_mask 2^24", E * 52
            elif name.startswith("k_step_features"):
                what, alg = "k_step_features 2^20 (+mask)", (1 << 20) * (47 + 720 + 36)
            elif name.startswith("k_rollout"):
                what = "k_rollout 1024x256" if k < 2 else "k_rollout 65536x256"
            elif name.startswith("k_sweep"):
                what = "k_sweep 1.25e8 games"
            elif name.startswith("k_mcts_run"):
                what = "k_mcts_run 1024 roots x 500 rollouts x 10 sims"
            elif name.startswith("k_env1"):
                what = "k_env1 single-env step"
            lines.append((what, g, alg, top))
    with open(os.path.join(OUT, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# ncu --set full, every kernel bench.py times ({tag})\n\n"
                "`bash profiles/capture.sh` on one B200, `--clock-control none`, after the same program exited 0 "
                "without ncu; per-launch times are cold-cache and serialised (compare shares; live numbers come "
                "from CUDA events in bench.py).  alg = algorithmic bytes of the launch (DESIGN.md section 4); "
                "dram = dram__bytes_read + write; conflicts = shared-memory bank-conflict wavefronts / all "
                "shared wavefronts; lanes = thread instructions per warp instruction; tensor pipe = 0 everywhere.\n\n"
                "| launch | us | dram MB | alg MB | dram % | ALU % | FMA % | LSU % | issue % | warps % | lanes | "
                "smem conflicts | inst/warp-of-32-games | regs | grid x block | top stalls (warps per issue) |\n"
                "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|\n")
        for what, g, alg, top in lines:
            conf = g["conf"] / g["wave"] if g["wave"] else float("nan")
            threads = g["grid"] * g["block"]
            f.write(f"| {what} | {g['dur']:.1f} | {(g['rd'] + g['wr']) / 1e6:.1f} | "
                    f"{(alg / 1e6 if alg else float('nan')):.1f} | {g['dram']:.1f} | {g['alu']:.1f} | {g['fma']:.1f} | "
                    f"{g['lsu']:.1f} | {g['issue']:.1f} | {g['warps']:.1f} | {g['lanes']:.1f} | {conf:.2f} | "
                    f"{g['inst'] / 1e6:.1f}M total | {int(g['regs'])} | {int(g['grid'])} x {int(g['block'])} | "
                    + ", ".join(f"{k} {v:.1f}" for k, v in top) + " |\n")
    if traffic:
        with open(os.path.join(OUT, "k_step_traffic.json"), "w") as jf:
            json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic),
                       "source": f"profiles/{tag}_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum, mean of "
                                 f"the {len(traffic)} k_step launches of plies 1-8 of one pass, 2^24 envs)"}, jf)


def hot(tag):
    path = os.path.join(GP, f"src_step_{tag}.csv.gz")
    if not os.path.exists(path):
        return
    txt = gzip.open(path, "rt").read()
    blocks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
    out = []
    for bi, b in enumerate(blocks):
        lines = b.split("\n")
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr = rows[0]
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        ie, isrc, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
        tot = sum(int(r[ie]) for r in data)
        out.append((lines[0][:70], tot, data, ie, isrc, ith))
    with open(os.path.join(OUT, f"{tag}_step_hot.md"), "w") as f:
        f.write(f"# k_step: executed warp instructions per launch and per opcode ({tag}, ncu source page)\n\n"
                "524,288 warps of 32 games per launch (2^24 envs).\n\n| launch # | kernel | warp instructions | per warp |\n|---|---|---:|---:|\n")
        seen = set()
        for bi, (name, tot, data, ie, isrc, ith) in enumerate(out):
            if (name, tot) in seen:
                continue
            seen.add((name, tot))
            f.write(f"| {bi} | `{name}` | {tot} | {tot / 524288:.1f} |\n")
        # opcode histogram of the most expensive lock-step launch
        name, tot, data, ie, isrc, ith = max(out, key=lambda t: t[1])
        hist = {}
        for r in data:
            op = r[isrc].strip().split()
            if not op:
                continue
            o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            o = o.split(".")[0]
            hist[o] = hist.get(o, 0) + int(r[ie])
        f.write(f"\n## opcode mix of the most expensive launch ({tot / 524288:.1f} instructions per warp)\n\n| opcode | per warp | share |\n|---|---:|---:|\n")
        for o, c in sorted(hist.items(), key=lambda kv: -kv[1])[:16]:
            f.write(f"| {o} | {c / 524288:.1f} | {100 * c / tot:.1f}% |\n")


def sass():
    lib = os.path.join(ROOT, "qtttgym_b200", "csrc", "libqttt_b200.so")
    if not os.path.exists(lib):
        return
    os.makedirs(os.path.join(OUT, "sass"), exist_ok=True)
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"(?m)^\s*Function : ", text)[1:]
    for want, fname in (("_ZN4qttt6k_stepILi0ELb0ELb1ELi0EEEvNS_8StepArgsE", "k_step_index_forced_full.sass"),
                        ("_ZN4qttt7k_sweepEllmPy", "k_sweep.sass")):
        for fn in funcs:
            if fn.startswith(want):
                body = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", fn)
                with open(os.path.join(OUT, "sass", fname), "w") as f:
                    f.write("Function : " + body)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    kernels(tag)
    hot(tag)
    sass()
    print("written to", OUT)
