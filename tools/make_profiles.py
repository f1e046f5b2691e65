"""Turns the raw artefacts of a measurement campaign (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/make_profiles.py [round-tag]   (default r1)"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(PROF, exist_ok=True)


def copy(src, dst):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
        return True
    return False


def ncu_raw(rep, kernel_filter, want, title, dst):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines, traffic = [title, ""], []
    for r in rows[2:]:
        if kernel_filter not in r[hdr.index("Kernel Name")]:
            continue
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for w in want:
            if w in hdr:
                lines.append(f"  {w:82s} {r[hdr.index(w)]:>22s} {units[hdr.index(w)]}")

        def val(name):
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index(name)]]
            return float(r[hdr.index(name)].replace(",", "")) * scale
        traffic.append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    open(os.path.join(PROF, dst), "w").write("\n".join(lines) + "\n")
    return sum(traffic) / len(traffic) if traffic else None


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct"]

copy("bench_r1_n1.json", f"{TAG}_bench_n1.json")
copy("bench_r1_ref.json", f"{TAG}_bench_reference_arm.json")
for n in (2, 4, 8):
    copy(f"bench_r1_n{n}.json", f"{TAG}_bench_n{n}.json")
copy("launches.csv", f"{TAG}_launches_bench_n1.csv")
copy("sweep.md", f"{TAG}_sweep_d64.md")
copy("cpu_baselines.md", f"{TAG}_cpu_baselines.md")
copy("fp64_peak.json", f"{TAG}_fp64_peak.json")
for f in os.listdir(OUT):
    if f.startswith("config") and f.endswith(".json"):
        copy(f, f"{TAG}_{f}")

t = ncu_raw("prof_de_generation.ncu-rep", "de_generation_kernel", WANT,
            "ncu --set full --clock-control none --import-source on -k regex:de_generation_kernel -s 5 -c 2 "
            "python bench.py --steps 5 --warmup 5 --skip-cpu-baseline   (B200; workload = BASELINE configs[1])",
            f"{TAG}_de_generation_ncu_full.txt")
if t:
    json.dump({"de_generation_kernel": {
        "dram_bytes_per_launch": t, "workload": "DE-random Rastrigin d=1000 P=1048576 fp64",
        "source": f"profiles/{TAG}_de_generation_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of the captured launches)"}},
        open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
ncu_raw("prof_pso_move.ncu-rep", "pso_move_kernel", WANT,
        "ncu --set full --clock-control none -k regex:pso_move_kernel -s 4 -c 1 python tests/tools/quick_time_pso.py 2097152 256 3 3 1 1"
        "   (B200; accelerated PSO, Ackley d=256, 2^21 particles, fp64 = one GPU's share of BASELINE configs[2])",
        f"{TAG}_pso_move_ncu_full.txt")

if os.path.exists(os.path.join(OUT, "launches.csv")):
    rows = [r for r in csv.reader(open(os.path.join(OUT, "launches.csv"))) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 5 --warmup 5 --skip-cpu-baseline   (B200)",
             "per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
             f"{'total ms':>10s} {'launches':>8s} {'share':>7s}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{v[1] / 1e6:10.3f} {v[0]:8d} {100 * v[1] / tot:6.1f}%  {k}")
    g = sum(v[1] for k, v in agg.items() if "de_generation_kernel" in k)
    rp = sum(v[1] for k, v in agg.items() if "de_repair_kernel" in k)
    c = sum(v[1] for k, v in agg.items() if "de_commit_kernel" in k)
    live = json.load(open(os.path.join(OUT, "bench_r1_n1.json")))["roofline"]["step_share"]["generation"]
    lines += ["", f"share of the generation step K2 / (K2 + K2r + K3) under ncu: {g / (g + rp + c):.4f}   "
                  f"(live CUDA-event step_share.generation in bench.py: {live:.4f})"]
    open(os.path.join(PROF, f"{TAG}_launches_bench_n1.txt"), "w").write("\n".join(lines) + "\n")
print("profiles/ updated:", sorted(os.listdir(PROF)))
