"""Round-2 evidence: turns the raw artefacts of the GPU calls (gpurun_out/r2_*) and of the local build into the
committed summaries under profiles/ (r2_*).   usage: python tools/make_profiles_r2.py"""
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def copy(src, dst):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
        return True
    return False


def launches(src_csv, bench_json, dst):
    rows = [r for r in csv.reader(open(src_csv)) if len(r) > 5]
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
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 5 --warmup 5 "
             "--skip-cpu-baseline --no-extras   (B200)",
             "per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
             f"{'total ms':>10s} {'launches':>8s} {'share':>7s}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{v[1] / 1e6:10.3f} {v[0]:8d} {100 * v[1] / tot:6.1f}%  {k}")
    g = sum(v[1] for k, v in agg.items() if "de_generation" in k)
    rp = sum(v[1] for k, v in agg.items() if "de_repair_kernel" in k)
    c = sum(v[1] for k, v in agg.items() if "de_commit_kernel" in k)
    b = last_json(bench_json)
    live = b["roofline"]["step_share"]["generation"] if b else float("nan")
    lines += ["", f"share of the generation step K2 / (K2 + K2r + K3) under ncu: {g / (g + rp + c):.4f}   "
                  f"(live CUDA-event step_share.generation of the same command: {live:.4f})"]
    open(dst, "w").write("\n".join(lines) + "\n")


def sass_summary(dst):
    """Which Blackwell / Hopper-class instructions the shipped cubins contain, per kernel family."""
    so = os.path.join(ROOT, "nlsolver_b200", "libnls_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fam = defaultdict(Counter)
    n_kernels = Counter()
    cur = None
    pat = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|SYNCS|UTCHMMA|UTCQMMA|LDTM|STTM|HMMA|LDGSTS|LDG\.E\.128|STG\.E\.128|LDS\.128|"
                     r"REDUX|ATOMG|ATOM|REDG|RED|UCGABAR_ARV|UCGABAR_WAIT|BAR\.SYNC|MUFU\.RSQ64H|MUFU\.RCP64H)\b")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"^void nls::", "", name).split("<")[0].split("(")[0]
            n_kernels[cur] += 1
            continue
        if cur:
            for hit in pat.findall(line.split("/*")[1] if line.count("/*") >= 2 else line):
                fam[cur][hit] += 1
    lines = ["cuobjdump -sass nlsolver_b200/libnls_b200.so   (sm_100a cubins of the shipped library), static instruction counts",
             "per kernel family (all template instantiations together).",
             "UBLKCP = cp.async.bulk (TMA unit, 1-D bulk copy global -> shared); SYNCS = mbarrier arrive / try_wait;",
             "UCGABAR = barrier.cluster (thread-block clusters); no tensor-core instruction is expected on this path.", ""]
    for k in sorted(fam, key=lambda k: -sum(fam[k].values())):
        lines.append(f"{k}  ({n_kernels[k]} instantiations)")
        lines.append("    " + ", ".join(f"{op} x{n}" for op, n in sorted(fam[k].items())))
    open(dst, "w").write("\n".join(lines) + "\n")


def spills(dst):
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, "nlsolver_b200", "csrc", "build", "*.ptxas.log"))):
        txt = open(f).read()
        for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
            name = b.split("'")[0]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
            r = re.search(r"Used (\d+) registers", b)
            if m and r:
                rows.append((name, int(r.group(1)), int(m.group(1)), int(m.group(2)), int(m.group(3))))
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    lines = ["nvcc -Xptxas -v (nlsolver_b200/csrc/build/*.ptxas.log): registers / stack / spill bytes of every kernel",
             f"{len(rows)} kernels, {sum(1 for r in rows if r[3] or r[4])} with spill traffic (forced by the launch bounds of 64 / 80 registers).",
             "The kernels of the five BASELINE configurations are listed first (*); tests/test_build_cpu.py asserts that the",
             "fp64 long-row / reduction kernels of configs 2-5 do not spill at all, the d = 64 short-row kernels at most 16",
             "bytes, and no kernel more than 192 bytes.", ""]
    key = [r"de_tiny_solve_kernel<double, 4>", r"de_generation_bulk_kernel<double, 2, ", r"de_generation_bulk_kernel<double, 1, ",
           r"pso_move_kernel<double, 3, 1, 32,", r"de_generation_kernel<(double|float), 0, (8|16), 2,",
           r"pso_move_kernel<(double|float), 0, [01], (8|16), 2,", r"de_repair_kernel<(double|float), [0-3], (16|32),",
           r"de_commit_kernel", r"pso_candidate", r"pso_apply", r"pso_gather_apply"]
    def is_key(n):
        return any(re.search(k, n) for k in key)
    table = sorted(zip(names, rows), key=lambda t: (not is_key(t[0]), t[0]))
    for n, r in table:
        if is_key(n) or r[3] or r[4]:
            lines.append(f"{'*' if is_key(n) else ' '} regs {r[1]:3d} stack {r[2]:3d} spill st/ld {r[3]:3d}/{r[4]:3d}  {n.replace('nls::', '').split('(')[0]}")
    open(dst, "w").write("\n".join(lines) + "\n")


def main():
    os.makedirs(PROF, exist_ok=True)
    for src, dst in (("r2_s2e/bench_n2.json", "r2_bench_n2.json"), ("r2_s2e/group_n2.json", "r2_device_group_n2.json"),
                     ("r2_s2e/multi_gpu_check_n2.txt", "r2_multi_gpu_check_n2.txt"),
                     ("r2_s2e/example_multi_gpu.txt", "r2_example_multi_gpu_n2.txt"),
                     ("r2_n8/bench_n8.json", "r2_bench_n8.json"), ("r2_n8/bench_n4.json", "r2_bench_n4.json"),
                     ("r2_n8/bench_n1.json", "r2_bench_n1_same_box_as_n8.json"),
                     ("r2_n8/multi_gpu_check_n8.txt", "r2_multi_gpu_check_n8.txt"), ("r2_n8/group_n8.json", "r2_device_group_n8.json"),
                     ("r2_final/bench_n1.json", "r2_bench_n1.json"), ("r2_final/bench_ref.json", "r2_bench_reference_arm.json"),
                     ("r2_final/bench_small.json", "r2_bench_small.json"), ("r2_final/sann.json", "r2_sann_bench_n1.json"),
                     ("r2_sweep/sweep.md", "r2_sweep_d64.md")):
        copy(src, dst)
    for d in ("r2_s2b", "r2_final_ncu"):
        src = os.path.join(OUT, d)
        if not os.path.isdir(src):
            continue
        if os.path.exists(os.path.join(src, "launches.csv")):
            shutil.copy(os.path.join(src, "launches.csv"), os.path.join(PROF, "r2_launches_bench_n1.csv"))
            launches(os.path.join(src, "launches.csv"), os.path.join(src, "bench_plain.json"),
                     os.path.join(PROF, "r2_launches_bench_n1.txt"))
        sys.argv = ["ncu_summary", src, PROF, "r2"]
        ncu_summary.main()
        raw = os.path.join(src, "k2_config2.raw.csv")
        if os.path.exists(raw):
            _, kernels = ncu_summary.raw_summary(raw)
            tr = [v["traffic"] for _, v in kernels if "traffic" in v]
            path = os.path.join(PROF, "roofline_traffic.json")
            cur = json.load(open(path)) if os.path.exists(path) else {}
            cur["de_generation_bulk_kernel"] = {
                "dram_bytes_per_launch": sum(tr) / len(tr), "workload": "DE-random Rastrigin d=1000 P=1048576 fp64",
                "source": "profiles/r2_k2_config2_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of the captured launches)"}
            json.dump(cur, open(path, "w"), indent=1)
    sass_summary(os.path.join(PROF, "r2_sass_summary.txt"))
    spills(os.path.join(PROF, "r2_ptxas_registers_spills.txt"))
    print("profiles/ updated:", sorted(f for f in os.listdir(PROF) if f.startswith("r2_")))


if __name__ == "__main__":
    main()
