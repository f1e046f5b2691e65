"""Summaries of the ncu pages exported on the GPU box (tools/r2_gpu_s2b.sh): for every `<name>.raw.csv` under the given
directory, the headline metrics per captured kernel; for `<name>.source.csv.gz`, the hottest SASS instructions by
stall samples and the static instruction mix of the hot loop.
usage: python tools/ncu_summary.py gpurun_out/r2_s2b profiles r2"""
import csv
import gzip
import os
import re
import sys
from collections import Counter

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out, kernels = [], []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        out.append("kernel: " + name)
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"  {w:88s} {r[i]:>20s} {units[i]}")
                try:
                    vals[w] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        if "dram__bytes_read.sum" in vals:
            t = vals.get("gpu__time_duration.sum")
            tr = vals["dram__bytes_read.sum"] + vals.get("dram__bytes_write.sum", 0.0)
            out.append(f"  -> DRAM traffic {tr / 1e9:.3f} GB per launch" + (f", {tr / t / 1e9:.0f} GB/s under ncu" if t else ""))
            vals["traffic"] = tr
        kernels.append((name, vals))
        out.append("")
    return out, kernels


def source_summary(path, top=14):
    rows = list(csv.reader(gzip.open(path, "rt")))
    out, seen = [], set()
    # the export holds one block per kernel: a "Kernel Name" line, a header line, then instructions
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name, hdr = rows[i][1], rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            i = j
            sig = (name, len(body), sum(int(b[hdr.index("# Samples")]) for b in body))
            if sig in seen:
                continue
            seen.add(sig)
            s_i, e_i, n_i = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            tot_s = sum(int(b[n_i]) for b in body) or 1
            tot_e = sum(int(b[e_i]) for b in body) or 1
            out.append(f"kernel: {name}")
            out.append(f"  {len(body)} SASS instructions, {tot_e} warp-instructions executed, {tot_s} stall samples")
            agg = Counter()
            for b in body:
                for k in stall_cols:
                    agg[hdr[k]] += int(b[k] or 0)
            st = sum(agg.values()) or 1
            out.append("  stall reasons (share of samples): " + ", ".join(f"{k[6:]} {100 * v / st:.1f}%" for k, v in agg.most_common(7)))
            mix = Counter()
            for b in body:
                parts = b[s_i].split()
                if not parts:
                    continue
                op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
                mix[op.split(".")[0]] += int(b[e_i])
            out.append("  executed instruction mix: " + ", ".join(f"{k} {100 * v / tot_e:.1f}%" for k, v in mix.most_common(12)))
            special = Counter()
            for b in body:
                m = re.match(r"(?:@!?U?P\d+\s+)?(UBLKCP|UTMALDG|UTMASTG|SYNCS|LDG|STG|LDS|STS|LDGSTS|ATOM|RED|BAR|MEMBAR|ERRBAR|CCTL)\S*", b[s_i].strip())
                if m:
                    special[m.group(0).split()[-1]] += 1
            out.append("  memory / sync instructions (static): " + ", ".join(f"{k} x{v}" for k, v in sorted(special.items())))
            out.append(f"  hottest instructions by stall samples (of {tot_s}):")
            for b in sorted(body, key=lambda b: -int(b[n_i]))[:top]:
                reasons = sorted(((int(b[k] or 0), hdr[k][6:]) for k in stall_cols), reverse=True)[:2]
                out.append(f"    {100 * int(b[n_i]) / tot_s:5.1f}%  {b[s_i].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in reasons if c))
            out.append("")
        else:
            i += 1
    return out


def main():
    src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    for f in sorted(os.listdir(src)):
        if not f.endswith(".raw.csv"):
            continue
        name = f[:-8]
        lines = [f"ncu --set full --clock-control none --import-source on  ({name}; command line in tools/r2_gpu_s2b.sh; B200)", ""]
        raw, _ = raw_summary(os.path.join(src, f))
        lines += raw
        sp = os.path.join(src, name + ".source.csv.gz")
        if os.path.exists(sp):
            lines += ["---- source page (SASS) ----"] + source_summary(sp)
        open(os.path.join(dst, f"{tag}_{name}_ncu_full.txt"), "w").write("\n".join(lines) + "\n")
        print("wrote", f"{tag}_{name}_ncu_full.txt")


if __name__ == "__main__":
    main()
