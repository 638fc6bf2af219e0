#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launches.md
  python tools/summarize_ncu.py raw gpurun_out/prof.ncu-rep profiles/rNN_kernel.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"<.*", "", r["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:70]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n{len(rows)} launches, {tot/1e6:.3f} ms summed device time "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n| share | launches | mean us | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {v[1]/tot*100:.2f}% | {v[0]} | {v[1]/v[0]/1e3:.1f} | `{k}` |\n")
    print("wrote", dst)


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]][:80] for c in cols])
    print("wrote", dst, len(rows) - 2, "launches")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
