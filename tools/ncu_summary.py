#!/usr/bin/env python
"""Summarise ncu outputs into the small text files committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.md
    python tools/ncu_summary.py kernel gpurun_out/prof_row_c.ncu-rep > profiles/r1_k_sweep_row.md
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("pm::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    print("| **all** | %d | %.3f | 100%% |" % (sum(v[0] for v in agg.values()), tot))


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print("## %s\n" % name.split("(")[0])
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEYS:
            if k in h:
                print("| %s | %s | %s |" % (k, r[h.index(k)], units[h.index(k)]))
        print("\nWarp stall reasons (cycles per issued instruction):\n")
        st = []
        for i, n in enumerate(h):
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") \
                    and "not_issued" not in n:
                try:
                    st.append((float(r[i]), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, n in sorted(st, reverse=True)[:8]:
            print("- %s: %.2f" % (n, v))
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
