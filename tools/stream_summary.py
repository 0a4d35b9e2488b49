#!/usr/bin/env python
"""gpurun_out/stream_dram_<tag>.csv (tools/profile_r2.sh) -> profiles/<rnd>_streaming_kernels.md: DRAM bytes,
time and GB/s of every streaming-kernel launch of the last step, against the measured copy bandwidth.
    python tools/stream_summary.py r4 r4"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "stream_dram_%s.csv" % tag))
                       if not l.startswith("==")))
h = rows[0]
ii, ki, mi, vi, ui = (h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launch = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("pm::", "")})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if "time" in r[mi]:
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[mi]] = v
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6531.6)
L = list(launch.values())
# the last step: launches after the last k_init_random (one per device pass)
last = max(i for i, d in enumerate(L) if "k_init_random" in d["name"])
first = max(i for i, d in enumerate(L[:last]) if "k_downscale2" in d["name"]) - 1
step = L[first:]
with open(os.path.join(ROOT, "profiles", "%s_streaming_kernels.md" % rnd), "w") as f:
    f.write("# Streaming kernels of one device pass (64 pairs of 1280x720, 2 levels): DRAM bytes and time per launch\n\n"
            "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none`\n"
            "(tools/profile_r2.sh %s); launch times under ncu are serialised and cold-cache. Peak = measured copy\n"
            "bandwidth %.1f GB/s (MEASURED_PEAKS.json).\n\n| kernel | ms | read MB | written MB | GB/s | of peak |\n"
            "|---|---:|---:|---:|---:|---:|\n" % (tag, peak))
    for d in step:
        rd, wr, ms = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0), d.get("gpu__time_duration.sum", 0)
        gbs = (rd + wr) / (ms * 1e-3) / 1e9 if ms else 0
        f.write("| %s | %.3f | %.0f | %.0f | %.0f | %.2f |\n" % (d["name"], ms, rd / 1e6, wr / 1e6, gbs, gbs / peak))
print(open(os.path.join(ROOT, "profiles", "%s_streaming_kernels.md" % rnd)).read())
