#!/usr/bin/env python
"""Turns the raw outputs of tools/profile_r2.sh <tag> (gpurun_out/) into the committed summaries:
profiles/%s_launch_list_64pairs.md, r2_launches_64pairs.csv, r2_sweep_dram_bytes.json,
r2_ncu_k_sweep_row3.md, r2_ncu_k_sweep_col.md.      python tools/profiles_refresh.py r2e"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"   # prefix of the committed files
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

rows = list(csv.reader(l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if not l.startswith("==")))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    seq.append((r[ki].split("(")[0].replace("void ", "").replace("pm::", ""), v))
last = max(i for i, (n, _) in enumerate(seq) if "k_fma_peak" in n)
step = [x for x in seq[last + 1:] if not x[0].startswith("at::")]
agg = collections.OrderedDict()
for n, v in step:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, "%s_launch_list_64pairs.md" % rnd), "w") as f:
    f.write("# Launch list of one device pass (64 pairs of 1280x720, D = 128, 3 iterations, 2 levels, random init)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --pairs-per-gpu 64 --steps 1\n"
            "--warmup 3 --no-cpu-baseline --no-e2e` (tools/profile_r2.sh %s); the launches after the FP32 probe = the\n"
            "timed step. Per-launch times under ncu are serialised and cold-cache: the SHARES are what agrees with\n"
            "the bench (`roofline.share_of_step`).\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n" % tag)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, v[0], v[1], 100 * v[1] / tot))
    f.write("| **all** | %d | %.3f | 100%% |\n" % (sum(v[0] for v in agg.values()), tot))
    sw = sum(v[1] for k, v in agg.items() if "k_sweep" in k)
    f.write("\nSweep kernels: %.3f ms = %.1f %% of the step.\n" % (sw, 100 * sw / tot))
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "%s_launches_64pairs.csv" % rnd))

rows = list(csv.reader(l for l in open(os.path.join(G, "sweep_dram_%s.csv" % tag)) if not l.startswith("==")))
h = rows[0]
ki, mi, vi, ui, ii = (h.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = {}
for r in rows[1:]:
    if len(r) <= vi:
        continue
    d = per.setdefault(r[ii], {"name": r[ki].split("(")[0]})
    v, u = float(r[vi].replace(",", "")), r[ui]
    if "byte" in u.lower():
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif u in ("ns", "us", "ms", "s"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[u]
    d[r[mi]] = v
ids = sorted(per, key=lambda x: int(x))[-24:]
out, total = [], 0.0
for i in ids:
    d = per[i]
    total += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    out.append({"kernel": d["name"].replace("void pm::", ""), "ms": round(d["gpu__time_duration.sum"], 4),
                "dram_read": d["dram__bytes_read.sum"], "dram_write": d["dram__bytes_write.sum"]})
json.dump({"workload": {"pairs_per_gpu": 64, "width": 1280, "height": 720, "pyramid_levels": 2, "iters": 3},
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control "
                     "none -k regex:k_sweep (tools/profile_r2.sh %s): the 24 sweep launches of one device pass of 64 pairs" % tag,
           "mean_dram_bytes_per_launch": total / 24, "launches": out},
          open(os.path.join(P, "%s_sweep_dram_bytes.json" % rnd), "w"), indent=1)
for what, name in (("row", "%s_ncu_k_sweep_row3.md" % rnd), ("col", "%s_ncu_k_sweep_col.md" % rnd)):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "kernel",
                          os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, what))], capture_output=True, text=True).stdout
    open(os.path.join(P, name), "w").write(txt)
print("mean DRAM GB per sweep launch: %.3f; sweeps %.1f %% of the step" % (total / 24 / 1e9, 100 * sw / tot))
