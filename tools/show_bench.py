#!/usr/bin/env python
"""Prints the headline numbers of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  ms/step %.2f  n_gpus %d  scaling %s" % (d["value"], d["unit"], d["ms_per_step"], d["n_gpus"], d["scaling"]))
e = d.get("e2e") or {}
print("e2e %.1f  sync-call %.1f  copy ceiling %.1f  equal %s" % (e.get("value", 0), e.get("sync_call_value", 0),
      e.get("copy_ceiling_pairs_per_s", 0), e.get("equals_device_resident_result")))
r = d.get("roofline") or {}
print("roofline hbm frac %.3f  dram frac %s  fp32 frac %s  share %.3f  avg launch %.3f ms" %
      (r.get("frac", 0), r.get("dram_frac"), r.get("fp32_frac"), r.get("share_of_step") or 0, r.get("avg_launch_ms") or 0))
print("alu", d.get("alu"))
print("stage", {k: round(v, 2) for k, v in d.get("stage_ms_per_step", {}).items()})
print("clocks", d.get("clocks"))
print("cpu_baseline", d.get("cpu_baseline"))
print("cpu_baseline_c1", d.get("cpu_baseline_c1"))
print("c5_band", d.get("c5_band"))
print("quality", d.get("quality"), "launches", d.get("gpu_launches"))
