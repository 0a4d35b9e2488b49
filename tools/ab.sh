#!/bin/bash
# A/B of kernel variants on the GPU box: each line = env settings; prints pairs/s and stage ms
while read -r envs; do
  [ -z "$envs" ] && continue
  out=$(env $envs python bench.py --pairs-per-gpu 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null)
  echo "$envs :: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); s=d["stage_ms_per_step"]; print("%.0f pairs/s row %.2f col %.2f noise %.2f copy %.2f" % (d["value"], s["sweep_row"], s["sweep_col"], s["noise_cost"], s["plane_copy"]))')"
done
