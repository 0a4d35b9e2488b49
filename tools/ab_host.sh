#!/bin/bash
# A/B of the host path on the GPU box: each line = env settings; prints value, async e2e and the blocking call
while read -r envs; do
  [ -z "$envs" ] && continue
  out=$(env $envs python bench.py --no-cpu-baseline 2>/dev/null)
  echo "$envs :: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("value %.0f  e2e %.0f  sync %.0f  ceiling %.0f  equal %s" % (d["value"], e["value"], e["sync_call_value"], e["copy_ceiling_pairs_per_s"], e["equals_device_resident_result"]))')"
done
