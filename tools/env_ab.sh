#!/bin/bash
# A/B of engine switches on config 4 (or a shard of it): each argument is "ENV=VAL ENV=VAL ... [-- bench args]".
#   tools/env_ab.sh "GB_OVERLAP=0" "GB_OVERLAP=1" "GB_OVERLAP=1 -- --voices 512 --groups 16"
for spec in "$@"; do
  envs="${spec%%--*}"; args=""
  case "$spec" in *--*) args="${spec#*-- }";; esac
  env $envs timeout 300 python bench.py --no-legs --no-cpu-baseline --steps 3 --warmup 2 $args 2>/dev/null > /tmp/ab.json
  SPEC="$spec" python - <<'PY'
import json, os
try:
    d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1]); r = d["roofline"]; m = d.get("mixdown") or {}
    print(os.environ["SPEC"], "| ms", round(d["ms_per_step"], 2), "e2e_ms", round(d["e2e"].get("ms_per_step", 0), 2), "rest_launch_ms", round(r["launch_ms"], 4),
          "ctas", r.get("ctas_per_launch"), "mixdown_ms", round(m.get("ms_per_step", 0), 3))
except Exception as ex:
    print(os.environ["SPEC"], "| failed:", ex)
PY
done
