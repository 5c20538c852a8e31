#!/bin/bash
# A/B of welsh_rest_kernel's voices-per-warp: config 4 with the default split (16-voice CTAs, two voices per warp,
# two CTAs per SM) against 32-voice CTAs with four voices per warp (GB_REST_NV=4, one CTA per SM).
mkdir -p gpurun_out
for spec in "0 2" "32 4" "0 2" "32 4"; do
  set -- $spec
  if [ "$1" = 0 ]; then unset GB_VPC; else export GB_VPC=$1; fi
  GB_REST_NV=$2 timeout 200 python bench.py --no-legs --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null > /tmp/nv.json
  VPC=$1 NV=$2 python - <<'PY'
import json, os
d = json.loads(open("/tmp/nv.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("vpc", os.environ["VPC"], "nv", os.environ["NV"], "ms", round(d["ms_per_step"], 2), "rest_launch_ms", round(r["launch_ms"], 4),
      "ctas", r.get("ctas_per_launch"), "mixdown", d.get("mixdown"))
PY
done
