#!/bin/bash
# Per-GPU time of config 4's strong-scaling shards on ONE GPU (4096/N voices as whole 32-voice instruments) by
# voices-per-CTA split; CTAs of <= 8 voices take welsh_rest_tp_kernel unless GB_REST_TP=0.
#   tools/strong_probe.sh "voices groups vpc tp" ...
for spec in "$@"; do
  set -- $spec
  GB_VPC=$3 GB_REST_TP=$4 timeout 100 python bench.py --voices $1 --groups $2 --no-legs --no-cpu-baseline --steps 3 --warmup 1 2>/dev/null > /tmp/sp.json
  V=$1 G=$2 VPC=$3 TP=$4 python - <<'PY'
import json, os
d = json.loads(open("/tmp/sp.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("voices", os.environ["V"], "groups", os.environ["G"], "vpc", os.environ["VPC"], "rest_tp", os.environ["TP"], "ms", round(d["ms_per_step"], 2),
      "rest_launch_ms", round(r["launch_ms"], 3), "ctas", r.get("ctas_per_launch"), "eff_vs_45.4", round(45.4 / (4096 / int(os.environ["V"]) * d["ms_per_step"]), 3))
PY
done
