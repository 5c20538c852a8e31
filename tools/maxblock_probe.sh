#!/bin/bash
# Config 4 at different chunk sizes: do the CTA partials stay in L2 between the voice kernels and the table sum?
for mb in 65536 32768 16384 8192; do
  timeout 100 python bench.py --max-block $mb --no-legs --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null > /tmp/mb.json
  MB=$mb python - <<'PY'
import json, os
d = json.loads(open("/tmp/mb.json").read().strip().splitlines()[-1]); r = d["roofline"]; m = d.get("mixdown", {})
print("max_block", os.environ["MB"], "ms", round(d["ms_per_step"], 2), "e2e_ms", round(d["e2e"]["ms_per_step"], 2), "rest_launch_ms", round(r["launch_ms"], 3),
      "launches", d["gpu_launches"], "mix_ms", round(m.get("ms_per_step", 0), 3), "mix_GB/s", round(m.get("achieved", 0)))
PY
done
