"""Times every compiled reference project (tests/golden/plans) on the GPU engine and on the CPU oracle."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, project
from tests import plans
from tests.oracle_binding import OracleEngine
rows = []
for name in plans.PLAN_NAMES:
    plan = plans.load_plan(name)
    o = OracleEngine(plan.sample_rate); project.build_plan(o, plan, plans.sample)
    t = time.perf_counter(); ref = o.render(plan.frames); t_cpu = time.perf_counter() - t
    best = 1e9
    for rep in range(3):
        g = Engine(plan.sample_rate); project.build_plan(g, plan, plans.sample)
        t = time.perf_counter(); out = g.render(plan.frames); dt = time.perf_counter() - t
        st = g.stats(); g.close(); best = min(best, dt)
    err = float(np.abs(out - ref).max())
    rows.append(dict(project=name, frames=plan.frames, seconds=plan.frames / plan.sample_rate, gpu_s=best, cpu_oracle_s=t_cpu,
                     gpu_realtime=plan.frames / plan.sample_rate / best, launches=int(st.kernel_launches), max_abs_err=err))
    print(json.dumps(rows[-1]), flush=True)
