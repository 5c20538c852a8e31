"""stream_fx workload (1024 chains) under GB_FX_MINB = 1 / 2 / 3 (development tool; run under gpurun)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from groove_b200 import Engine, workloads
frames = 1 << 16
for minb in sys.argv[1:] or ["1", "2", "3"]:
    os.environ["GB_FX_MINB"] = minb
    out = []
    for rep in range(3):
        e = Engine(48000.0, device=0, max_block=frames)
        e.set_timing(True)
        e.push_events(workloads.build_fx_chains(e, 1024, frames))
        e.render_device(frames)
        st = e.stats()
        out.append((round(st.fx_kernel_ms, 3), round(st.render_ms, 2), st.kernel_launches, st.fx_batched_nodes))
        e.close()
    print("minb", minb, out, "GB/s (32 B/frame/chain)", round(32.0 * frames * 1024 / (out[-1][0] * 1e-3) / 1e9, 1), flush=True)
