"""A/B timing of built variants of libgroove_b200.so on the config-4 recipe (run under gpurun).

    python tools/ab_bench.py [--seconds 6] [--voices 4096] [--env K=V ...] lib_a.so lib_b.so ...

Each library is loaded in its own process (set before `groove_b200.engine` loads it), renders the
config-4 recipe `--reps` times with engine timing on, and reports the mean Welsh-kernel launch time and
render time.  The first library's output is the reference for the max-abs difference of the others.
Development tool only — never part of the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(lib: str, seconds: float, voices: int, reps: int, out_npy: str, filter_decay: float):
    sys.path.insert(0, ROOT)
    import numpy as np
    from groove_b200 import engine as eng_mod
    eng_mod.LIB_PATH = os.path.abspath(lib)
    from groove_b200 import Engine, workloads
    frames = int(round(seconds * 48000))
    cfg = workloads.Cfg4(total_voices=voices, frames=frames, note_off_base=int(frames * 2_400_000 / 2_880_000),
                         groups=min(128, voices))
    if filter_decay > 0:
        orig = workloads.cello_params

        def patched(*a, **k):
            p = orig(*a, **k)
            p.filter_envelope.decay = filter_decay
            p.filter_envelope.release = filter_decay
            return p
        workloads.cello_params = patched
    best = None
    out = np.empty((frames, 2))
    for _ in range(reps + 1):
        e = Engine(48000.0, device=0, max_block=1 << 16)
        e.set_timing(True)
        workloads.build_cfg4(e, cfg)
        e.render(frames, out)
        st = e.stats()
        r = {"render_ms": st.render_ms, "voice_kernel_ms": st.voice_kernel_ms, "launches": st.voice_kernel_launches,
             "rest_ms": getattr(st, "rest_kernel_ms", 0.0), "rest_launches": getattr(st, "rest_kernel_launches", 0),
             "rest_vs": getattr(st, "rest_voice_samples", 0), "fx_ms": st.fx_kernel_ms, "all_launches": st.kernel_launches}
        e.close()
        if best is None or r["voice_kernel_ms"] < best["voice_kernel_ms"]:
            best = r
    np.save(out_npy, out)
    best["vs_per_s"] = voices * frames / (best["render_ms"] * 1e-3)
    best["launch_ms"] = best["voice_kernel_ms"] / max(best["launches"], 1)
    print("RESULT " + json.dumps(best), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--seconds", type=float, default=6.0)
    ap.add_argument("--voices", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--filter-decay", type=float, default=0.0,
                    help="override the filter envelope decay (s): longer than the note = cutoff always moving")
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--child", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.child:
        child(a.child, a.seconds, a.voices, a.reps, a.out, a.filter_decay)
        return
    import numpy as np
    ref = None
    for i, lib in enumerate(a.libs):
        out = f"/tmp/ab_out_{i}.npy"
        env = dict(os.environ)
        for kv in a.env:
            k, v = kv.split("=", 1)
            env[k] = v
        p = subprocess.run([sys.executable, __file__, "--child", lib, "--out", out, "--seconds", str(a.seconds),
                            "--voices", str(a.voices), "--reps", str(a.reps), "--filter-decay", str(a.filter_decay)],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        res = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        if not res:
            print(f"{lib}: FAILED\n{p.stdout[-2000:]}")
            continue
        r = json.loads(res[0][7:])
        y = np.load(out)
        if ref is None:
            ref = y
        diff = float(np.abs(y - ref).max())
        print(f"{os.path.basename(lib):28s} launch {r['launch_ms']:.3f} ms  render {r['render_ms']:.2f} ms  "
              f"{r['vs_per_s']:.3e} vs/s  rest {r.get('rest_ms', 0) / max(r.get('rest_launches', 0), 1):.3f} ms x {r.get('rest_launches', 0)}  voice {r['voice_kernel_ms']:.2f} fx {r.get('fx_ms', 0):.2f} gap {r['render_ms'] - r['voice_kernel_ms'] - r.get('fx_ms', 0):.2f} ms ({r.get('all_launches', 0)} launches)  peak {np.abs(y).max():.4f}  maxdiff_vs_first {diff:.3e}", flush=True)


if __name__ == "__main__":
    main()
