"""groove-b200 CLI — the shape of `groove-cli` (src/bin/groove-cli.rs:24-53,115-152): load a project,
render it offline on the GPU, optionally write a 16-bit stereo WAV and print the --perf report.

    python -m groove_b200.cli PROJECT.json --assets /path/to/assets [--wav [OUT.wav]] [--perf] [--quiet]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

from . import project
from .engine import Engine


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="groove-b200", description=__doc__.split("\n")[0])
    ap.add_argument("input", nargs="+", help="project file(s), JSON or JSON5")
    ap.add_argument("--assets", required=True, help="directory holding patches/ and samples/ (reference layout)")
    ap.add_argument("-w", "--wav", nargs="?", const="", default=None, help="write a WAV (default name: input with .wav)")
    ap.add_argument("-p", "--perf", action="store_true", help="print performance information")
    ap.add_argument("-q", "--quiet", action="store_true")
    ap.add_argument("--sample-rate", type=float, default=44100.0)  # SampleRate::DEFAULT (src/lib.rs:30)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    loader = project.ProjectLoader(a.assets)
    for path in a.input:
        plan = loader.load(path, a.sample_rate)
        for note in plan.skipped:
            print(f"Warning: {note}", file=sys.stderr)
        eng = Engine(plan.sample_rate, device=a.device)
        project.build_plan(eng, plan, loader.sample)
        t0 = time.perf_counter()
        pcm = eng.render_pcm16(plan.frames)
        dt = time.perf_counter() - t0
        eng.close()
        if not a.quiet:
            print(f"{plan.title or os.path.basename(path)}: {plan.frames} frames ({plan.frames / plan.sample_rate:.2f} s)")
        if a.perf:  # groove-cli.rs:123-139
            ms = dt * 1e3
            print(f"Tempo: {plan.bpm}")
            print(f"Sample count: {plan.frames}")
            print(f"Elapsed    : {dt:.3f}s")
            print(f"Samples per msec    : {plan.frames / ms:.2f} (goal >{plan.sample_rate / 1000.0:.2f})")
            print(f"usec per sample     : {ms * 1e3 / max(plan.frames, 1):.2f} (goal <{1e6 / plan.sample_rate:.2f})")
        if a.wav is not None:
            out = a.wav or os.path.splitext(path)[0] + ".wav"
            project.write_wav16(out, pcm, plan.sample_rate)
            if not a.quiet:
                print(f"wrote {out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
