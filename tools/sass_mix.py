"""Static SASS instruction mix of an inlined call site, from the line info in the built library.

    python tools/sass_mix.py [--lib groove_b200/libgroove_b200.so] --kernel 'welsh_kernelILi8ELi2ELb0' --site 1026

Disassembles the library's cubin with `nvdisasm -gi`, keeps the instructions of the kernel whose
mangled name contains --kernel and whose inline chain ends at source line --site (e.g. the line of
the `welsh_block_simple<true>` call), and prints the opcode histogram plus a per-source-line
breakdown.  The specialised Welsh block is straight-line code, so the static count is the executed
count per 256-frame block of one voice (divide by kT = 8 for "per voice-sample").  No GPU needed.
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble(lib: str) -> str:
    tmp = tempfile.mkdtemp(prefix="sassmix_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], check=True, stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "groove_b200", "libgroove_b200.so"))
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--site", type=int, required=True, help="outermost source line of the inlined call")
    ap.add_argument("--per", type=float, default=8.0, help="divide counts by this (frames per lane)")
    ap.add_argument("--lines", action="store_true", help="per innermost source line")
    ap.add_argument("--all-runs", action="store_true", help="print every clone of the site, not just the first")
    a = ap.parse_args()
    text = disassemble(a.lib)
    in_kernel = False
    chain: list[tuple[str, int]] = []
    runs: list[dict] = []   # contiguous address runs (the compiler may clone the site, e.g. by loop peeling)
    last_addr = None
    line_re = re.compile(r'File "([^"]+)", line (\d+)')
    ins_re = re.compile(r"/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
    pending: list[tuple[str, int]] = []
    for ln in text.splitlines():
        if ln.startswith(".text."):
            in_kernel = a.kernel in ln
            continue
        if not in_kernel:
            continue
        if "//## File" in ln:
            m = line_re.findall(ln)
            # a run of annotation lines describes one chain, innermost first
            pending.append((os.path.basename(m[0][0]), int(m[0][1])))
            continue
        m = ins_re.search(ln)
        if not m:
            continue
        if pending:
            chain = pending
            pending = []
        if chain and chain[-1][1] == a.site:
            addr = int(m.group(1), 16)
            if last_addr is None or addr - last_addr > 0x400:
                runs.append({"start": addr, "ops": collections.Counter(), "by_line": collections.Counter()})
            last_addr = addr
            runs[-1]["ops"][m.group(2)] += 1
            runs[-1]["by_line"][chain[0]] += 1
            runs[-1]["end"] = addr
    for r in runs if a.all_runs else runs[:1]:
        ops, by_line = r["ops"], r["by_line"]
        total = sum(ops.values())
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        print(f"site line {a.site} @ {r['start']:#x}..{r['end']:#x} ({len(runs)} clone(s)): {total} instructions "
              f"({total / a.per:.1f} per unit), FP64 {fp64} ({fp64 / a.per:.1f} per unit)")
        for k, v in ops.most_common():
            print(f"  {k:10s} {v:6d}  {v / a.per:7.2f}")
        if a.lines:
            for (f, l), v in sorted(by_line.items()):
                print(f"  {f}:{l:<5d} {v}")


if __name__ == "__main__":
    main()
