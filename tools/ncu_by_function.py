"""Stall samples / executed instructions of one kernel by device function (development tool).

    python tools/ncu_by_function.py <report.ncu-rep> <kernel-name-substring> [lib.so]

ncu's source page gives per-instruction samples by address; nvdisasm gives the offsets of the out-of-line device
functions inside the kernel's .text section.  Prints one line per function: share of samples, of executed warp
instructions, and the dominant stall reasons."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "groove_b200", "libgroove_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the kernel's section, then labels of the form $kernel$function: with the offset of the next instruction
labels, in_kernel, pending = [], False, None
for line in dis:
    m = re.match(r"^(\$?[_A-Za-z][^\s:]*):", line)
    if m:
        name = m.group(1)
        if not name.startswith("$") and not name.startswith(".L"):
            in_kernel = kname in name
            if in_kernel:
                pending = "<kernel body>"
        elif in_kernel and name.startswith("$"):
            pending = name.split("$")[-1]
        continue
    if in_kernel and pending:
        m = re.search(r"/\*([0-9a-f]{4,})\*/", line)
        if m:
            labels.append((int(m.group(1), 16), pending))
            pending = None
labels.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = {}
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    off = a - base
    fn = "<kernel body>"
    for o, n in labels:
        if o <= off:
            fn = n
        else:
            break
    d = agg.setdefault(fn, {"samples": 0, "inst": 0, "n": 0, **{s: 0 for s in stall_cols}})
    d["samples"] += int(r[col["# Samples"]] or 0)
    d["inst"] += int(r[col["Instructions Executed"]] or 0)
    d["n"] += 1
    for s in stall_cols:
        d[s] += int(r[col[s]] or 0)
ts = sum(d["samples"] for d in agg.values()) or 1
ti = sum(d["inst"] for d in agg.values()) or 1
print(f"{'function':60s} {'static':>7s} {'samples':>8s} {'inst':>7s}  top stalls")
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((d[s], s[6:]) for s in stall_cols), reverse=True)[:4]
    short = re.sub(r"^_ZN3gbk\d+", "", fn)[:60]
    print(f"{short:60s} {d['n']:7d} {100 * d['samples'] / ts:7.1f}% {100 * d['inst'] / ti:6.1f}%  " +
          ", ".join(f"{n} {100 * v / max(d['samples'], 1):.0f}%" for v, n in top if v))
