"""Where a kernel's stall samples sit: the hottest SASS instructions and a histogram over address buckets.

    python tools/ncu_hot_spots.py <report.ncu-rep> [bucket_instructions=64] [top=40]

Reads `ncu -i <report> --page source --csv` (one captured kernel), prints per bucket of consecutive instructions
the share of samples, the executed warp instructions, the dominant stall reason and the opcodes that make up the
bucket (so the phases of a straight-line block — passes, scans, output stage, reduce — can be told apart), then
the `top` single instructions by samples.  Development tool; run on the GPU box (reports are tens of MB)."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        smp = int(r[col["# Samples"]] or 0)
        ex = int(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    op = (m.group(1) if m else "?").split(".")[0]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols}
    ins.append((r[col["Address"]], op, smp, ex, st, r[col["Source"]].strip()[:60]))
total = sum(i[2] for i in ins) or 1
print(f"{len(ins)} instructions, {total} samples")
for b in range(0, len(ins), bucket):
    chunk = ins[b:b + bucket]
    s = sum(i[2] for i in chunk)
    if s * 200 < total:
        continue
    ex = max(i[3] for i in chunk)
    st = collections.Counter()
    for i in chunk:
        st.update(i[4])
    ops = collections.Counter(i[1] for i in chunk)
    print(f"[{b:5d}..{b + len(chunk):5d}) samples {s / total * 100:5.1f}%  executed x{ex:<9d} stalls "
          + ", ".join(f"{k} {v / max(s, 1) * 100:.0f}%" for k, v in st.most_common(3))
          + "  | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(5)))
print("-- hottest instructions")
for a, op, smp, ex, st, src in sorted(ins, key=lambda i: -i[2])[:top]:
    best = max(st.items(), key=lambda kv: kv[1]) if st else ("", 0)
    print(f"{a:>8s} {smp / total * 100:5.2f}%  x{ex:<9d} {best[0]:18s} {src}")
