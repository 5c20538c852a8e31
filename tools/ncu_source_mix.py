"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions by opcode, stall samples,
and the hottest address ranges."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); smp = collections.Counter(); tot = 0; tsm = 0
live = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex]); sm = int(r[ismp] or 0)
    except ValueError: continue
    m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', r[isrc])
    op = m.group(1) if m else '?'
    ops[op] += ex; smp[op] += sm; tot += ex; tsm += sm
    if ex > 0: live += 1
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp-instr {tot:.4g}; live static instrs {live}; samples {tsm}")
for op, v in ops.most_common(28):
    extra = f"  {v*32/units:7.1f}/unit" if units else ""
    print(f"{op:12s} {v:14d} {v/tot*100:5.1f}%  stall-samples {smp[op]/max(tsm,1)*100:5.1f}%{extra}")
