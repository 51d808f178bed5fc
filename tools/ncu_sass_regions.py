"""Per-region totals of an `ncu --page source --csv` dump (SASS view): regions are cut at BAR.SYNC, loop back edges and
EXIT; prints executed warp instructions, stall samples and the opcode mix of each region.
usage: python tools/ncu_sass_regions.py dump.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    ins.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i in stall_cols]))
base = ins[0][0]
tot_s = sum(x[2] for x in ins)
tot_e = sum(x[3] for x in ins)
print("total samples", tot_s, "executed", tot_e)
cuts = {0}
for k, (a, src, s, e, st) in enumerate(ins):
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    if op.startswith("BAR") or op.startswith("EXIT"):
        cuts.add(k + 1)
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", src)
        if m:
            t = int(m.group(1), 16)
            if t <= a - base:
                cuts.add(k + 1)
                for j, x in enumerate(ins):
                    if x[0] - base == t:
                        cuts.add(j)
cuts = sorted(cuts) + [len(ins)]
for c0, c1 in zip(cuts[:-1], cuts[1:]):
    seg = ins[c0:c1]
    if not seg:
        continue
    s = sum(x[2] for x in seg)
    e = sum(x[3] for x in seg)
    if s < tot_s * 0.003 and e < tot_e * 0.003:
        continue
    mix = collections.Counter()
    for x in seg:
        src = x[1]
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        mix[op] += 1
    st = [sum(x[4][i] for x in seg) for i in range(len(stall_cols))]
    top = sorted(((v, hdr[stall_cols[i]][6:]) for i, v in enumerate(st)), reverse=True)[:4]
    print(f"[{seg[0][0]-base:#06x}..{seg[-1][0]-base:#06x}] n={len(seg):4d} exec {e/tot_e*100:5.1f}% samples {s/tot_s*100:5.1f}%  "
          f"stalls {', '.join(f'{n} {v/max(s,1)*100:.0f}%' for v, n in top)}  mix {dict(mix.most_common(8))}")
