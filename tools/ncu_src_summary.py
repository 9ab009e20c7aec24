"""Summarise an `ncu --page source --csv` dump: stall samples per opcode and per stall reason (design aid)."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; byop = Counter(); cnt = Counter(); bystall = Counter(); opstall = {}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = int(r[idx['# Samples']]); tot += s
    src = r[idx['Source']].split()
    op = src[1] if src and src[0].startswith('@') else (src[0] if src else '')
    op = op.rstrip(';')
    byop[op] += s; cnt[op] += 1
    for st in stalls:
        v = int(r[idx[st]]); bystall[st] += v
        opstall.setdefault(op, Counter())[st] += v
print('total samples', tot)
for st, v in bystall.most_common(8): print(f"  {st:22s} {100*v/tot:5.1f}%")
for op, s in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 14):
    top = ', '.join(f"{k[6:]}={100*v/max(s,1):.0f}%" for k, v in opstall[op].most_common(3))
    print(f"{op:24s} samples={s:7d} ({100*s/tot:5.1f}%) n={cnt[op]:4d} per-instr={s/cnt[op]:6.0f}  [{top}]")
