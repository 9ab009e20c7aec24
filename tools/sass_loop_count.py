"""Static instruction count of a kernel's loops from `cuobjdump -sass` output (design aid, CPU only).

    cuobjdump -sass -fun <mangled> file.o | python tools/sass_loop_count.py [inner_trip_count]

Finds backward branches, treats [target, branch] as a loop body and prints per-opcode counts; nested loops are
weighted by `inner_trip_count` (default 3 = gadget levels of BASELINE config 5).
"""
import re, sys
from collections import Counter
trip = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ins = []
for line in sys.stdin:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        addr = int(m.group(1), 16); txt = m.group(2).strip()
        parts = txt.split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        ins.append((addr, op, txt))
loops = []
for a, op, txt in ins:
    if op.startswith('BRA'):
        m = re.search(r"0x([0-9a-f]+)\s*$", txt)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) != a:
            loops.append((int(m.group(1), 16), a))
loops.sort(key=lambda l: l[1] - l[0])
print("loops (start, end, instrs):", [(hex(s), hex(e), sum(1 for a, _, _ in ins if s <= a <= e)) for s, e in loops])
if loops:
    outer = max(loops, key=lambda l: l[1] - l[0])
    inner = [l for l in loops if l != outer and outer[0] <= l[0] and l[1] <= outer[1]]
    cnt = Counter(); total = 0
    for a, op, _ in ins:
        if outer[0] <= a <= outer[1]:
            w = trip if any(s <= a <= e for s, e in inner) else 1
            cnt[op] += w; total += w
    print(f"outer loop weighted instruction count (inner x{trip}): {total}")
    heavy = 0
    for op, n in cnt.most_common(40):
        print(f"  {op:24s} {n}")
    w = {'IMAD': 2, 'IMAD.HI.U32': 4, 'IMAD.WIDE.U32': 5.6, 'IMAD.IADD': 2, 'IMAD.MOV.U32': 2, 'IMAD.MOV': 2, 'IMAD.SHL.U32': 2, 'IMAD.X': 2, 'IMAD.U32': 2}
    for op, n in cnt.items():
        heavy += w.get(op, 0) * n
    print(f"fma-heavy pipe cycles per warp (IMAD 2, IMAD.HI 4, IMAD.WIDE 5.6): {heavy:.0f}")
