"""Summarise an `ncu --page source --csv` dump: total samples per stall reason and the hottest instructions."""
import csv, sys
src = list(csv.reader(open(sys.argv[1])))
h = src[1]
isrc, isamp = h.index('Source'), h.index('# Samples')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = {c: 0 for _, c in stall_cols}
rows = [r for r in src[2:] if len(r) > isamp and r[isamp].isdigit()]
for r in rows:
    for i, c in stall_cols:
        try: tot[c] += int(r[i])
        except ValueError: pass
total = sum(tot.values())
print('total samples', total)
for c, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print(f'  {c:26s} {v:9d} {100.0*v/total:5.1f}%')
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print('hottest instructions (samples, dominant stall, sass):')
for r in sorted(rows, key=lambda r: -int(r[isamp]))[:n]:
    dom = max(stall_cols, key=lambda ic: int(r[ic[0]]) if r[ic[0]].isdigit() else 0)
    print(f'  {r[isamp]:>7s} {dom[1]:22s} {r[isrc][:80]}')
