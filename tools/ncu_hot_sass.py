"""Top stall-sample SASS lines of an `ncu --page source --print-source sass --csv` export.
usage: python tools/ncu_hot_sass.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci['# Samples']]) for r in body)
print(rows[0][1][:100], "total samples", tot)
stall_cols = [h for h in hdr if h.startswith('stall_')]
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci['# Samples']]))[:top]
for i in sorted(idx):
    r = body[i]
    n = int(r[ci['# Samples']])
    st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100.0 * n / tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {st}")
