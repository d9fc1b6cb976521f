import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; data = []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
agg = collections.defaultdict(lambda: [0, 0.0])
for d in data:
    name = re.sub(r'\(.*', '', d['Kernel Name'])
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches, total {tot:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v[1]:9.3f} ms {v[0]:4d}x {v[1]/v[0]:8.3f} ms/launch {100*v[1]/tot:5.1f}%  {k[:90]}")
