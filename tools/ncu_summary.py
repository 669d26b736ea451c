"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(',', ''))
        if r[ui] == 'ns': v /= 1000
        d[r[ki].split('(')[0]].append(v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:30s} n={len(v):4d} avg={sum(v)/len(v):8.2f} us  max={max(v):8.2f}  share={100*sum(v)/tot:5.1f}%")
