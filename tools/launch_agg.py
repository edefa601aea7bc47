"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) per kernel over one step between two marker kernels."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
marker = sys.argv[2] if len(sys.argv) > 2 else 'box_kernel'
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
marks = [i for i, (n, v) in enumerate(data) if marker in n]
seg = data[marks[-2]:marks[-1]]
agg = collections.OrderedDict()
for n, v in seg:
    a = agg.setdefault(n.split('(')[0][:64], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, (c, v) in agg.items())
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:66s} x{c:3d} {v / 1e3:9.1f} us  {100 * v / tot:5.1f}%')
print('total', round(tot / 1e3, 1), 'us over', len(seg), 'launches')
