"""Per-source-line instruction counts and stall samples from `ncu --page source --csv --print-source cuda,sass` output.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])   # (file, line) -> [inst, samples, src]
tot_i = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(range(len(hdr)), r))
    line = r[0]; src = r[1]
    try:
        ie = int(r[hdr.index("Instructions Executed")] or 0)
        ss = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    if line == "":
        continue
    a = agg[(cur_file, int(line))]
    a[0] += ie; a[1] += ss
    if src and not a[2]:
        a[2] = src.strip()
    tot_i += ie; tot_s += ss
print(f"total inst {tot_i}  samples {tot_s}")
for (f, l), (ie, ss, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:18s}:{l:4d}  inst {100 * ie / max(tot_i, 1):5.1f}%  samples {100 * ss / max(tot_s, 1):5.1f}%  {src[:110]}")
