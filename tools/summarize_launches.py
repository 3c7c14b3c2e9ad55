"""Per-kernel totals of one training step from an ncu launch list (gpu__time_duration.sum): python tools/summarize_launches.py <csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h, rows = rows[0], rows[1:]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
names = [r[ki] for r in rows]
ad = [i for i, n in enumerate(names) if "adam_step_inc" in n]
a, b = ad[-2] + 1, ad[-1] + 1
agg = collections.OrderedDict()
for r in rows[a:b]:
    n = r[ki].split("(")[0][-58:]
    t = float(r[vi].replace(",", "")) / 1000
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print(f"one step: {b - a} launches, {tot:.1f} us summed")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us {100 * t / tot:5.1f}% {c:3d}x {t / c:7.1f}  {n}")
