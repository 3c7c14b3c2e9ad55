#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the whole bench process:
warm-up + timed steps, un-captured).  Usage: launch_summary.py launches.csv [n_steps_in_the_list]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][:70]
    tot[name][0] += 1
    tot[name][1] += float(r[14].replace(",", "")) / 1e3
grand = sum(v[1] for v in tot.values())
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:72s} {n:5d} launches {us:10.1f} us {us / n:9.2f} us/launch {100 * us / grand:5.1f}%")
steps = sys.argv[2] if len(sys.argv) > 2 else "?"
print(f"total {grand:.1f} us over {sum(v[0] for v in tot.values())} launches ({steps} un-captured steps incl. warm-up)")
