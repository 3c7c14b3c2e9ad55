#!/usr/bin/env python
"""Per-CUDA-source-line digest of an .ncu-rep: stall samples and instructions executed, top lines."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
data, cur_file, h = [], "", None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        iN, iI = h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h is None or len(r) < len(h) or not r[0].isdigit():
        continue
    try:
        data.append((int(r[iN] or 0), int(r[iI] or 0), f"{cur_file}:{r[0]}", r[1].strip()[:120]))
    except ValueError:
        pass
ts = sum(d[0] for d in data); ti = sum(d[1] for d in data)
print(f"total samples {ts}, total warp instr {ti}")
print("--- by samples")
for s, i, a, src in sorted(data, reverse=True)[:top]:
    print(f"{100*s/max(ts,1):5.1f}%  inst {100*i/max(ti,1):5.1f}%  {a:24s} {src}")
print("--- by instructions")
for s, i, a, src in sorted(data, key=lambda d: -d[1])[:top // 2]:
    print(f"{100*s/max(ts,1):5.1f}%  inst {100*i/max(ti,1):5.1f}%  {a:24s} {src}")
