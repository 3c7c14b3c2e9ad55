#!/usr/bin/env python
"""Assemble profiles/parity_report_<tag>.txt from what `pytest tests -m gpu` leaves under gpurun_out/ (parity_fullsize.txt,
parity_<mode>_<case>.txt, callers_train_single_epoch.txt):   python tools/parity_report.py r2_final "252 passed" """
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
note = sys.argv[2] if len(sys.argv) > 2 else ""
from tests.gpu_util import TOLERANCES  # noqa: E402

HEAD_BIAS = ("coord_mlp_r_virtual.0.bias", "coord_mlp_v_virtual.0.bias")
out = [f"Parity of the CUDA path against the fp64 oracle, {tag} (B200, `pytest tests -m gpu`{': ' + note if note else ''}).",
       "Outputs are judged on the UPDATE (x' - x, Z' - Z) relative to the largest entry of the reference update, after subtracting",
       "4 ulp(fp32) of max|x'|; gradients relative to the largest entry of each tensor.  gb = the first-Linear bias gradients of the",
       "two virtual coordinate heads (cancellation-dominated column sums, tests/gpu_util.py).  Stated tolerances (out / gin / gw / gb):",
       "  " + "    ".join(f"{m} {t.out:g} / {t.gin:g} / {t.gw:g} / {t.gb:g}" for m, t in TOLERANCES.items() if m != "tf32_all"), ""]
fs = os.path.join(ROOT, "gpurun_out", "parity_fullsize.txt")
if os.path.exists(fs):
    out.append("== BASELINE.json configs at the benchmarked sizes (tests/test_gpu_fullsize.py); last run of every case")
    last = {}
    for line in open(fs):
        m = re.match(r"(\S+ gain=\S+ \[\w+\])", line)
        if m:
            last[m.group(1)] = line.rstrip()
    out += list(last.values()) + [""]
out.append("== seeded batches (tests/test_gpu_model.py::test_seeded_batches_against_oracle): worst per class")
out.append(f"{'case':22s} {'mode':8s} {'update':>9s} {'in grads':>9s} {'w grads':>9s} {'head bias':>9s}  worst weight-gradient tensor")
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_*_*.txt"))):
    m = re.match(r"parity_(fp32|tf32x3|tf32)_(.+)\.txt", os.path.basename(f))
    if not m:
        continue
    mode, case = m.groups()
    upd = gin = gw = gb = 0.0
    worst = ""
    for line in open(f):
        mm = re.match(r"\S+ \[\w+\] (.+?): gpu (\S+)", line)
        if not mm:
            continue
        name, e = mm.group(1), float(mm.group(2))
        if "(update)" in name:
            upd = max(upd, e)
        elif name.startswith("gin."):
            gin = max(gin, e)
        elif name.endswith(HEAD_BIAS):
            gb = max(gb, e)
        elif name.startswith("gp.") and e > gw:
            gw, worst = e, name
    rows.append((case, mode, upd, gin, gw, gb, worst))
for case, mode, upd, gin, gw, gb, worst in sorted(rows):
    out.append(f"{case:22s} {mode:8s} {upd:9.2e} {gin:9.2e} {gw:9.2e} {gb:9.2e}  {worst}")
for mode in ("fp32", "tf32x3", "tf32"):
    sel = [r for r in rows if r[1] == mode]
    if sel:
        out.append(f"worst over the cases [{mode}]: update {max(r[2] for r in sel):.2e}  in grads {max(r[3] for r in sel):.2e}  "
                   f"w grads {max(r[4] for r in sel):.2e}  head bias {max(r[5] for r in sel):.2e}")
cf = os.path.join(ROOT, "gpurun_out", "callers_train_single_epoch.txt")
if os.path.exists(cf):
    out += ["", "== the reference's own callers on the drop-in (tests/test_gpu_callers.py): average epoch losses returned by "
            "utils/train.py::train_single_epoch"] + [l.rstrip() for l in open(cf)]
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "equivariance_*.txt"))):
    out += ["", f"== {os.path.basename(f)}"] + [l.rstrip() for l in open(f)][:12]
path = os.path.join(ROOT, "profiles", f"parity_report_{tag}.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path)
