#!/usr/bin/env python
"""SASS audit for programmatic dependent launch: in every kernel that contains griddepcontrol.wait (ACQBULK), list the
global loads that sit BEFORE it.  Loads of weights are expected there; an LDG.CONSTANT (ld.global.nc) is a load the compiler
was free to hoist above the wait -- none may read a predecessor's output (launch chain).
    python tools/pdl_audit.py [fastegnn_b200/_C/libfegnn.so]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "fastegnn_b200/_C/libfegnn.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, info = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        info[fn] = dict(pre=[], nc=[], acq=False, pre_exit=False)
        continue
    if fn is None:
        continue
    if "PREEXIT" in line:
        info[fn]["pre_exit"] = True
    if "ACQBULK" in line:
        info[fn]["acq"] = True
    if not info[fn]["acq"] and "LDG" in line:
        info[fn]["pre"].append(line.strip())
        if "CONSTANT" in line:
            info[fn]["nc"].append(line.strip())
bad = 0
for k, v in info.items():
    if v["acq"]:
        print(f"{k[:90]:90s} trigger={'yes' if v['pre_exit'] else 'NO '} loads before wait: {len(v['pre']):3d}  ld.global.nc before wait: {len(v['nc'])}")
        bad += len(v["nc"])
print("OK: no non-coherent load ahead of a griddepcontrol.wait" if bad == 0 else f"CHECK: {bad} non-coherent loads ahead of a wait")
sys.exit(1 if bad else 0)
