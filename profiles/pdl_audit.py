#!/usr/bin/env python
"""List the global-memory instructions every kernel of libfsg_dense.so issues BEFORE its griddepcontrol.wait (SASS:
ACQBULK).  Under programmatic dependent launch only the step's INPUTS may be read there; a load of something the
preceding kernel writes is a race (nvcc hoists `const __restrict__` loads above the wait -- see
common.cuh:produced_by_dependency).  Usage: python profiles/pdl_audit.py [path/to/libfsg_dense.so]"""
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "full_scale_gambler_for_object_detection_b200",
    "libfsg_dense.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
total = 0
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    waits = [i for i, l in enumerate(lines) if "ACQBULK" in l]
    if not waits:
        continue
    pre = [re.sub(r"/\*.*?\*/", "", l).strip() for l in lines[:waits[0]] if re.search(r"\bLDG|\bLD\.|ATOM|\bRED|\bSTG", l)]
    dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print("%2d  %s" % (len(pre), dn[:110]))
    for l in pre:
        print("        " + l)
    total += len(pre)
print("global-memory instructions in front of a wait: %d" % total)
