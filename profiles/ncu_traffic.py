#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture: DRAM bytes of ONE launch of the dominant kernel.

    python profiles/ncu_traffic.py gpurun_out/<capture>.ncu-rep|<raw page>.csv loss_main_kernel [profiles/<committed name>]

bench.py reports the numbers as roofline.traffic together with the capture they came from."""
import csv
import io
import json
import os
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
name = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
if rep.endswith(".csv"):     # the raw page already exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
kn, rd, wr, du = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                         "gpu__time_duration.sum"))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    if pat in r[kn]:
        res = {"capture": name, "kernel": r[kn][:100], "grid": r[hdr.index("Grid Size")],
               "dram_bytes_read": int(float(r[rd]) * scale[units[rd]]),
               "dram_bytes_write": int(float(r[wr]) * scale[units[wr]]),
               "duration_us_under_ncu": float(r[du]) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
                                                         "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[units[du]]}
        here = os.path.dirname(os.path.abspath(__file__))
        with open(os.path.join(here, "traffic.json"), "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
        break
else:
    sys.exit("no kernel matching %r in %s" % (pat, rep))
