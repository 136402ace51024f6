#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing of one kernel with the CUDA source lines of the built library.

    python profiles/sass_lines.py <report.ncu-rep> <kernel name regex> [top N]

ncu's CSV source page is SASS-only; the line table comes from `nvdisasm -g` on the cubin embedded in
libfsg_dense.so (built with -lineinfo).  Prints stall samples and executed instructions aggregated per source
line (inlined callees are attributed to their own lines)."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname, hdr, sass = None, None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if kname is not None:
            break                     # first matching launch only
        kname = r[1]
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr):
        sass.append(r)
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
mangled_hint = re.sub(r"[<(].*", "", kname.replace("void ", "")).split("::")[-1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "full_scale_gambler_for_object_detection_b200",
                                                          "libfsg_dense.so")], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if mangled_hint not in dis:
        continue
    cur, infn, per = None, False, []
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = mangled_hint in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            per.append(cur)
    if per and abs(len(per) - len(sass)) <= 4:
        lines = per
        break
    # several instantiations of a template live in one cubin: take the one whose length matches
    cur, infn, per, best = None, False, [], None
    for ln in dis.splitlines() + [".text.END:"]:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if infn and abs(len(per) - len(sass)) <= 4:
                best = per
            infn, per = mangled_hint in m.group(1), []
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            per.append(cur)
    if best:
        lines = best
        break
if lines is None or abs(len(lines) - len(sass)) > 4:
    print("could not align SASS (%d) with the line table (%s)" % (len(sass), None if lines is None else len(lines)))
    sys.exit(1)
agg = collections.defaultdict(lambda: [0, 0])
for i, r in enumerate(sass):
    key = lines[min(i, len(lines) - 1)]
    agg[key][0] += int(r[si] or 0)
    agg[key][1] += int(r[ii] or 0)
tot = sum(v[0] for v in agg.values()) or 1
srcs = {}
print("%s: %d SASS instructions, %d samples" % (kname[:70], len(sass), tot))
for key, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key:
        p = os.path.join(ROOT, "full_scale_gambler_for_object_detection_b200", "csrc", key[0])
        if key[0] not in srcs and os.path.isfile(p):
            srcs[key[0]] = open(p).read().splitlines()
        if key[0] in srcs and key[1] - 1 < len(srcs[key[0]]):
            text = srcs[key[0]][key[1] - 1].strip()[:90]
    print("%5.1f%% %9d inst  %s:%s  %s" % (100.0 * s / tot, n, key[0] if key else "?", key[1] if key else "?", text))
