#!/usr/bin/env python
"""Register / spill table of every kernel in one translation unit: python profiles/ptxas_table.py csrc/<file>.cu"""
import os
import re
import subprocess
import sys

src = sys.argv[1]
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
       "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v", "-c", "-o", "/tmp/ptxas_table.o", src] + sys.argv[2:]
txt = subprocess.run(cmd, capture_output=True, text=True).stderr
# one block per entry function: from its "Compiling entry function" line to the next one (device functions that are
# kept out of line -- __noinline__ -- report their own "Function properties" inside the block of the kernel that follows
# their compilation; only the numbers after the kernel's own "Function properties for <mangled>" line are taken)
blocks = re.split(r"(?=ptxas info\s*: Compiling entry function ')", txt)
for b in blocks:
    m = re.match(r"ptxas info\s*: Compiling entry function '(\S+)'", b)
    if not m:
        continue
    mangled = m.group(1)
    p = re.search(r"Function properties for " + re.escape(mangled) + r"\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                  r"(\d+) bytes spill loads\n", b)
    u = re.search(r"Used (\d+) registers(?:.*?(\d+) bytes smem)?", b[p.end():] if p else b)
    if not p or not u:
        continue
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)
    print("%-70s regs %3s  stack %4s  spill %4s/%4s  smem %6s" % (name[:70], u.group(1), p.group(1), p.group(2),
                                                                   p.group(3), u.group(2) or "0"))
