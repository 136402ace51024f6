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
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*?Used (\d+) registers.*?(\d+) bytes smem", re.S)
for m in pat.finditer(txt):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)
    print("%-70s regs %3s  stack %4s  spill %4s/%4s  smem %6s" % (name[:70], m.group(5), m.group(2), m.group(3),
                                                                   m.group(4), m.group(6)))
