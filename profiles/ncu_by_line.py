#!/usr/bin/env python3
"""Aggregate an ncu report's warp-stall samples / executed instructions per CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <library.so> <kernel-mangled-substring> [top]
Needs ncu, cuobjdump and nvdisasm on PATH (works without a GPU).  The SASS page of the report gives samples per
instruction address; nvdisasm --print-line-info on the cubin inside the library gives address -> file:line.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
amap = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    on, cur, stack = False, None, []
    for ln in txt.split("\n"):
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            on = kern in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S+)", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = next(r for r in rows if r and r[0] == "Address")
iA, iS, iI, iT = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]; base = None
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= iT or not r[iA]:
        continue
    a = int(r[iA], 16)
    base = a if base is None else base
    k = amap.get(a - base, (("?", 0), ""))[0] or ("?", 0)
    for j, i in enumerate((iS, iI, iT)):
        v = int(r[i] or 0); agg[k][j] += v; tot[j] += v
print("total: samples %d, warp-instructions %d, thread-instructions %d (%.1f threads/instr)" % (tot[0], tot[1], tot[2], tot[2] / max(1, tot[1])))
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = k
    if f not in src:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "**", f), recursive=True)
        src[f] = open(c[0]).read().split("\n") if c else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print("%-16s %4d  samples %6d (%4.1f%%)  instr %9d  thr/instr %4.1f  %s" % (f, l, v[0], 100.0 * v[0] / max(1, tot[0]), v[1], v[2] / max(1, v[1]), text))
