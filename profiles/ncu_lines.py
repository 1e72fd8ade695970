#!/usr/bin/env python3
"""Executed warp instructions and stall samples of an ncu report per CUDA source line, in source order, summed over the kernel results of
the report (SASS page joined with nvdisasm line info of the library).  usage: ncu_lines.py <report.ncu-rep> <library.so> <kernel-substring> [file-filter]
NCU_LINES_DEPTH=k attributes an inlined instruction to the k-th frame from the outside (0: the line of the kernel body, the default)."""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
flt = sys.argv[4] if len(sys.argv) > 4 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
amap = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    on, frames, fresh = False, [], True
    for ln in txt.split("\n"):
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            on = kern in m.group(1); continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # one "File" line per inline frame, innermost first; a new group starts after an instruction
            if fresh: frames = []; fresh = False
            frames.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            fresh = True
            if frames:
                depth = int(os.environ.get("NCU_LINES_DEPTH", "0"))
                outer = frames[max(0, len(frames) - 1 - depth)]
                amap[int(m.group(1), 16)] = (outer, frames[0])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
agg = collections.defaultdict(lambda: [0, 0]); hdr = None; base = None; nres = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r; base = None; nres += 1; continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"): continue
    a = int(r[0], 16); base = a if base is None else base
    cur = amap.get(a - base)
    key = cur[0] if cur else ("?", 0)
    agg[key][0] += int(r[hdr.index("Instructions Executed")] or 0); agg[key][1] += int(r[hdr.index("# Samples")] or 0)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("results %d, warp instructions %d, samples %d" % (nres, ti, ts))
for (f, l), (ni, ns) in sorted(agg.items()):
    if flt and flt not in f: continue
    print("%-16s %5d  instr %9d  samples %7d" % (f, l, ni, ns))
