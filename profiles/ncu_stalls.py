#!/usr/bin/env python3
"""Top instructions of an ncu report by warp-stall samples, summed over all kernel results in the report, with the stall reasons
and the CUDA source line.  usage: ncu_stalls.py <report.ncu-rep> <library.so> <kernel-mangled-substring> [top]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
amap = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    on, cur = False, None
    for ln in txt.split("\n"):
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            on = kern in m.group(1); continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m: amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
agg = collections.defaultdict(lambda: collections.Counter()); hdr = None; base = None
for r in rows:
    if r and r[0] == "Address":
        hdr = r; base = None; continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"): continue
    a = int(r[0], 16); base = a if base is None else base
    d = agg[a - base]
    d["samples"] += int(r[hdr.index("# Samples")] or 0); d["instr"] += int(r[hdr.index("Instructions Executed")] or 0)
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i]: d[h] += int(r[i])
tot = sum(d["samples"] for d in agg.values()); ti = sum(d["instr"] for d in agg.values())
print("total samples %d, warp instructions %d" % (tot, ti))
allst = collections.Counter()
for d in agg.values():
    for k, v in d.items():
        if k.startswith("stall_"): allst[k] += v
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, tot)) for k, v in allst.most_common(8)))
for off, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    (cur, sass) = amap.get(off, (("?", 0), "?"))
    st = ", ".join("%s %d" % (k[6:], v) for k, v in d.most_common(5) if k.startswith("stall_"))
    print("%5.2f%%  x%-7d %-14s:%-4d %-60s %s" % (100.0 * d["samples"] / max(1, tot), d["instr"], (cur or ("?", 0))[0], (cur or ("?", 0))[1], sass[:60], st))
