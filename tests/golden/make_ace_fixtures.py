#!/usr/bin/env python3
"""Writes data/ace/*.acebin + data/ace/aceLib from the reference's bundled ACE files
(run where /root/reference exists; the GPU box has no /root/reference, so the card arrays travel as fixtures).

An .acebin file is the ACE card itself (ZAID, AW, TZ, NXS(16), JXS(32), XSS(:) as aceCard holds them after readFromFile,
NuclearData/DataDecks/ACE/aceCard_class.f90:1454-1534) in binary: "SBACE1\\0\\0", ZAID[16], AW f64, TZ f64, NXS[16] i32,
JXS[32] i32, n i64, XSS[n] f64.  The numbers are parsed from the text exactly as a Fortran list-directed read would
(Python float() = correctly rounded decimal -> binary64)."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/IntegrationTestFiles/"
FILES = [("1001.03", "1001JEF311.ace", 1779), ("92233.03", "92233JEF311.ace", 1), ("52126.03", "52126JEF311.ace", 1),
         ("91231.03", "91231JEF311.ace", 1), ("91232.03", "91232JEF311.ace", 1)]


def read_card(path, line):
    with open(path) as f:
        L = f.read().split("\n")
    L = L[line - 1:]
    zaid = L[0][:10].strip()
    aw, tz = float(L[0][10:22]), float(L[0][22:34])
    toks = " ".join(L[6:12]).split()
    nxs = [int(t) for t in toks[:16]]
    jxs = [int(t) for t in toks[16:48]]
    n = nxs[0]
    vals = []
    i = 12
    while len(vals) < n:
        vals.extend(L[i].split())
        i += 1
    xss = np.array([float(v) for v in vals[:n]], np.float64)
    return zaid, aw, tz, nxs, jxs, xss


def main():
    out = os.path.join(ROOT, "data", "ace")
    os.makedirs(out, exist_ok=True)
    lib = ["! ACE library of the fixtures (aceLibrary_mod.f90 format: NAME; LINE; PATH;) - paths relative to this file"]
    for name, fn, line in FILES:
        zaid, aw, tz, nxs, jxs, xss = read_card(REF + fn, line)
        b = fn.replace(".ace", ".acebin")
        with open(os.path.join(out, b), "wb") as f:
            f.write(b"SBACE1\0\0")
            f.write(zaid.encode().ljust(16, b"\0"))
            f.write(struct.pack("<dd", aw, tz))
            f.write(struct.pack("<16i", *nxs))
            f.write(struct.pack("<32i", *jxs))
            f.write(struct.pack("<q", len(xss)))
            f.write(xss.tobytes())
        lib.append("%sc; 1; %s;" % (name, b))
        print(name, zaid, aw, tz, len(xss))
    with open(os.path.join(out, "aceLib"), "w") as f:
        f.write("\n".join(lib) + "\n")


if __name__ == "__main__":
    sys.exit(main())
