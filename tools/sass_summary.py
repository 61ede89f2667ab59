#!/usr/bin/env python3
"""SASS opcode summary per kernel of libb381.so: the evidence that the hot path is IMAD.WIDE carry-chain code using
tensor memory as a scratchpad (LDTM / STTM) and no tensor-core or TMA instructions.
usage: python tools/sass_summary.py [libb381.so] > profiles/sass_summary_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plonky2-bls12-381-pairing_b200", "libb381.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
fn = None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        mm = re.search(r"\d+(k_\w+?)(?:E|I[A-Z])", name)
        fn = mm.group(1) if mm else name[:40]
        per.setdefault(fn, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        per[fn][m.group(1)] += 1
cols = ["IMAD.WIDE", "IMAD.HI", "IADD3.X", "LDTM", "STTM", "LDS", "STS", "LDG/LD", "STG/ST", "LDL", "STL", "BAR", "UTC*MMA", "UTMA*", "HMMA"]
print("%-22s %8s " % ("kernel", "total") + " ".join("%9s" % c for c in cols))
for k, c in per.items():
    tot = sum(c.values())
    def cnt(pred):
        return sum(v for o, v in c.items() if pred(o))
    vals = [cnt(lambda o: o.startswith("IMAD.WIDE")), cnt(lambda o: o.startswith("IMAD.HI")), cnt(lambda o: o.startswith("IADD3.X")),
            cnt(lambda o: o.startswith("LDTM")), cnt(lambda o: o.startswith("STTM")), cnt(lambda o: o.startswith("LDS")), cnt(lambda o: o.startswith("STS")),
            cnt(lambda o: o.startswith("LDG") or o.startswith("LD.")), cnt(lambda o: o.startswith("STG") or o.startswith("ST.")),
            cnt(lambda o: o.startswith("LDL")), cnt(lambda o: o.startswith("STL")), cnt(lambda o: o.startswith("BAR")),
            cnt(lambda o: o.startswith("UTC") and "MMA" in o), cnt(lambda o: o.startswith("UTMA") or o.startswith("UBLKCP")), cnt(lambda o: o.startswith("HMMA"))]
    print("%-22s %8d " % (k, tot) + " ".join("%9d" % v for v in vals))
