#!/usr/bin/env python3
"""Attribute an ncu source-page profile of a kernel to its __noinline__ sub-functions using the
ELF symbol table of the profiled .so.  Usage: ncu_by_function.py file.ncu-rep lib.so kernel_substr"""
import collections, csv, io, re, subprocess, sys

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
funcs = []   # (offset, size, name)
for line in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+\S+\s+(\S+)", line)
    if not m:
        continue
    name = m.group(3)
    if kname not in name:
        continue
    off = int(m.group(1), 16) if m.group(1) != "0" else 0
    size = int(m.group(2), 16)
    short = name.split("$")[-1] if name.startswith("$") else "<kernel body>"
    funcs.append((off, size, short, name.startswith("$")))
subs = sorted([f for f in funcs if f[3]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {x: i for i, x in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h)]
base = int(data[0][ix["Address"]], 16)
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: collections.Counter())
for r in data:
    off = int(r[ix["Address"]], 16) - base
    name = "<kernel body>"
    for o, s, n, _ in subs:
        if o <= off < o + s:
            name = n
            break
    a = agg[name]
    a["samples"] += int(r[ix["# Samples"]]); a["inst"] += int(r[ix["Instructions Executed"]])
    op = r[ix["Source"]].strip(); op = (op.split()[1] if op.startswith("@") else op.split()[0])
    if op.startswith("IMAD"): a["imad"] += int(r[ix["Instructions Executed"]])
    for c in stall_cols:
        a[c] += int(r[ix[c]])
ts = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
import subprocess as sp
def dem(n):
    try:
        return sp.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
print("%-34s %7s %7s %7s %7s   top stalls" % ("function", "time%", "inst%", "imad%", "ipc*"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((c[6:], a[c]) for c in stall_cols), key=lambda x: -x[1])[:4]
    print("%-34s %6.1f%% %6.1f%% %6.1f%% %7.2f   %s" % (dem(name)[:34], 100.0 * a["samples"] / ts, 100.0 * a["inst"] / ti, 100.0 * a["imad"] / max(1, a["inst"]),
          (a["inst"] / ti) / max(1e-9, a["samples"] / ts), ", ".join("%s %.0f%%" % (n, 100.0 * v / max(1, a["samples"])) for n, v in st)))
