#!/usr/bin/env python3
"""Static schedule summary of the __noinline__ primitives inside one kernel of libb381.so.

    cuobjdump -xelf all libb381.so ; nvdisasm -hex -c kernels.sm_100a.cubin > kernels.sass
    python tools/sass_sched.py kernels.sass k_pairing [f2_sop f2_cyc_fp4 ...]

For every sub-function of the kernel (the `$kernel$function:` symbols nvcc emits for noinline device
functions) it prints the instruction count, the sum of the stall counts ptxas encoded (bits 41..44 of the
high control word: the issue-time of ONE warp running alone, ignoring scoreboard waits and loops), the
multiplier-pipe time (4 cycles per IMAD.WIDE / IMAD.HI, 2 per other fma-pipe instruction) and the opcode mix.
single-warp stall sum >> pipe time means the primitive is latency-bound and needs more warps (or more
independent chains) to fill the pipe.
"""
import collections
import re
import sys


def parse(path, kernel):
    funcs = collections.OrderedDict()
    cur = None
    in_kernel = False
    pend = None
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    hi_re = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    with open(path) as f:
        for line in f:
            if line.startswith(".text."):
                in_kernel = kernel in line
                cur = None
                continue
            if not in_kernel:
                continue
            m = re.match(r"^(\$?[\w$]+):\s*$", line)
            if m:
                name = m.group(1)
                if name.startswith(".L"):
                    continue
                short = name.split("$")[-1] if "$" in name else "<kernel body>"
                mm = re.search(r"b381\d+(\w+?)E", short)
                cur = mm.group(1) if mm else short
                # strip the itanium length prefix artefacts
                cur = re.sub(r"^_ZN4b381\d+", "", cur)
                funcs.setdefault(cur, [])
                continue
            m = ins_re.match(line)
            if m and cur is not None:
                pend = (m.group(2).strip(), int(m.group(3), 16))
                continue
            m = hi_re.match(line)
            if m and pend is not None and cur is not None:
                hi = int(m.group(1), 16)
                funcs[cur].append((pend[0], pend[1], hi))
                pend = None
    return funcs


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0]


def summarize(name, ins):
    n = len(ins)
    stall = 0
    ops = collections.Counter()
    for text, lo, hi in ins:
        st = (hi >> 41) & 0xF
        stall += max(st, 1)
        ops[opcode(text)] += 1
    wide = sum(c for o, c in ops.items() if o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))
    fma_other = sum(c for o, c in ops.items() if (o.startswith("IMAD") and not (o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))) or o.startswith("HFMA2") or o.startswith("FFMA"))
    pipe = 4 * wide + 2 * fma_other
    print("%-22s inst %6d  stall-sum %7d  imad.wide %5d  fma-pipe cycles %6d  pipe/stall %.2f" % (name, n, stall, wide, pipe, pipe / max(stall, 1)))
    top = ", ".join("%s %d" % (o, c) for o, c in ops.most_common(12))
    print("    " + top)


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    want = sys.argv[3:]
    funcs = parse(path, kernel)
    for name, ins in funcs.items():
        if want and name not in want:
            continue
        if ins:
            summarize(name, ins)


if __name__ == "__main__":
    main()
