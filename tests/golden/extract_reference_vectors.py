#!/usr/bin/env python3
"""Extract the reference's own fixed vectors / constants for the hot path into
tests/golden/reference_vectors.json.  Run in the authoring container (reads /root/reference, which
does not exist on the GPU box); the JSON is committed.

Every `Bls12_381Base([...])` literal (6 x u64 little-endian limbs) inside the cited line range is
taken in source order.  The limbs are Montgomery-form (R = 2^384) values copied by the reference from
zkcrypto/bls12_381; they are stored here exactly as written.
"""
import json
import os
import re

REF = "/root/reference/src"
SPECS = [
    # name, file, first line, last line, expected number of Fp literals, labels
    ("fq2_add", "fields_as_trees/fq2_target_tree.rs", 220, 307, 6, ["a.c0", "a.c1", "b.c0", "b.c1", "c.c0", "c.c1"]),
    ("fq2_sub", "fields_as_trees/fq2_target_tree.rs", 335, 420, 6, ["a.c0", "a.c1", "b.c0", "b.c1", "c.c0", "c.c1"]),
    ("g1_generator", "fields_as_trees/g1_curve.rs", 50, 76, 2, ["x", "y"]),
    ("g2_generator", "fields_as_trees/g2_curve.rs", 60, 119, 4, ["x.c0", "x.c1", "y.c0", "y.c1"]),
    ("frob_fq12_c1", "fields_as_trees/fq12_target_tree.rs", 96, 124, 2, ["c0", "c1"]),
    ("frob_fq6_c1_c2", "fields_as_trees/fq6_target_tree.rs", 134, 166, None, None),
    ("fq6_arith_abc", "fields_as_trees/fq6_target_tree.rs", 391, 647, 18, None),
    ("fq12_arith_abc", "fields_as_trees/fq12_target_tree.rs", 447, 942, 36, None),
]

LIT = re.compile(r"Bls12_381Base\(\s*\[(.*?)\]\s*,?\s*\)", re.S)


def extract(path, lo, hi):
    lines = open(path).read().split("\n")[lo - 1:hi]
    text = "\n".join(lines)
    out = []
    for m in LIT.finditer(text):
        limbs = [int(x.strip().replace("_", ""), 16) for x in m.group(1).split(",") if x.strip()]
        assert len(limbs) == 6, (path, lo, m.group(0)[:60])
        out.append(["0x%016x" % v for v in limbs])
    return out


def main():
    res = {"_comment": "extracted by tests/golden/extract_reference_vectors.py from /root/reference/src; 6 x u64 LE limbs per Fp, as written in the reference"}
    for name, rel, lo, hi, count, labels in SPECS:
        vals = extract(os.path.join(REF, rel), lo, hi)
        if count is not None:
            assert len(vals) == count, (name, len(vals), count)
        res[name] = {"source": "%s:%d-%d" % ("src/" + rel, lo, hi), "labels": labels, "fp": vals}
        print(name, len(vals))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    json.dump(res, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
