#!/usr/bin/env python3
"""Write tests/golden/ark_inputs.txt: the first 16 pairs of pairs_256.npz (pair 0 = the generators) as canonical
hex, the input of tools/ark_vectors (the arkworks 0.4 golden-vector harness)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import b381_oracle as o  # noqa: E402

N = 16


def main():
    z = np.load(os.path.join(HERE, "pairs_256.npz"))
    with open(os.path.join(HERE, "ark_inputs.txt"), "w") as f:
        f.write("# px py qx.c0 qx.c1 qy.c0 qy.c1 (canonical, big-endian hex); pair i = (a_i G1, b_i G2) of pairs_256.npz\n")
        for i in range(N):
            g1, g2 = z["g1"][i].tolist(), z["g2"][i].tolist()
            vals = [o.fp_from_limbs32(g1[:12]), o.fp_from_limbs32(g1[12:])] + [o.fp_from_limbs32(g2[12 * k:12 * k + 12]) for k in range(4)]
            f.write(" ".join("0x%096x" % v for v in vals) + "\n")


if __name__ == "__main__":
    main()
