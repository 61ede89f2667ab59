#!/usr/bin/env python3
"""big_parity_dump.py log2_n seed out.json -- like big_parity.py (pairing only), but dumps every mismatching pair:
index, the scalars that made it, the points, the Miller value, the GPU and the C-port pairing values."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b381  # noqa: E402
import util  # noqa: E402
import test_gpu_distinct as T  # noqa: E402

logn, seed, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
n = 1 << logn
L = b381._lib
lib = L.init(0)
a, b, p, q = T.device_points(L, lib, n, seed)
ref = util.load_ref_lib()
threads = os.cpu_count() or 8
out = np.zeros(n * 144, dtype=np.uint32)
chk = np.zeros(n * 144, dtype=np.uint32)
L.check(lib.b381_pairing(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
assert ref.ref_pairing(util.p32(p), util.p32(q), None, util.p32(chk), n, threads) == 0
bad = np.nonzero((out.reshape(n, 144) != chk.reshape(n, 144)).any(axis=1))[0]
mil = np.zeros(n * 144, dtype=np.uint32)
L.check(lib.b381_miller_loop(L.u32(p)[1], L.u32(q)[1], None, L.u32(mil)[1], n, L.MODE_ARK))
dump = []
for i in bad.tolist():
    one = {"index": i, "a": a[i].tolist(), "b": b[i].tolist(), "g1": p[24 * i:24 * i + 24].tolist(), "g2": q[48 * i:48 * i + 48].tolist(),
           "miller": mil[144 * i:144 * i + 144].tolist(), "gpu": out[144 * i:144 * i + 144].tolist(), "cport": chk[144 * i:144 * i + 144].tolist()}
    # the same pair alone, through the separate kernels
    o1 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(p[24 * i:24 * i + 24])[1], L.u32(q[48 * i:48 * i + 48])[1], None, L.u32(o1)[1], 1, L.MODE_ARK))
    one["gpu_alone"] = o1.tolist()
    L.check(lib.b381_final_exp(L.u32(mil[144 * i:144 * i + 144])[1], L.u32(o1)[1], 1))
    one["gpu_final_exp_of_miller"] = o1.tolist()
    dump.append(one)
json.dump({"pairs": n, "seed": seed, "mismatches": dump}, open(path, "w"))
print("mismatches:", bad.tolist())
