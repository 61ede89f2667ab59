#!/usr/bin/env python3
"""big_parity.py [log2_n] [seed] -- one-off parity run on MANY distinct pairs (default 2^19): (a_i G1, b_i G2) made on the
device, b381_pairing and b381_miller_loop compared IN FULL with the oracle's C port on all host cores, the packed
prepared stage and the multi-Miller product on the same data.  Prints one JSON line (kept under profiles/).
Test infrastructure: this is the checker side (tests/test_gpu_distinct.py at a larger size), not a product path."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b381  # noqa: E402
import util  # noqa: E402
import test_gpu_distinct as T  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 19
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0xB16
n = 1 << logn
L = b381._lib
lib = L.init(0)
t0 = time.time()
a, b, p, q = T.device_points(L, lib, n, seed)
res = {"pairs": n, "seed": seed, "distinct_g1": int(len(np.unique(p.reshape(n, 24), axis=0))), "distinct_g2": int(len(np.unique(q.reshape(n, 48), axis=0)))}
# the generated points themselves: every one must lie in the prime-order subgroup (a wrong scalar multiplication would not)
f8 = np.zeros(n, dtype=np.uint8)
u8 = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
L.check(lib.b381_g1_in_subgroup(L.u32(p)[1], None, u8(f8), n)); res["g1_in_subgroup"] = int(f8.sum())
L.check(lib.b381_g2_in_subgroup(L.u32(q)[1], None, u8(f8), n)); res["g2_in_subgroup"] = int(f8.sum())
ref = util.load_ref_lib()
threads = os.cpu_count() or 8
out = np.zeros(n * 144, dtype=np.uint32)
chk = np.zeros(n * 144, dtype=np.uint32)
L.check(lib.b381_pairing(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
assert ref.ref_pairing(util.p32(p), util.p32(q), None, util.p32(chk), n, threads) == 0
res["pairing_mismatches"] = int((out.reshape(n, 144) != chk.reshape(n, 144)).any(axis=1).sum())
L.check(lib.b381_miller_loop(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
assert ref.ref_miller_loop(util.p32(p), util.p32(q), None, util.p32(chk), n, threads) == 0
res["miller_mismatches"] = int((out.reshape(n, 144) != chk.reshape(n, 144)).any(axis=1).sum())
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
dp = torch.from_numpy(p.view(np.int32)).to(dev); dq = torch.from_numpy(q.view(np.int32)).to(dev)
pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
dout = torch.empty(n * 144, dtype=torch.int32, device=dev)
L.check(lib.b381_g2_prepare_packed_dev(dq.data_ptr(), pk.data_ptr(), n, L.MODE_ARK, st))
L.check(lib.b381_miller_loop_packed_dev(dp.data_ptr(), pk.data_ptr(), None, dout.data_ptr(), n, L.MODE_ARK, 0, st))
L.check(lib.b381_check_dev(st))
res["packed_miller_mismatches"] = int((dout.cpu().numpy().view(np.uint32).reshape(n, 144) != chk.reshape(n, 144)).any(axis=1).sum())
o144 = torch.empty(144, dtype=torch.int32, device=dev); c144 = np.zeros(144, dtype=np.uint32)
assert ref.ref_multi_miller_loop(util.p32(p), util.p32(q), None, util.p32(c144), n, threads) == 0
L.check(lib.b381_multi_miller_loop_dev(dp.data_ptr(), dq.data_ptr(), None, o144.data_ptr(), n, L.MODE_ARK, st))
L.check(lib.b381_check_dev(st))
res["multi_miller_equal"] = bool(np.array_equal(o144.cpu().numpy().view(np.uint32), c144))
L.check(lib.b381_multi_miller_loop_packed_dev(dp.data_ptr(), pk.data_ptr(), None, o144.data_ptr(), n, 0, st))
L.check(lib.b381_check_dev(st))
res["multi_miller_packed_equal"] = bool(np.array_equal(o144.cpu().numpy().view(np.uint32), c144))
res["seconds"] = round(time.time() - t0, 1)
res["host_threads"] = threads
print(json.dumps(res))
sys.exit(0 if (res["g1_in_subgroup"] == n and res["g2_in_subgroup"] == n and res["pairing_mismatches"] == 0 and res["miller_mismatches"] == 0 and res["packed_miller_mismatches"] == 0 and res["multi_miller_equal"] and res["multi_miller_packed_equal"]) else 1)
