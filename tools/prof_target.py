#!/usr/bin/env python3
"""prof_target.py -- one short pass over the kernels of interest, for ncu (launch list / --set full captures):
    pairing (2 rounds), Miller loop, prepared Miller loop, multi-Miller, G1 MSM at 2^16 and 2^20, subgroup checks.
Run plain first (exit code 0), then under ncu; numbers printed under ncu are not bench values."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b381  # noqa: E402

L = b381._lib
lib = L.init(0)
z = np.load(os.path.join(ROOT, "tests", "golden", "pairs_256.npz"))
what = sys.argv[1] if len(sys.argv) > 1 else "all"
sm = torch.cuda.get_device_properties(0).multi_processor_count
st = torch.cuda.current_stream().cuda_stream
n = sm * 256 * 2
perm = np.random.default_rng(4).integers(0, 256, size=1 << 20)
d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][perm]).reshape(-1).view(np.int32)).cuda()
d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][perm]).reshape(-1).view(np.int32)).cuda()
dout = torch.empty(n * 144, dtype=torch.int32, device="cuda")
if what in ("all", "pairing"):
    L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st))
if what in ("all", "miller"):
    L.check(lib.b381_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st))
    co = torch.empty(n * L.G2PREP_WORDS, dtype=torch.int32, device="cuda")
    L.check(lib.b381_g2_prepare_dev(d2.data_ptr(), co.data_ptr(), n, 0, st))
    L.check(lib.b381_miller_loop_prepared_dev(d1.data_ptr(), co.data_ptr(), None, dout.data_ptr(), n, 0, 0, st))
    L.check(lib.b381_multi_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), 2 * n, 0, st))
if what in ("all", "packed"):
    n4 = 4 * n // 2                                  # one round of the four-pairs-per-thread multi loop
    pk = torch.empty(lib.b381_g2_packed_words(n4), dtype=torch.int32, device="cuda")
    L.check(lib.b381_g2_prepare_packed_dev(d2.data_ptr(), pk.data_ptr(), n4, 0, st))
    L.check(lib.b381_miller_loop_packed_dev(d1.data_ptr(), pk.data_ptr(), None, dout.data_ptr(), n, 0, 0, st))
    L.check(lib.b381_multi_miller_loop_packed_dev(d1.data_ptr(), pk.data_ptr(), None, dout.data_ptr(), n4, 0, st))
    sc = torch.randint(-(1 << 31), (1 << 31) - 1, ((1 << 20) * 8,), dtype=torch.int32, device="cuda")
    r2 = torch.empty(48, dtype=torch.int32, device="cuda"); rf2 = torch.empty(1, dtype=torch.uint8, device="cuda")
    L.check(lib.b381_g2_msm_dev(d2.data_ptr(), None, sc.data_ptr(), r2.data_ptr(), rf2.data_ptr(), 1 << 20, st))
if what in ("all", "groups"):
    nn = 1 << 20
    sc = torch.randint(-(1 << 31), (1 << 31) - 1, (nn * 8,), dtype=torch.int32, device="cuda")
    mp = torch.empty(nn * 24, dtype=torch.int32, device="cuda"); mf = torch.empty(nn, dtype=torch.uint8, device="cuda")
    L.check(lib.b381_g1_scalar_mul_dev(d1.data_ptr(), None, sc.data_ptr(), mp.data_ptr(), mf.data_ptr(), nn, st))
    r1 = torch.empty(24, dtype=torch.int32, device="cuda"); rf = torch.empty(1, dtype=torch.uint8, device="cuda")
    sc2 = torch.randint(-(1 << 31), (1 << 31) - 1, (nn * 8,), dtype=torch.int32, device="cuda")
    for logn in (16, 20):
        L.check(lib.b381_g1_msm_dev(mp.data_ptr(), None, sc2.data_ptr(), r1.data_ptr(), rf.data_ptr(), 1 << logn, st))
    f8 = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    L.check(lib.b381_g1_in_subgroup_dev(mp.data_ptr(), None, f8.data_ptr(), 1 << 16, st))
    L.check(lib.b381_g2_in_subgroup_dev(d2.data_ptr(), None, f8.data_ptr(), 1 << 16, st))
L.check(lib.b381_check_dev(st))
torch.cuda.synchronize()
print("prof_target ok:", what)
