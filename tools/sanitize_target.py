#!/usr/bin/env python3
"""sanitize_target.py -- small invocations of the kernels added in round 2 (packed prepared stage, multi-Miller loops,
G2 bucket MSM, streaming Fp12 products), sized for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
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
st = torch.cuda.current_stream().cuda_stream
dev = "cuda"
for n in (1, 131, 300):
    idx = np.arange(n) % 256
    d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][idx]).reshape(-1).view(np.int32)).to(dev)
    d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][idx]).reshape(-1).view(np.int32)).to(dev)
    out = torch.empty(n * 144, dtype=torch.int32, device=dev)
    o144 = torch.empty(144, dtype=torch.int32, device=dev)
    pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(d2.data_ptr(), pk.data_ptr(), n, 0, st))
    L.check(lib.b381_miller_loop_packed_dev(d1.data_ptr(), pk.data_ptr(), None, out.data_ptr(), n, 0, 1, st))
    ok = np.array_equal(out.cpu().numpy().view(np.uint32).reshape(n, 144), z["pairing"][idx])
    L.check(lib.b381_miller_loop_packed_one_dev(d1.data_ptr(), pk.data_ptr(), None, out.data_ptr(), n, 0, 0, st))
    L.check(lib.b381_multi_miller_loop_packed_dev(d1.data_ptr(), pk.data_ptr(), None, o144.data_ptr(), n, 0, st))
    a = o144.clone()
    L.check(lib.b381_multi_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, o144.data_ptr(), n, 0, st))
    ok = ok and bool(torch.equal(a, o144))
    L.check(lib.b381_fp12_mul_dev(out.data_ptr(), out.data_ptr(), out.data_ptr(), n, st))
    if n >= 256:
        sc = torch.randint(-(1 << 31), (1 << 31) - 1, (n * 8,), dtype=torch.int32, device=dev)
        r2 = torch.empty(48, dtype=torch.int32, device=dev); rf = torch.empty(1, dtype=torch.uint8, device=dev)
        L.check(lib.b381_g2_msm_dev(d2.data_ptr(), None, sc.data_ptr(), r2.data_ptr(), rf.data_ptr(), n, st))
    L.check(lib.b381_check_dev(st))
    print("n = %d ok = %s" % (n, ok), flush=True)
print("sanitize_target done")
