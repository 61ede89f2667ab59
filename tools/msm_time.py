import sys, time, ctypes, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import b381
L = b381._lib; L.init(0); lib = L.lib()
z = np.load('tests/golden/pairs_256.npz')
u8p = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
for kp in (1 << 10, 1 << 15):
    idx = np.arange(kp) % 256
    p1 = np.ascontiguousarray(z['g1'][idx]).reshape(-1)
    sc = np.random.default_rng(7).integers(0, 1 << 32, size=kp * 8, dtype=np.uint64).astype(np.uint32)
    m1 = np.zeros(24, dtype=np.uint32); f1 = np.zeros(1, dtype=np.uint8)
    for name, fn in (("sum", lambda: lib.b381_g1_sum(L.u32(p1)[1], None, L.u32(m1)[1], u8p(f1), kp)),
                     ("msm", lambda: lib.b381_g1_msm(L.u32(p1)[1], None, L.u32(sc)[1], L.u32(m1)[1], u8p(f1), kp))):
        fn()
        t0 = time.perf_counter(); rc = fn(); dt = time.perf_counter() - t0
        print(kp, name, rc, round(dt * 1e3, 2), "ms", lib.b381_kernel_launches())
