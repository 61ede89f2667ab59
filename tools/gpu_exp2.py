"""scratch: time pairing / miller kernels of an experimental libb381 build (B381_LIB) and check parity."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import b381
L = b381._lib; lib = L.init(0)
z = np.load("tests/golden/pairs_256.npz")
dev = torch.device("cuda:0")
def run(n, what):
    perm = np.random.default_rng(1).integers(0, 256, size=n)
    d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][perm]).reshape(-1).view(np.int32)).to(dev)
    d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][perm]).reshape(-1).view(np.int32)).to(dev)
    out = torch.empty(n * 144, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    if what == "pairing":
        f = lambda: L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, out.data_ptr(), n, 0, st)); gold = z["pairing"]
    else:
        f = lambda: L.check(lib.b381_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, out.data_ptr(), n, 0, st)); gold = z["miller_ark"]
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    L.check(lib.b381_check_dev(st))
    ok = np.array_equal(out.view(n, 144)[:512].cpu().numpy().view(np.uint32), gold[perm[:512]])
    print("%s %s n=%d: %.1f ms  %.0f /s parity=%s" % (os.environ.get("B381_LIB", "default"), what, n, ms, n / ms * 1e3, ok), flush=True)
sm = 148
for mult in sys.argv[1:]:
    blk = int(os.environ.get("BLK", "128"))
    run(sm * blk * int(mult), "pairing")
    run(sm * blk * int(mult), "miller")
