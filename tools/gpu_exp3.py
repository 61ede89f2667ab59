"""scratch: ms per round as a function of rounds per launch."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import b381
L = b381._lib; lib = L.init(0)
z = np.load("tests/golden/pairs_256.npz")
dev = torch.device("cuda:0")
blk = int(os.environ.get("BLK", "256"))
nmax = 148 * blk * 32
perm = np.random.default_rng(1).integers(0, 256, size=nmax)
d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][perm]).reshape(-1).view(np.int32)).to(dev)
d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][perm]).reshape(-1).view(np.int32)).to(dev)
out = torch.empty(nmax * 144, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rounds in (1, 2, 4, 8, 16, 28, 32, 1, 4):
    n = 148 * blk * rounds
    f = lambda: L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, out.data_ptr(), n, 0, st))
    if rounds == 1: f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rounds=%2d n=%7d: %8.1f ms  %.2f ms/round  %.0f /s" % (rounds, n, ms, ms / rounds, n / ms * 1e3), flush=True)
n = 148 * blk * 4
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(8):
    L.check(lib.b381_pairing_dev(d1.data_ptr() + k * n * 96, d2.data_ptr() + k * n * 192, None, out.data_ptr() + k * n * 576, n, 0, st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("8 launches x 4 rounds: %.1f ms  %.2f ms/round" % (ms, ms / 32))
