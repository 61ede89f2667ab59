import sys, os, time, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch, ctypes
import b381
L = b381._lib; lib = L.init(0)
z = np.load("tests/golden/pairs_256.npz")
dev = torch.device("cuda:0")
def run(n, with_smi, label):
    perm = np.random.default_rng(1).integers(0, 256, size=n)
    d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][perm]).reshape(-1).view(np.int32)).to(dev)
    d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][perm]).reshape(-1).view(np.int32)).to(dev)
    out = torch.empty(n * 144, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, out.data_ptr(), n, 0, st))
    f(); torch.cuda.synchronize()
    p = None
    if with_smi:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", str(with_smi)], stdout=subprocess.PIPE, text=True)
        time.sleep(0.3)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    txt = ""
    if p:
        p.terminate(); txt = " | ".join(p.stdout.read().strip().split("\n")[:6])
    print("%s n=%d smi=%s: %.1f ms  %.0f pairings/s  %s" % (label, n, with_smi, ms, n / ms * 1e3, txt), flush=True)
run(1 << 17, 0, "a")
run(1 << 20, 0, "b")
run(1 << 20, 100, "c")
run(1 << 20, 500, "d")
run(1 << 20, 0, "e")
run(148 * 128 * 8, 0, "f")
