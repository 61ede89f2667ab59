#!/usr/bin/env python3
"""quick_bench.py -- A/B timing of experiment builds of libb381.so on one GPU.

    python tools/quick_bench.py _variants/lib_a.so _variants/lib_b.so ...        (run under gpurun)

Each library is loaded in its own subprocess (B381_LIB), checked against the golden fixture on 1024 tiled
pairs (pairing, Miller loop), then timed with CUDA events on device-resident buffers: full pairings
(8 rounds of #SM x 256 pairs), Miller loops, prepared Miller loops, multi-Miller.  One JSON line per library.
A number printed here is an experiment, not a bench value (bench.py is the contract).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(path):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import b381
    L = b381._lib
    lib = L.init(0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "pairs_256.npz"))
    res = {"lib": os.path.basename(path)}
    path = path.partition("@")[0]
    n0 = 1024
    perm = np.random.default_rng(3).integers(0, 256, size=n0)
    g1 = np.ascontiguousarray(z["g1"][perm]).reshape(-1); g2 = np.ascontiguousarray(z["g2"][perm]).reshape(-1)
    out = np.zeros(n0 * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(g1)[1], L.u32(g2)[1], None, L.u32(out)[1], n0, 0))
    res["pairing_ok"] = bool(np.array_equal(out.reshape(n0, 144), z["pairing"][perm]))
    L.check(lib.b381_miller_loop(L.u32(g1)[1], L.u32(g2)[1], None, L.u32(out)[1], n0, 0))
    res["miller_ok"] = bool(np.array_equal(out.reshape(n0, 144), z["miller_ark"][perm]))
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    rounds = int(os.environ.get("QB_ROUNDS", "8"))
    n = sm * 256 * rounds
    perm = np.random.default_rng(4).integers(0, 256, size=n)
    d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][perm]).reshape(-1).view(np.int32)).cuda()
    d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][perm]).reshape(-1).view(np.int32)).cuda()
    dout = torch.empty(n * 144, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=2):
        fn(); torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best

    ms = timed(lambda: L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st)))
    L.check(lib.b381_check_dev(st))
    res["pairings_per_s"] = round(n / ms * 1e3)
    res["ms_per_round"] = round(ms / rounds, 3)
    got = dout[:144 * 4096].cpu().numpy().view(np.uint32).reshape(4096, 144)
    res["pairing_dev_ok"] = bool(np.array_equal(got, z["pairing"][perm[:4096]]))
    ms = timed(lambda: L.check(lib.b381_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st)))
    res["miller_per_s"] = round(n / ms * 1e3)
    if os.environ.get("QB_MORE", "1") == "1":
        m = sm * 256 * 2
        co = torch.empty(m * L.G2PREP_WORDS, dtype=torch.int32, device="cuda")
        L.check(lib.b381_g2_prepare_dev(d2.data_ptr(), co.data_ptr(), m, 0, st))
        ms = timed(lambda: L.check(lib.b381_miller_loop_prepared_dev(d1.data_ptr(), co.data_ptr(), None, dout.data_ptr(), m, 0, 0, st)))
        res["miller_prepared_per_s"] = round(m / ms * 1e3)
        got = dout[:144 * 512].cpu().numpy().view(np.uint32).reshape(512, 144)
        res["prepared_ok"] = bool(np.array_equal(got, z["miller_ark"][perm[:512]]))
        d144 = torch.empty(144, dtype=torch.int32, device="cuda")
        ms = timed(lambda: L.check(lib.b381_multi_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, d144.data_ptr(), n, 0, st)))
        res["multi_miller_pairs_per_s"] = round(n / ms * 1e3)
    L.check(lib.b381_check_dev(st))
    print(json.dumps(res), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    for spec in sys.argv[1:]:
        # lib.so or lib.so@VAR=value[,VAR=value...] (environment of the child: run-time experiment switches)
        path, _, extra = spec.partition("@")
        env = dict(os.environ, B381_LIB=os.path.abspath(path))
        for kv in filter(None, extra.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", spec], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            print(json.dumps({"lib": os.path.basename(path), "error": r.stderr[-600:]}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
