"""scratch: first GPU run -- parity of every entry point on small inputs + first timings."""
import sys, os, time, random
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import b381_oracle as o
import b381
L = b381._lib
lib = L.init(0)
rnd = random.Random(7)
def rfp(): return rnd.randrange(o.P)
def rf2(): return (rfp(), rfp())
def rf12(): return o.f12_unflat([rfp() for _ in range(12)])
def f2l(a): return o.fp_to_limbs32(a[0]) + o.fp_to_limbs32(a[1])
def npa(l): return np.array(l, dtype=np.uint32)
ok = True
def chk(name, cond):
    global ok
    print(("PASS " if cond else "FAIL ") + name, flush=True)
    ok = ok and cond

n = 300
A = [rfp() for _ in range(n)]; B = [rfp() for _ in range(n)]
A[0] = 0; A[1] = 1; A[2] = o.P - 1; B[2] = o.P - 1
a = npa(sum((o.fp_to_limbs32(x) for x in A), [])); b = npa(sum((o.fp_to_limbs32(x) for x in B), []))
out = np.zeros(n * 12, dtype=np.uint32)
L.check(lib.b381_fp_mul(L.u32(a)[1], L.u32(b)[1], L.u32(out)[1], n))
chk("fp_mul", all(o.fp_from_limbs32(out[12*i:12*i+12]) == A[i] * B[i] % o.P for i in range(n)))
L.check(lib.b381_fp_mul_chain(L.u32(a)[1], L.u32(b)[1], L.u32(out)[1], n, 5))
chk("fp_mul_chain", all(o.fp_from_limbs32(out[12*i:12*i+12]) == A[i] * pow(B[i], 5, o.P) % o.P for i in range(n)))
n2 = 100
A2 = [rf2() for _ in range(n2)]; B2 = [rf2() for _ in range(n2)]
a = npa(sum((f2l(x) for x in A2), [])); b = npa(sum((f2l(x) for x in B2), [])); out = np.zeros(n2 * 24, dtype=np.uint32)
L.check(lib.b381_fp2_mul(L.u32(a)[1], L.u32(b)[1], L.u32(out)[1], n2))
chk("fp2_mul", all((o.fp_from_limbs32(out[24*i:24*i+12]), o.fp_from_limbs32(out[24*i+12:24*i+24])) == o.f2_mul(A2[i], B2[i]) for i in range(n2)))
n3 = 20
A3 = [rf12() for _ in range(n3)]; B3 = [rf12() for _ in range(n3)]
a = npa(sum((o.f12_to_limbs32(x) for x in A3), [])); b = npa(sum((o.f12_to_limbs32(x) for x in B3), [])); out = np.zeros(n3 * 144, dtype=np.uint32)
L.check(lib.b381_fp12_mul(L.u32(a)[1], L.u32(b)[1], L.u32(out)[1], n3))
chk("fp12_mul", all(o.f12_eq(o.f12_from_limbs32(out[144*i:144*i+144]), o.f12_mul(A3[i], B3[i])) for i in range(n3)))
am = [o.myfq12_from_fq12(x) for x in A3]; bm = [o.myfq12_from_fq12(x) for x in B3]
a = npa(sum((sum((o.fp_to_limbs32(v) for v in x), []) for x in am), [])); b = npa(sum((sum((o.fp_to_limbs32(v) for v in x), []) for x in bm), []))
L.check(lib.b381_fp12_mul_wbasis(L.u32(a)[1], L.u32(b)[1], L.u32(out)[1], n3))
chk("fp12_mul_wbasis", all([o.fp_from_limbs32(out[144*i+12*j:144*i+12*j+12]) for j in range(12)] == o.myfq12_mul(am[i], bm[i]) for i in range(n3)))
# pairs
npairs = 6
pairs = [(o.G1_GEN, o.G2_GEN)]
for i in range(npairs - 2):
    pairs.append((o.g1_mul(o.G1_GEN, rnd.randrange(1, o.R_ORDER)), o.g2_mul(o.G2_GEN, rnd.randrange(1, o.R_ORDER))))
pairs.append((None, o.G2_GEN))
g1 = npa(sum((o.g1_to_limbs32(p) for p, q in pairs), [])); g2 = npa(sum((o.g2_to_limbs32(q) for p, q in pairs), []))
inf = np.array([(1 if p is None else 0) | (2 if q is None else 0) for p, q in pairs], dtype=np.uint8)
out = np.zeros(npairs * 144, dtype=np.uint32)
t = time.time()
L.check(lib.b381_miller_loop(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(out)[1], npairs, 0))
print("miller call s", time.time() - t)
ml = [o.f12_from_limbs32(out[144*i:144*i+144]) for i in range(npairs)]
chk("miller ARK", all(o.f12_eq(ml[i], o.ark_miller_loop(*pairs[i])) for i in range(npairs)))
print("ark sha", o.f12_sha256(ml[0]))
L.check(lib.b381_miller_loop(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(out)[1], npairs, 1))
chk("miller ZK", all(o.f12_eq(o.f12_from_limbs32(out[144*i:144*i+144]), o.zk_miller_loop(*pairs[i])) for i in range(npairs)))
fin = npa(sum((o.f12_to_limbs32(x) for x in ml), []))
L.check(lib.b381_final_exp(L.u32(fin)[1], L.u32(out)[1], npairs))
es = [o.ark_final_exponentiation(x) for x in ml]
chk("final_exp", all(o.f12_eq(o.f12_from_limbs32(out[144*i:144*i+144]), es[i]) for i in range(npairs)))
L.check(lib.b381_pairing(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(out)[1], npairs, 0))
chk("pairing ARK", all(o.f12_eq(o.f12_from_limbs32(out[144*i:144*i+144]), es[i]) for i in range(npairs)))
print("e(G1,G2) sha", o.f12_sha256(o.f12_from_limbs32(out[:144])))
L.check(lib.b381_pairing(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(out)[1], npairs, 1))
chk("pairing ZK", all(o.f12_eq(o.f12_from_limbs32(out[144*i:144*i+144]), es[i]) for i in range(npairs)))
o144 = np.zeros(144, dtype=np.uint32)
L.check(lib.b381_multi_miller_loop(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(o144)[1], npairs, 0))
chk("multi_miller ARK", o.f12_eq(o.f12_from_limbs32(o144), o.ark_multi_miller_loop(pairs)))
L.check(lib.b381_multi_pairing(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(o144)[1], npairs, 0))
chk("multi_pairing ARK", o.f12_eq(o.f12_from_limbs32(o144), o.ark_multi_pairing(pairs)))
L.check(lib.b381_multi_miller_loop(L.u32(g1)[1], L.u32(g2)[1], L.u8(inf)[1], L.u32(o144)[1], npairs, 2))
chk("multi_miller LITERAL", o.f12_eq(o.f12_from_limbs32(o144), o.F12_ONE))
L.check(lib.b381_fp12_product(L.u32(fin)[1], L.u32(o144)[1], npairs))
pr = o.F12_ONE
for x in ml: pr = o.f12_mul(pr, x)
chk("fp12_product", o.f12_eq(o.f12_from_limbs32(o144), pr))
g1p = npa(o.fp_to_limbs32(o.G1_X) + o.fp_to_limbs32(o.G1_Y) + o.fp_to_limbs32(1))
g2p = npa(f2l(o.G2_X) + f2l(o.G2_Y) + f2l((1, 0)))
L.check(lib.b381_literal_optimized(L.u32(g1p)[1], L.u32(g2p)[1], L.u32(o144)[1], 1))
chk("literal", o.f12_eq(o.f12_from_limbs32(o144), o.literal_optimized_miller_loop((o.G1_X, o.G1_Y, 1), (o.G2_X, o.G2_Y, (1, 0)))))
# error paths
bad = npa([(o.P >> (32 * i)) & 0xffffffff for i in range(12)] * 2)
rc = lib.b381_fp_mul(L.u32(bad)[1], L.u32(bad)[1], L.u32(np.zeros(24, dtype=np.uint32))[1], 2)
chk("non-canonical -> error", rc == -3)
z = np.zeros(144, dtype=np.uint32)
rc = lib.b381_final_exp(L.u32(z)[1], L.u32(o144)[1], 1)
chk("final_exp(0) -> error", rc == -4)

# timings (device resident)
import torch
dev = torch.device("cuda:0")
def time_dev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
sms = 148
for mult in (1, 2, 4):
    N = sms * 128 * mult
    G1 = torch.from_numpy(np.tile(g1[:24 * 4], (N + 3) // 4)[:N * 24].astype(np.int32)).to(dev)
    G2 = torch.from_numpy(np.tile(g2[:48 * 4], (N + 3) // 4)[:N * 48].astype(np.int32)).to(dev)
    OUT = torch.zeros(N * 144, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ms = time_dev(lambda: L.check(lib.b381_pairing_dev(G1.data_ptr(), G2.data_ptr(), None, OUT.data_ptr(), N, 0, st)))
    print("pairing_dev N=%d  %.2f ms  %.0f pairings/s" % (N, ms, N / ms * 1e3), flush=True)
    ms = time_dev(lambda: L.check(lib.b381_miller_loop_dev(G1.data_ptr(), G2.data_ptr(), None, OUT.data_ptr(), N, 0, st)))
    print("miller_dev  N=%d  %.2f ms  %.0f loops/s" % (N, ms, N / ms * 1e3), flush=True)
    L.check(lib.b381_check_dev(st))
    h = OUT[:144 * 4].cpu().numpy().astype(np.uint32)
    chk("miller_dev parity", all(o.f12_eq(o.f12_from_limbs32(h[144*i:144*i+144]), ml[i]) for i in range(4)))
N = 1 << 22
Ad = torch.from_numpy(np.tile(a[:12], N).astype(np.int32)).to(dev); Bd = torch.from_numpy(np.tile(b[:12], N).astype(np.int32)).to(dev); Od = torch.zeros(N * 12, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = time_dev(lambda: L.check(lib.b381_fp_mul_dev(Ad.data_ptr(), Bd.data_ptr(), Od.data_ptr(), N, st)))
print("fp_mul_dev N=%d %.3f ms %.2f G mul/s  %.0f GB/s" % (N, ms, N / ms / 1e6, N * 144 / ms / 1e6))
K = 256
ms = time_dev(lambda: L.check(lib.b381_fp_mul_chain_dev(Ad.data_ptr(), Bd.data_ptr(), Od.data_ptr(), N, K, st)))
print("fp_mul_chain_dev N=%d k=%d %.3f ms %.2f G mul/s" % (N, K, ms, N * (K + 3) / ms / 1e6))
pk = (ctypes := __import__("ctypes")).c_double(); mh = ctypes.c_double()
L.check(lib.b381_imad_peak(ctypes.byref(pk), ctypes.byref(mh)))
print("imad peak %.1f Ginst/s at %.0f MHz" % (pk.value, mh.value))
print("launches", lib.b381_kernel_launches())
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
