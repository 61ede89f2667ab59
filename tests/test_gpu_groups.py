"""GPU tests of the group layer and the boundary added in round 2 (SURVEY 8f ranks 3-4, 8b "Threading"):
endomorphism subgroup tests and cofactor clearing (vs the oracle's restatement of ark-bls12-381 0.4, pinned by the
RFC 9380 effective cofactors), the native-Fq G1 path, the G1 bucket MSM at 2^16 / 2^20 (sum a_i k_i identity),
contexts driven from two host threads, every `_dev` entry point, and device-side ordering of `_dev` calls issued
on different streams."""
import ctypes
import threading

import numpy as np
import pytest

import b381_oracle as o
import util

pytestmark = pytest.mark.gpu
u8 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


@pytest.fixture(scope="module")
def L():
    import b381
    b381._lib.init(0)
    return b381._lib


@pytest.fixture(scope="module")
def lib(L):
    return L.lib()


@pytest.fixture(scope="module")
def z():
    return util.pairs_256()


def _outside_points():
    g1s, g2s = [], []
    x = 1
    while len(g1s) < 6:
        res = o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)
        if res[0] == "ok":
            g1s.append(res[1])
        x += 1
    x = 1
    while len(g2s) < 4:
        res = o.g2_deserialize(bytes([0x80]) + bytes(47) + x.to_bytes(48, "big"), True)
        if res[0] == "ok":
            g2s.append(res[1])
        x += 1
    return g1s, g2s


def _scal(ks):
    return np.array([[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for k in ks], dtype=np.uint32).reshape(-1)


def test_endomorphism_subgroup_checks_and_cofactor_clearing(L, lib, z):
    g1s, g2s = _outside_points()
    m = 300
    idx = np.arange(m) % 256
    g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1).copy()
    g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1).copy()
    want1 = np.ones(m, dtype=np.uint8); want2 = np.ones(m, dtype=np.uint8)
    for j, P in enumerate(g1s):
        g1[24 * (10 + j):24 * (11 + j)] = o.g1_to_limbs32(P)
        want1[10 + j] = 1 if o.g1_mul(P, o.R_ORDER) is None else 0
    for j, Q in enumerate(g2s):
        g2[48 * (20 + j):48 * (21 + j)] = o.g2_to_limbs32(Q)
        want2[20 + j] = 1 if o.g2_mul(Q, o.R_ORDER) is None else 0
    assert want1[10:16].min() == 0 and want2[20:24].min() == 0
    inf = np.zeros(m, dtype=np.uint8); inf[299] = 1
    o1 = np.zeros(m, dtype=np.uint8); o2 = np.zeros(m, dtype=np.uint8)
    L.check(lib.b381_g1_in_subgroup(L.u32(g1)[1], u8(inf), u8(o1), m))
    L.check(lib.b381_g2_in_subgroup(L.u32(g2)[1], u8(inf), u8(o2), m))
    assert np.array_equal(o1, want1) and np.array_equal(o2, want2)
    # cofactor clearing: [h_eff] P for points outside AND inside the subgroups, identity flags
    c1 = np.zeros(m * 24, dtype=np.uint32); c2 = np.zeros(m * 48, dtype=np.uint32)
    f1 = np.zeros(m, dtype=np.uint8); f2 = np.zeros(m, dtype=np.uint8)
    L.check(lib.b381_g1_clear_cofactor(L.u32(g1)[1], u8(inf), L.u32(c1)[1], u8(f1), m))
    L.check(lib.b381_g2_clear_cofactor(L.u32(g2)[1], u8(inf), L.u32(c2)[1], u8(f2), m))
    assert f1[299] == 1 and f2[299] == 1 and f1[:299].sum() == 0 and f2[:299].sum() == 0
    for i in (0, 10, 11, 15, 200):
        P = (o.fp_from_limbs32(g1[24 * i:24 * i + 12].tolist()), o.fp_from_limbs32(g1[24 * i + 12:24 * i + 24].tolist()))
        assert c1[24 * i:24 * i + 24].tolist() == o.g1_to_limbs32(o.g1_mul(P, o.G1_H_EFF)), i
    for i in (0, 20, 23):
        Q = (util.f2_from_words(g2[48 * i:48 * i + 24].tolist()), util.f2_from_words(g2[48 * i + 24:48 * i + 48].tolist()))
        assert c2[48 * i:48 * i + 48].tolist() == o.g2_to_limbs32(o.g2_clear_cofactor(Q)), i
    # everything that came out is in the subgroup, and pairs
    L.check(lib.b381_g1_in_subgroup(L.u32(c1)[1], u8(f1), u8(o1), m))
    L.check(lib.b381_g2_in_subgroup(L.u32(c2)[1], u8(f2), u8(o2), m))
    assert o1.min() == 1 and o2.min() == 1
    out = np.zeros(2 * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(c1[24 * 10:24 * 12])[1], L.u32(c2[48 * 20:48 * 22])[1], None, L.u32(out)[1], 2, L.MODE_ARK))
    P = (o.fp_from_limbs32(c1[240:252].tolist()), o.fp_from_limbs32(c1[252:264].tolist()))
    Q = (util.f2_from_words(c2[960:984].tolist()), util.f2_from_words(c2[984:1008].tolist()))
    assert o.f12_eq(o.f12_from_limbs32(out[:144].tolist()), o.ark_pairing(P, Q))


def _device_points(L, lib, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x3FFFFFFF; a[:, 0] |= 1
    g1 = np.tile(np.array(o.g1_to_limbs32(o.G1_GEN), dtype=np.uint32), n)
    p = np.zeros(n * 24, dtype=np.uint32); f = np.zeros(n, dtype=np.uint8)
    L.check(lib.b381_g1_scalar_mul(L.u32(g1)[1], None, L.u32(a.reshape(-1))[1], L.u32(p)[1], u8(f), n))
    assert f.sum() == 0
    return a, p


def _ints(words):
    """n x 8 u32 words -> list of Python ints"""
    w = words.astype(object)
    return [sum(int(w[i, j]) << (32 * j) for j in range(8)) for i in range(w.shape[0])]


@pytest.mark.parametrize("logn", [6, 12, 16])
def test_g1_bucket_msm_against_identity(L, lib, logn):
    """sum_i [k_i] (a_i G) == [sum a_i k_i mod r] G with DISTINCT device-generated points and random 256-bit scalars;
    the right-hand side is one oracle scalar multiplication.  Also: identity flags and zero scalars drop out."""
    n = 1 << logn
    a, p = _device_points(L, lib, n, 0x500 + logn)
    rng = np.random.default_rng(0x600 + logn)
    k = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    inf = np.zeros(n, dtype=np.uint8)
    inf[3] = 1
    k[5] = 0
    ai, ki = _ints(a), _ints(k)
    tot = sum(x * y for i, (x, y) in enumerate(zip(ai, ki)) if i != 3) % o.R_ORDER
    out = np.zeros(24, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g1_msm(L.u32(p)[1], u8(inf), L.u32(k.reshape(-1))[1], L.u32(out)[1], u8(f), n))
    assert f[0] == 0 and out.tolist() == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, tot))
    # plain sum of the same points
    L.check(lib.b381_g1_sum(L.u32(p)[1], u8(inf), L.u32(out)[1], u8(f), n))
    tot = sum(x for i, x in enumerate(ai) if i != 3) % o.R_ORDER
    assert f[0] == 0 and out.tolist() == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, tot))


def test_g1_msm_2p20_identity(L, lib):
    """SURVEY 8f rank 4 at the BASELINE batch size: 2^20 points (2^16 distinct device-generated points tiled, scalars
    all distinct), sum k_i P_i checked through the scalar identity."""
    n0, n = 1 << 16, 1 << 20
    a, p = _device_points(L, lib, n0, 0x900)
    rng = np.random.default_rng(0x901)
    k = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    k[:, 7] &= 0x7FFFFFFF
    pts = np.tile(p.reshape(n0, 24), (n // n0, 1)).reshape(-1)
    ai, ki = _ints(a), _ints(k)
    tot = sum(ai[i % n0] * ki[i] for i in range(n)) % o.R_ORDER
    out = np.zeros(24, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g1_msm(L.u32(pts)[1], None, L.u32(k.reshape(-1))[1], L.u32(out)[1], u8(f), n))
    assert f[0] == 0 and out.tolist() == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, tot))


def test_g2_sum_and_msm(L, lib, z):
    r = util.rng(111)
    for m in (1, 5, 17):
        g2 = np.ascontiguousarray(z["g2"][:m]).reshape(-1)
        inf = np.zeros(m, dtype=np.uint8)
        if m > 3:
            inf[3] = 1
        o2 = np.zeros(48, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
        L.check(lib.b381_g2_sum(L.u32(g2)[1], u8(inf), L.u32(o2)[1], u8(f), m))
        acc = None
        for i in range(m):
            if not inf[i]:
                acc = o.g2_add(acc, (util.f2_from_words(z["g2"][i][:24].tolist()), util.f2_from_words(z["g2"][i][24:].tolist())))
        assert f[0] == 0 and o2.tolist() == o.g2_to_limbs32(acc), m
    m = 100
    ks = [r.randrange(1, 1 << 256) for _ in range(m)]
    g2 = np.tile(np.array(o.g2_to_limbs32(o.G2_GEN), dtype=np.uint32), m)
    o2 = np.zeros(48, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g2_msm(L.u32(g2)[1], None, L.u32(_scal(ks))[1], L.u32(o2)[1], u8(f), m))
    assert f[0] == 0 and o2.tolist() == o.g2_to_limbs32(o.g2_mul(o.G2_GEN, sum(ks) % o.R_ORDER))
    k = ks[0] % o.R_ORDER
    L.check(lib.b381_g2_msm(L.u32(g2[:96])[1], None, L.u32(_scal([k, o.R_ORDER - k]))[1], L.u32(o2)[1], u8(f), 2))
    assert f[0] == 1


@pytest.mark.parametrize("logn", [8, 12, 17])
def test_g2_bucket_msm_against_identity(L, lib, logn):
    """G2 bucket method (n >= 256; window 4 / 8 / 16 bits): sum_i [k_i] (a_i G2) == [sum a_i k_i mod r] G2 with DISTINCT
    device-generated points and random 256-bit scalars; identity flags and zero scalars drop out; repeated points and
    P, -P with equal scalars meet inside buckets."""
    n = 1 << logn
    rng = np.random.default_rng(0xA00 + logn)
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x3FFFFFFF; a[:, 0] |= 1
    a[11] = a[10]                                                  # the same point twice ...
    g2 = np.tile(np.array(o.g2_to_limbs32(o.G2_GEN), dtype=np.uint32), n)
    p = np.zeros(n * 48, dtype=np.uint32); f = np.zeros(n, dtype=np.uint8)
    L.check(lib.b381_g2_scalar_mul(L.u32(g2)[1], None, L.u32(a.reshape(-1))[1], L.u32(p)[1], u8(f), n))
    assert f.sum() == 0
    k = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    k[11] = k[10]                                                  # ... with the same scalar: doubling inside every bucket it hits
    inf = np.zeros(n, dtype=np.uint8)
    inf[3] = 1
    k[5] = 0
    ai, ki = _ints(a), _ints(k)
    ai[21] = o.R_ORDER - ai[20]                                    # P_21 = -P_20 with equal scalars: the buckets cancel
    pm = p.reshape(n, 48)
    q20 = (util.f2_from_words(pm[20][:24].tolist()), util.f2_from_words(pm[20][24:].tolist()))
    pm[21] = np.array(o.g2_to_limbs32(o.g2_neg(q20)), dtype=np.uint32)
    k[21] = k[20]; ki[21] = ki[20]
    tot = sum(x * y for i, (x, y) in enumerate(zip(ai, ki)) if i != 3) % o.R_ORDER
    out = np.zeros(48, dtype=np.uint32); fo = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g2_msm(L.u32(p)[1], u8(inf), L.u32(k.reshape(-1))[1], L.u32(out)[1], u8(fo), n))
    assert fo[0] == 0 and out.tolist() == o.g2_to_limbs32(o.g2_mul(o.G2_GEN, tot))


def test_two_contexts_from_two_threads(L, lib, z):
    """b381_ctx_create / b381_ctx_set_current: two contexts (on the same GPU here: the box has one) driven by two
    host threads at the same time give the fixture values; the default context stays usable."""
    n = 600
    idx = np.arange(n) % 256
    g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1); g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1)
    ctxs = [ctypes.c_void_p(), ctypes.c_void_p()]
    for c in ctxs:
        L.check(lib.b381_ctx_create(0, ctypes.byref(c)))
    results, errors = [None, None], []

    def work(t):
        try:
            L.check(lib.b381_ctx_set_current(ctxs[t]))
            cur = ctypes.c_void_p()
            lib.b381_ctx_get_current(ctypes.byref(cur))
            assert cur.value == ctxs[t].value
            out = np.zeros(n * 144, dtype=np.uint32)
            for _ in range(2):
                L.check(lib.b381_pairing(L.u32(g1)[1], L.u32(g2)[1], None, L.u32(out)[1], n, L.MODE_ARK))
            results[t] = out
            lib.b381_ctx_set_current(None)
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for t in range(2):
        assert np.array_equal(results[t].reshape(n, 144), z["pairing"][idx])
    out = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(g1[:24])[1], L.u32(g2[:48])[1], None, L.u32(out)[1], 1, L.MODE_ARK))
    assert np.array_equal(out, z["pairing"][0])
    for c in ctxs:
        L.check(lib.b381_ctx_destroy(c))
    assert lib.b381_ctx_destroy(None) == -2


def test_dev_calls_on_different_streams_are_ordered(L, lib, z):
    """two small `_dev` batches on two streams share the context's scratch arena; the library serialises them on the
    device (without that they overwrite each other's slots)."""
    import torch
    n = 64                                            # one CTA each: they would run concurrently on a 148-SM chip
    d1 = torch.from_numpy(np.ascontiguousarray(z["g1"][:n]).reshape(-1).view(np.int32)).cuda()
    d2 = torch.from_numpy(np.ascontiguousarray(z["g2"][:n]).reshape(-1).view(np.int32)).cuda()
    e1 = torch.from_numpy(np.ascontiguousarray(z["g1"][n:2 * n]).reshape(-1).view(np.int32)).cuda()
    e2 = torch.from_numpy(np.ascontiguousarray(z["g2"][n:2 * n]).reshape(-1).view(np.int32)).cuda()
    oa = torch.zeros(n * 144, dtype=torch.int32, device="cuda"); ob = torch.zeros(n * 144, dtype=torch.int32, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):
        L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, oa.data_ptr(), n, 0, sa.cuda_stream))
        L.check(lib.b381_pairing_dev(e1.data_ptr(), e2.data_ptr(), None, ob.data_ptr(), n, 0, sb.cuda_stream))
    L.check(lib.b381_check_dev(sa.cuda_stream)); L.check(lib.b381_check_dev(sb.cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(oa.cpu().numpy().view(np.uint32).reshape(n, 144), z["pairing"][:n])
    assert np.array_equal(ob.cpu().numpy().view(np.uint32).reshape(n, 144), z["pairing"][n:2 * n])


def test_every_dev_entry_point(L, lib, z):
    """each `_dev` symbol gives the same result as its host-pointer twin"""
    import torch
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype == np.uint32 else np.uint8)).cuda()
    back = lambda t, dt: t.cpu().numpy().view(dt)
    r = util.rng(7)
    n = 200
    vals = [util.rfp(r) or 1 for _ in range(n)]
    a = np.array(sum((o.fp_to_limbs32(v) for v in vals), []), dtype=np.uint32)
    da = dev(a)
    # field helpers
    for name, w, extra in (("b381_fp_inv", 12, ()), ("b381_fp_to_u32_digits", 12, ()), ("b381_fp2_inv", 24, ())):
        m = n * 12 // w
        h = np.zeros(m * w, dtype=np.uint32); d = torch.zeros(m * w, dtype=torch.int32, device="cuda")
        L.check(getattr(lib, name)(L.u32(a)[1], L.u32(h)[1], m))
        L.check(getattr(lib, name + "_dev")(da.data_ptr(), d.data_ptr(), m, st))
        L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d, np.uint32), h), name
    dig = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_to_u32_digits(L.u32(a)[1], L.u32(dig)[1], n))
    d = torch.zeros(n * 12, dtype=torch.int32, device="cuda")
    ddig = dev(dig)                                    # (device inputs are kept in named tensors: a temporary would be freed, and its memory reused, before the kernel runs)
    L.check(lib.b381_fp_from_u32_digits_dev(ddig.data_ptr(), d.data_ptr(), n, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d, np.uint32), a)
    sqv = [v * v % o.P for v in vals]
    sq = np.array(sum((o.fp_to_limbs32(v) for v in sqv), []), dtype=np.uint32)
    sgn = np.array([i & 1 for i in range(n)], dtype=np.uint8)
    h = np.zeros(n * 12, dtype=np.uint32); d = torch.zeros(n * 12, dtype=torch.int32, device="cuda")
    L.check(lib.b381_fp_sqrt(L.u32(sq)[1], u8(sgn), L.u32(h)[1], n))
    dsq, dsgn = dev(sq), dev(sgn)
    L.check(lib.b381_fp_sqrt_dev(dsq.data_ptr(), dsgn.data_ptr(), d.data_ptr(), n, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d, np.uint32), h)
    hb = np.zeros(n, dtype=np.uint8); db = torch.zeros(n, dtype=torch.uint8, device="cuda")
    L.check(lib.b381_fp_is_square(L.u32(a)[1], u8(hb), n))
    L.check(lib.b381_fp_is_square_dev(da.data_ptr(), db.data_ptr(), n, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(db, np.uint8), hb) and 0 < hb.sum() < n
    exp = (ctypes.c_uint64 * 2)(0x123456789ABCDEF, 77)
    L.check(lib.b381_fp_pow(L.u32(a)[1], exp, 2, L.u32(h)[1], n))
    L.check(lib.b381_fp_pow_dev(da.data_ptr(), exp, 2, d.data_ptr(), n, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d, np.uint32), h)
    assert o.fp_from_limbs32(h[:12].tolist()) == o.pow_fq(vals[0], [0x123456789ABCDEF, 77])
    a2 = a[:n * 12 // 24 * 24]
    m2 = len(a2) // 24
    sq2 = np.array(sum((util.f2_words(o.f2_sqr((vals[2 * i], vals[2 * i + 1]))) for i in range(m2)), []), dtype=np.uint32)
    h2 = np.zeros(m2 * 24, dtype=np.uint32); d2 = torch.zeros(m2 * 24, dtype=torch.int32, device="cuda")
    L.check(lib.b381_fp2_sqrt(L.u32(sq2)[1], None, L.u32(h2)[1], m2))
    dsq2 = dev(sq2)
    L.check(lib.b381_fp2_sqrt_dev(dsq2.data_ptr(), None, d2.data_ptr(), m2, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d2, np.uint32), h2)
    hb2 = np.zeros(m2, dtype=np.uint8); db2 = torch.zeros(m2, dtype=torch.uint8, device="cuda")
    L.check(lib.b381_fp2_is_square(L.u32(a2)[1], u8(hb2), m2))
    da2 = dev(a2)
    L.check(lib.b381_fp2_is_square_dev(da2.data_ptr(), db2.data_ptr(), m2, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(db2, np.uint8), hb2)
    # Fq6 / Fq12 inverse, w-basis product, witness limbs
    k = 40
    f12 = np.ascontiguousarray(z["pairing"][:k]).reshape(-1)
    df12 = dev(f12)
    for name, w in (("b381_fp12_inv", 144), ("b381_fp6_inv", 72), ("b381_fp12_to_witness_limbs", 144)):
        m = k * 144 // w
        h = np.zeros(m * w, dtype=np.uint32); d = torch.zeros(m * w, dtype=torch.int32, device="cuda")
        L.check(getattr(lib, name)(L.u32(f12)[1], L.u32(h)[1], m))
        L.check(getattr(lib, name + "_dev")(df12.data_ptr(), d.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d, np.uint32), h), name
    g12 = np.ascontiguousarray(z["miller_ark"][:k]).reshape(-1)
    h = np.zeros(k * 144, dtype=np.uint32); d = torch.zeros(k * 144, dtype=torch.int32, device="cuda")
    L.check(lib.b381_fp12_mul_wbasis(L.u32(f12)[1], L.u32(g12)[1], L.u32(h)[1], k))
    dg12 = dev(g12)
    L.check(lib.b381_fp12_mul_wbasis_dev(df12.data_ptr(), dg12.data_ptr(), d.data_ptr(), k, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d, np.uint32), h)
    # points: serialise / deserialise, subgroup, cofactor, scalar mul, sums, MSM, multi-pairing
    m = 128
    g1 = np.ascontiguousarray(z["g1"][:m]).reshape(-1); g2 = np.ascontiguousarray(z["g2"][:m]).reshape(-1)
    inf = np.zeros(m, dtype=np.uint8); inf[5] = 1
    dg1, dg2, dinf = dev(g1), dev(g2), dev(inf)
    for grp, w, pts, dpts in (("g1", 24, g1, dg1), ("g2", 48, g2, dg2)):
        for c in (1, 0):
            nb = (w * 2 if c else w * 4)
            e = np.zeros(m * nb, dtype=np.uint8); de = torch.zeros(m * nb, dtype=torch.uint8, device="cuda")
            L.check(getattr(lib, "b381_%s_serialize" % grp)(L.u32(pts)[1], u8(inf), c, u8(e), m))
            L.check(getattr(lib, "b381_%s_serialize_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), c, de.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
            assert np.array_equal(back(de, np.uint8), e), (grp, c)
            dp = torch.zeros(m * w, dtype=torch.int32, device="cuda"); df = torch.zeros(m, dtype=torch.uint8, device="cuda")
            L.check(getattr(lib, "b381_%s_deserialize_dev" % grp)(de.data_ptr(), c, dp.data_ptr(), df.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
            keep = np.repeat(inf == 0, w)
            assert np.array_equal(back(df, np.uint8), inf) and np.array_equal(back(dp, np.uint32)[keep], pts[keep]), (grp, c)
        hb = np.zeros(m, dtype=np.uint8); db = torch.zeros(m, dtype=torch.uint8, device="cuda")
        L.check(getattr(lib, "b381_%s_in_subgroup" % grp)(L.u32(pts)[1], u8(inf), u8(hb), m))
        L.check(getattr(lib, "b381_%s_in_subgroup_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), db.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(db, np.uint8), hb) and hb.min() == 1
        h = np.zeros(m * w, dtype=np.uint32); hf = np.zeros(m, dtype=np.uint8)
        d = torch.zeros(m * w, dtype=torch.int32, device="cuda"); df = torch.zeros(m, dtype=torch.uint8, device="cuda")
        L.check(getattr(lib, "b381_%s_clear_cofactor" % grp)(L.u32(pts)[1], u8(inf), L.u32(h)[1], u8(hf), m))
        L.check(getattr(lib, "b381_%s_clear_cofactor_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), d.data_ptr(), df.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d, np.uint32), h) and np.array_equal(back(df, np.uint8), hf)
        sc = _scal([r.randrange(1, 1 << 256) for _ in range(m)])
        dsc = dev(sc)
        L.check(getattr(lib, "b381_%s_scalar_mul" % grp)(L.u32(pts)[1], u8(inf), L.u32(sc)[1], L.u32(h)[1], u8(hf), m))
        L.check(getattr(lib, "b381_%s_scalar_mul_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), dsc.data_ptr(), d.data_ptr(), df.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d, np.uint32), h) and np.array_equal(back(df, np.uint8), hf)
        h1 = np.zeros(w, dtype=np.uint32); hf1 = np.zeros(1, dtype=np.uint8)
        d1 = torch.zeros(w, dtype=torch.int32, device="cuda"); df1 = torch.zeros(1, dtype=torch.uint8, device="cuda")
        L.check(getattr(lib, "b381_%s_sum" % grp)(L.u32(pts)[1], u8(inf), L.u32(h1)[1], u8(hf1), m))
        L.check(getattr(lib, "b381_%s_sum_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), d1.data_ptr(), df1.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d1, np.uint32), h1) and back(df1, np.uint8)[0] == hf1[0] == 0
        L.check(getattr(lib, "b381_%s_msm" % grp)(L.u32(pts)[1], u8(inf), L.u32(sc)[1], L.u32(h1)[1], u8(hf1), m))
        L.check(getattr(lib, "b381_%s_msm_dev" % grp)(dpts.data_ptr(), dinf.data_ptr(), dsc.data_ptr(), d1.data_ptr(), df1.data_ptr(), m, st)); L.check(lib.b381_check_dev(st))
        assert np.array_equal(back(d1, np.uint32), h1) and back(df1, np.uint8)[0] == hf1[0] == 0
    h144 = np.zeros(144, dtype=np.uint32); d144 = torch.zeros(144, dtype=torch.int32, device="cuda")
    L.check(lib.b381_multi_pairing(L.u32(g1)[1], L.u32(g2)[1], u8(inf), L.u32(h144)[1], m, L.MODE_ARK))
    L.check(lib.b381_multi_pairing_dev(dg1.data_ptr(), dg2.data_ptr(), dinf.data_ptr(), d144.data_ptr(), m, L.MODE_ARK, st)); L.check(lib.b381_check_dev(st))
    assert np.array_equal(back(d144, np.uint32), h144)


def test_single_process_device_set(L, lib, z):
    """distributed.DeviceSet: one process, one context per GPU (all visible GPUs; the same GPU twice on a 1-GPU box),
    a host thread per context: sharded pairings and the multi-pairing product equal the fixture values."""
    import torch
    import b381
    ng = torch.cuda.device_count()
    devs = list(range(ng)) if ng > 1 else [0, 0]
    ds = b381.distributed.DeviceSet(devs)
    try:
        n = 1000
        idx = np.arange(n) % 256
        g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1); g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1)
        inf = np.zeros(n, dtype=np.uint8); inf[17] = 1
        out = ds.pairing(g1, g2, inf, L.MODE_ARK).reshape(n, 144)
        one = np.array(o.f12_to_limbs32(o.F12_ONE), dtype=np.uint32)
        for i in (0, 16, 17, 499, 500, 999):
            assert np.array_equal(out[i], one if inf[i] else z["pairing"][idx[i]]), i
        m = 64
        res = ds.multi_pairing(g1[:24 * m], g2[:48 * m], None, L.MODE_ARK)
        pr = o.F12_ONE
        for i in range(m):
            pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i].tolist()))
        assert o.f12_eq(o.f12_from_limbs32(res.tolist()), o.ark_final_exponentiation(pr))
    finally:
        ds.close()
