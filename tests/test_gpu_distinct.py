"""GPU parity on DISTINCT data at the BASELINE sizes ("what's weak" 2): the other GPU tests
tile 256 fixture pairs; here every pair is different.

* 2^16 pairs (a_i G1, b_i G2) generated ON THE DEVICE with b381_g1/g2_scalar_mul from seeded 255-bit
  scalars; b381_pairing and b381_miller_loop compared IN FULL with the oracle's C port (a few seconds on the
  host cores), a sample of the generated points against the Python oracle's scalar multiplication.
* adversarial field operands (0, 1, p-1, p-2, 2^k - 1 patterns, all-ones low words, R mod p, values next to
  the conditional-subtraction boundaries) through b381_fp_mul / b381_fp2_mul / b381_fp12_mul, all pairs.
* BASELINE config #2 once at its stated size: 2^26 random Fp products compared in full with the C port.
* the LITERAL loop's vertical-line branch (optimized_line_function's third case,
  /root/reference/src/miller_loop_native_optimized.rs:62-77) and the (1, 1, 0) identity of ark-ec 0.4.
"""
import ctypes
import os

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


def _scalars(seed, n):
    """n seeded scalars in [1, r) as n x 8 little-endian u32 words (and as Python ints for a sample)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    w[:, 7] &= 0x3FFFFFFF                      # < 2^254 < r
    w[:, 0] |= 1                               # never zero
    return w


def _to_int(words):
    return sum(int(v) << (32 * i) for i, v in enumerate(words))


def device_points(L, lib, n, seed):
    """(P_i, Q_i) = (a_i G1, b_i G2), computed by the library itself on the GPU."""
    a, b = _scalars(seed, n), _scalars(seed + 1, n)
    g1 = np.tile(np.array(o.g1_to_limbs32(o.G1_GEN), dtype=np.uint32), n)
    g2 = np.tile(np.array(o.g2_to_limbs32(o.G2_GEN), dtype=np.uint32), n)
    p = np.zeros(n * 24, dtype=np.uint32); q = np.zeros(n * 48, dtype=np.uint32)
    i1 = np.zeros(n, dtype=np.uint8); i2 = np.zeros(n, dtype=np.uint8)
    L.check(lib.b381_g1_scalar_mul(L.u32(g1)[1], None, L.u32(a.reshape(-1))[1], L.u32(p)[1], u8(i1), n))
    L.check(lib.b381_g2_scalar_mul(L.u32(g2)[1], None, L.u32(b.reshape(-1))[1], L.u32(q)[1], u8(i2), n))
    assert i1.sum() == 0 and i2.sum() == 0
    return a, b, p, q


def test_distinct_pairs_2p16_full_compare(L, lib):
    n = 1 << 16
    a, b, p, q = device_points(L, lib, n, 0x381)
    # the inputs really are distinct and really are the points they claim to be
    assert len(np.unique(p.reshape(n, 24), axis=0)) == n and len(np.unique(q.reshape(n, 48), axis=0)) == n
    for i in (0, 1, 777, n - 1):
        assert p[24 * i:24 * i + 24].tolist() == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, _to_int(a[i])))
        assert q[48 * i:48 * i + 48].tolist() == o.g2_to_limbs32(o.g2_mul(o.G2_GEN, _to_int(b[i])))
    ref = util.load_ref_lib()
    threads = os.cpu_count() or 8
    out = np.zeros(n * 144, dtype=np.uint32)
    chk = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
    assert ref.ref_pairing(util.p32(p), util.p32(q), None, util.p32(chk), n, threads) == 0
    bad = np.nonzero((out.reshape(n, 144) != chk.reshape(n, 144)).any(axis=1))[0]
    assert len(bad) == 0, "pairing differs from the C port at %d of %d pairs, first %s" % (len(bad), n, bad[:8])
    L.check(lib.b381_miller_loop(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
    assert ref.ref_miller_loop(util.p32(p), util.p32(q), None, util.p32(chk), n, threads) == 0
    bad = np.nonzero((out.reshape(n, 144) != chk.reshape(n, 144)).any(axis=1))[0]
    assert len(bad) == 0, "Miller loop differs from the C port at %d of %d pairs, first %s" % (len(bad), n, bad[:8])
    # the Python oracle agrees with both on a sample (the C port is itself pinned to it by tests/test_oracle.py)
    for i in (0, 4242):
        P = (o.fp_from_limbs32(p[24 * i:24 * i + 12].tolist()), o.fp_from_limbs32(p[24 * i + 12:24 * i + 24].tolist()))
        Q = (util.f2_from_words(q[48 * i:48 * i + 24].tolist()), util.f2_from_words(q[48 * i + 24:48 * i + 48].tolist()))
        assert o.f12_eq(o.f12_from_limbs32(out[144 * i:144 * i + 144].tolist()), o.ark_miller_loop(P, Q))
    # prepared and shared-squaring paths on the same distinct data (first 2^12 pairs)
    m = 1 << 12
    co = np.zeros(m * L.G2PREP_WORDS, dtype=np.uint32)
    L.check(lib.b381_g2_prepare(L.u32(q[:48 * m])[1], L.u32(co)[1], m, L.MODE_ARK))
    o2 = np.zeros(m * 144, dtype=np.uint32)
    L.check(lib.b381_miller_loop_prepared(L.u32(p[:24 * m])[1], L.u32(co)[1], None, L.u32(o2)[1], m, L.MODE_ARK))
    assert np.array_equal(o2, out[:144 * m])
    # packed (internal-format) prepared stage on ALL 2^16 distinct pairs, device-resident
    import torch
    dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
    dp = torch.from_numpy(p.view(np.int32)).to(dev); dq = torch.from_numpy(q.view(np.int32)).to(dev)
    pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
    dout = torch.empty(n * 144, dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(dq.data_ptr(), pk.data_ptr(), n, L.MODE_ARK, st))
    L.check(lib.b381_miller_loop_packed_dev(dp.data_ptr(), pk.data_ptr(), None, dout.data_ptr(), n, L.MODE_ARK, 0, st))
    L.check(lib.b381_check_dev(st))
    assert np.array_equal(dout.cpu().numpy().view(np.uint32), out), "packed prepared Miller loop differs from the unprepared one"
    del dp, dq, pk, dout
    o144 = np.zeros(144, dtype=np.uint32); c144 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_multi_miller_loop(L.u32(p)[1], L.u32(q)[1], None, L.u32(o144)[1], n, L.MODE_ARK))
    assert ref.ref_multi_miller_loop(util.p32(p), util.p32(q), None, util.p32(c144), n, threads) == 0
    assert np.array_equal(o144, c144)


def test_bilinearity_on_distinct_pairs(L, lib):
    """e(a_i G1, b_i G2) = e(G1, G2)^(a_i b_i): the right-hand side from the oracle's Fq12 exponentiation."""
    n = 64
    a, b, p, q = device_points(L, lib, n, 0x777)
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(p)[1], L.u32(q)[1], None, L.u32(out)[1], n, L.MODE_ARK))
    e = o.ark_pairing(o.G1_GEN, o.G2_GEN)
    for i in (0, 31, 63):
        k = _to_int(a[i]) * _to_int(b[i]) % o.R_ORDER
        assert o.f12_eq(o.f12_from_limbs32(out[144 * i:144 * i + 144].tolist()), o.f12_pow(e, k))


SPECIAL = None


def special_values():
    """canonical field values that stress carries and the reduction boundaries"""
    global SPECIAL
    if SPECIAL is None:
        P = o.P
        vals = [0, 1, 2, 3, P - 1, P - 2, P - 3, (P - 1) // 2, (P + 1) // 2, o.MONT_R_MOD_P, o.MONT_R2_MOD_P, P - o.MONT_R_MOD_P,
                (1 << 380) - 1, (1 << 380), (1 << 380) + 1,
                (1 << 352) - 1, (1 << 352), (1 << 320) - 1, (1 << 256) - 1, (1 << 192) - 1, (1 << 64) - 1, (1 << 32) - 1, 1 << 32, (1 << 33) - 1,
                0xFFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF % P,
                int("aaaaaaaa" * 12, 16) % P, int("55555555" * 12, 16) % P, int("80000000" * 12, 16) % P, int("7fffffff" * 12, 16) % P]
        # Montgomery pre-images: values whose R = 2^384 representation has all-ones low words / sits next to p
        rinv = o.MONT_RINV
        for rep in ((1 << 352) - 1, (1 << 380) - 1, P - 1, P - 2, 1, 2, (1 << 96) - 1, ((1 << 381) - 1) % P, 0xFFFFFFFF, (P >> 1), (P >> 1) + 1):
            vals.append(rep * rinv % P)
        SPECIAL = sorted(set(v % P for v in vals))
    return SPECIAL


def test_adversarial_field_operands(L, lib):
    S = special_values()
    k = len(S)
    A = [x for x in S for _ in S]
    B = [y for _ in S for y in S]
    n = len(A)
    a = util.arr(sum((o.fp_to_limbs32(x) for x in A), [])); b = util.arr(sum((o.fp_to_limbs32(x) for x in B), []))
    out = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_mul(util.p32(a), util.p32(b), util.p32(out), n))
    got = [o.fp_from_limbs32(out[12 * i:12 * i + 12].tolist()) for i in range(n)]
    assert got == [x * y % o.P for x, y in zip(A, B)]
    for kk in (2, 5):                          # the register-resident chain runs the 13-word internal format
        L.check(lib.b381_fp_mul_chain(util.p32(a), util.p32(b), util.p32(out), n, kk))
        assert all(o.fp_from_limbs32(out[12 * i:12 * i + 12].tolist()) == A[i] * pow(B[i], kk, o.P) % o.P for i in range(n))
    # Fp2: all pairs of (s_i, s_j) x (s_j', s_i') on a deterministic shuffle
    r = util.rng(5)
    n2 = 2048
    A2 = [(r.choice(S), r.choice(S)) for _ in range(n2)]; B2 = [(r.choice(S), r.choice(S)) for _ in range(n2)]
    a2 = util.arr(sum((util.f2_words(x) for x in A2), [])); b2 = util.arr(sum((util.f2_words(x) for x in B2), []))
    o2 = np.zeros(n2 * 24, dtype=np.uint32)
    L.check(lib.b381_fp2_mul(util.p32(a2), util.p32(b2), util.p32(o2), n2))
    assert all(util.f2_from_words(o2[24 * i:24 * i + 24].tolist()) == o.f2_mul(A2[i], B2[i]) for i in range(n2))
    # Fp12 (tower and w-basis): coefficients drawn from the special set, including all-(p-1) and sparse elements
    n12 = 300
    X = [o.f12_unflat([r.choice(S) for _ in range(12)]) for _ in range(n12)]
    Y = [o.f12_unflat([r.choice(S) for _ in range(12)]) for _ in range(n12)]
    X[0] = o.f12_unflat([o.P - 1] * 12); Y[0] = o.f12_unflat([o.P - 1] * 12)
    X[1] = o.f12_unflat([o.P - 1] * 12); Y[1] = o.f12_unflat([1] + [0] * 11)
    x = util.arr(sum((o.f12_to_limbs32(t) for t in X), [])); y = util.arr(sum((o.f12_to_limbs32(t) for t in Y), []))
    o12 = np.zeros(n12 * 144, dtype=np.uint32)
    L.check(lib.b381_fp12_mul(util.p32(x), util.p32(y), util.p32(o12), n12))
    assert all(o.f12_eq(f, o.f12_mul(X[i], Y[i])) for i, f in enumerate(util.f12s(o12, n12)))
    # ... and through the final exponentiation's cyclotomic chain: FE of special-valued inputs vs the C port
    ref = util.load_ref_lib()
    fe = np.zeros(n12 * 144, dtype=np.uint32); ck = np.zeros(n12 * 144, dtype=np.uint32)
    L.check(lib.b381_final_exp(util.p32(x), util.p32(fe), n12))
    assert ref.ref_final_exp(util.p32(x), util.p32(ck), n12, os.cpu_count() or 8) == 0
    assert np.array_equal(fe, ck)
    # non-canonical limbs (all ones) are refused, not reduced
    ones = np.full(12, 0xFFFFFFFF, dtype=np.uint32)
    assert lib.b381_fp_mul(util.p32(ones), util.p32(ones), util.p32(out), 1) == -3


def test_config2_full_size_2p26(L, lib):
    """BASELINE config #2 as stated: 2^26 random Fp elements per operand, every product compared with the C port."""
    n = 1 << 26
    rng = np.random.default_rng(0x381)
    a = rng.integers(0, 1 << 32, size=n * 12, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, size=n * 12, dtype=np.uint64).astype(np.uint32)
    a[11::12] &= 0x0FFFFFFF; b[11::12] &= 0x0FFFFFFF          # < 2^380 < p: canonical Montgomery limbs
    out = np.empty(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_mul(util.p32(a), util.p32(b), util.p32(out), n))
    ref = util.load_ref_lib()
    chk = np.empty(n * 12, dtype=np.uint32)
    assert ref.ref_fp_mul(util.p32(a), util.p32(b), util.p32(chk), n, os.cpu_count() or 8) == 0
    assert np.array_equal(out, chk)
    for i in (0, n // 3, n - 1):
        x, y = o.fp_from_limbs32(a[12 * i:12 * i + 12].tolist()), o.fp_from_limbs32(b[12 * i:12 * i + 12].tolist())
        assert o.fp_from_limbs32(out[12 * i:12 * i + 12].tolist()) == x * y % o.P


def test_literal_vertical_line_and_identity(L, lib):
    """A point with x = 0 has order 3 under the a = 0 doubling formula (2 (0, y) = (0, -y)), so after 17
    doublings R = -Q and the addition at bit 16 takes optimized_line_function's vertical-line branch
    (den == 0, num != 0); R + Q is then the identity, the next tangent has den = 2 y z = 0 and f_den
    collapses to 0: the reference panics in `/`, the oracle raises, the library reports ZERO_DIVISION.
    The ark-ec 0.4 identity is (1, 1, 0) in both groups."""
    from b381.curves import G1Affine, G2Affine, G1Projective, G2Projective
    qq = ((0, 0), (5, 7), (1, 0))
    pp = (o.G1_X, o.G1_Y, 1)
    # the oracle's line function really takes the vertical branch on (−Q, Q)
    neg = (qq[0], o.f2_neg(qq[1]), qq[2])
    A = o.f2_sub(o.f2_mul((pp[0], 0), neg[2]), o.f2_mul(neg[0], (pp[2], 0)))
    n_, d_ = o.literal_line_function(neg, qq, pp)
    assert n_ == A and d_ == o.f2_mul(neg[2], (pp[2], 0))
    with pytest.raises(ZeroDivisionError):
        o.literal_optimized_miller_loop(pp, qq)
    g1p = util.arr(sum((o.fp_to_limbs32(v) for v in pp), []))
    g2p = util.arr(sum((util.f2_words(v) for v in qq), []))
    out = np.zeros(144, dtype=np.uint32)
    assert lib.b381_literal_optimized(util.p32(g1p), util.p32(g2p), util.p32(out), 1) == -4
    # a batch mixing that input with a regular one: the regular lane is still bit-exact
    reg1 = util.arr(o.fp_to_limbs32(o.G1_X) + o.fp_to_limbs32(o.G1_Y) + o.fp_to_limbs32(1))
    reg2 = util.arr(util.f2_words(o.G2_X) + util.f2_words(o.G2_Y) + util.f2_words((1, 0)))
    out2 = np.zeros(2 * 144, dtype=np.uint32)
    assert lib.b381_literal_optimized(util.p32(np.concatenate([g1p, reg1])), util.p32(np.concatenate([g2p, reg2])), util.p32(out2), 2) == -4
    want = o.literal_optimized_miller_loop((o.G1_X, o.G1_Y, 1), (o.G2_X, o.G2_Y, (1, 0)))
    assert o.f12_eq(o.f12_from_limbs32(out2[144:].tolist()), want)
    i1, i2 = G1Projective.from_affine(G1Affine.identity()), G2Projective.from_affine(G2Affine.identity())
    assert (i1.x.v, i1.y.v, i1.z.v) == (1, 1, 0)
    assert (i2.x.c0.v, i2.x.c1.v, i2.y.c0.v, i2.y.c1.v, i2.z.c0.v, i2.z.c1.v) == (1, 0, 1, 0, 0, 0)
    import b381
    with pytest.raises(Exception):
        b381.optimized_miller_loop(i1, G2Projective.generator())      # zp = 0 -> f_den = 0 (reference panics)
