"""GPU parity tests proper: every entry point of the C ABI (include/b381.h) called through ctypes
on a real B200, compared bit-for-bit with the oracle (Python restatement, its C port) and the
committed golden fixtures; ragged / empty / identity / invalid inputs; and size-independent
properties at the full BASELINE sizes (2^16 Miller loops, 2^20 pairings)."""
import ctypes
import hashlib

import numpy as np
import pytest

import b381_oracle as o
import util

pytestmark = pytest.mark.gpu


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


def _pairs(z, idx):
    g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1)
    g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1)
    return g1, g2


def test_device_is_blackwell(L, lib):
    sm, maj, mnr, scratch = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    L.check(lib.b381_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(scratch)))
    assert maj.value == 10 and sm.value >= 100


def test_fp_mul_parity_and_edges(L, lib):
    r = util.rng(41)
    n = 1000
    A = [util.rfp(r) for _ in range(n)]
    B = [util.rfp(r) for _ in range(n)]
    A[:6] = [0, 1, o.P - 1, o.P - 1, o.MONT_R_MOD_P, o.MONT_R2_MOD_P]
    B[:6] = [5, 1, o.P - 1, 1, o.MONT_R_MOD_P, 2]
    a = util.arr(sum((o.fp_to_limbs32(x) for x in A), []))
    b = util.arr(sum((o.fp_to_limbs32(x) for x in B), []))
    out = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_mul(util.p32(a), util.p32(b), util.p32(out), n))
    assert all(o.fp_from_limbs32(out[12 * i:12 * i + 12]) == A[i] * B[i] % o.P for i in range(n))
    ref = util.load_ref_lib()
    chk = np.zeros_like(out)
    ref.ref_fp_mul(util.p32(a), util.p32(b), util.p32(chk), n, 4)
    assert np.array_equal(out, chk)
    for k in (0, 1, 7, 64):
        L.check(lib.b381_fp_mul_chain(util.p32(a), util.p32(b), util.p32(out), 64, k))
        assert all(o.fp_from_limbs32(out[12 * i:12 * i + 12]) == A[i] * pow(B[i], k, o.P) % o.P for i in range(64))


def test_fp2_fp12_wbasis_parity(L, lib):
    r = util.rng(42)
    n = 257
    A = [util.rf2(r) for _ in range(n)]; B = [util.rf2(r) for _ in range(n)]
    A[0], B[0] = (0, 0), (1, 2); A[1], B[1] = (o.P - 1, o.P - 1), (o.P - 1, o.P - 1)
    a = util.arr(sum((util.f2_words(x) for x in A), [])); b = util.arr(sum((util.f2_words(x) for x in B), []))
    out = np.zeros(n * 24, dtype=np.uint32)
    L.check(lib.b381_fp2_mul(util.p32(a), util.p32(b), util.p32(out), n))
    assert all(util.f2_from_words(out[24 * i:24 * i + 24]) == o.f2_mul(A[i], B[i]) for i in range(n))
    n = 40
    X = [util.rf12(r) for _ in range(n)]; Y = [util.rf12(r) for _ in range(n)]
    X[0] = o.F12_ONE; X[1] = o.F12_ZERO
    x = util.arr(sum((o.f12_to_limbs32(t) for t in X), [])); y = util.arr(sum((o.f12_to_limbs32(t) for t in Y), []))
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_fp12_mul(util.p32(x), util.p32(y), util.p32(out), n))
    assert all(o.f12_eq(f, o.f12_mul(X[i], Y[i])) for i, f in enumerate(util.f12s(out, n)))
    xm = util.arr(sum((sum((o.fp_to_limbs32(v) for v in o.myfq12_from_fq12(t)), []) for t in X), []))
    ym = util.arr(sum((sum((o.fp_to_limbs32(v) for v in o.myfq12_from_fq12(t)), []) for t in Y), []))
    L.check(lib.b381_fp12_mul_wbasis(util.p32(xm), util.p32(ym), util.p32(out), n))        # helpers.rs:248-267 (test_myfq12)
    for i in range(n):
        got = [o.fp_from_limbs32(out[144 * i + 12 * j:144 * i + 12 * j + 12]) for j in range(12)]
        assert got == o.myfq12_mul(o.myfq12_from_fq12(X[i]), o.myfq12_from_fq12(Y[i]))


def test_reference_fixed_fq12_inputs_on_gpu(L, lib):
    """fq12_target_tree.rs:447-942 inputs: a^2 == a*a and (a+b) c^2 == c^2 a + c^2 b, on the GPU."""
    v = util.ref_vectors()["fq12_arith_abc"]["fp"]
    words = [util.arr(sum((util.limbs64_to_words(l) for l in v[12 * k:12 * k + 12]), [])) for k in range(3)]
    a, b, c = [o.f12_from_limbs32(w) for w in words]
    out = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_fp12_mul(util.p32(words[2]), util.p32(words[2]), util.p32(out), 1))
    c2 = out.copy()
    assert o.f12_eq(o.f12_from_limbs32(c2), o.f12_sqr(c))
    ab = util.arr(o.f12_to_limbs32(o.f12_add(a, b)))
    L.check(lib.b381_fp12_mul(util.p32(ab), util.p32(c2), util.p32(out), 1)); lhs = o.f12_from_limbs32(out)
    L.check(lib.b381_fp12_mul(util.p32(c2), util.p32(words[0]), util.p32(out), 1)); t1 = o.f12_from_limbs32(out)
    L.check(lib.b381_fp12_mul(util.p32(c2), util.p32(words[1]), util.p32(out), 1)); t2 = o.f12_from_limbs32(out)
    assert o.f12_eq(lhs, o.f12_add(t1, t2))


def test_known_answer_generators(L, lib, z):
    kv = util.pairing_vectors()
    g1, g2 = _pairs(z, [0])
    out = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_ARK))
    assert o.f12_sha256(o.f12_from_limbs32(out)) == kv["ark_miller_g1_g2_sha256"]
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_ZK))
    assert o.f12_sha256(o.f12_from_limbs32(out)) == kv["zk_miller_g1_g2_sha256"]
    for mode in (L.MODE_ARK, L.MODE_ZK):
        L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 1, mode))
        assert o.f12_sha256(o.f12_from_limbs32(out)) == kv["e_g1_g2_sha256"]
        assert o.f12_flat(o.f12_from_limbs32(out)) == [int(h, 16) for h in kv["e_g1_g2"]]
    L.check(lib.b381_multi_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_LITERAL))
    assert list(out) == o.f12_to_limbs32(o.F12_ONE)                # src/miller_loop_native.rs:163-188 as written
    g1p = util.arr(o.fp_to_limbs32(o.G1_X) + o.fp_to_limbs32(o.G1_Y) + o.fp_to_limbs32(1))
    g2p = util.arr(util.f2_words(o.G2_X) + util.f2_words(o.G2_Y) + util.f2_words((1, 0)))
    L.check(lib.b381_literal_optimized(util.p32(g1p), util.p32(g2p), util.p32(out), 1))
    lit = o.f12_from_limbs32(out)
    assert [lit[0][0][0], lit[0][0][1]] == [int(h, 16) for h in kv["literal_g1_g2_c00"]] and all(v == 0 for v in o.f12_flat(lit)[2:])


@pytest.mark.parametrize("n", [1, 2, 31, 127, 128, 129, 256])
def test_golden_fixture_ragged_sizes(L, lib, z, n):
    idx = list(range(n))
    g1, g2 = _pairs(z, idx)
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
    assert np.array_equal(out.reshape(n, 144), z["miller_ark"][:n])
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
    assert np.array_equal(out.reshape(n, 144), z["pairing"][:n])
    fin = np.ascontiguousarray(z["miller_ark"][:n]).reshape(-1)
    L.check(lib.b381_final_exp(util.p32(fin), util.p32(out), n))
    assert np.array_equal(out.reshape(n, 144), z["pairing"][:n])


def test_tail_launch_shapes(L, lib, z):
    """Batches around the launch-shape boundaries of the pair kernels (host_api.inc pair_cfg): up to #SM x 128 pairs
    run as CTAs of 128 threads, more as a full round of 256-thread CTAs, and the last launch of a larger batch is a
    tail of either shape.  Every output must equal the fixture value of its source pair."""
    sm = ctypes.c_int(0)
    L.check(lib.b381_device_info(ctypes.byref(sm), None, None, None))
    half, full = sm.value * 128, sm.value * 256
    for n in (half - 1, half, half + 1, full, full + 1, full + half, 2 * full + 77):
        perm = np.random.default_rng(n).integers(0, 256, size=n)
        g1, g2 = _pairs(z, perm)
        out = np.zeros(n * 144, dtype=np.uint32)
        L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
        assert np.array_equal(out.reshape(n, 144), z["pairing"][perm]), n
    for n in (half, half + 1, full + 3):
        perm = np.random.default_rng(n + 1).integers(0, 256, size=n)
        g1, g2 = _pairs(z, perm)
        out = np.zeros(n * 144, dtype=np.uint32)
        L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
        assert np.array_equal(out.reshape(n, 144), z["miller_ark"][perm]), n
        fin = np.ascontiguousarray(z["miller_ark"][perm]).reshape(-1)
        L.check(lib.b381_final_exp(util.p32(fin), util.p32(out), n))
        assert np.array_equal(out.reshape(n, 144), z["pairing"][perm]), n


def test_zk_mode_batch(L, lib, z):
    n = 16
    g1, g2 = _pairs(z, list(range(n)))
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ZK))
    for i in (0, 7, 15):
        P = (o.fp_from_limbs32(z["g1"][i][:12]), o.fp_from_limbs32(z["g1"][i][12:]))
        Q = (util.f2_from_words(z["g2"][i][:24]), util.f2_from_words(z["g2"][i][24:]))
        assert o.f12_eq(o.f12_from_limbs32(out[144 * i:144 * i + 144]), o.zk_miller_loop(P, Q))
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ZK))
    assert np.array_equal(out.reshape(n, 144), z["pairing"][:n])     # ZK and ARK agree after final exponentiation


def test_identity_pairs_and_multi(L, lib, z):
    n = 37
    g1, g2 = _pairs(z, list(range(n)))
    inf = np.zeros(n, dtype=np.uint8)
    inf[[2, 9, 30]] = [1, 2, 3]
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(out), n, L.MODE_ARK))
    one = util.arr(o.f12_to_limbs32(o.F12_ONE))
    for i in range(n):
        assert np.array_equal(out[144 * i:144 * i + 144], one if inf[i] else z["pairing"][i])
    o144 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_multi_miller_loop(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(o144), n, L.MODE_ARK))
    pr = o.F12_ONE
    for i in range(n):
        if not inf[i]:
            pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i]))
    assert o.f12_eq(o.f12_from_limbs32(o144), pr)
    L.check(lib.b381_multi_pairing(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(o144), n, L.MODE_ARK))
    assert o.f12_eq(o.f12_from_limbs32(o144), o.ark_final_exponentiation(pr))
    fin = np.ascontiguousarray(z["miller_ark"][:n]).reshape(-1)
    L.check(lib.b381_fp12_product(util.p32(fin), util.p32(o144), n))
    pr2 = o.F12_ONE
    for i in range(n):
        pr2 = o.f12_mul(pr2, o.f12_from_limbs32(z["miller_ark"][i]))
    assert o.f12_eq(o.f12_from_limbs32(o144), pr2)


def test_literal_random_projective(L, lib):
    r = util.rng(43)
    pairs = util.random_pairs(44, 3)
    g1p, g2p, exp = [], [], []
    for P, Q in pairs:
        z1, z2 = util.rfp(r) or 1, util.rf2(r)
        pj = (P[0] * z1 * z1 % o.P, P[1] * z1 * z1 * z1 % o.P, z1)
        z22 = o.f2_sqr(z2)
        qj = (o.f2_mul(Q[0], z22), o.f2_mul(Q[1], o.f2_mul(z22, z2)), z2)
        g1p += sum((o.fp_to_limbs32(x) for x in pj), [])
        g2p += sum((util.f2_words(x) for x in qj), [])
        exp.append(o.literal_optimized_miller_loop(pj, qj))
    out = np.zeros(3 * 144, dtype=np.uint32)
    L.check(lib.b381_literal_optimized(util.p32(util.arr(g1p)), util.p32(util.arr(g2p)), util.p32(out), 3))
    assert all(o.f12_eq(f, e) for f, e in zip(util.f12s(out, 3), exp))


def test_error_behaviour(L, lib, z):
    g1, g2 = _pairs(z, [0, 1])
    out = np.zeros(2 * 144, dtype=np.uint32)
    assert lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 0, 0) == -2          # empty batch
    assert lib.b381_pairing(None, util.p32(g2), None, util.p32(out), 2, 0) == -2                  # null pointer
    assert lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 2, 7) == -2          # bad mode
    bad = g1.copy()
    bad[24:36] = [(o.P >> (32 * i)) & 0xFFFFFFFF for i in range(12)]                              # x = p (not canonical)
    assert lib.b381_pairing(util.p32(bad), util.p32(g2), None, util.p32(out), 2, 0) == -3
    assert b"canonical" in lib.b381_last_error()
    zero = np.zeros(144, dtype=np.uint32)
    assert lib.b381_final_exp(util.p32(zero), util.p32(out), 1) == -4                             # final_exponentiation(0) = None
    g1p = util.arr(o.fp_to_limbs32(o.G1_X) + o.fp_to_limbs32(o.G1_Y) + o.fp_to_limbs32(1))
    qinf = util.arr(util.f2_words((0, 0)) + util.f2_words((1, 0)) + util.f2_words((0, 0)))
    assert lib.b381_literal_optimized(util.p32(g1p), util.p32(qinf), util.p32(out), 1) == -4      # reference panics (f_den = 0)
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 2, 0))              # library still healthy
    assert np.array_equal(out.reshape(2, 144), z["pairing"][:2])


def test_c_port_cross_check_4096(L, lib, z):
    """4096 pairs (tiled, permuted fixture) against the oracle's C port run on the host cores."""
    n = 4096
    perm = np.random.default_rng(45).integers(0, 256, size=n)
    g1, g2 = _pairs(z, perm)
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
    ref = util.load_ref_lib()
    sample = np.arange(0, n, 16)
    s1, s2 = np.ascontiguousarray(g1.reshape(n, 24)[sample]).reshape(-1), np.ascontiguousarray(g2.reshape(n, 48)[sample]).reshape(-1)
    chk = np.zeros(len(sample) * 144, dtype=np.uint32)
    assert ref.ref_pairing(util.p32(s1), util.p32(s2), None, util.p32(chk), len(sample), 8) == 0
    assert np.array_equal(out.reshape(n, 144)[sample].reshape(-1), chk)
    assert np.array_equal(out.reshape(n, 144), z["pairing"][perm])


def test_full_size_miller_2p16(L, lib, z):
    """BASELINE config #3 size.  Properties: every output equals the fixture value of its source pair
    (tiling map), and the product of all Miller values (multi_miller_loop) equals the product of the
    per-pair outputs (checksum of checksums, via b381_fp12_product on a folded copy)."""
    n = 1 << 16
    perm = np.random.default_rng(46).integers(0, 256, size=n)
    g1, g2 = _pairs(z, perm)
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
    assert np.array_equal(out.reshape(n, 144), z["miller_ark"][perm])
    m = 4096
    o144 = np.zeros(144, dtype=np.uint32); p144 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_multi_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(o144), m, L.MODE_ARK))
    L.check(lib.b381_fp12_product(util.p32(out), util.p32(p144), m))
    assert np.array_equal(o144, p144)


def test_full_size_pairing_2p20_and_bls_shape(L, lib, z):
    """BASELINE config #4/#5 size (2^20).  (a) tiled parity against the fixture, checked through a
    SHA-256 of the whole output; (b) BLS batch-verify shape: pairs (P_i, Q_i) and (-P_i, Q_i) in
    equal numbers multiply to 1 after the final exponentiation, at any size."""
    n = 1 << 20
    perm = np.random.default_rng(47).integers(0, 256, size=n)
    g1, g2 = _pairs(z, perm)
    out = np.zeros(n * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, L.MODE_ARK))
    exp = z["pairing"][perm]
    assert hashlib.sha256(out.tobytes()).hexdigest() == hashlib.sha256(np.ascontiguousarray(exp).tobytes()).hexdigest()
    del out, exp
    half = n // 2
    g1n = g1.reshape(n, 24).copy()
    neg_y = np.zeros((256, 12), dtype=np.uint32)
    for i in range(256):
        y = o.fp_from_limbs32(z["g1"][i][12:])
        neg_y[i] = o.fp_to_limbs32((-y) % o.P)
    g1n[half:, :12] = g1n[:half, :12]                     # second half: (-P_i, Q_i)
    g1n[half:, 12:] = neg_y[perm[:half]]
    g2n = g2.reshape(n, 48).copy()
    g2n[half:] = g2n[:half]
    o144 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_multi_pairing(util.p32(g1n.reshape(-1)), util.p32(g2n.reshape(-1)), None, util.p32(o144), n, L.MODE_ARK))
    assert list(o144) == o.f12_to_limbs32(o.F12_ONE)


def test_field_microbench_sizes_against_c_port(L, lib):
    """BASELINE config #2 shape at 2^22 elements: Fp products compared in full with the oracle's C
    port; Fp12 products (2^14) likewise; plus commutativity as a size-independent property."""
    rng = np.random.default_rng(49)
    n = 1 << 22
    base = rng.integers(0, 1 << 32, size=(4096, 12), dtype=np.uint64).astype(np.uint32)
    base[:, 11] &= 0x0FFFFFFF                                 # < 2^380 < p: canonical Montgomery limbs
    a = np.ascontiguousarray(base[rng.integers(0, 4096, size=n)]).reshape(-1)
    b = np.ascontiguousarray(base[rng.integers(0, 4096, size=n)]).reshape(-1)
    out = np.zeros(n * 12, dtype=np.uint32)
    out2 = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_mul(util.p32(a), util.p32(b), util.p32(out), n))
    L.check(lib.b381_fp_mul(util.p32(b), util.p32(a), util.p32(out2), n))
    assert np.array_equal(out, out2)
    ref = util.load_ref_lib()
    chk = np.zeros(n * 12, dtype=np.uint32)
    ref.ref_fp_mul(util.p32(a), util.p32(b), util.p32(chk), n, 8)
    assert np.array_equal(out, chk)
    m = 1 << 14
    x = np.ascontiguousarray(base[rng.integers(0, 4096, size=m * 12)]).reshape(-1)
    y = np.ascontiguousarray(base[rng.integers(0, 4096, size=m * 12)]).reshape(-1)
    o12 = np.zeros(m * 144, dtype=np.uint32)
    c12 = np.zeros(m * 144, dtype=np.uint32)
    L.check(lib.b381_fp12_mul(util.p32(x), util.p32(y), util.p32(o12), m))
    ref.ref_fp12_mul(util.p32(x), util.p32(y), util.p32(c12), m, 8)
    assert np.array_equal(o12, c12)
    f2n = 1 << 18
    xa = np.ascontiguousarray(base[rng.integers(0, 4096, size=f2n * 2)]).reshape(-1)
    xb = np.ascontiguousarray(base[rng.integers(0, 4096, size=f2n * 2)]).reshape(-1)
    o2 = np.zeros(f2n * 24, dtype=np.uint32)
    L.check(lib.b381_fp2_mul(util.p32(xa), util.p32(xb), util.p32(o2), f2n))
    for i in rng.integers(0, f2n, size=64):
        A = util.f2_from_words(xa[24 * i:24 * i + 24]); B = util.f2_from_words(xb[24 * i:24 * i + 24])
        assert util.f2_from_words(o2[24 * i:24 * i + 24]) == o.f2_mul(A, B)


def test_device_pointer_api_and_launch_counter(L, lib, z):
    import torch
    n = 300
    g1, g2 = _pairs(z, np.arange(n) % 256)
    d1 = torch.from_numpy(g1.astype(np.int32)).cuda(); d2 = torch.from_numpy(g2.astype(np.int32)).cuda()
    dout = torch.zeros(n * 144, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    before = lib.b381_kernel_launches()
    L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st))
    L.check(lib.b381_check_dev(st))
    assert lib.b381_kernel_launches() == before + 1
    got = dout.cpu().numpy().astype(np.uint32).reshape(n, 144)
    assert np.array_equal(got, z["pairing"][np.arange(n) % 256])
    L.check(lib.b381_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, dout.data_ptr(), n, 0, st))
    L.check(lib.b381_final_exp_dev(dout.data_ptr(), dout.data_ptr(), n, st))
    L.check(lib.b381_check_dev(st))
    assert np.array_equal(dout.cpu().numpy().astype(np.uint32).reshape(n, 144), got)
    d144 = torch.zeros(144, dtype=torch.int32, device="cuda")
    L.check(lib.b381_multi_miller_loop_dev(d1.data_ptr(), d2.data_ptr(), None, d144.data_ptr(), n, 0, st))
    L.check(lib.b381_check_dev(st))
    pr = o.F12_ONE
    for i in range(n):
        pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i % 256]))
    assert o.f12_eq(o.f12_from_limbs32(d144.cpu().numpy().astype(np.uint32)), pr)


def test_python_host_api(L):
    import b381
    from b381.curves import G1Affine, G2Affine, G1Projective, G2Projective
    from b381.fields import Fq12, MyFq12
    kv = util.pairing_vectors()
    P, Q = G1Affine.generator(), G2Affine.generator()
    res = b381.multi_miller_loop([(P, Q)])
    assert o.f12_sha256(o.f12_unflat(res.f.flat())) == kv["ark_miller_g1_g2_sha256"]
    assert res.final_exponentiation().flat() == [int(h, 16) for h in kv["e_g1_g2"]]
    assert b381.multi_miller_loop([(P, Q)], mode=b381.miller_loop_native.MODE_LITERAL).f == Fq12.one()
    assert b381.multi_miller_loop([]).f == Fq12.one()
    assert b381.pairing_batch([(P, Q), (G1Affine.identity(), Q)])[1] == Fq12.one()
    from b381.miller_loop_native import G2Prepared, miller_loop_prepared_batch, pairing_prepared_batch
    qp = G2Prepared.from_affine(Q)
    assert miller_loop_prepared_batch([(P, qp)])[0].f == res.f
    assert pairing_prepared_batch([(P, qp), (G1Affine.identity(), qp)]) == [res.final_exponentiation(), Fq12.one()]
    from b381.miller_loop_native import G2PreparedBatch
    qb = G2PreparedBatch([Q, Q])
    assert qb.miller_loop([P, P])[1].f == res.f
    assert qb.pairing([P, G1Affine.identity()]) == [res.final_exponentiation(), Fq12.one()]
    from b381.fields import helpers as H
    from b381.fields.types import Fq, Fq2
    xs = [Fq(3), Fq(4), Fq(o.P - 1)]
    assert [x.v for x in H.inverse_fq_batch(xs)] == [o.fp_inv(x.v) for x in xs]
    assert H.is_square_fq_batch(xs) == [o.fp_legendre_is_square(x.v) for x in xs]
    assert H.sqrt_with_sgn_fq_batch([Fq(4), Fq(4)], [False, True])[0].v == 2
    assert H.pow_fq_batch(xs, [5])[0].v == 243
    z2 = Fq2(Fq(5), Fq(7))
    inv2 = H.inverse_fq2_batch([z2])[0]
    assert (inv2.c0.v, inv2.c1.v) == o.f2_inv((5, 7))
    lit = b381.optimized_miller_loop(G1Projective.generator(), G2Projective.generator())
    assert [lit.c0.c0.c0.v, lit.c0.c0.c1.v] == [int(h, 16) for h in kv["literal_g1_g2_c00"]]
    r = util.rng(48)
    a, b = util.rf12(r), util.rf12(r)
    am, bm = MyFq12.from_fq12(Fq12.from_flat(o.f12_flat(a))), MyFq12.from_fq12(Fq12.from_flat(o.f12_flat(b)))
    assert (am * bm).to_fq12().flat() == o.f12_flat(o.f12_mul(a, b))       # test_myfq12, helpers.rs:248-267


def test_g2_prepared_stage(L, lib, z):
    """b381_g2_prepare / b381_miller_loop_prepared / b381_pairing_prepared: coefficients equal the
    oracle's G2Prepared, prepared loops equal the unprepared ones bit for bit (golden fixture),
    ragged batch, identity flags, device-pointer variants, invalid input."""
    n = 300                                           # > one CTA, not a multiple of anything
    idx = np.arange(n) % 256
    g1, g2 = _pairs(z, idx)
    W = L.G2PREP_WORDS
    for mode, prep in ((L.MODE_ARK, o.ark_g2_prepare), (L.MODE_ZK, o.zk_g2_prepare)):
        co = np.zeros(n * W, dtype=np.uint32)
        L.check(lib.b381_g2_prepare(L.u32(g2)[1], L.u32(co)[1], n, mode))
        for i in (0, 5, 299):
            Q = (util.f2_from_words(z["g2"][idx[i]][:24]), util.f2_from_words(z["g2"][idx[i]][24:]))
            want = sum((util.f2_words(c) for t in prep(Q) for c in t), [])
            assert co[i * W:(i + 1) * W].tolist() == want, (mode, i)
        out = np.zeros(n * 144, dtype=np.uint32)
        ref = np.zeros(n * 144, dtype=np.uint32)
        L.check(lib.b381_miller_loop_prepared(L.u32(g1)[1], L.u32(co)[1], None, L.u32(out)[1], n, mode))
        L.check(lib.b381_miller_loop(L.u32(g1)[1], L.u32(g2)[1], None, L.u32(ref)[1], n, mode))
        assert np.array_equal(out, ref)
        if mode == L.MODE_ARK:
            assert np.array_equal(out.reshape(n, 144), z["miller_ark"][idx])
        L.check(lib.b381_pairing_prepared(L.u32(g1)[1], L.u32(co)[1], None, L.u32(out)[1], n, mode))
        assert np.array_equal(out.reshape(n, 144), z["pairing"][idx])
        inf = np.zeros(n, dtype=np.uint8); inf[3] = 1; inf[4] = 2; inf[299] = 3
        L.check(lib.b381_pairing_prepared(L.u32(g1)[1], L.u32(co)[1], inf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), L.u32(out)[1], n, mode))
        one = np.array(o.f12_to_limbs32(o.F12_ONE), dtype=np.uint32)
        got = out.reshape(n, 144)
        for i in range(n):
            assert np.array_equal(got[i], one if inf[i] else z["pairing"][idx[i]]), i
    # one prepared Q reused against many P (the BLS-verify reuse the stage exists for)
    import torch
    dev = torch.device("cuda:0")
    co1 = np.zeros(W, dtype=np.uint32)
    L.check(lib.b381_g2_prepare(L.u32(np.ascontiguousarray(z["g2"][9]))[1], L.u32(co1)[1], 1, L.MODE_ARK))
    m = 64
    dco = torch.from_numpy(np.tile(co1, m).view(np.int32)).to(dev)
    dg1 = torch.from_numpy(np.ascontiguousarray(z["g1"][:m]).reshape(-1).view(np.int32)).to(dev)
    dout = torch.empty(m * 144, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b381_miller_loop_prepared_dev(dg1.data_ptr(), dco.data_ptr(), None, dout.data_ptr(), m, L.MODE_ARK, 1, st))
    L.check(lib.b381_check_dev(st))
    got = dout.cpu().numpy().view(np.uint32).reshape(m, 144)
    for i in (0, 9, 63):
        P = (o.fp_from_limbs32(z["g1"][i][:12].tolist()), o.fp_from_limbs32(z["g1"][i][12:].tolist()))
        Q = (util.f2_from_words(z["g2"][9][:24]), util.f2_from_words(z["g2"][9][24:]))
        assert o.f12_eq(o.f12_from_limbs32(got[i].tolist()), o.ark_pairing(P, Q)), i
    # invalid arguments / non-canonical coefficients
    assert lib.b381_g2_prepare(None, L.u32(co1)[1], 1, L.MODE_ARK) == -2
    assert lib.b381_g2_prepare(L.u32(g2)[1], L.u32(co1)[1], 1, L.MODE_LITERAL) == -2
    bad = np.full(W, 0xFFFFFFFF, dtype=np.uint32)
    o1 = np.zeros(144, dtype=np.uint32)
    assert lib.b381_miller_loop_prepared(L.u32(np.ascontiguousarray(z["g1"][0]))[1], L.u32(bad)[1], None, L.u32(o1)[1], 1, L.MODE_ARK) == -3


def test_witness_helpers_batched(L, lib):
    """b381_fp_inv / sqrt / is_square / pow, b381_fp2_inv / sqrt / is_square, b381_fp6_inv, b381_fp12_inv on
    ragged batches against the oracle, plus the error codes where the reference panics."""
    r = util.rng(77)
    n = 1000
    vals = [1, 2, 3, 4, o.P - 1] + [util.rfp(r) for _ in range(n - 5)]
    a = np.array(sum((o.fp_to_limbs32(v) for v in vals), []), dtype=np.uint32)
    out = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_inv(L.u32(a)[1], L.u32(out)[1], n))
    for i in list(range(8)) + [999]:
        assert o.fp_from_limbs32(out[12 * i:12 * i + 12].tolist()) == o.fp_inv(vals[i])
    sq = np.zeros(n, dtype=np.uint8)
    L.check(lib.b381_fp_is_square(L.u32(a)[1], sq.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n))
    assert sq.tolist() == [1 if o.fp_legendre_is_square(v) else 0 for v in vals]
    # square roots of the squares among them, alternating sign request
    sqv = [v for v, f in zip(vals, sq) if f]
    m = len(sqv)
    sgn = np.array([i & 1 for i in range(m)], dtype=np.uint8)
    b = np.array(sum((o.fp_to_limbs32(v) for v in sqv), []), dtype=np.uint32)
    ro = np.zeros(m * 12, dtype=np.uint32)
    L.check(lib.b381_fp_sqrt(L.u32(b)[1], sgn.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), L.u32(ro)[1], m))
    for i in range(m):
        assert o.fp_from_limbs32(ro[12 * i:12 * i + 12].tolist()) == o.fp_sqrt_sgn(sqv[i], sgn[i]), i
    nonsq = next(v for v, f in zip(vals, sq) if not f)
    assert lib.b381_fp_sqrt(L.u32(np.array(o.fp_to_limbs32(nonsq), dtype=np.uint32))[1], None, L.u32(ro)[1], 1) == -6
    assert lib.b381_fp_inv(L.u32(np.zeros(12, dtype=np.uint32))[1], L.u32(ro)[1], 1) == -4
    exp = (ctypes.c_uint64 * 2)(0xFFFFFFFFFFFFFFFF, 3)
    L.check(lib.b381_fp_pow(L.u32(a)[1], exp, 2, L.u32(out)[1], n))
    for i in (0, 4, 500, 999):
        assert o.fp_from_limbs32(out[12 * i:12 * i + 12].tolist()) == o.pow_fq(vals[i], [0xFFFFFFFFFFFFFFFF, 3])
    # Fq2
    n2 = 400
    v2 = [(1, 0), (0, 1), (3, 0), (o.P - 1, 0)] + [util.rf2(r) for _ in range(n2 - 4 - 50)] + [o.f2_sqr(util.rf2(r)) for _ in range(50)]
    a2 = np.array(sum((util.f2_words(v) for v in v2), []), dtype=np.uint32)
    o2 = np.zeros(n2 * 24, dtype=np.uint32)
    L.check(lib.b381_fp2_inv(L.u32(a2)[1], L.u32(o2)[1], n2))
    for i in (0, 1, 2, 3, 100, 399):
        assert util.f2_from_words(o2[24 * i:24 * i + 24].tolist()) == o.f2_inv(v2[i])
    s2 = np.zeros(n2, dtype=np.uint8)
    L.check(lib.b381_fp2_is_square(L.u32(a2)[1], s2.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n2))
    assert s2.tolist() == [1 if o.f2_is_square(v) else 0 for v in v2]
    q2 = [v for v, f in zip(v2, s2) if f]
    m2 = len(q2)
    sg2 = np.array([(i >> 1) & 1 for i in range(m2)], dtype=np.uint8)
    b2 = np.array(sum((util.f2_words(v) for v in q2), []), dtype=np.uint32)
    r2 = np.zeros(m2 * 24, dtype=np.uint32)
    L.check(lib.b381_fp2_sqrt(L.u32(b2)[1], sg2.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), L.u32(r2)[1], m2))
    for i in range(m2):
        assert util.f2_from_words(r2[24 * i:24 * i + 24].tolist()) == o.f2_sqrt_sgn(q2[i], sg2[i]), i
    # Fq6 / Fq12 inverse, batch larger than one CTA
    k = 300
    v12 = [util.rf12(r) for _ in range(4)]
    a12 = np.array(sum((o.f12_to_limbs32(v12[i % 4]) for i in range(k)), []), dtype=np.uint32)
    o12 = np.zeros(k * 144, dtype=np.uint32)
    L.check(lib.b381_fp12_inv(L.u32(a12)[1], L.u32(o12)[1], k))
    for i in (0, 1, 2, 3, 299):
        assert o.f12_eq(o.f12_from_limbs32(o12[144 * i:144 * i + 144].tolist()), o.f12_inv(v12[i % 4]))
    v6 = [tuple(util.rf2(r) for _ in range(3)) for _ in range(3)]
    a6 = np.array(sum((sum((util.f2_words(c) for c in v6[i % 3]), []) for i in range(k)), []), dtype=np.uint32)
    o6 = np.zeros(k * 72, dtype=np.uint32)
    L.check(lib.b381_fp6_inv(L.u32(a6)[1], L.u32(o6)[1], k))
    for i in (0, 1, 2, 299):
        got = tuple(util.f2_from_words(o6[72 * i + 24 * j:72 * i + 24 * j + 24].tolist()) for j in range(3))
        assert got == o.f6_inv(v6[i % 3])
    assert lib.b381_fp12_inv(L.u32(np.zeros(144, dtype=np.uint32))[1], L.u32(o12)[1], 1) == -4


def test_wire_formats_batched(L, lib, z):
    """b381_fp_to_u32_digits / from / b381_fp12_to_witness_limbs and the point (de)serialisers on ragged
    batches against the oracle; invalid encodings give the documented error codes."""
    r = util.rng(55)
    n = 700
    vals = [0, 1, o.P - 1] + [util.rfp(r) for _ in range(n - 3)]
    a = np.array(sum((o.fp_to_limbs32(v) for v in vals), []), dtype=np.uint32)
    d = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_to_u32_digits(L.u32(a)[1], L.u32(d)[1], n))
    assert d.tolist() == sum((o.fp_to_u32_digits(v) for v in vals), [])
    back = np.zeros(n * 12, dtype=np.uint32)
    L.check(lib.b381_fp_from_u32_digits(L.u32(d)[1], L.u32(back)[1], n))
    assert np.array_equal(back, a)
    assert lib.b381_fp_from_u32_digits(L.u32(np.array([(o.P >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32))[1], L.u32(back)[1], 1) == -3
    k = 40
    f12 = np.ascontiguousarray(z["pairing"][:k]).reshape(-1)
    w = np.zeros(k * 144, dtype=np.uint32)
    L.check(lib.b381_fp12_to_witness_limbs(L.u32(f12)[1], L.u32(w)[1], k))
    for i in (0, 7, 39):
        assert w[144 * i:144 * i + 144].tolist() == o.f12_to_witness_limbs(o.f12_from_limbs32(z["pairing"][i].tolist()))
    u8 = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    m = 256
    g1 = np.ascontiguousarray(z["g1"][:m]).reshape(-1)
    g2 = np.ascontiguousarray(z["g2"][:m]).reshape(-1)
    inf = np.zeros(m, dtype=np.uint8); inf[5] = 1
    for c, b1, b2 in ((1, 48, 96), (0, 96, 192)):
        e1 = np.zeros(m * b1, dtype=np.uint8)
        e2 = np.zeros(m * b2, dtype=np.uint8)
        L.check(lib.b381_g1_serialize(L.u32(g1)[1], u8(inf), c, u8(e1), m))
        L.check(lib.b381_g2_serialize(L.u32(g2)[1], u8(inf), c, u8(e2), m))
        for i in (0, 5, 100, 255):
            P1 = None if inf[i] else (o.fp_from_limbs32(z["g1"][i][:12].tolist()), o.fp_from_limbs32(z["g1"][i][12:].tolist()))
            Q = None if inf[i] else (util.f2_from_words(z["g2"][i][:24].tolist()), util.f2_from_words(z["g2"][i][24:].tolist()))
            assert bytes(e1[b1 * i:b1 * (i + 1)]) == o.g1_serialize(P1, bool(c))
            assert bytes(e2[b2 * i:b2 * (i + 1)]) == o.g2_serialize(Q, bool(c))
        d1 = np.zeros(m * 24, dtype=np.uint32); d2 = np.zeros(m * 48, dtype=np.uint32)
        i1 = np.zeros(m, dtype=np.uint8); i2 = np.zeros(m, dtype=np.uint8)
        L.check(lib.b381_g1_deserialize(u8(e1), c, L.u32(d1)[1], u8(i1), m))
        L.check(lib.b381_g2_deserialize(u8(e2), c, L.u32(d2)[1], u8(i2), m))
        assert np.array_equal(i1, inf) and np.array_equal(i2, inf)
        keep = np.repeat(inf == 0, 24); keep2 = np.repeat(inf == 0, 48)
        assert np.array_equal(d1[keep], g1[keep]) and np.array_equal(d2[keep2], g2[keep2])
    # decompressed points feed the pairing directly
    out = np.zeros(m * 144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(d1)[1], L.u32(d2)[1], u8(inf), L.u32(out)[1], m, L.MODE_ARK))
    got = out.reshape(m, 144)
    assert np.array_equal(got[0], z["pairing"][0]) and np.array_equal(got[255], z["pairing"][255])
    bad = bytearray(o.g1_serialize(o.G1_GEN, True)); bad[0] &= 0x7F
    bb = np.frombuffer(bytes(bad), dtype=np.uint8).copy()
    assert lib.b381_g1_deserialize(u8(bb), 1, L.u32(d1)[1], u8(i1), 1) == -7
    x = 1
    while o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)[0] == "ok":
        x += 1
    bb = np.frombuffer(bytes([0x80]) + x.to_bytes(47, "big"), dtype=np.uint8).copy()
    assert lib.b381_g1_deserialize(u8(bb), 1, L.u32(d1)[1], u8(i1), 1) == -6


def test_subgroup_check_batched(L, lib, z):
    """b381_g1_in_subgroup / b381_g2_in_subgroup: the fixture points are in the subgroups, curve points
    decompressed from small x are (almost surely) not -- expectation from the oracle's scalar multiplication."""
    u8 = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    m = 300
    idx = np.arange(m) % 256
    g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1).copy()
    g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1).copy()
    want1 = np.ones(m, dtype=np.uint8); want2 = np.ones(m, dtype=np.uint8)
    x, put = 1, 0
    while put < 3:
        res = o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)
        if res[0] == "ok":
            g1[24 * (10 + put):24 * (11 + put)] = o.g1_to_limbs32(res[1])
            want1[10 + put] = 1 if o.g1_mul(res[1], o.R_ORDER) is None else 0
            put += 1
        x += 1
    x, put = 1, 0
    while put < 2:
        res = o.g2_deserialize(bytes([0x80]) + bytes(47) + x.to_bytes(48, "big"), True)
        if res[0] == "ok":
            g2[48 * (20 + put):48 * (21 + put)] = o.g2_to_limbs32(res[1])
            want2[20 + put] = 1 if o.g2_mul(res[1], o.R_ORDER) is None else 0
            put += 1
        x += 1
    inf = np.zeros(m, dtype=np.uint8); inf[299] = 1
    o1 = np.zeros(m, dtype=np.uint8); o2 = np.zeros(m, dtype=np.uint8)
    L.check(lib.b381_g1_in_subgroup(L.u32(g1)[1], u8(inf), u8(o1), m))
    L.check(lib.b381_g2_in_subgroup(L.u32(g2)[1], u8(inf), u8(o2), m))
    assert np.array_equal(o1, want1) and np.array_equal(o2, want2)
    assert want1[10:13].min() == 0 and want2[20:22].min() == 0


def test_scalar_mul_batched(L, lib, z):
    """b381_g1_scalar_mul / b381_g2_scalar_mul with per-element scalars against the oracle; the products
    feed the pairing: e([a] P, [b] Q) computed on the device equals e(P, Q)^(ab) (bilinearity)."""
    u8 = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    r = util.rng(99)
    m = 300
    ks = [1, 2, o.R_ORDER - 1, o.R_ORDER] + [r.randrange(1, 1 << 256) for _ in range(m - 4)]
    sc = np.array([[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for k in ks], dtype=np.uint32).reshape(-1)
    g1 = np.tile(np.array(o.g1_to_limbs32(o.G1_GEN), dtype=np.uint32), m)
    g2 = np.tile(np.array(o.g2_to_limbs32(o.G2_GEN), dtype=np.uint32), m)
    o1 = np.zeros(m * 24, dtype=np.uint32); o2 = np.zeros(m * 48, dtype=np.uint32)
    i1 = np.zeros(m, dtype=np.uint8); i2 = np.zeros(m, dtype=np.uint8)
    L.check(lib.b381_g1_scalar_mul(L.u32(g1)[1], None, L.u32(sc)[1], L.u32(o1)[1], u8(i1), m))
    L.check(lib.b381_g2_scalar_mul(L.u32(g2)[1], None, L.u32(sc)[1], L.u32(o2)[1], u8(i2), m))
    for i in (0, 1, 2, 3, 4, 150, 299):
        w1, w2 = o.g1_mul(o.G1_GEN, ks[i]), o.g2_mul(o.G2_GEN, ks[i])
        assert (i1[i] == 1 and w1 is None) or (i1[i] == 0 and o1[24 * i:24 * i + 24].tolist() == o.g1_to_limbs32(w1)), i
        assert (i2[i] == 1 and w2 is None) or (i2[i] == 0 and o2[48 * i:48 * i + 48].tolist() == o.g2_to_limbs32(w2)), i
    assert i1[3] == 1 and i2[3] == 1 and i1.sum() == 1
    # bilinearity through the device: e([a]G1, [b]G2) == e([ab]G1, G2)
    a, b = ks[10] % o.R_ORDER, ks[11] % o.R_ORDER
    ab = np.array([(a * b % o.R_ORDER >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
    pab = np.zeros(24, dtype=np.uint32); iab = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g1_scalar_mul(L.u32(g1[:24])[1], None, L.u32(ab)[1], L.u32(pab)[1], u8(iab), 1))
    e1 = np.zeros(144, dtype=np.uint32); e2 = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_pairing(L.u32(o1[240:264])[1], L.u32(o2[48 * 11:48 * 12])[1], None, L.u32(e1)[1], 1, L.MODE_ARK))
    L.check(lib.b381_pairing(L.u32(pab)[1], L.u32(g2[:48])[1], None, L.u32(e2)[1], 1, L.MODE_ARK))
    assert np.array_equal(e1, e2)


def test_point_sum_and_msm(L, lib, z):
    """b381_g1/g2_sum and b381_g1/g2_msm (scalar multiplications + 16-ary reduction tree on the device)
    against the oracle; sizes that exercise one, two and three tree levels; identity handling."""
    u8 = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    r = util.rng(111)
    for m in (1, 5, 17, 300):
        idx = np.arange(m) % 256
        g1 = np.ascontiguousarray(z["g1"][idx]).reshape(-1)
        g2 = np.ascontiguousarray(z["g2"][idx]).reshape(-1)
        inf = np.zeros(m, dtype=np.uint8)
        if m > 3:
            inf[3] = 1
        o1 = np.zeros(24, dtype=np.uint32); o2 = np.zeros(48, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
        L.check(lib.b381_g1_sum(L.u32(g1)[1], u8(inf), L.u32(o1)[1], u8(f), m))
        acc = None
        for i in range(m):
            if not inf[i]:
                acc = o.g1_add(acc, (o.fp_from_limbs32(z["g1"][idx[i]][:12].tolist()), o.fp_from_limbs32(z["g1"][idx[i]][12:].tolist())))
        assert f[0] == 0 and o1.tolist() == o.g1_to_limbs32(acc), m
        if m <= 17:
            L.check(lib.b381_g2_sum(L.u32(g2)[1], u8(inf), L.u32(o2)[1], u8(f), m))
            acc2 = None
            for i in range(m):
                if not inf[i]:
                    acc2 = o.g2_add(acc2, (util.f2_from_words(z["g2"][idx[i]][:24].tolist()), util.f2_from_words(z["g2"][idx[i]][24:].tolist())))
            assert f[0] == 0 and o2.tolist() == o.g2_to_limbs32(acc2), m
    # MSM: sum [k_i] G = [sum k_i] G
    m = 100
    ks = [r.randrange(1, 1 << 256) for _ in range(m)]
    sc = np.array([[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for k in ks], dtype=np.uint32).reshape(-1)
    g1 = np.tile(np.array(o.g1_to_limbs32(o.G1_GEN), dtype=np.uint32), m)
    g2 = np.tile(np.array(o.g2_to_limbs32(o.G2_GEN), dtype=np.uint32), m)
    o1 = np.zeros(24, dtype=np.uint32); o2 = np.zeros(48, dtype=np.uint32); f = np.zeros(1, dtype=np.uint8)
    L.check(lib.b381_g1_msm(L.u32(g1)[1], None, L.u32(sc)[1], L.u32(o1)[1], u8(f), m))
    assert f[0] == 0 and o1.tolist() == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, sum(ks) % o.R_ORDER))
    L.check(lib.b381_g2_msm(L.u32(g2)[1], None, L.u32(sc)[1], L.u32(o2)[1], u8(f), m))
    assert f[0] == 0 and o2.tolist() == o.g2_to_limbs32(o.g2_mul(o.G2_GEN, sum(ks) % o.R_ORDER))
    # the scalars r - k and k cancel: the sum is the identity
    k = ks[0] % o.R_ORDER
    sc2 = np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in (k, o.R_ORDER - k)], dtype=np.uint32).reshape(-1)
    L.check(lib.b381_g1_msm(L.u32(g1[:48])[1], None, L.u32(sc2)[1], L.u32(o1)[1], u8(f), 2))
    assert f[0] == 1


@pytest.mark.gpu
def test_g2_packed_stage(L, lib, z):
    """b381_g2_prepare_packed_dev / b381_miller_loop_packed_dev (G2Prepared in the internal, tile-interleaved layout):
    values equal the golden fixture of the unprepared path bit for bit, at ragged sizes around the tile (256), the
    half-round (148 x 128) and the round (148 x 256) boundaries, in both modes, with identity flags, with final exp."""
    import torch
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(77)
    assert lib.b381_g2_packed_words(1) == 256 * L.G2PREP_WORDS and lib.b381_g2_packed_words(257) == 512 * L.G2PREP_WORDS
    for n in (1, 255, 257, sm * 128 + 3, sm * 256 + 129):
        idx = rng.integers(0, 256, size=n)
        g1 = torch.from_numpy(np.ascontiguousarray(z["g1"][idx]).reshape(-1).view(np.int32)).to(dev)
        g2 = torch.from_numpy(np.ascontiguousarray(z["g2"][idx]).reshape(-1).view(np.int32)).to(dev)
        pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
        out = torch.empty(n * 144, dtype=torch.int32, device=dev)
        for mode in ((L.MODE_ARK, L.MODE_ZK) if n < 1000 else (L.MODE_ARK,)):
            L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr(), n, mode, st))
            L.check(lib.b381_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), None, out.data_ptr(), n, mode, 0, st))
            L.check(lib.b381_check_dev(st))
            got = out.cpu().numpy().view(np.uint32).reshape(n, 144)
            if mode == L.MODE_ARK:
                assert np.array_equal(got, z["miller_ark"][idx]), n
            else:
                ref = torch.empty_like(out)
                L.check(lib.b381_miller_loop_dev(g1.data_ptr(), g2.data_ptr(), None, ref.data_ptr(), n, mode, st))
                L.check(lib.b381_check_dev(st))
                assert torch.equal(out, ref), n
            L.check(lib.b381_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), None, out.data_ptr(), n, mode, 1, st))
            L.check(lib.b381_check_dev(st))
            assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(n, 144), z["pairing"][idx]), (n, mode)
        if n == 257:
            inf = rng.integers(0, 4, size=n).astype(np.uint8)
            dinf = torch.from_numpy(inf).to(dev)
            L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr(), n, L.MODE_ARK, st))     # pk held the ZK-mode lines
            L.check(lib.b381_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), dinf.data_ptr(), out.data_ptr(), n, L.MODE_ARK, 1, st))
            L.check(lib.b381_check_dev(st))
            got = out.cpu().numpy().view(np.uint32).reshape(n, 144)
            one = np.array(o.f12_to_limbs32(o.F12_ONE), dtype=np.uint32)
            for i in range(n):
                assert np.array_equal(got[i], one if inf[i] else z["pairing"][idx[i]]), i
    # a coordinate that is not canonical (all-ones words >= p) is reported through the context's error word; a Q flagged
    # as the identity is never read
    bad2 = torch.full((48,), -1, dtype=torch.int32, device=dev)
    pk1 = torch.empty(lib.b381_g2_packed_words(1), dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(bad2.data_ptr(), pk1.data_ptr(), 1, L.MODE_ARK, st))
    assert lib.b381_check_dev(st) == -3
    bad1 = torch.full((24,), -1, dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk1.data_ptr(), 1, L.MODE_ARK, st))
    o1 = torch.empty(144, dtype=torch.int32, device=dev)
    L.check(lib.b381_miller_loop_packed_dev(bad1.data_ptr(), pk1.data_ptr(), None, o1.data_ptr(), 1, L.MODE_ARK, 0, st))
    assert lib.b381_check_dev(st) == -3
    f2 = torch.tensor([1], dtype=torch.uint8, device=dev)             # P flagged as the identity: its words are ignored
    L.check(lib.b381_miller_loop_packed_dev(bad1.data_ptr(), pk1.data_ptr(), f2.data_ptr(), o1.data_ptr(), 1, L.MODE_ARK, 0, st))
    assert lib.b381_check_dev(st) == 0
    assert o1.cpu().numpy().view(np.uint32).tolist() == o.f12_to_limbs32(o.F12_ONE)
    # argument checks: null / misaligned buffer, LITERAL mode
    assert lib.b381_g2_prepare_packed_dev(g2.data_ptr(), None, 1, L.MODE_ARK, st) == -2
    assert lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr() + 4, 1, L.MODE_ARK, st) == -2
    assert lib.b381_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), None, out.data_ptr(), 1, L.MODE_LITERAL, 0, st) == -2


@pytest.mark.gpu
def test_multi_miller_packed(L, lib, z):
    """b381_multi_miller_loop_packed_dev (four pairs per thread, shared squarings, lines from the packed stage) equals
    b381_multi_miller_loop_dev on the same pairs: ragged sizes (partly filled groups of four tiles, several launches with
    accumulation), identity flags; with the final exponentiation a batch made of (P_i, Q_i), (-P_i, Q_i) gives one."""
    import torch
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(78)
    for n in (1, 3, 257, 1025, 4 * sm * 256 + 777):
        idx = rng.integers(0, 256, size=n)
        g1 = torch.from_numpy(np.ascontiguousarray(z["g1"][idx]).reshape(-1).view(np.int32)).to(dev)
        g2 = torch.from_numpy(np.ascontiguousarray(z["g2"][idx]).reshape(-1).view(np.int32)).to(dev)
        pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
        L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr(), n, L.MODE_ARK, st))
        a = torch.empty(144, dtype=torch.int32, device=dev); b = torch.empty(144, dtype=torch.int32, device=dev)
        for inf in (None, rng.integers(0, 4, size=n).astype(np.uint8) * (rng.integers(0, 3, size=n) == 0)):
            dinf = None if inf is None else torch.from_numpy(inf.astype(np.uint8)).to(dev)
            ip = None if inf is None else dinf.data_ptr()
            L.check(lib.b381_multi_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), ip, a.data_ptr(), n, 0, st))
            L.check(lib.b381_multi_miller_loop_dev(g1.data_ptr(), g2.data_ptr(), ip, b.data_ptr(), n, L.MODE_ARK, st))
            L.check(lib.b381_check_dev(st))
            assert torch.equal(a, b), (n, inf is None)
            if n == 3 and inf is None:
                want = o.F12_ONE
                for i in idx:
                    want = o.f12_mul(want, o.f12_from_limbs32(z["miller_ark"][i]))
                assert o.f12_eq(o.f12_from_limbs32(a.cpu().numpy().view(np.uint32).tolist()), want)
    # BLS shape: product of e(P_i, Q_i) e(-P_i, Q_i) = 1
    n = 2048
    idx = rng.integers(0, 256, size=n // 2)
    h1 = np.ascontiguousarray(z["g1"][idx]).copy()
    neg = h1.copy()
    for i in range(n // 2):
        y = o.fp_from_limbs32(h1[i][12:].tolist())
        neg[i][12:] = o.fp_to_limbs32((o.P - y) % o.P)
    g1 = torch.from_numpy(np.concatenate([h1, neg]).reshape(-1).view(np.int32)).to(dev)
    g2 = torch.from_numpy(np.ascontiguousarray(np.concatenate([z["g2"][idx], z["g2"][idx]])).reshape(-1).view(np.int32)).to(dev)
    pk = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr(), n, L.MODE_ARK, st))
    a = torch.empty(144, dtype=torch.int32, device=dev)
    L.check(lib.b381_multi_miller_loop_packed_dev(g1.data_ptr(), pk.data_ptr(), None, a.data_ptr(), n, 1, st))
    L.check(lib.b381_check_dev(st))
    assert a.cpu().numpy().view(np.uint32).tolist() == o.f12_to_limbs32(o.F12_ONE)
    assert lib.b381_multi_miller_loop_packed_dev(g1.data_ptr(), None, None, a.data_ptr(), n, 1, st) == -2


@pytest.mark.gpu
def test_multi_miller_ragged_counts_both_modes(L, lib, z):
    """the four-pairs-per-thread multi-Miller loop at every residue of the batch size mod 4 (padding pairs contribute 1),
    ARK and ZK modes, with identity flags: equals the product (b381_fp12_product) of the individual Miller values."""
    rng = np.random.default_rng(79)
    for mode in (L.MODE_ARK, L.MODE_ZK):
        for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 1023):
            idx = rng.integers(0, 256, size=n)
            g1, g2 = _pairs(z, idx)
            inf = (rng.integers(0, 4, size=n) * (rng.integers(0, 4, size=n) == 0)).astype(np.uint8)
            each = np.zeros(n * 144, dtype=np.uint32)
            L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(each), n, mode))
            want = np.zeros(144, dtype=np.uint32); got = np.zeros(144, dtype=np.uint32)
            L.check(lib.b381_fp12_product(util.p32(each), util.p32(want), n))
            L.check(lib.b381_multi_miller_loop(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(got), n, mode))
            assert np.array_equal(got, want), (mode, n)


@pytest.mark.gpu
def test_one_cached_q_against_many_p(L, lib, z):
    """b381_miller_loop_packed_one_dev: one packed prepared Q, n points P_i (broadcast line reads): equals the pairing /
    Miller loop of (P_i, Q) computed the ordinary way, across a round boundary, with identity flags on P."""
    import torch
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    n = sm * 256 + 300
    rng = np.random.default_rng(80)
    idx = rng.integers(0, 256, size=n)
    q = np.ascontiguousarray(z["g2"][17])
    g1 = torch.from_numpy(np.ascontiguousarray(z["g1"][idx]).reshape(-1).view(np.int32)).to(dev)
    g2 = torch.from_numpy(np.tile(q, n).view(np.int32)).to(dev)
    pk = torch.empty(lib.b381_g2_packed_words(1), dtype=torch.int32, device=dev)
    L.check(lib.b381_g2_prepare_packed_dev(g2.data_ptr(), pk.data_ptr(), 1, L.MODE_ARK, st))
    inf = (rng.integers(0, 8, size=n) == 0).astype(np.uint8)
    dinf = torch.from_numpy(inf).to(dev)
    a = torch.empty(n * 144, dtype=torch.int32, device=dev); b = torch.empty(n * 144, dtype=torch.int32, device=dev)
    for fe in (0, 1):
        L.check(lib.b381_miller_loop_packed_one_dev(g1.data_ptr(), pk.data_ptr(), dinf.data_ptr(), a.data_ptr(), n, L.MODE_ARK, fe, st))
        if fe:
            L.check(lib.b381_pairing_dev(g1.data_ptr(), g2.data_ptr(), dinf.data_ptr(), b.data_ptr(), n, L.MODE_ARK, st))
        else:
            L.check(lib.b381_miller_loop_dev(g1.data_ptr(), g2.data_ptr(), dinf.data_ptr(), b.data_ptr(), n, L.MODE_ARK, st))
        L.check(lib.b381_check_dev(st))
        assert torch.equal(a, b), fe
    P = (o.fp_from_limbs32(z["g1"][idx[1]][:12].tolist()), o.fp_from_limbs32(z["g1"][idx[1]][12:].tolist()))
    Q = (util.f2_from_words(q[:24].tolist()), util.f2_from_words(q[24:].tolist()))
    if not inf[1]:
        assert o.f12_eq(o.f12_from_limbs32(a[144:288].cpu().numpy().view(np.uint32).tolist()), o.ark_pairing(P, Q))


@pytest.mark.gpu
def test_regression_redc_carry_pair(L, lib):
    """the pair (one of 2^20 distinct ones, tools/big_parity.py seed 7) whose final exponentiation hit word 12 = 0xffffffff
    plus a carry in row 0 of the Montgomery reduction: pairing, Miller loop, final exponentiation and the prepared paths
    against the Python oracle's values (tests/golden/regress_redc_carry.json)."""
    import json
    import os
    g = json.load(open(os.path.join(util.ROOT, "tests", "golden", "regress_redc_carry.json")))
    g1 = np.array(g["g1"], dtype=np.uint32); g2 = np.array(g["g2"], dtype=np.uint32)
    out = np.zeros(144, dtype=np.uint32)
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_ARK))
    assert out.tolist() == g["pairing"]
    L.check(lib.b381_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_ARK))
    assert out.tolist() == g["miller_ark"]
    mil = np.array(g["miller_ark"], dtype=np.uint32)
    L.check(lib.b381_final_exp(util.p32(mil), util.p32(out), 1))
    assert out.tolist() == g["pairing"]
    L.check(lib.b381_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), 1, L.MODE_ZK))
    assert out.tolist() == g["pairing"]
    co = np.zeros(L.G2PREP_WORDS, dtype=np.uint32)
    L.check(lib.b381_g2_prepare(util.p32(g2), util.p32(co), 1, L.MODE_ARK))
    L.check(lib.b381_pairing_prepared(util.p32(g1), util.p32(co), None, util.p32(out), 1, L.MODE_ARK))
    assert out.tolist() == g["pairing"]
    # a full round of copies of the pair: every lane of every warp takes the rare path at once
    import torch
    n = 4096
    d1 = torch.from_numpy(np.tile(g1, n).view(np.int32)).cuda(); d2 = torch.from_numpy(np.tile(g2, n).view(np.int32)).cuda()
    do = torch.empty(n * 144, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b381_pairing_dev(d1.data_ptr(), d2.data_ptr(), None, do.data_ptr(), n, L.MODE_ARK, st))
    L.check(lib.b381_check_dev(st))
    assert np.array_equal(do.cpu().numpy().view(np.uint32).reshape(n, 144), np.tile(np.array(g["pairing"], dtype=np.uint32), (n, 1)))
