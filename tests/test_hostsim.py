"""The device arithmetic (fp32.cuh or fp28.cuh / tower.cuh / programs.cuh), compiled for the host
by tests/hostsim/hostsim.cpp, checked bit-for-bit against the oracle -- and, in the
-DB381_TRACK_BOUNDS build, with every magnitude / sign bound of the lazy
arithmetic asserted along the executed path (control flow is input-independent in ARK/ZK mode, so
one run covers the worst case of every operation site)."""
import ctypes
import os

import numpy as np
import pytest

import b381_oracle as o
import util


def u(n):
    return (ctypes.c_uint32 * n)()


def A(l):
    return (ctypes.c_uint32 * len(l))(*[int(x) for x in l])


@pytest.fixture(scope="module", params=[False, True], ids=["plain", "track_bounds"])
def hs(request):
    lib = util.load_hostsim(track=request.param)
    assert lib.hs_tracking() == (1 if request.param else 0)
    return lib


def test_fp_roundtrip_and_mul_edges(hs):
    r = util.rng(21)
    edge = [0, 1, 2, o.P - 1, o.P - 2, o.MONT_R_MOD_P, o.MONT_R2_MOD_P, (1 << 380), o.P >> 1]
    for v in edge + [util.rfp(r) for _ in range(40)]:
        out = u(12)
        assert hs.hs_fp_roundtrip(A(o.fp_to_limbs32(v)), out) == 0
        assert o.fp_from_limbs32(list(out)) == v
    vals = edge + [util.rfp(r) for _ in range(30)]
    for a in vals:
        for b in (vals[0], vals[3], vals[-1], util.rfp(r)):
            out = u(12)
            hs.hs_fp_mul(A(o.fp_to_limbs32(a)), A(o.fp_to_limbs32(b)), out)
            assert o.fp_from_limbs32(list(out)) == a * b % o.P
    # limbs >= p are rejected (the reference panics in Fq::from_bigint(..).unwrap())
    for bad in (o.P, o.P + 1, (1 << 384) - 1):
        out = u(12)
        assert hs.hs_fp_roundtrip(A([(bad >> (32 * i)) & 0xFFFFFFFF for i in range(12)]), out) == 1


def test_fp2_ops(hs):
    r = util.rng(22)
    cases = [((0, 0), (0, 0)), ((1, 0), (0, 1)), ((o.P - 1, o.P - 1), (o.P - 1, o.P - 1))] + [(util.rf2(r), util.rf2(r)) for _ in range(20)]
    for a, b in cases:
        out = u(24)
        hs.hs_fp2_mul(A(util.f2_words(a)), A(util.f2_words(b)), out)
        assert util.f2_from_words(list(out)) == o.f2_mul(a, b)
        if a != (0, 0):
            hs.hs_fp2_inv(A(util.f2_words(a)), out)
            assert util.f2_from_words(list(out)) == o.f2_inv(a)


def test_fp12_ops(hs):
    r = util.rng(23)
    for _ in range(4):
        a, b = util.rf12(r), util.rf12(r)
        wa, wb = A(o.f12_to_limbs32(a)), A(o.f12_to_limbs32(b))
        out = u(144)
        hs.hs_fp12_mul(wa, wb, out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_mul(a, b))
        hs.hs_fp12_sqr(wa, out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_sqr(a))
        hs.hs_fp12_inv(wa, out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_inv(a))
        for k in (1, 2, 3):
            hs.hs_fp12_frobenius(wa, k, out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_frobenius(a, k))
        c0, c1, c4 = util.rf2(r), util.rf2(r), util.rf2(r)
        hs.hs_fp12_mul_by_014(wa, A(util.f2_words(c0)), A(util.f2_words(c1)), A(util.f2_words(c4)), out)
        assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_mul_by_014(a, c0, c1, c4))
        am, bm = o.myfq12_from_fq12(a), o.myfq12_from_fq12(b)
        hs.hs_fp12_mul_wbasis(A(sum((o.fp_to_limbs32(x) for x in am), [])), A(sum((o.fp_to_limbs32(x) for x in bm), [])), out)
        assert [o.fp_from_limbs32(list(out)[12 * i:12 * i + 12]) for i in range(12)] == o.myfq12_mul(am, bm)


def test_reference_fixed_fq12_inputs(hs):
    """the reference's fixed Fq12 inputs (fq12_target_tree.rs:447-942) through the device arithmetic:
    a^2 = a a and (a+b) c^2 = c^2 a + c^2 b."""
    v = util.ref_vectors()["fq12_arith_abc"]["fp"]
    words = [sum((util.limbs64_to_words(l) for l in v[12 * k:12 * k + 12]), []) for k in range(3)]
    vals = [o.f12_from_limbs32(w) for w in words]
    out = u(144)
    for w, t in zip(words, vals):
        hs.hs_fp12_sqr(A(w), out)
        sq = list(out)
        hs.hs_fp12_mul(A(w), A(w), out)
        assert sq == list(out) and o.f12_eq(o.f12_from_limbs32(sq), o.f12_sqr(t))
    a, b, c = vals
    hs.hs_fp12_sqr(A(words[2]), out); c2 = list(out)
    hs.hs_fp12_mul(A(o.f12_to_limbs32(o.f12_add(a, b))), A(c2), out); lhs = list(out)
    hs.hs_fp12_mul(A(c2), A(words[0]), out); t1 = o.f12_from_limbs32(list(out))
    hs.hs_fp12_mul(A(c2), A(words[1]), out); t2 = o.f12_from_limbs32(list(out))
    assert o.f12_eq(o.f12_from_limbs32(lhs), o.f12_add(t1, t2))


def test_miller_final_exp_pairing_all_modes(hs):
    kv = util.pairing_vectors()
    z = util.pairs_256()
    out = u(144)
    g1, g2 = A(z["g1"][0]), A(z["g2"][0])
    assert hs.hs_miller_loop(g1, g2, 0, out, 0) == 0
    assert o.f12_sha256(o.f12_from_limbs32(list(out))) == kv["ark_miller_g1_g2_sha256"]
    assert hs.hs_miller_loop(g1, g2, 0, out, 1) == 0
    assert o.f12_sha256(o.f12_from_limbs32(list(out))) == kv["zk_miller_g1_g2_sha256"]
    for i in (0, 1, 2, 100, 255):
        g1, g2 = A(z["g1"][i]), A(z["g2"][i])
        assert hs.hs_miller_loop(g1, g2, 0, out, 0) == 0 and list(out) == list(z["miller_ark"][i])
        assert hs.hs_final_exp(A(z["miller_ark"][i]), out) == 0 and list(out) == list(z["pairing"][i])
        assert hs.hs_pairing(g1, g2, 0, out, 0) == 0 and list(out) == list(z["pairing"][i])
        assert hs.hs_pairing(g1, g2, 0, out, 1) == 0 and list(out) == list(z["pairing"][i])     # ZK == ARK after final exp
    one = o.f12_to_limbs32(o.F12_ONE)
    for inf in (1, 2, 3):
        assert hs.hs_pairing(g1, g2, inf, out, 0) == 0 and list(out) == one
        assert hs.hs_miller_loop(g1, g2, inf, out, 1) == 0 and list(out) == one
    assert hs.hs_final_exp(A([0] * 144), out) == 2               # final_exponentiation(0) is None in ark
    # cyclotomic squaring and exp_by_x on a cyclotomic element
    e = o.f12_from_limbs32(z["pairing"][5])
    hs.hs_fp12_cyclotomic_square(A(z["pairing"][5]), out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.f12_sqr(e))
    hs.hs_fp12_exp_by_x(A(z["pairing"][5]), out); assert o.f12_eq(o.f12_from_limbs32(list(out)), o.ark_exp_by_x(e))


def test_multi_miller_and_literal(hs):
    z = util.pairs_256()
    n = 5
    g1 = A(np.ascontiguousarray(z["g1"][:n]).reshape(-1)); g2 = A(np.ascontiguousarray(z["g2"][:n]).reshape(-1))
    out = u(144)
    assert hs.hs_multi_miller(g1, g2, None, ctypes.c_size_t(n), out, 0) == 0
    pr = o.F12_ONE
    for i in range(n):
        pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i]))
    assert o.f12_eq(o.f12_from_limbs32(list(out)), pr)
    # shared squarings, ZK mode, identity flags in both positions of a pair-of-pairs, odd count
    pairs = []
    for i in range(n):
        P = (o.fp_from_limbs32(z["g1"][i][:12]), o.fp_from_limbs32(z["g1"][i][12:]))
        Q = (util.f2_from_words(z["g2"][i][:24]), util.f2_from_words(z["g2"][i][24:]))
        pairs.append((P, Q))
    inf = (ctypes.c_uint8 * n)(0, 1, 2, 0, 0)
    live = [pq for k, pq in enumerate(pairs) if inf[k] == 0]
    for mode, fn in ((0, o.ark_multi_miller_loop), (1, o.zk_multi_miller_loop)):
        assert hs.hs_multi_miller(g1, g2, inf, ctypes.c_size_t(n), out, mode) == 0
        assert o.f12_eq(o.f12_from_limbs32(list(out)), fn(live)), mode
    kv = util.pairing_vectors()
    g1p = o.fp_to_limbs32(o.G1_X) + o.fp_to_limbs32(o.G1_Y) + o.fp_to_limbs32(1)
    g2p = util.f2_words(o.G2_X) + util.f2_words(o.G2_Y) + util.f2_words((1, 0))
    assert hs.hs_literal(A(g1p), A(g2p), out) == 0
    lit = o.f12_from_limbs32(list(out))
    assert [lit[0][0][0], lit[0][0][1]] == [int(h, 16) for h in kv["literal_g1_g2_c00"]]
    r = util.rng(24)
    P, Q = util.random_pairs(25, 1)[0]
    z1, z2 = util.rfp(r) or 1, util.rf2(r)
    pj = (P[0] * z1 * z1 % o.P, P[1] * z1 * z1 * z1 % o.P, z1)
    z22 = o.f2_sqr(z2)
    qj = (o.f2_mul(Q[0], z22), o.f2_mul(Q[1], o.f2_mul(z22, z2)), z2)
    assert hs.hs_literal(A(sum((o.fp_to_limbs32(x) for x in pj), [])), A(sum((util.f2_words(x) for x in qj), [])), out) == 0
    assert o.f12_eq(o.f12_from_limbs32(list(out)), o.literal_optimized_miller_loop(pj, qj))
    # Q at infinity (z = 0): f_den becomes 0 and the reference panics -> error bit 2
    qinf = ((0, 0), (1, 0), (0, 0))
    assert hs.hs_literal(A(g1p), A(sum((util.f2_words(x) for x in qinf), [])), out) == 2


def test_g2_packed_stage(hs):
    """G2Prepared in the packed (internal-format) layout: the Miller loop / pairing over it equals the on-the-fly
    loop and the golden fixture; identity flags give one."""
    import numpy as np
    z = util.pairs_256()
    for i in (0, 11):
        g1, g2 = A(z["g1"][i]), A(z["g2"][i])
        for mode in (0, 1):
            buf = np.zeros(68 * 72 + 4, dtype=np.uint32)
            off = (-buf.ctypes.data // 4) % 4                   # 16-byte alignment of the uint4 view
            pk = buf[off:off + 68 * 72]
            pkp = pk.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
            assert hs.hs_g2_prepare_packed(g2, pkp, mode) == 0
            out, ref = u(144), u(144)
            assert hs.hs_miller_packed(g1, pkp, 0, out, mode, 0) == 0
            assert hs.hs_miller_loop(g1, g2, 0, ref, mode) == 0
            assert list(out) == list(ref), (i, mode)
            if mode == 0:
                assert list(out) == list(z["miller_ark"][i])
            assert hs.hs_miller_packed(g1, pkp, 0, out, mode, 1) == 0 and list(out) == list(z["pairing"][i])
            one = o.f12_to_limbs32(o.F12_ONE)
            for inf in (1, 2, 3):
                assert hs.hs_miller_packed(g1, pkp, inf, out, mode, 0) == 0 and list(out) == one


def test_multi_miller_packed(hs):
    """four pairs per thread against packed prepared Q's with shared squarings: the product equals the product of the
    golden Miller values; identity flags and a count that is not a multiple of four (padding pairs contribute 1)."""
    import numpy as np
    z = util.pairs_256()
    n = 6
    buf = np.zeros(n * 68 * 72 + 4, dtype=np.uint32)
    off = (-buf.ctypes.data // 4) % 4
    pk = buf[off:off + n * 68 * 72]
    for i in range(n):
        pki = pk[68 * 72 * i:].ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        assert hs.hs_g2_prepare_packed(A(z["g2"][i]), pki, 0) == 0
    pkp = pk.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    g1 = A(np.ascontiguousarray(z["g1"][:n]).reshape(-1))
    out = u(144)
    for inf in (None, (0, 1, 2, 0, 3, 0)):
        want = o.F12_ONE
        for i in range(n):
            if inf is None or inf[i] == 0:
                want = o.f12_mul(want, o.f12_from_limbs32(z["miller_ark"][i]))
        cinf = None if inf is None else (ctypes.c_uint8 * n)(*inf)
        assert hs.hs_multi_miller_packed(g1, pkp, cinf, ctypes.c_size_t(n), out) == 0
        assert o.f12_eq(o.f12_from_limbs32(list(out)), want), inf


def test_redc_input_word_patterns(hs):
    """Montgomery reduction of double-width values whose words are all-ones / zero patterns, word 12 in particular: the
    rows consume words 0..12, and word 12 = 0xffffffff plus the carry of row 0 must ripple into word 13 (it did not:
    one wrong pairing in 2^20 distinct pairs, tests/golden/regress_redc_carry.json).  Reference: Python integers."""
    r = util.rng(4242)
    R = 1 << 416
    Rinv = pow(R, -1, o.P)
    out = u(13)
    n_bad = 0
    for trial in range(4000):
        words = [r.randrange(1 << 32) for _ in range(26)]
        for k in range(24, 26):
            words[k] = 0
        words[23] = r.randrange(1 << 18)                     # t < 2^754: a sum of a few products of values below 2^384 / 4
        for k in range(26):                                  # sprinkle the patterns that break "no carry out" assumptions
            c = r.randrange(8)
            if c == 0 and k < 24:
                words[k] = 0xFFFFFFFF
            elif c == 1 and k < 24:
                words[k] = 0
        if trial % 2 == 0:
            words[12] = 0xFFFFFFFF
        if trial % 8 == 0:
            words[11] = 0xFFFFFFFF; words[13] = 0xFFFFFFFF
        t = sum(w << (32 * k) for k, w in enumerate(words))
        assert hs.hs_redc(A(words), out) == 0, "acc_redc and acc_redc2 disagree"
        got = sum(int(w) << (32 * k) for k, w in enumerate(out))
        if got % o.P != t * Rinv % o.P or got > 2 * o.P + (t >> 416):
            n_bad += 1
    assert n_bad == 0, "%d of 4000 reductions wrong" % n_bad


def test_product_word_patterns(hs):
    """the even / odd carry chains of the 12- and 13-word products and of the fused three-product accumulation on
    operands made of all-ones / zero / random words (the patterns that stress every "this chain end cannot carry out"
    argument), against Python integers"""
    r = util.rng(4343)
    out = u(26)

    def operand(nwords):
        w = []
        for k in range(13):
            c = r.randrange(4)
            w.append(0 if k >= nwords else (0xFFFFFFFF if c == 0 else 0 if c == 1 else r.randrange(1 << 32)))
        if nwords == 13:
            w[12] = r.randrange(1 << 6)                          # 13-word operands stay far below 2^415
        return w

    val = lambda w: sum(x << (32 * k) for k, x in enumerate(w))
    for trial in range(1500):
        n12 = trial % 2
        nw = 12 if n12 else 13
        a, b = operand(nw), operand(nw)
        if trial % 10 == 0:
            a = [0xFFFFFFFF] * nw + [0] * (13 - nw) if n12 else a
            b = [0xFFFFFFFF] * nw + [0] * (13 - nw) if n12 else b
        hs.hs_mul_raw(A(a), A(b), n12, out)
        assert val(list(out)) == val(a) * val(b), ("product", trial)
        ops = [operand(nw) for _ in range(6)]
        if trial % 10 == 0 and n12:
            ops = [[0xFFFFFFFF] * 12 + [0]] * 6
        hs.hs_mul3_raw(A(sum(ops, [])), n12, out)
        assert val(list(out)) == sum(val(ops[2 * i]) * val(ops[2 * i + 1]) for i in range(3)), ("three products", trial)


def test_weak_reduction_word_patterns(hs):
    """fp_wreduce on values of either sign up to 2^20 p whose words are all-ones / zero patterns, and on the multiples of
    p and their neighbours: the result is the same residue in [0, 1.02 p)"""
    r = util.rng(4444)
    out = u(13)
    M = 1 << 416
    cases = []
    for k in range(-(1 << 20) + 1, 1 << 20, 65521):
        for d in (-1, 0, 1):
            cases.append(k * o.P + d)
    for _ in range(3000):
        w = []
        for k in range(13):
            c = r.randrange(4)
            w.append(0xFFFFFFFF if c == 0 else 0 if c == 1 else r.randrange(1 << 32))
        v = sum(x << (32 * k) for k, x in enumerate(w)) % (1 << 400)
        v = v % ((1 << 20) * o.P)
        cases.append(v if r.randrange(2) else -v)
    for v in cases:
        words = [((v % M) >> (32 * k)) & 0xFFFFFFFF for k in range(13)]
        hs.hs_wreduce_raw(A(words), out)
        got = sum(int(x) << (32 * k) for k, x in enumerate(out))
        assert got % o.P == v % o.P and 0 <= got < o.P * 102 // 100, hex(v)


def test_regression_redc_carry_pair(hs):
    """the pair and the cyclotomic-square input that exposed the reduction carry, against the Python oracle's values"""
    import json
    g = json.load(open(os.path.join(util.ROOT, "tests", "golden", "regress_redc_carry.json")))
    out = u(144)
    hs.hs_fp12_cyclotomic_square(A(g["cyclotomic_square_input"]), out)
    assert list(out) == g["cyclotomic_square"]
    assert hs.hs_final_exp(A(g["miller_ark"]), out) == 0 and list(out) == g["pairing"]
    assert hs.hs_miller_loop(A(g["g1"]), A(g["g2"]), 0, out, 0) == 0 and list(out) == g["miller_ark"]
    assert hs.hs_pairing(A(g["g1"]), A(g["g2"]), 0, out, 0) == 0 and list(out) == g["pairing"]


def _triples_words(co):
    return sum((util.f2_words(c) for t in co for c in t), [])


def test_g2_prepared_stage(hs):
    """G2Prepared coefficients (68 triples) against the oracle's ark_g2_prepare / zk_g2_prepare, and the
    prepared Miller loop / pairing against the golden fixture of the unprepared path."""
    z = util.pairs_256()
    for i in (0, 7):
        g1, g2 = A(z["g1"][i]), A(z["g2"][i])
        Q = (util.f2_from_words(z["g2"][i][:24]), util.f2_from_words(z["g2"][i][24:]))
        for mode, prep in ((0, o.ark_g2_prepare), (1, o.zk_g2_prepare)):
            co = u(68 * 72)
            assert hs.hs_g2_prepare(g2, co, mode) == 0
            want = prep(Q)
            assert len(want) == 68
            assert list(co) == _triples_words(want), (i, mode)
            out = u(144)
            assert hs.hs_miller_prepared(g1, co, 0, out, mode, 0) == 0
            ref = u(144)
            assert hs.hs_miller_loop(g1, g2, 0, ref, mode) == 0
            assert list(out) == list(ref)
            if mode == 0:
                assert list(out) == list(z["miller_ark"][i])
            assert hs.hs_miller_prepared(g1, co, 0, out, mode, 1) == 0 and list(out) == list(z["pairing"][i])
            one = o.f12_to_limbs32(o.F12_ONE)
            for inf in (1, 2, 3):
                assert hs.hs_miller_prepared(g1, co, inf, out, mode, 0) == 0 and list(out) == one
    # a coefficient that is not canonical is reported (bit 1), identity pairs ignore their coefficients
    bad = u(68 * 72)
    for k in range(68 * 72):
        bad[k] = 0xFFFFFFFF
    out = u(144)
    assert hs.hs_miller_prepared(A(z["g1"][0]), bad, 0, out, 0, 0) == 1
    assert hs.hs_miller_prepared(A(z["g1"][0]), bad, 2, out, 0, 0) == 0


def test_witness_helpers(hs):
    """Batched witness helpers (helpers.cuh) against the oracle: inverse, sqrt with prescribed sgn0,
    Legendre / is_square, pow_fq, Fq2 / Fq6 / Fq12 inverse; edge cases the reference panics on."""
    r = util.rng(31)
    out, out2, b8 = u(12), u(24), (ctypes.c_uint8 * 1)()
    vals = [1, 2, 3, 4, o.P - 1, o.P - 2, (o.P - 1) // 2] + [util.rfp(r) for _ in range(12)]
    for a in vals:
        wa = A(o.fp_to_limbs32(a))
        assert hs.hs_fp_inv(wa, out) == 0 and o.fp_from_limbs32(list(out)) == o.fp_inv(a)
        assert hs.hs_fp_is_square(wa, b8) == 0 and bool(b8[0]) == o.fp_legendre_is_square(a)
        for sgn in (0, 1):
            want = o.fp_sqrt_sgn(a, sgn)
            rc = hs.hs_fp_sqrt(wa, sgn, out)
            if want is None:
                assert rc == 4
            else:
                got = o.fp_from_limbs32(list(out))
                assert rc == 0 and got == want and got * got % o.P == a and o.sgn0_fq(got) == bool(sgn)
        for exp in ([5], [0], [1], [0xFFFFFFFFFFFFFFFF, 3], [(o.P - 1) // 2 & (2**64 - 1), 7, 0]):
            e32 = sum(([x & 0xFFFFFFFF, x >> 32] for x in exp), [])
            assert hs.hs_fp_pow(wa, A(e32), len(e32), out) == 0
            assert o.fp_from_limbs32(list(out)) == o.pow_fq(a, exp), exp
    assert hs.hs_fp_inv(A([0] * 12), out) == 2                       # inverse of zero: the reference unwraps None
    assert hs.hs_fp_sqrt(A([0] * 12), 0, out) == 0 and list(out) == [0] * 12
    assert hs.hs_fp_sqrt(A([0] * 12), 1, out) == 4                   # sqrt(0) cannot have sgn0 = 1
    assert hs.hs_fp_is_square(A([0] * 12), b8) == 0 and b8[0] == 0
    assert hs.hs_fp_inv(A([0xFFFFFFFF] * 12), out) & 1              # not canonical
    cases = [(1, 0), (0, 1), (4, 0), (3, 0), (o.P - 1, 0), (0, o.P - 4), (5, 7)] + [util.rf2(r) for _ in range(10)]
    cases += [o.f2_sqr(util.rf2(r)) for _ in range(6)]
    for a in cases:
        wa = A(util.f2_words(a))
        assert hs.hs_fp2_inv_ext(wa, out2) == 0 and util.f2_from_words(list(out2)) == o.f2_inv(a)
        assert hs.hs_fp2_is_square(wa, b8) == 0 and bool(b8[0]) == o.f2_is_square(a)
        for sgn in (0, 1):
            want = o.f2_sqrt_sgn(a, sgn)
            rc = hs.hs_fp2_sqrt(wa, sgn, out2)
            if want is None:
                assert rc == 4, (a, sgn)
            else:
                got = util.f2_from_words(list(out2))
                assert rc == 0 and got == want and o.f2_sqr(got) == a and o.sgn0_fq2(got) == bool(sgn)
    assert hs.hs_fp2_inv_ext(A([0] * 24), out2) == 2
    assert hs.hs_fp2_sqrt(A([0] * 24), 0, out2) == 0 and list(out2) == [0] * 24
    assert hs.hs_fp2_sqrt(A([0] * 24), 1, out2) == 4
    o72, o144 = u(72), u(144)
    for _ in range(3):
        a6 = util.rf6(r) if hasattr(util, "rf6") else tuple(util.rf2(r) for _ in range(3))
        assert hs.hs_fp6_inv_ext(A(sum((util.f2_words(c) for c in a6), [])), o72) == 0
        got6 = tuple(util.f2_from_words(list(o72)[24 * i:24 * i + 24]) for i in range(3))
        assert got6 == o.f6_inv(a6)
        a12 = util.rf12(r)
        assert hs.hs_fp12_inv_ext(A(o.f12_to_limbs32(a12)), o144) == 0
        assert o.f12_eq(o.f12_from_limbs32(list(o144)), o.f12_inv(a12))
    assert hs.hs_fp12_inv_ext(A([0] * 144), o144) == 2
    assert hs.hs_fp6_inv_ext(A([0] * 72), o72) == 2


def test_safegcd_inversion(hs):
    """fp_inv_safegcd (Bernstein-Yang division steps, fp32.cuh) against the modular inverse: structured values
    (powers of two and their neighbours, p - 2^k, values whose Montgomery form is all ones / sparse) and 400
    random ones; inverse(0) = 0 with the zero-division flag."""
    r = util.rng(977)
    out = u(12)
    vals = [1, 2, 3, o.P - 1, o.P - 2, o.P - 3, (o.P - 1) // 2, (o.P + 1) // 2]
    vals += [v % o.P for k in range(1, 381, 13) for v in (1 << k, (1 << k) - 1, o.P - (1 << k))]
    rinv = pow(1 << 384, -1, o.P)
    vals += [(m * rinv) % o.P for m in ((1 << 381) - 1, (1 << 380) + 1, 0x5555555555555555 * (1 << 300), o.P - 1, 1)]
    vals += [util.rfp(r) for _ in range(400)]
    for a in vals:
        if a == 0:
            continue
        assert hs.hs_fp_inv(A(o.fp_to_limbs32(a)), out) == 0
        got = o.fp_from_limbs32(list(out))
        assert got == pow(a, -1, o.P), hex(a)
    assert hs.hs_fp_inv(A([0] * 12), out) == 2 and list(out) == [0] * 12


def test_compressed_exp_by_x(hs):
    """The final exponentiation's exp_by_x runs 57 of its 63 squarings in Karabina's compressed form and
    decompresses three values with one shared inversion (tower.cuh f12_exp_by_x): same field element as the plain
    Granger-Scott ladder and as the oracle, on cyclotomic elements; the identity (z2 = 0) takes the fall-back."""
    z = util.pairs_256()
    out, out2 = u(144), u(144)
    for i in (0, 5, 77, 200, 255):
        a = A(z["pairing"][i])
        e = o.f12_from_limbs32(list(z["pairing"][i]))
        assert hs.hs_fp12_exp_by_x(a, out) == 0 and hs.hs_fp12_exp_by_x_plain(a, out2) == 0
        assert list(out) == list(out2) and o.f12_eq(o.f12_from_limbs32(list(out)), o.ark_exp_by_x(e))
    one = A(o.f12_to_limbs32(o.F12_ONE))
    assert hs.hs_fp12_exp_by_x(one, out) == 0 and list(out) == list(one)
    # compressed squarings alone: coefficients z4, z3, z2, z5 (slots 1, 2, 3, 5) of a^(2^n)
    e = o.f12_from_limbs32(list(z["pairing"][9]))
    w = e
    for n in range(1, 20):
        w = o.f12_cyclotomic_square(w)
        if n in (1, 2, 7, 19):
            assert hs.hs_fp12_compressed_squarings(A(z["pairing"][9]), n, out) == 0
            got = o.f12_from_limbs32(list(out))
            assert (got[0][1], got[0][2], got[1][0], got[1][2]) == (w[0][1], w[0][2], w[1][0], w[1][2])


def test_wire_formats(hs):
    """Witness digits (fq_target.rs:300-313, fq12_target.rs:408-416) and the ZCash / IETF point encodings
    against the oracle: round trips, both y signs, infinity, uncompressed, and every rejection path."""
    r = util.rng(33)
    out = u(12)
    for v in [0, 1, o.P - 1, (1 << 380) + 12345] + [util.rfp(r) for _ in range(6)]:
        assert hs.hs_fp_to_digits(A(o.fp_to_limbs32(v)), out) == 0 and list(out) == o.fp_to_u32_digits(v)
        assert hs.hs_fp_from_digits(A(o.fp_to_u32_digits(v)), out) == 0 and list(out) == o.fp_to_limbs32(v)
    assert hs.hs_fp_from_digits(A([(o.P >> (32 * i)) & 0xFFFFFFFF for i in range(12)]), out) == 1
    f = util.rf12(r)
    o144 = u(144)
    assert hs.hs_fp12_to_witness(A(o.f12_to_limbs32(f)), o144) == 0 and list(o144) == o.f12_to_witness_limbs(f)
    B = lambda b: (ctypes.c_uint8 * len(b))(*b)
    inf = (ctypes.c_uint8 * 1)()
    pts1 = [o.G1_GEN] + [o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER)) for _ in range(3)]
    pts1 += [(p[0], (o.P - p[1]) % o.P) for p in pts1]
    for pt in pts1:
        for c in (1, 0):
            enc = o.g1_serialize(pt, bool(c))
            g1 = u(24)
            assert hs.hs_g1_deserialize(B(enc), c, g1, inf) == 0 and inf[0] == 0
            assert list(g1) == o.g1_to_limbs32(pt)
            ob = (ctypes.c_uint8 * len(enc))()
            assert hs.hs_g1_serialize(g1, 0, c, ob) == 0 and bytes(ob) == enc
    pts2 = [o.G2_GEN] + [o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER)) for _ in range(3)]
    pts2 += [(p[0], o.f2_neg(p[1])) for p in pts2]
    for pt in pts2:
        for c in (1, 0):
            enc = o.g2_serialize(pt, bool(c))
            g2 = u(48)
            assert hs.hs_g2_deserialize(B(enc), c, g2, inf) == 0 and inf[0] == 0
            assert list(g2) == o.g2_to_limbs32(pt)
            ob = (ctypes.c_uint8 * len(enc))()
            assert hs.hs_g2_serialize(g2, 0, c, ob) == 0 and bytes(ob) == enc
    # infinity
    for c, n1, n2 in ((1, 48, 96), (0, 96, 192)):
        g1, g2 = u(24), u(48)
        assert hs.hs_g1_deserialize(B(o.g1_serialize(None, bool(c))), c, g1, inf) == 0 and inf[0] == 1
        assert hs.hs_g2_deserialize(B(o.g2_serialize(None, bool(c))), c, g2, inf) == 0 and inf[0] == 1
        ob = (ctypes.c_uint8 * n1)()
        assert hs.hs_g1_serialize(g1, 1, c, ob) == 0 and bytes(ob) == o.g1_serialize(None, bool(c))
        ob = (ctypes.c_uint8 * n2)()
        assert hs.hs_g2_serialize(g2, 1, c, ob) == 0 and bytes(ob) == o.g2_serialize(None, bool(c))
    # rejections: wrong compression flag, infinity with stray bits, x >= p, x not on the curve, y off the curve
    g1 = u(24)
    good = bytearray(o.g1_serialize(o.G1_GEN, True))
    bad = bytearray(good); bad[0] &= 0x7F
    assert hs.hs_g1_deserialize(B(bytes(bad)), 1, g1, inf) & 8
    bad = bytearray(o.g1_serialize(None, True)); bad[47] = 1
    assert hs.hs_g1_deserialize(B(bytes(bad)), 1, g1, inf) & 8
    bad = bytearray(o.P.to_bytes(48, "big")); bad[0] |= 0x80
    assert hs.hs_g1_deserialize(B(bytes(bad)), 1, g1, inf) & 1
    x = 1
    while o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)[0] == "ok":
        x += 1
    assert hs.hs_g1_deserialize(B(bytes([0x80]) + x.to_bytes(47, "big")), 1, g1, inf) & 4
    unc = bytearray(o.g1_serialize(o.G1_GEN, False)); unc[95] ^= 1
    assert hs.hs_g1_deserialize(B(bytes(unc)), 0, g1, inf) & 4
    g2 = u(48)
    unc = bytearray(o.g2_serialize(o.G2_GEN, False)); unc[191] ^= 1
    assert hs.hs_g2_deserialize(B(bytes(unc)), 0, g2, inf) & 4


def test_subgroup_check(hs):
    """[r] P == infinity on the host simulation: subgroup points, curve points outside the subgroup
    (found by decompressing small x), the identity flag; G1 (embedded in Fq2) and G2."""
    b = (ctypes.c_uint8 * 1)()
    r = util.rng(35)
    for _ in range(2):
        P1 = o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER))
        Q = o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER))
        assert hs.hs_subgroup_check(A(o.g1_to_limbs32(P1)), 0, 0, b) == 0 and b[0] == 1
        assert hs.hs_subgroup_check(A(o.g2_to_limbs32(Q)), 1, 0, b) == 0 and b[0] == 1
    x = 1
    found = 0
    while found < 2:
        res = o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)
        if res[0] == "ok":
            want = o.g1_mul(res[1], o.R_ORDER) is None
            assert hs.hs_subgroup_check(A(o.g1_to_limbs32(res[1])), 0, 0, b) == 0 and bool(b[0]) == want
            found += 0 if want else 1
        x += 1
    x = 1
    while True:
        res = o.g2_deserialize(bytes([0x80]) + bytes(47) + x.to_bytes(48, "big"), True)
        if res[0] == "ok":
            want = o.g2_mul(res[1], o.R_ORDER) is None
            assert hs.hs_subgroup_check(A(o.g2_to_limbs32(res[1])), 1, 0, b) == 0 and bool(b[0]) == want
            if not want:
                break
        x += 1
    assert hs.hs_subgroup_check(A([0] * 24), 0, 1, b) == 0 and b[0] == 1


def test_scalar_mul(hs):
    """[k] P on the host simulation against the oracle: small, order-related and random 256-bit scalars."""
    b = (ctypes.c_uint8 * 1)()
    r = util.rng(36)
    base1 = o.g1_mul(o.G1_GEN, 7)
    base2 = o.g2_mul(o.G2_GEN, 11)
    for k in [1, 2, 3, o.R_ORDER - 1, o.R_ORDER, o.R_ORDER + 5, r.randrange(1, 1 << 255), (1 << 256) - 1]:
        kw = A([(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
        o1, o2 = u(24), u(48)
        assert hs.hs_scalar_mul(A(o.g1_to_limbs32(base1)), 0, 0, kw, o1, b) == 0
        w = o.g1_mul(base1, k)
        assert (b[0] == 1 and w is None) or (b[0] == 0 and list(o1) == o.g1_to_limbs32(w)), k
        assert hs.hs_scalar_mul(A(o.g2_to_limbs32(base2)), 1, 0, kw, o2, b) == 0
        w = o.g2_mul(base2, k)
        assert (b[0] == 1 and w is None) or (b[0] == 0 and list(o2) == o.g2_to_limbs32(w)), k
    assert hs.hs_scalar_mul(A(o.g1_to_limbs32(base1)), 0, 0, A([0] * 8), u(24), b) == 0 and b[0] == 1
    assert hs.hs_scalar_mul(A(o.g1_to_limbs32(base1)), 0, 1, A([5] + [0] * 7), u(24), b) == 0 and b[0] == 1


def test_point_sum(hs):
    """One level of the G2 point-reduction tree against the oracle's group law: distinct points, a repeated
    point (doubling case), P + (-P), identity flags."""
    r = util.rng(37)
    pts = [o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER)) for _ in range(4)]
    pts += [pts[0], (pts[1][0], o.f2_neg(pts[1][1]))]
    words = sum((o.g2_to_limbs32(p) for p in pts), [])
    inf = (ctypes.c_uint8 * len(pts))(*[1 if i == 2 else 0 for i in range(len(pts))])
    out, f = u(48), (ctypes.c_uint8 * 1)()
    assert hs.hs_g2_point_sum(A(words), inf, ctypes.c_size_t(len(pts)), out, f) == 0
    acc = None
    for i, p in enumerate(pts):
        if i != 2:
            acc = o.g2_add(acc, p)
    assert f[0] == 0 and list(out) == o.g2_to_limbs32(acc)
    q = o.g2_mul(o.G2_GEN, 5)
    w2 = o.g2_to_limbs32(q) + o.g2_to_limbs32((q[0], o.f2_neg(q[1])))
    assert hs.hs_g2_point_sum(A(w2), None, ctypes.c_size_t(2), out, f) == 0 and f[0] == 1


def test_g1_bucket_msm(hs):
    """the G1 bucket method staged as on the device (host simulation of every kernel), against the oracle:
    random points and 256-bit scalars, repeated points (doubling inside a bucket), P and -P in one bucket,
    identity flags, zero scalars, several window widths / chunk sizes."""
    r = util.rng(38)
    f = (ctypes.c_uint8 * 1)()
    for n, c, ch in ((1, 4, 4), (9, 3, 2), (40, 5, 8), (40, 7, 32)):
        ks = [r.randrange(0, 1 << 256) for _ in range(n)]
        pts = [o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER)) for _ in range(n)]
        inf = [0] * n
        if n >= 9:
            pts[3] = pts[2]; ks[3] = ks[2]                         # same point, same digits: doubling inside buckets
            pts[5] = o.g1_neg(pts[4]); ks[5] = ks[4]               # P and -P with equal scalars: buckets cancel
            ks[6] = 0
            inf[7] = 1
            ks[8] = (1 << 256) - 1
        want = None
        for p_, k_, i_ in zip(pts, ks, inf):
            if not i_:
                want = o.g1_add(want, o.g1_mul(p_, k_))
        words = sum((o.g1_to_limbs32(p_) for p_ in pts), [])
        sc = sum(([(k_ >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for k_ in ks), [])
        out = u(24)
        assert hs.hs_g1_msm(A(words), (ctypes.c_uint8 * n)(*inf), A(sc), ctypes.c_size_t(n), c, ch, out, f) == 0
        assert (f[0] == 1 and want is None) or (f[0] == 0 and list(out) == o.g1_to_limbs32(want)), (n, c, ch)


def test_g2_bucket_msm(hs):
    """the G2 bucket method staged as on the device, against the oracle: the cases of test_g1_bucket_msm over Fq2."""
    r = util.rng(39)
    f = (ctypes.c_uint8 * 1)()
    for n, c, ch in ((1, 4, 4), (9, 3, 2), (24, 5, 8)):
        ks = [r.randrange(0, 1 << 256) for _ in range(n)]
        pts = [o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER)) for _ in range(n)]
        inf = [0] * n
        if n >= 9:
            pts[3] = pts[2]; ks[3] = ks[2]
            pts[5] = o.g2_neg(pts[4]); ks[5] = ks[4]
            ks[6] = 0
            inf[7] = 1
            ks[8] = (1 << 256) - 1
        want = None
        for p_, k_, i_ in zip(pts, ks, inf):
            if not i_:
                want = o.g2_add(want, o.g2_mul(p_, k_))
        words = sum((o.g2_to_limbs32(p_) for p_ in pts), [])
        sc = sum(([(k_ >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for k_ in ks), [])
        out = u(48)
        assert hs.hs_g2_msm(A(words), (ctypes.c_uint8 * n)(*inf), A(sc), ctypes.c_size_t(n), c, ch, out, f) == 0
        assert (f[0] == 1 and want is None) or (f[0] == 0 and list(out) == o.g2_to_limbs32(want)), (n, c, ch)


def test_literal_vertical_line_branch(hs):
    """x = 0 gives 2 (0, y) = (0, -y) under the a = 0 doubling formula: at bit 16 the running point is -Q and
    optimized_line_function takes its vertical-line branch (src/miller_loop_native_optimized.rs:62-77); the sum is
    the identity, the next tangent has den = 0, f_den = 0: reference panics, oracle raises, error bit 2 here.
    The identity of ark-ec 0.4, (1, 1, 0), as P gives zp = 0 and the same outcome."""
    out = u(144)
    pp = (o.G1_X, o.G1_Y, 1)
    qq = ((0, 0), (5, 7), (1, 0))
    with pytest.raises(ZeroDivisionError):
        o.literal_optimized_miller_loop(pp, qq)
    assert hs.hs_literal(A(sum((o.fp_to_limbs32(x) for x in pp), [])), A(sum((util.f2_words(x) for x in qq), [])), out) == 2
    ident = (1, 1, 0)
    g2p = util.f2_words(o.G2_X) + util.f2_words(o.G2_Y) + util.f2_words((1, 0))
    with pytest.raises(ZeroDivisionError):
        o.literal_optimized_miller_loop(ident, (o.G2_X, o.G2_Y, (1, 0)))
    assert hs.hs_literal(A(sum((o.fp_to_limbs32(x) for x in ident), [])), A(g2p), out) == 2


def test_adversarial_operands(hs):
    """carry-stressing canonical values (all-ones words, 2^k - 1, p - 1, Montgomery pre-images of such patterns)
    through the Fp / Fp2 / Fp12 products, the cyclotomic chain and a Miller loop of the host simulation."""
    P = o.P
    rinv = o.MONT_RINV
    S = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 380) - 1, (1 << 380) + 1, (1 << 352) - 1, (1 << 256) - 1, (1 << 32) - 1, 1 << 32,
         int("aaaaaaaa" * 12, 16) % P, int("7fffffff" * 12, 16) % P]
    S += [rep * rinv % P for rep in ((1 << 352) - 1, (1 << 380) - 1, P - 1, (1 << 96) - 1, 0xFFFFFFFF, (P >> 1) + 1)]
    out = u(12)
    for a in S:
        for b in S:
            hs.hs_fp_mul(A(o.fp_to_limbs32(a)), A(o.fp_to_limbs32(b)), out)
            assert o.fp_from_limbs32(list(out)) == a * b % P
    r = util.rng(31)
    o2 = u(24)
    for _ in range(200):
        a, b = (r.choice(S), r.choice(S)), (r.choice(S), r.choice(S))
        hs.hs_fp2_mul(A(util.f2_words(a)), A(util.f2_words(b)), o2)
        assert util.f2_from_words(list(o2)) == o.f2_mul(a, b)
    o12 = u(144)
    for k in range(12):
        x = o.f12_unflat([r.choice(S) for _ in range(12)]) if k else o.f12_unflat([P - 1] * 12)
        y = o.f12_unflat([r.choice(S) for _ in range(12)]) if k else o.f12_unflat([P - 1] * 12)
        hs.hs_fp12_mul(A(o.f12_to_limbs32(x)), A(o.f12_to_limbs32(y)), o12)
        assert o.f12_eq(o.f12_from_limbs32(list(o12)), o.f12_mul(x, y))
        if k < 3:
            assert hs.hs_final_exp(A(o.f12_to_limbs32(x)), o12) == 0
            assert o.f12_eq(o.f12_from_limbs32(list(o12)), o.ark_final_exponentiation(x))


def _curve_points_outside_subgroups():
    """curve points decompressed from small x: (almost surely) outside the prime-order subgroups"""
    g1s, g2s = [], []
    x = 1
    while len(g1s) < 3:
        res = o.g1_deserialize(bytes([0x80]) + x.to_bytes(47, "big"), True)
        if res[0] == "ok":
            g1s.append(res[1])
        x += 1
    x = 1
    while len(g2s) < 2:
        res = o.g2_deserialize(bytes([0x80]) + bytes(47) + x.to_bytes(48, "big"), True)
        if res[0] == "ok":
            g2s.append(res[1])
        x += 1
    return g1s, g2s


def test_endomorphism_subgroup_checks_and_cofactor_clearing(hs):
    """the endomorphism tests (G1: (BETA x, y) == -[x^2] P; G2: psi(P) == [x] P) agree with [r] P == identity, and
    cofactor clearing equals multiplication by the RFC 9380 effective cofactors (oracle: g1/g2_clear_cofactor)."""
    b = (ctypes.c_uint8 * 1)()
    g1s, g2s = _curve_points_outside_subgroups()
    r = util.rng(37)
    g1s += [o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER)), o.G1_GEN]
    g2s += [o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER))]
    for P1 in g1s:
        want = o.g1_mul(P1, o.R_ORDER) is None
        assert o.g1_in_subgroup_fast(P1) == want
        assert hs.hs_subgroup_check(A(o.g1_to_limbs32(P1)), 0, 0, b) == 0 and bool(b[0]) == want
        out = u(24)
        assert hs.hs_clear_cofactor(A(o.g1_to_limbs32(P1)), 0, 0, out, b) == 0
        C = o.g1_clear_cofactor(P1)
        assert C == o.g1_mul(P1, o.G1_H_EFF)
        assert (b[0] == 1 and C is None) or (b[0] == 0 and list(out) == o.g1_to_limbs32(C))
    for Q in g2s:
        want = o.g2_mul(Q, o.R_ORDER) is None
        assert o.g2_in_subgroup_fast(Q) == want
        assert hs.hs_subgroup_check(A(o.g2_to_limbs32(Q)), 1, 0, b) == 0 and bool(b[0]) == want
        out = u(48)
        assert hs.hs_clear_cofactor(A(o.g2_to_limbs32(Q)), 1, 0, out, b) == 0
        C = o.g2_clear_cofactor(Q)
        assert C == o.g2_mul(Q, o.G2_H_EFF)
        assert b[0] == 0 and list(out) == o.g2_to_limbs32(C)
        assert o.g2_in_subgroup_fast(C)
    assert hs.hs_clear_cofactor(A([0] * 24), 0, 1, u(24), b) == 0 and b[0] == 1
