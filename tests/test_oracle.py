"""The oracle pinned against the reference's own fixed vectors / constants and the survey-derived
known answers; the C port pinned against the Python restatement."""
import ctypes
import hashlib

import numpy as np

import b381_oracle as o
import util


# ---- reference constants (pin Montgomery R = 2^384 and the tower) ---------------------------------
def test_generators_are_montgomery_form_of_canonical_generators():
    v = util.ref_vectors()
    g1 = [util.mont_decode(l) for l in v["g1_generator"]["fp"]]
    assert tuple(g1) == (o.G1_X, o.G1_Y)                     # src/fields_as_trees/g1_curve.rs:53-74
    g2 = [util.mont_decode(l) for l in v["g2_generator"]["fp"]]
    assert (g2[0], g2[1]) == o.G2_X and (g2[2], g2[3]) == o.G2_Y   # g2_curve.rs:63-117
    assert o.g1_on_curve(o.G1_GEN) and o.g2_on_curve(o.G2_GEN)
    # raw limbs are NOT on the curve: the limbs are Montgomery form
    raw = [util.limbs64_to_int(l) for l in v["g1_generator"]["fp"]]
    assert not o.g1_on_curve((raw[0], raw[1]))
    assert o.g1_mul(o.G1_GEN, o.R_ORDER) is None and o.g2_mul(o.G2_GEN, o.R_ORDER) is None


def test_frobenius_constants_match_reference():
    v = util.ref_vectors()
    c = [util.mont_decode(l) for l in v["frob_fq12_c1"]["fp"]]
    assert (c[0], c[1]) == o.FROB_GAMMA[1][1]                # fq12_target_tree.rs:96-124: xi^((p-1)/6)
    d = [util.mont_decode(l) for l in v["frob_fq6_c1_c2"]["fp"]]
    assert o.FROB_GAMMA[1][2] == (0, d[0])                   # fq6_target_tree.rs:139-146: xi^((p-1)/3)
    assert o.FROB_GAMMA[1][4] == (d[1], 0)                   # fq6_target_tree.rs:155-162: xi^((2p-2)/3)


def test_fq2_add_sub_fixed_vectors():
    v = util.ref_vectors()
    for name, fn in (("fq2_add", o.f2_add), ("fq2_sub", o.f2_sub)):
        x = [util.mont_decode(l) for l in v[name]["fp"]]
        a, b, c = (x[0], x[1]), (x[2], x[3]), (x[4], x[5])
        assert fn(a, b) == c, name                            # fq2_target_tree.rs:220-307, :335-420
        raw = [util.limbs64_to_int(l) for l in v[name]["fp"]]  # linear => also valid on the raw limbs
        assert fn((raw[0], raw[1]), (raw[2], raw[3])) == (raw[4], raw[5])


def _fq6(vals):
    return tuple((vals[2 * i], vals[2 * i + 1]) for i in range(3))


def test_fq6_fq12_fixed_input_identities():
    v = util.ref_vectors()
    x = [util.mont_decode(l) for l in v["fq6_arith_abc"]["fp"]]    # fq6_target_tree.rs:391-647
    a, b, c = _fq6(x[0:6]), _fq6(x[6:12]), _fq6(x[12:18])
    assert o.f6_sqr(a) == o.f6_mul(a, a) and o.f6_sqr(b) == o.f6_mul(b, b) and o.f6_sqr(c) == o.f6_mul(c, c)
    c2 = o.f6_sqr(c)
    assert o.f6_mul(o.f6_add(a, b), c2) == o.f6_add(o.f6_mul(c2, a), o.f6_mul(c2, b))
    assert o.f6_mul(a, o.f6_inv(a)) == o.F6_ONE
    y = [util.mont_decode(l) for l in v["fq12_arith_abc"]["fp"]]   # fq12_target_tree.rs:447-942
    A, B, C = o.f12_unflat(y[0:12]), o.f12_unflat(y[12:24]), o.f12_unflat(y[24:36])
    for t in (A, B, C):
        assert o.f12_eq(o.f12_sqr(t), o.f12_mul(t, t))
        assert o.f12_eq(o.f12_mul(t, o.f12_inv(t)), o.F12_ONE)
    C2 = o.f12_sqr(C)
    assert o.f12_eq(o.f12_mul(o.f12_add(A, B), C2), o.f12_add(o.f12_mul(C2, A), o.f12_mul(C2, B)))


def test_constants():
    assert o.PSEUDO_BINARY_ENCODING == [
        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1]   # src/global_constants.rs:3-6
    assert o.ATE_LOOP_COUNT == o.BLS_X and o.BLS_X.bit_length() - 2 == o.LOG_ATE_LOOP_COUNT
    x = -o.BLS_X
    assert o.R_ORDER == x ** 4 - x ** 2 + 1 and o.P == (x - 1) ** 2 * o.R_ORDER // 3 + x
    assert o.N0_32 == 0xFFFCFFFD and o.N0_64 == 0x89F3FFFCFFFCFFFD


# ---- native helper types (the reference's only native-vs-ark test: src/fields/helpers.rs:248-267) ----
def test_myfq12_matches_tower():
    r = util.rng(11)
    for _ in range(5):
        a, b = util.rf12(r), util.rf12(r)
        am, bm = o.myfq12_from_fq12(a), o.myfq12_from_fq12(b)
        assert o.f12_eq(o.myfq12_to_fq12(o.myfq12_add(am, bm)), o.f12_add(a, b))
        assert o.f12_eq(o.myfq12_to_fq12(o.myfq12_mul(am, bm)), o.f12_mul(a, b))
        assert o.myfq12_to_fq12(am) == a
    a6 = _fq6([util.rfp(r) for _ in range(6)])
    assert o.myfq6_to_fq6(o.myfq6_from_fq6(a6)) == a6


def test_naf_pow_sgn0():
    r = util.rng(12)
    for _ in range(10):
        e = [r.getrandbits(64) for _ in range(r.randrange(1, 4))]
        naf = o.get_naf(e)
        assert sum(d << i for i, d in enumerate(naf)) == sum(v << (64 * i) for i, v in enumerate(e))
        assert all(d in (-1, 0, 1) for d in naf)      # (the reference's per-limb carry can break non-adjacency at limb boundaries)
        a = util.rfp(r) or 1
        if any(e):
            assert o.pow_fq(a, e) == pow(a, sum(v << (64 * i) for i, v in enumerate(e)), o.P)
    assert o.get_naf([(1 << 64) - 1])[-1] == 1
    assert o.sgn0_fq(1) and not o.sgn0_fq(2) and not o.sgn0_fq(0)
    assert o.sgn0_fq2((0, 1)) and not o.sgn0_fq2((2, 1)) and o.sgn0_fq2((1, 0))


# ---- pairing-level known answers (survey-derived; the reference pins none) ----------------------------
def test_known_answers():
    kv = util.pairing_vectors()
    m = o.ark_miller_loop(o.G1_GEN, o.G2_GEN)
    e = o.ark_final_exponentiation(m)
    assert o.f12_sha256(m) == kv["ark_miller_g1_g2_sha256"] == "71f4207fad85a47d5aa8e07f04e8a6ea40675cc06175c95ef4e571d904c0ba2d"
    assert o.f12_sha256(e) == kv["e_g1_g2_sha256"] == "ff9912603bb02b77bc6ec1deaeddf9d1fee40ac17a781fb13c9c6e7a9f74d22b"
    assert [int(h, 16) for h in kv["e_g1_g2"]] == o.f12_flat(e)
    assert o.f12_flat(e)[0] >> 320 == 0x1250EBD871FC0A92
    lit = o.literal_optimized_miller_loop((o.G1_X, o.G1_Y, 1), (o.G2_X, o.G2_Y, (1, 0)))
    assert [lit[0][0][0], lit[0][0][1]] == [int(h, 16) for h in kv["literal_g1_g2_c00"]]
    assert all(v == 0 for v in o.f12_flat(lit)[2:])
    assert o.f12_eq(o.literal_multi_miller_loop([(o.G1_GEN, o.G2_GEN)]), o.F12_ONE)
    assert len(o.ark_g2_prepare(o.G2_GEN)) == 68
    o.reset_counter(); o.ark_miller_loop(o.G1_GEN, o.G2_GEN); assert o.fp_muls() == kv["fp_muls_miller"] == 6952
    o.reset_counter(); o.ark_final_exponentiation(m); assert o.fp_muls() == kv["fp_muls_final_exp"] == 7675


def test_three_constructions_agree_and_group_laws():
    r = util.rng(13)
    a, b = r.randrange(1, o.R_ORDER), r.randrange(1, o.R_ORDER)
    P, Q = o.g1_mul(o.G1_GEN, a), o.g2_mul(o.G2_GEN, b)
    e = o.ark_pairing(P, Q)
    assert o.f12_eq(o.zk_final_exponentiation(o.zk_miller_loop(P, Q)), e)
    assert o.f12_eq(o.ark_final_exponentiation(o.textbook_miller_loop(P, Q)), e)
    assert o.f12_eq(o.final_exponentiation_plain(o.ark_miller_loop(P, Q)), e)
    assert not o.f12_eq(o.zk_miller_loop(P, Q), o.ark_miller_loop(P, Q))          # raw values differ (F8)
    base = o.ark_pairing(o.G1_GEN, o.G2_GEN)
    assert o.f12_eq(e, o.f12_pow(base, a * b % o.R_ORDER))                        # bilinearity
    assert o.f12_eq(o.f12_pow(e, o.R_ORDER), o.F12_ONE) and not o.f12_eq(e, o.F12_ONE)
    assert o.f12_eq(o.ark_pairing(None, Q), o.F12_ONE) and o.f12_eq(o.ark_pairing(P, None), o.F12_ONE)
    negP = (P[0], (-P[1]) % o.P)
    assert o.f12_eq(o.ark_multi_pairing([(P, Q), (negP, Q)]), o.F12_ONE)          # BLS-verify shape
    assert o.f12_eq(o.ark_multi_miller_loop([(P, Q), (o.G1_GEN, o.G2_GEN)]),
                    o.f12_mul(o.ark_miller_loop(P, Q), o.ark_miller_loop(o.G1_GEN, o.G2_GEN)))


def test_golden_fixture_is_consistent():
    z = util.pairs_256()
    assert hashlib.sha256(z["pairing"].tobytes()).hexdigest() == "dbd021d2be5dc7037b7bbe161bc13ec02951d6b4b753d642ce34e630966ab112"
    assert o.f12_sha256(o.f12_from_limbs32(z["pairing"][0])) == "ff9912603bb02b77bc6ec1deaeddf9d1fee40ac17a781fb13c9c6e7a9f74d22b"
    for i in (1, 77, 255):
        a, b = int(str(z["scalars_a"][i]), 16), int(str(z["scalars_b"][i]), 16)
        assert list(z["g1"][i]) == o.g1_to_limbs32(o.g1_mul(o.G1_GEN, a))
        assert list(z["g2"][i]) == o.g2_to_limbs32(o.g2_mul(o.G2_GEN, b))


# ---- C port vs Python restatement and the fixture ------------------------------------------------------
def test_c_port_matches_fixture_and_python():
    lib = util.load_ref_lib()
    z = util.pairs_256()
    n = 64
    g1 = np.ascontiguousarray(z["g1"][:n]).reshape(-1)
    g2 = np.ascontiguousarray(z["g2"][:n]).reshape(-1)
    out = np.zeros(n * 144, dtype=np.uint32)
    assert lib.ref_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(out), n, 4) == 0
    assert np.array_equal(out.reshape(n, 144), z["miller_ark"][:n])
    assert lib.ref_pairing(util.p32(g1), util.p32(g2), None, util.p32(out), n, 4) == 0
    assert np.array_equal(out.reshape(n, 144), z["pairing"][:n])
    fin = np.ascontiguousarray(z["miller_ark"][:8]).reshape(-1)
    out8 = np.zeros(8 * 144, dtype=np.uint32)
    assert lib.ref_final_exp(util.p32(fin), util.p32(out8), 8, 2) == 0
    assert np.array_equal(out8.reshape(8, 144), z["pairing"][:8])
    inf = np.zeros(n, dtype=np.uint8); inf[3] = 1; inf[5] = 2
    assert lib.ref_pairing(util.p32(g1), util.p32(g2), util.p8(inf), util.p32(out), n, 4) == 0
    one = util.arr(o.f12_to_limbs32(o.F12_ONE))
    assert np.array_equal(out[3 * 144:4 * 144], one) and np.array_equal(out[5 * 144:6 * 144], one)
    o144 = np.zeros(144, dtype=np.uint32)
    assert lib.ref_multi_miller_loop(util.p32(g1), util.p32(g2), None, util.p32(o144), 4, 2) == 0
    pr = o.F12_ONE
    for i in range(4):
        pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i]))
    assert o.f12_eq(o.f12_from_limbs32(o144), pr)
    r = util.rng(14)
    A, B = [util.rfp(r) for _ in range(50)], [util.rfp(r) for _ in range(50)]
    A[0], A[1], B[1] = 0, o.P - 1, o.P - 1
    a, b = util.arr(sum((o.fp_to_limbs32(x) for x in A), [])), util.arr(sum((o.fp_to_limbs32(x) for x in B), []))
    c = np.zeros(50 * 12, dtype=np.uint32)
    lib.ref_fp_mul(util.p32(a), util.p32(b), util.p32(c), 50, 1)
    assert all(o.fp_from_limbs32(c[12 * i:12 * i + 12]) == A[i] * B[i] % o.P for i in range(50))
    X, Y = util.rf12(r), util.rf12(r)
    x, y = util.arr(o.f12_to_limbs32(X)), util.arr(o.f12_to_limbs32(Y))
    lib.ref_fp12_mul(util.p32(x), util.p32(y), util.p32(o144), 1, 1)
    assert o.f12_eq(o.f12_from_limbs32(o144), o.f12_mul(X, Y))


def test_point_encodings_known_answers():
    """The standard compressed encodings of the BLS12-381 generators (IETF / ZCash serialisation, the
    format ark-bls12-381 0.4 implements) pin the oracle's serialiser; round trips pin the parser."""
    g1c = o.g1_serialize(o.G1_GEN, True)
    assert g1c.hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                         "6c55e83ff97a1aeffb3af00adb22c6bb")
    g2c = o.g2_serialize(o.G2_GEN, True)
    assert g2c.hex() == ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
                         "334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051"
                         "c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    assert o.g1_serialize(None, True).hex() == "c0" + "00" * 47
    assert o.g2_serialize(None, False).hex() == "40" + "00" * 191
    r = util.rng(91)
    for _ in range(4):
        p1 = o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER))
        p2 = o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER))
        for c in (True, False):
            assert o.g1_deserialize(o.g1_serialize(p1, c), c) == ("ok", p1)
            assert o.g2_deserialize(o.g2_serialize(p2, c), c) == ("ok", p2)
            neg1 = (p1[0], (o.P - p1[1]) % o.P)
            assert o.g1_deserialize(o.g1_serialize(neg1, c), c) == ("ok", neg1)
    assert o.g1_deserialize(o.g1_serialize(None, True), True) == ("ok", None)
    assert o.g1_deserialize(bytes([0x80]) + bytes(47), True)[0] in ("ok", "err")
    f = o.f12_mul(util.rf12(r), util.rf12(r))
    w = o.f12_to_witness_limbs(f)
    assert len(w) == 144 and w[12:24] == o.fp_to_u32_digits(f[1][0][0]) and w[72:84] == o.fp_to_u32_digits(f[0][0][1])
