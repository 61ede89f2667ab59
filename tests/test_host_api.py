"""Host-side mirror of the reference's native API surface (types, constants, helpers)."""
import pytest

import b381
import b381_oracle as o
import util
from b381.fields import Fq, Fq2, Fq6, Fq12, MyFq12, MyFq6, Bls12_381Base, from_biguint_to_fq, sgn0_fq, sgn0_fq2, pow_fq, get_naf
from b381.curves import G1Affine, G2Affine, G1Projective, G2Projective


def test_constants_mirror_reference():
    gc = b381.global_constants
    assert gc.LOG_ATE_LOOP_COUNT == 62 and gc.ATE_LOOP_COUNT == 15132376222941642752     # src/global_constants.rs:1-2
    assert gc.BLS_X == 0xD201000000010000 and gc.BLS_X_IS_NEGATIVE                        # :7-8
    assert gc.PSEUDO_BINARY_ENCODING == o.PSEUDO_BINARY_ENCODING                          # :3-6
    assert b381.utils.constants.BLS_X == gc.BLS_X and b381.utils.constants.BLS_X_IS_NEGATIVE


def test_wire_layout_matches_oracle():
    r = util.rng(31)
    f = util.rf12(r)
    t = Fq12.from_flat(o.f12_flat(f))
    assert t.limbs() == o.f12_to_limbs32(f)
    assert Fq12.from_limbs(t.limbs()) == t and t.flat() == o.f12_flat(f)
    assert Fq12.one().limbs() == o.f12_to_limbs32(o.F12_ONE)
    assert G1Affine.generator().limbs() == o.g1_to_limbs32(o.G1_GEN)
    assert G2Affine.generator().limbs() == o.g2_to_limbs32(o.G2_GEN)
    v = util.ref_vectors()
    assert G1Affine.generator().limbs() == sum((util.limbs64_to_words(l) for l in v["g1_generator"]["fp"]), [])   # g1_curve.rs:53-74
    assert G2Affine.generator().limbs() == sum((util.limbs64_to_words(l) for l in v["g2_generator"]["fp"]), [])   # g2_curve.rs:63-117
    assert G1Projective.generator().z == Fq(1) and G2Projective.from_affine(G2Affine.identity()).z.is_zero()
    with pytest.raises(ValueError):
        Fq(o.P)
    with pytest.raises(ValueError):
        Fq.from_limbs([(o.P >> (32 * i)) & 0xFFFFFFFF for i in range(12)])


def test_myfq12_myfq6_permutations():
    r = util.rng(32)
    f = util.rf12(r)
    t = Fq12.from_flat(o.f12_flat(f))
    m = MyFq12.from_fq12(t)
    assert [c.v for c in m.coeffs] == o.myfq12_from_fq12(f)                               # helpers.rs:39-41
    assert m.to_fq12() == t
    g = util.rf12(r)
    s = (m + MyFq12.from_fq12(Fq12.from_flat(o.f12_flat(g)))).to_fq12()
    assert s.flat() == o.f12_flat(o.f12_add(f, g))
    f6 = t.c0
    m6 = MyFq6.from_fq6(f6)
    assert [c.v for c in m6.coeffs] == o.myfq6_from_fq6(f[0]) and m6.to_fq6() == f6       # my_fq6.rs:25
    assert (m6 + m6).to_fq6().c1.c0.v == 2 * f[0][1][0] % o.P


def test_helpers():
    assert from_biguint_to_fq(5) == Fq(5)
    with pytest.raises(ValueError):
        from_biguint_to_fq(o.P)
    assert sgn0_fq(Fq(1)) and not sgn0_fq(Fq(0)) and sgn0_fq2(Fq2(Fq(0), Fq(1))) and not sgn0_fq2(Fq2(Fq(2), Fq(1)))
    r = util.rng(33)
    for _ in range(5):
        e = [r.getrandbits(64) for _ in range(2)]
        assert get_naf(e) == o.get_naf(e)
        a = util.rfp(r) or 1
        assert pow_fq(Fq(a), e).v == o.pow_fq(a, e)
    b = Bls12_381Base.from_fq(Fq(o.G1_X))
    assert b.to_int() == o.G1_X and b.to_u32_digits() == [(o.G1_X >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    assert Bls12_381Base.order_u32() == [0xffffaaab, 0xb9feffff, 0xb153ffff, 0x1eabfffe, 0xf6b0f624, 0x6730d2a0,
                                         0xf38512bf, 0x64774b84, 0x434bacd7, 0x4b1ba7b6, 0x397fe69a, 0x1a0111ea]   # bls12_381base.rs:109-112


def test_shard_bounds():
    sb = b381.distributed.shard_bounds
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 5):
        for w in (1, 2, 3, 4, 8):
            spans = [sb(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sb(4, 2, 2)
