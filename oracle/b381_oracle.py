"""b381_oracle.py -- CPU restatement (Python big integers) of the reference's native
BLS12-381 pairing path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (plonky2-bls12-381-pairing_b200/) never does.

Reference = /root/reference (NikolayKostadinov21/plonky2-bls12-381-pairing); all file:line
citations are relative to it.  The reference's native path is glue over arkworks 0.4
(ark-bls12-381 = "0.4.0", ark-ff = "0.4.2", ark-ec = "0.4.2"; Cargo.toml:11-14, caret ranges, no
Cargo.lock), whose source is NOT vendored; the `ARK` functions below restate arkworks' published
algorithms (ark-ec 0.4 models/bls12/{mod,g2}.rs) and are anchored on the reference's call sites.

Pinning status:
  * tower / Montgomery conventions: pinned by the reference's own fixed vectors and constants
    (tests/golden/reference_vectors.json, extracted by tests/golden/extract_reference_vectors.py);
  * Miller-loop / final-exponentiation / pairing VALUES: the reference holds no golden value
      ->  "parity unpinned" at that level; pinned instead by three independent
    constructions agreeing after final exponentiation, bilinearity and the well-known
    e(G1,G2) generator of GT (tests/golden/pairing_vectors.json).

Three parity modes:
  ARK      ark_bls12_381::Bls12_381::{multi_miller_loop, final_exponentiation} semantics
  ZK       the zkcrypto-structured loop spelled out at src/miller_loop_native.rs:27-116 (+ ell)
  LITERAL  the code exactly as written (multi_miller_loop -> 1; optimized_miller_loop in Fq2)
"""

import hashlib

# ----------------------------------------------------------------------------------------------
# constants  (src/miller_loop_native_optimized.rs:104,110; src/global_constants.rs:1-8)
# ----------------------------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000            # |x| ; src/global_constants.rs:7, src/utils/constants.rs:1
BLS_X_IS_NEGATIVE = True               # src/global_constants.rs:8
LOG_ATE_LOOP_COUNT = 62                # src/global_constants.rs:1
ATE_LOOP_COUNT = 15132376222941642752  # src/global_constants.rs:2
PSEUDO_BINARY_ENCODING = [(BLS_X >> i) & 1 for i in range(64)]   # src/global_constants.rs:3-6 (LSB first)

MONT_R = 1 << 384                      # ark Fp384 / zkcrypto Montgomery radix
MONT_R_MOD_P = MONT_R % P
MONT_R2_MOD_P = (MONT_R * MONT_R) % P
MONT_RINV = pow(MONT_R, -1, P)
N0_32 = (-pow(P, -1, 1 << 32)) % (1 << 32)
N0_64 = (-pow(P, -1, 1 << 64)) % (1 << 64)

G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2_X = (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)
G2_Y = (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)

# Fp-multiplication counter (roofline accounting).  Squarings count as
# multiplications; Fp2-by-Fp scalings count 2; additions are free.
FP_MULS = 0


def reset_counter():
    global FP_MULS
    FP_MULS = 0


def fp_muls():
    return FP_MULS


def _m(a, b):
    global FP_MULS
    FP_MULS += 1
    return a * b % P


# ----------------------------------------------------------------------------------------------
# Fp
# ----------------------------------------------------------------------------------------------
def fp_inv(a):
    if a % P == 0:
        raise ZeroDivisionError("Fp inverse of zero")
    return pow(a, -1, P)


def to_mont(a):
    return a * MONT_R % P


def from_mont(a):
    return a * MONT_RINV % P


def fp_to_limbs32(a):
    """canonical integer -> 12 little-endian u32 limbs of its Montgomery form (ark in-memory Fp384)."""
    m = to_mont(a)
    return [(m >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def fp_from_limbs32(l):
    m = 0
    for i, v in enumerate(l):
        m |= int(v) << (32 * i)
    return from_mont(m)


# ----------------------------------------------------------------------------------------------
# Fp2 = Fp[u]/(u^2+1)          src/fields_as_trees/fq2_target_tree.rs:66-142
# ----------------------------------------------------------------------------------------------
F2_ZERO = (0, 0)
F2_ONE = (1, 0)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_dbl(a):
    return ((2 * a[0]) % P, (2 * a[1]) % P)


def f2_mul(a, b):
    """(a0b0 - a1b1, a0b1 + a1b0); fq2_target_tree.rs:97-115.  Counted as 3 Fp muls (Karatsuba)."""
    global FP_MULS
    FP_MULS += 3
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    """((a0+a1)(a0-a1), 2 a0 a1); fq2_target_tree.rs:80-91.  2 Fp muls."""
    global FP_MULS
    FP_MULS += 2
    return ((a[0] + a[1]) * (a[0] - a[1]) % P, (2 * a[0] * a[1]) % P)


def f2_mul_fp(a, s):
    global FP_MULS
    FP_MULS += 2
    return (a[0] * s % P, a[1] * s % P)


def f2_mul_xi(a):
    """multiply by the Fp6 non-residue xi = 1+u; fq2_target_tree.rs:137-142."""
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def f2_conj(a):
    return (a[0], (-a[1]) % P)


def f2_inv(a):
    """(a0, -a1)/(a0^2+a1^2); fq2_target_tree.rs:66-78."""
    global FP_MULS
    FP_MULS += 4
    n = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2_is_zero(a):
    return a[0] % P == 0 and a[1] % P == 0


# ----------------------------------------------------------------------------------------------
# Fp6 = Fp2[v]/(v^3 - xi)      src/fields_as_trees/fq6_target_tree.rs:59-293
# ----------------------------------------------------------------------------------------------
F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b):
    return tuple(f2_add(x, y) for x, y in zip(a, b))


def f6_sub(a, b):
    return tuple(f2_sub(x, y) for x, y in zip(a, b))


def f6_neg(a):
    return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    """Karatsuba, 6 Fp2 muls; fq6_target_tree.rs:172-214."""
    aa = f2_mul(a[0], b[0])
    bb = f2_mul(a[1], b[1])
    cc = f2_mul(a[2], b[2])
    t1 = f2_sub(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), bb), cc)
    c0 = f2_add(f2_mul_xi(t1), aa)
    t2 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), aa), bb)
    c1 = f2_add(t2, f2_mul_xi(cc))
    t3 = f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), aa)
    c2 = f2_sub(f2_add(t3, bb), cc)
    return (c0, c1, c2)


def f6_sqr(a):
    """CH-SQR2; fq6_target_tree.rs:270-293."""
    s0 = f2_sqr(a[0])
    ab = f2_mul(a[0], a[1])
    s1 = f2_dbl(ab)
    s2 = f2_sqr(f2_add(f2_sub(a[0], a[1]), a[2]))
    bc = f2_mul(a[1], a[2])
    s3 = f2_dbl(bc)
    s4 = f2_sqr(a[2])
    c0 = f2_add(f2_mul_xi(s3), s0)
    c1 = f2_add(f2_mul_xi(s4), s1)
    c2 = f2_sub(f2_sub(f2_add(f2_add(s1, s2), s3), s0), s4)
    return (c0, c1, c2)


def f6_mul_by_v(a):
    """(xi*c2, c0, c1); fq6_target_tree.rs:219-230."""
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_mul_by_01(a, c0, c1):
    """fq6_target_tree.rs:232-259 (= ark Fp6::mul_by_01), 5 Fp2 muls."""
    a_a = f2_mul(a[0], c0)
    b_b = f2_mul(a[1], c1)
    t1 = f2_add(f2_mul_xi(f2_sub(f2_mul(c1, f2_add(a[1], a[2])), b_b)), a_a)
    t3 = f2_sub(f2_sub(f2_mul(f2_add(c0, c1), f2_add(a[0], a[1])), a_a), b_b)
    t2 = f2_add(f2_sub(f2_mul(c0, f2_add(a[0], a[2])), a_a), b_b)
    return (t1, t3, t2)


def f6_mul_by_1(a, c1):
    """fq6_target_tree.rs:261-268 (= ark Fp6::mul_by_1), 3 Fp2 muls."""
    return (f2_mul_xi(f2_mul(a[2], c1)), f2_mul(a[0], c1), f2_mul(a[1], c1))


def f6_inv(a):
    """fq6_target_tree.rs:59-89."""
    c0 = f2_sub(f2_sqr(a[0]), f2_mul_xi(f2_mul(a[1], a[2])))
    c1 = f2_sub(f2_mul_xi(f2_sqr(a[2])), f2_mul(a[0], a[1]))
    c2 = f2_sub(f2_sqr(a[1]), f2_mul(a[0], a[2]))
    t = f2_add(f2_mul_xi(f2_add(f2_mul(a[2], c1), f2_mul(a[1], c2))), f2_mul(a[0], c0))
    t = f2_inv(t)
    return (f2_mul(t, c0), f2_mul(t, c1), f2_mul(t, c2))


# ----------------------------------------------------------------------------------------------
# Fp12 = Fp6[w]/(w^2 - v)      src/fields_as_trees/fq12_target_tree.rs:53-176
# ----------------------------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)
F12_ZERO = (F6_ZERO, F6_ZERO)


def f12_mul(a, b):
    """Karatsuba, 3 Fp6 muls; fq12_target_tree.rs:130-141."""
    aa = f6_mul(a[0], b[0])
    bb = f6_mul(a[1], b[1])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), aa), bb)
    c0 = f6_add(f6_mul_by_v(bb), aa)
    return (c0, c1)


def f12_sqr(a):
    """complex squaring; fq12_target_tree.rs:143-155."""
    ab = f6_mul(a[0], a[1])
    c0c1 = f6_add(a[0], a[1])
    c0 = f6_add(f6_mul_by_v(a[1]), a[0])
    c0 = f6_mul(c0, c0c1)
    c0 = f6_sub(c0, ab)
    c1 = f6_add(ab, ab)
    c0 = f6_sub(c0, f6_mul_by_v(ab))
    return (c0, c1)


def f12_conj(a):
    """(c0, -c1); fq12_target_tree.rs:53-58; used at src/miller_loop_native.rs:194-200."""
    return (a[0], f6_neg(a[1]))


def f12_inv(a):
    """(c0, -c1)/(c0^2 - v c1^2); fq12_target_tree.rs:77-90."""
    t = f6_sub(f6_sqr(a[0]), f6_mul_by_v(f6_sqr(a[1])))
    t = f6_inv(t)
    return (f6_mul(a[0], t), f6_neg(f6_mul(a[1], t)))


def f12_mul_by_014(f, c0, c1, c4):
    """sparse multiply by c0 + c1*v + c4*v*w; fq12_target_tree.rs:157-176 and the commented
    native twin src/miller_loop_native.rs:118-137."""
    aa = f6_mul_by_01(f[0], c0, c1)
    bb = f6_mul_by_1(f[1], c4)
    o = f2_add(c1, c4)
    n1 = f6_mul_by_01(f6_add(f[1], f[0]), c0, o)
    n1 = f6_sub(f6_sub(n1, aa), bb)
    n0 = f6_add(f6_mul_by_v(bb), aa)
    return (n0, n1)


def f12_flat(a):
    """tower order c0.c0.c0, c0.c0.c1, c0.c1.c0, ..., c1.c2.c1 (12 ints)."""
    return [a[i][j][k] for i in range(2) for j in range(3) for k in range(2)]


def f12_unflat(l):
    return tuple(tuple((l[i * 6 + j * 2], l[i * 6 + j * 2 + 1]) for j in range(3)) for i in range(2))


def f12_eq(a, b):
    return [x % P for x in f12_flat(a)] == [x % P for x in f12_flat(b)]


def f12_sha256(a):
    """SHA-256 over the 12 canonical values as 48-byte little-endian."""
    h = hashlib.sha256()
    for v in f12_flat(a):
        h.update(int(v % P).to_bytes(48, "little"))
    return h.hexdigest()


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


# Frobenius.  General rule (fq6_target_tree.rs:129-169, fq12_target_tree.rs:92-128): frob^k
# multiplies the coefficient of w^j by xi^(j (p^k - 1)/6) after conjugating Fp2 parts k times.
def _f2_pow(a, e):
    r = F2_ONE
    for bit in bin(e)[2:]:
        r = f2_sqr(r)
        if bit == "1":
            r = f2_mul(r, a)
    return r


XI = (1, 1)
FROB_GAMMA = {}
for _k in (1, 2, 3):
    FROB_GAMMA[_k] = [_f2_pow(XI, j * (P ** _k - 1) // 6) for j in range(6)]
FP_MULS = 0


def f12_frobenius(a, k):
    """a^(p^k) for k in 1..3."""
    g = FROB_GAMMA[k]
    out = [[None] * 3, [None] * 3]
    for i in range(2):
        for j in range(3):
            c = a[i][j]
            if k & 1:
                c = f2_conj(c)
            out[i][j] = f2_mul(c, g[2 * j + i])      # coefficient of w^(2j+i)
    return (tuple(out[0]), tuple(out[1]))


# ----------------------------------------------------------------------------------------------
# native w-basis types  (src/fields/helpers.rs:8-152, src/fields/my_fq6.rs:6-56)
# ----------------------------------------------------------------------------------------------
def myfq12_from_fq12(a):
    """helpers.rs:14-44: [c000,c100,c010,c110,c020,c120, c001,c101,c011,c111,c021,c121]."""
    c = a
    return [c[0][0][0], c[1][0][0], c[0][1][0], c[1][1][0], c[0][2][0], c[1][2][0],
            c[0][0][1], c[1][0][1], c[0][1][1], c[1][1][1], c[0][2][1], c[1][2][1]]


def myfq12_to_fq12(m):
    """helpers.rs:47-76."""
    c0 = ((m[0], m[6]), (m[2], m[8]), (m[4], m[10]))
    c1 = ((m[1], m[7]), (m[3], m[9]), (m[5], m[11]))
    return (c0, c1)


def myfq12_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def myfq12_mul(a, b):
    """schoolbook over w with w^6 = 1+u, 144 Fp muls; helpers.rs:90-152."""
    global FP_MULS
    FP_MULS += 144
    re00 = [0] * 11
    im01 = [0] * 11
    im10 = [0] * 11
    re11 = [0] * 11
    for i in range(6):
        for j in range(6):
            re00[i + j] += a[i] * b[j]
            im01[i + j] += a[i] * b[j + 6]
            im10[i + j] += a[i + 6] * b[j]
            re11[i + j] += a[i + 6] * b[j + 6]
    re = [(re00[i] - re11[i]) % P for i in range(11)]
    im = [(im01[i] + im10[i]) % P for i in range(11)]
    out = []
    for i in range(6):
        out.append((re[i] + re[i + 6] - im[i + 6]) % P if i < 5 else re[i])
    for i in range(6):
        out.append((im[i] + re[i + 6] + im[i + 6]) % P if i < 5 else im[i])
    return out


def myfq6_from_fq6(a):
    """my_fq6.rs:12-28: [c00,c10,c20,c01,c11,c21]."""
    return [a[0][0], a[1][0], a[2][0], a[0][1], a[1][1], a[2][1]]


def myfq6_to_fq6(m):
    """my_fq6.rs:31-44."""
    return ((m[0], m[3]), (m[1], m[4]), (m[2], m[5]))


# helpers.rs:154-239
def from_biguint_to_fq(x):
    if not (0 <= x < P):
        raise ValueError("not a canonical Fq")     # Fq::from_bigint(..).unwrap() panics
    return x


def sgn0_fq(x):
    return (x % P) & 1 == 1


def sgn0_fq2(x):
    return sgn0_fq(x[0]) or (x[0] % P == 0 and sgn0_fq(x[1]))


def fp_legendre_is_square(a):
    """fq_target.rs:269-280: legendre(a) = a^((p-1)/2), is_square = (legendre == 1); zero is not a square."""
    return pow(a % P, (P - 1) // 2, P) == 1


def fp_sqrt_sgn(a, sgn):
    """FqSqrtGenerator::run_once (fq_target.rs:316-343): sqrt(a), negated when its sgn0 differs from sgn.
    None where the reference panics (non-residue; zero with sgn = True)."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    if s * s % P != a:
        return None
    if sgn0_fq(s) != bool(sgn):
        if s == 0:
            return None
        s = P - s
    return s


def f2_is_square(a):
    """a^((p^2-1)/2) == 1 <=> the norm a0^2 + a1^2 is a non-zero square in Fq."""
    return fp_legendre_is_square((a[0] * a[0] + a[1] * a[1]) % P)


def f2_sqrt_sgn(a, sgn):
    """Fq2 square root with sgn0_fq2(result) == sgn (fq2_target.rs:373-410); both roots +-s of a non-zero
    element have opposite sgn0, so the result does not depend on the root-finding algorithm."""
    a = (a[0] % P, a[1] % P)
    if a == (0, 0):
        return None if sgn else (0, 0)
    if a[1] == 0:
        s = pow(a[0], (P + 1) // 4, P)
        if s * s % P == a[0]:
            r = (s, 0)
        else:
            s = pow(P - a[0], (P + 1) // 4, P)
            r = (0, s)
    else:
        n = (a[0] * a[0] + a[1] * a[1]) % P
        s = pow(n, (P + 1) // 4, P)
        if s * s % P != n:
            return None
        half = (P + 1) // 2
        d = (a[0] + s) * half % P
        t = pow(d, (P + 1) // 4, P)
        if t * t % P != d:
            d = (a[0] - s) * half % P
            t = pow(d, (P + 1) // 4, P)
        r = (t, a[1] * pow(2 * t, P - 2, P) % P)
    if f2_sqr(r) != a:
        return None
    if sgn0_fq2(r) != bool(sgn):
        r = ((P - r[0]) % P, (P - r[1]) % P)
    return r


# ---- wire formats ----------------------------------------------------------------
def fp_to_u32_digits(a):
    """`BigUint::from(fq).to_u32_digits()` padded to 12 (src/fields/fq_target.rs:300-313): canonical integer."""
    return [(a % P >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def f12_to_witness_limbs(f):
    """Fq12Target::set_witness (src/fields/fq12_target.rs:408-416): MyFq12 order (helpers.rs:39-41), 12 digits each."""
    out = []
    for k in range(2):
        for j in range(3):
            for i in range(2):
                out += fp_to_u32_digits(f[i][j][k])
    return out


def _lex_largest_fp(y):
    return y % P > (P - 1) // 2


def _lex_largest_f2(y):
    return _lex_largest_fp(y[1]) or (y[1] % P == 0 and _lex_largest_fp(y[0]))


def g1_serialize(pt, compressed=True):
    """ZCash / IETF encoding as implemented by ark-bls12-381 0.4: big-endian, flags in the top bits of byte 0."""
    n = 48 if compressed else 96
    if pt is None:
        return bytes([(0x80 if compressed else 0) | 0x40]) + bytes(n - 1)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    if compressed:
        b[0] |= 0x80 | (0x20 if _lex_largest_fp(y) else 0)
        return bytes(b)
    return bytes(b) + y.to_bytes(48, "big")


def g1_deserialize(data, compressed=True):
    """-> ("ok", point-or-None) or ("err", reason); no subgroup check."""
    f = data[0]
    fc, fi, fs = bool(f & 0x80), bool(f & 0x40), bool(f & 0x20)
    if fc != compressed:
        return ("err", "encoding")
    x = int.from_bytes(bytes([f & 0x1F]) + data[1:48], "big")
    if fi:
        if fs or x != 0 or any(data[48:]):
            return ("err", "encoding")
        return ("ok", None)
    if x >= P:
        return ("err", "canonical")
    rhs = (x * x * x + 4) % P
    if compressed:
        y = pow(rhs, (P + 1) // 4, P)
        if y * y % P != rhs:
            return ("err", "curve")
        if _lex_largest_fp(y) != fs:
            y = (P - y) % P
        return ("ok", (x, y))
    if fs:
        return ("err", "encoding")
    y = int.from_bytes(data[48:96], "big")
    if y >= P:
        return ("err", "canonical")
    if y * y % P != rhs:
        return ("err", "curve")
    return ("ok", (x, y))


def g2_serialize(pt, compressed=True):
    n = 96 if compressed else 192
    if pt is None:
        return bytes([(0x80 if compressed else 0) | 0x40]) + bytes(n - 1)
    x, y = pt
    b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
    if compressed:
        b[0] |= 0x80 | (0x20 if _lex_largest_f2(y) else 0)
        return bytes(b)
    return bytes(b) + y[1].to_bytes(48, "big") + y[0].to_bytes(48, "big")


def g2_deserialize(data, compressed=True):
    f = data[0]
    fc, fi, fs = bool(f & 0x80), bool(f & 0x40), bool(f & 0x20)
    if fc != compressed:
        return ("err", "encoding")
    x1 = int.from_bytes(bytes([f & 0x1F]) + data[1:48], "big")
    x0 = int.from_bytes(data[48:96], "big")
    if fi:
        if fs or x1 != 0 or any(data[48:]):
            return ("err", "encoding")
        return ("ok", None)
    if x0 >= P or x1 >= P:
        return ("err", "canonical")
    x = (x0, x1)
    rhs = f2_add(f2_mul(f2_sqr(x), x), (4, 4))
    if compressed:
        y = f2_sqrt_sgn(rhs, False)
        if y is None:
            return ("err", "curve")
        if _lex_largest_f2(y) != fs:
            y = f2_neg(y)
        return ("ok", (x, y))
    if fs:
        return ("err", "encoding")
    y = (int.from_bytes(data[144:192], "big"), int.from_bytes(data[96:144], "big"))
    if y[0] >= P or y[1] >= P:
        return ("err", "canonical")
    if f2_sqr(y) != rhs:
        return ("err", "curve")
    return ("ok", (x, y))


def get_naf(exp_limbs):
    """helpers.rs:197-239 (u64 limbs, little-endian) -> NAF digits, LSB first."""
    exp = list(exp_limbs)
    naf = []
    n = len(exp)
    for idx in range(n):
        e = exp[idx]
        for _ in range(64):
            if e & 1:
                z = 2 - (e % 4)
                e //= 2
                if z == -1:
                    e += 1
                naf.append(z)
            else:
                naf.append(0)
                e //= 2
        if e != 0:
            assert e == 1
            j = idx + 1
            while j < len(exp) and exp[j] == (1 << 64) - 1:
                exp[j] = 0
                j += 1
            if j < len(exp):
                exp[j] += 1
            else:
                exp.append(1)
    if len(exp) != n:
        naf.append(1)
    return naf


def pow_fq(a, exp_limbs):
    """helpers.rs:176-195 (NAF ladder)."""
    res = a
    started = False
    for z in reversed(get_naf(exp_limbs)):
        if started:
            res = res * res % P
        if z != 0:
            if started:
                res = res * a % P if z == 1 else res * fp_inv(a) % P
            else:
                assert z == 1
                started = True
    return res


# ----------------------------------------------------------------------------------------------
# curve helpers (test-point generation; affine, plain integers).  None = point at infinity.
# ----------------------------------------------------------------------------------------------
def g1_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * fp_inv(2 * y1) % P
    else:
        lam = (y2 - y1) * fp_inv((x2 - x1) % P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = g1_add(r, p)
        p = g1_add(p, p)
        k >>= 1
    return r


def g2_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if f2_sub(x1, x2) == F2_ZERO:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_mul_fp(f2_sqr(x1), 3), f2_inv(f2_dbl(y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = g2_add(r, p)
        p = g2_add(p, p)
        k >>= 1
    return r


def g1_on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - 4) % P == 0


def g2_on_curve(q):
    if q is None:
        return True
    lhs = f2_sqr(q[1])
    rhs = f2_add(f2_mul(f2_sqr(q[0]), q[0]), (4, 4))
    return f2_sub(lhs, rhs) == F2_ZERO


G1_GEN = (G1_X, G1_Y)
G2_GEN = (G2_X, G2_Y)


# ----------------------------------------------------------------------------------------------
# Endomorphisms, fast subgroup membership and cofactor clearing: ark-bls12-381 0.4 `curves/g1.rs`, `curves/g2.rs`
# (third-party, restated from the published algorithms; pinned below by the RFC 9380 effective cofactors and by
# agreement with the plain [r] P test).  The reference's circuit side needs them before untrusted points can be
# paired (/root/reference/src/fields/fq_target.rs:288-313, src/fields/fq12_target.rs:408-416).
# ----------------------------------------------------------------------------------------------
# BETA: the non-trivial cube root of unity ark-bls12-381 uses for (x, y) -> (BETA x, y)   (g1.rs `BETA`)
G1_BETA = 793479390729215512621379701633421447060886740281060493010456487427281649075476305620758731620350
assert pow(G1_BETA, 3, P) == 1 and G1_BETA != 1
# effective cofactors of RFC 9380 section 8.8 (BLS12381G1: 1 - x; BLS12381G2: Budroni-Pintore h_eff)
G1_H_EFF = 0xD201000000010001
G2_H_EFF = 0xBC69F08F2EE75B3584C6A0EA91B352888E2A8E9145AD7689986FF031508FFE1329C2F178731DB956D82BF015D1212B02EC0EC69D7477C1AE954CBC06689F6A359894C0ADEBBF6B4E8020005AAA95551
assert G1_H_EFF == 1 + BLS_X


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % P)


def g2_neg(q):
    return None if q is None else (q[0], f2_neg(q[1]))


def g1_endomorphism(p):
    """ark g1.rs endomorphism(): (x, y) -> (BETA x, y)"""
    return None if p is None else (G1_BETA * p[0] % P, p[1])


def g1_in_subgroup_fast(p):
    """ark g1.rs is_in_correct_subgroup_assuming_on_curve (eprint 2021/1130 section 6):
    endomorphism(P) == -[X^2] P with X = |x|, plus the early-out [X] P == P for P != identity."""
    if p is None:
        return True
    xp = g1_mul(p, BLS_X)
    if xp == p:
        return False
    return g1_neg(g1_mul(xp, BLS_X)) == g1_endomorphism(p)


def g1_clear_cofactor(p):
    """ark g1.rs clear_cofactor: multiplication by the effective cofactor 1 - x = 1 + |x| (eprint 2019/403 section 5)"""
    return g1_add(g1_mul(p, BLS_X), p)


# psi = twist o Frobenius o untwist on E'(Fq2): (x, y) -> (conj(x) / xi^((p-1)/3), conj(y) / xi^((p-1)/2))   (g2.rs
# P_POWER_ENDOMORPHISM_COEFF_0 / _1);  psi^2: (x, y) -> (x / xi^((p^2-1)/3), -y)   (DOUBLE_P_POWER_ENDOMORPHISM_COEFF_0)
def _f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


PSI_CX = f2_inv(_f2_pow((1, 1), (P - 1) // 3))
PSI_CY = f2_inv(_f2_pow((1, 1), (P - 1) // 2))
PSI2_CX = f2_inv(_f2_pow((1, 1), (P * P - 1) // 3))
assert PSI2_CX[1] == 0 and _f2_pow((1, 1), (P * P - 1) // 2) == ((-1) % P, 0)


def g2_psi(q):
    if q is None:
        return None
    return (f2_mul(f2_conj(q[0]), PSI_CX), f2_mul(f2_conj(q[1]), PSI_CY))


def g2_psi2(q):
    if q is None:
        return None
    return (f2_mul_fp(q[0], PSI2_CX[0]), f2_neg(q[1]))


def g2_in_subgroup_fast(q):
    """ark g2.rs is_in_correct_subgroup_assuming_on_curve (eprint 2021/1130 section 4): psi(P) == [x] P, x < 0"""
    if q is None:
        return True
    return g2_neg(g2_mul(q, BLS_X)) == g2_psi(q)


def g2_clear_cofactor(q):
    """ark g2.rs clear_cofactor (Budroni-Pintore, eprint 2017/419 section 4.1):
    [x^2 - x - 1] P + [x - 1] psi(P) + psi^2(2 P), evaluated as ark does with X = |x| and negations."""
    if q is None:
        return None
    x_p = g2_neg(g2_mul(q, BLS_X))                       # [x] P
    psi_p = g2_psi(q)
    psi2_p2 = g2_psi2(g2_add(q, q))
    tmp2 = g2_neg(g2_mul(g2_add(x_p, psi_p), BLS_X))     # [x^2] P + [x] psi(P)
    acc = g2_add(psi2_p2, tmp2)
    acc = g2_add(acc, g2_neg(x_p))
    acc = g2_add(acc, g2_neg(psi_p))
    return g2_add(acc, g2_neg(q))

# ----------------------------------------------------------------------------------------------
# ARK mode: ark-ec 0.4 models/bls12/g2.rs (G2Prepared) + models/bls12/mod.rs; SURVEY A.2-A.5.
# Truth the reference defers to: src/miller_loop_native_optimized.rs:131-132,151,163.
# ----------------------------------------------------------------------------------------------
TWO_INV = fp_inv(2)
X_BITS_AFTER_LEADING = [int(c) for c in bin(BLS_X)[3:]]        # 63 bits, MSB-first, leading 1 skipped


def ark_double_step(r):
    """G2HomProjective::double_in_place; returns (new_r, (c0, c1, c2)) for the M-twist."""
    X, Y, Z = r
    a = f2_mul_fp(f2_mul(X, Y), TWO_INV)
    b = f2_sqr(Y)
    c = f2_sqr(Z)
    c3 = f2_add(f2_dbl(c), c)
    e = f2_mul_xi(f2_dbl(f2_dbl(c3)))                 # COEFF_B(4,4) * 3c = 4*xi*3c  (additions only)
    f = f2_add(f2_dbl(e), e)
    g = f2_mul_fp(f2_add(b, f), TWO_INV)
    h = f2_sub(f2_sqr(f2_add(Y, Z)), f2_add(b, c))
    i = f2_sub(e, b)
    j = f2_sqr(X)
    e2 = f2_sqr(e)
    nx = f2_mul(a, f2_sub(b, f))
    ny = f2_sub(f2_sqr(g), f2_add(f2_dbl(e2), e2))
    nz = f2_mul(b, h)
    return (nx, ny, nz), (i, f2_add(f2_dbl(j), j), f2_neg(h))


def ark_add_step(r, q):
    """G2HomProjective::add_in_place (mixed addition with affine q)."""
    X, Y, Z = r
    qx, qy = q
    theta = f2_sub(Y, f2_mul(qy, Z))
    lam = f2_sub(X, f2_mul(qx, Z))
    c = f2_sqr(theta)
    d = f2_sqr(lam)
    e = f2_mul(lam, d)
    f = f2_mul(Z, c)
    g = f2_mul(X, d)
    h = f2_sub(f2_add(e, f), f2_dbl(g))
    nx = f2_mul(lam, h)
    ny = f2_sub(f2_mul(theta, f2_sub(g, h)), f2_mul(e, Y))
    nz = f2_mul(Z, e)
    j = f2_sub(f2_mul(theta, qx), f2_mul(lam, qy))
    return (nx, ny, nz), (j, f2_neg(theta), lam)


def ark_g2_prepare(q):
    """G2Prepared::from(q): 68 coefficient triples."""
    if q is None:
        return None
    r = (q[0], q[1], F2_ONE)
    coeffs = []
    for bit in X_BITS_AFTER_LEADING:
        r, c = ark_double_step(r)
        coeffs.append(c)
        if bit:
            r, c = ark_add_step(r, q)
            coeffs.append(c)
    return coeffs


def ark_ell(f, coeffs, p):
    """Bls12::ell for TwistType::M: c2 *= py; c1 *= px; f.mul_by_014(c0, c1, c2)."""
    c0, c1, c2 = coeffs
    return f12_mul_by_014(f, c0, f2_mul_fp(c1, p[0]), f2_mul_fp(c2, p[1]))


def ark_multi_miller_loop(pairs):
    """Bls12::multi_miller_loop; pairs = [(P, Q)], None = identity (dropped => contributes 1)."""
    live = [(p, ark_g2_prepare(q)) for p, q in pairs if p is not None and q is not None]
    idx = [0] * len(live)
    f = F12_ONE
    for bit in X_BITS_AFTER_LEADING:
        f = f12_sqr(f)
        for n, (p, co) in enumerate(live):
            f = ark_ell(f, co[idx[n]], p)
            idx[n] += 1
        if bit:
            for n, (p, co) in enumerate(live):
                f = ark_ell(f, co[idx[n]], p)
                idx[n] += 1
    if BLS_X_IS_NEGATIVE:
        f = f12_conj(f)
    return f


def ark_miller_loop(p, q):
    return ark_multi_miller_loop([(p, q)])


def f2_fp4_square(a, b):
    """src/fields_as_trees/miller_loop.rs:29-44: (a^2 + xi b^2, (a+b)^2 - a^2 - b^2)."""
    t0 = f2_sqr(a)
    t1 = f2_sqr(b)
    c0 = f2_add(f2_mul_xi(t1), t0)
    c1 = f2_sub(f2_sub(f2_sqr(f2_add(a, b)), t0), t1)
    return c0, c1


def f12_cyclotomic_square(f):
    """Granger-Scott; src/fields_as_trees/miller_loop.rs:46-104 (valid after the easy part)."""
    z0, z4, z3 = f[0]
    z2, z1, z5 = f[1]
    t0, t1 = f2_fp4_square(z0, z1)
    z0 = f2_add(f2_dbl(f2_sub(t0, z0)), t0)
    z1 = f2_add(f2_dbl(f2_add(t1, z1)), t1)
    t0, t1 = f2_fp4_square(z2, z3)
    t2, t3 = f2_fp4_square(z4, z5)
    z4 = f2_add(f2_dbl(f2_sub(t0, z4)), t0)
    z5 = f2_add(f2_dbl(f2_add(t1, z5)), t1)
    t0 = f2_mul_xi(t3)
    z2 = f2_add(f2_dbl(f2_add(t0, z2)), t0)
    z3 = f2_add(f2_dbl(f2_sub(t2, z3)), t2)
    return ((z0, z4, z3), (z2, z1, z5))


def f12_cyclotomic_exp_abs_x(a):
    """a^|x| by square-and-multiply with cyclotomic squarings (ark Fp12::cyclotomic_exp)."""
    r = a
    for bit in X_BITS_AFTER_LEADING:
        r = f12_cyclotomic_square(r)
        if bit:
            r = f12_mul(r, a)
    return r


def ark_exp_by_x(a):
    """Bls12::exp_by_x: a^x with x negative -> conj(a^|x|)."""
    r = f12_cyclotomic_exp_abs_x(a)
    return f12_conj(r) if BLS_X_IS_NEGATIVE else r


def ark_final_exponentiation(f):
    """Bls12::final_exponentiation (eprint 2020/875 chain); SURVEY A.5.  None for f = 0."""
    if f12_eq(f, F12_ZERO):
        return None
    f1 = f12_conj(f)
    f2 = f12_inv(f)
    r = f12_mul(f1, f2)
    f2 = r
    r = f12_frobenius(r, 2)
    r = f12_mul(r, f2)
    y0 = f12_cyclotomic_square(r)
    y1 = ark_exp_by_x(r)
    y2 = f12_conj(r)
    y1 = f12_mul(y1, y2)
    y2 = ark_exp_by_x(y1)
    y1 = f12_conj(y1)
    y1 = f12_mul(y1, y2)
    y2 = ark_exp_by_x(y1)
    y1 = f12_frobenius(y1, 1)
    y1 = f12_mul(y1, y2)
    r = f12_mul(r, y0)
    y0 = ark_exp_by_x(y1)
    y2 = ark_exp_by_x(y0)
    y0 = f12_frobenius(y1, 2)
    y1 = f12_conj(y1)
    y1 = f12_mul(y1, y2)
    y1 = f12_mul(y1, y0)
    r = f12_mul(r, y1)
    return r


FINAL_EXP_POWER = 3 * (P ** 12 - 1) // R_ORDER        # exponent of both chains


def final_exponentiation_plain(f):
    """f^(3 (p^12-1)/r) by plain square-and-multiply (independent cross-check)."""
    return f12_pow(f, FINAL_EXP_POWER)


def ark_pairing(p, q):
    return ark_final_exponentiation(ark_miller_loop(p, q))


def ark_multi_pairing(pairs):
    return ark_final_exponentiation(ark_multi_miller_loop(pairs))


# ----------------------------------------------------------------------------------------------
# ZK mode: the loop the reference spells out; src/miller_loop_native.rs:27-116 and the
# commented ell/mul_by_014 at :118-152 (uncommented here); final exponentiation per the circuit
# twin src/fields_as_trees/miller_loop.rs:106-178 with its bugs fixed.
# ----------------------------------------------------------------------------------------------
def zk_double_step(r):
    """_point_doubling_and_line_evaluation, src/miller_loop_native.rs:27-55 (Alg. 26)."""
    x, y, z = r
    tmp0 = f2_sqr(x)
    tmp1 = f2_sqr(y)
    tmp2 = f2_sqr(tmp1)
    tmp3 = f2_sub(f2_sub(f2_sqr(f2_add(tmp1, x)), tmp0), tmp2)
    tmp3 = f2_dbl(tmp3)
    tmp4 = f2_add(f2_dbl(tmp0), tmp0)
    tmp6 = f2_add(x, tmp4)
    tmp5 = f2_sqr(tmp4)
    zsq = f2_sqr(z)
    nx = f2_sub(f2_sub(tmp5, tmp3), tmp3)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(z, y)), tmp1), zsq)
    ny = f2_mul(f2_sub(tmp3, nx), tmp4)
    tmp2 = f2_dbl(f2_dbl(f2_dbl(tmp2)))
    ny = f2_sub(ny, tmp2)
    tmp3 = f2_neg(f2_dbl(f2_mul(tmp4, zsq)))
    tmp6 = f2_sub(f2_sub(f2_sqr(tmp6), tmp0), tmp5)
    tmp1 = f2_dbl(f2_dbl(tmp1))
    tmp6 = f2_sub(tmp6, tmp1)
    tmp0 = f2_dbl(f2_mul(nz, zsq))
    return (nx, ny, nz), (tmp0, tmp3, tmp6)


def zk_add_step(r, q):
    """_point_addition_and_line_evaluation, src/miller_loop_native.rs:58-87 (Alg. 27)."""
    x, y, z = r
    qx, qy = q
    zsq = f2_sqr(z)
    ysq = f2_sqr(qy)
    t0 = f2_mul(zsq, qx)
    t1 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(qy, z)), ysq), zsq), zsq)
    t2 = f2_sub(t0, x)
    t3 = f2_sqr(t2)
    t4 = f2_dbl(f2_dbl(t3))
    t5 = f2_mul(t4, t2)
    t6 = f2_sub(f2_sub(t1, y), y)
    t9 = f2_mul(t6, qx)
    t7 = f2_mul(t4, x)
    nx = f2_sub(f2_sub(f2_sub(f2_sqr(t6), t5), t7), t7)
    nz = f2_sub(f2_sub(f2_sqr(f2_add(z, t2)), zsq), t3)
    t10 = f2_add(qy, nz)
    t8 = f2_mul(f2_sub(t7, nx), t6)
    t0 = f2_dbl(f2_mul(y, t5))
    ny = f2_sub(t8, t0)
    t10 = f2_sub(f2_sqr(t10), ysq)
    ztsq = f2_sqr(nz)
    t10 = f2_sub(t10, ztsq)
    t9 = f2_sub(f2_dbl(t9), t10)
    t10 = f2_dbl(nz)
    t6 = f2_neg(t6)
    t1 = f2_dbl(t6)
    return (nx, ny, nz), (t10, t1, t9)


def zk_ell(f, coeffs, p):
    """src/miller_loop_native.rs:139-152: c0 *= py; c1 *= px; f.mul_by_014(coeffs.2, c1, c0)."""
    c0 = f2_mul_fp(coeffs[0], p[1])
    c1 = f2_mul_fp(coeffs[1], p[0])
    return f12_mul_by_014(f, coeffs[2], c1, c0)


def zk_multi_miller_loop(pairs):
    """driver src/miller_loop_native.rs:89-116 with the per-term bodies at :163-188 wired in
    (identity pairs skipped, as the comments at :167-170 intend)."""
    live = [[p, q, (q[0], q[1], F2_ONE)] for p, q in pairs if p is not None and q is not None]
    f = F12_ONE

    def dbl(f):
        for t in live:
            t[2], c = zk_double_step(t[2])
            f = zk_ell(f, c, t[0])
        return f

    def add(f):
        for t in live:
            t[2], c = zk_add_step(t[2], t[1])
            f = zk_ell(f, c, t[0])
        return f

    found_one = False
    for b in range(63, -1, -1):
        i = (((BLS_X >> 1) >> b) & 1) == 1
        if not found_one:
            found_one = i
            continue
        f = dbl(f)
        if i:
            f = add(f)
        f = f12_sqr(f)
    f = dbl(f)
    if BLS_X_IS_NEGATIVE:
        f = f12_conj(f)
    return f


def zk_g2_prepare(q):
    """The 68 coefficient triples of the ZK loop in the order the driver consumes them
    (src/miller_loop_native.rs:89-116): the G2Prepared shape of src/miller_loop_target.rs:23-76."""
    if q is None:
        return None
    r = (q[0], q[1], F2_ONE)
    coeffs = []
    found_one = False
    for b in range(63, -1, -1):
        i = (((BLS_X >> 1) >> b) & 1) == 1
        if not found_one:
            found_one = i
            continue
        r, c = zk_double_step(r)
        coeffs.append(c)
        if i:
            r, c = zk_add_step(r, q)
            coeffs.append(c)
    r, c = zk_double_step(r)
    coeffs.append(c)
    return coeffs


def zk_miller_loop(p, q):
    return zk_multi_miller_loop([(p, q)])


def zk_final_exponentiation(f):
    """zkcrypto chain; src/fields_as_trees/miller_loop.rs:128-178 (dropped mul at :121 restored)."""
    t0 = f12_conj(f)                                  # frob^6 = conjugation
    t1 = f12_inv(f)
    t2 = f12_mul(t0, t1)
    t1 = t2
    t2 = f12_mul(f12_frobenius(t2, 2), t1)
    t1 = f12_conj(f12_cyclotomic_square(t2))
    t3 = ark_exp_by_x(t2)
    t4 = f12_cyclotomic_square(t3)
    t5 = f12_mul(t1, t3)
    t1 = ark_exp_by_x(t5)
    t0 = ark_exp_by_x(t1)
    t6 = f12_mul(ark_exp_by_x(t0), t4)
    t4 = ark_exp_by_x(t6)
    t5 = f12_conj(t5)
    t4 = f12_mul(f12_mul(t4, t5), t2)
    t5 = f12_conj(t2)
    t1 = f12_frobenius(f12_mul(t1, t2), 3)
    t6 = f12_frobenius(f12_mul(t6, t5), 1)
    t3 = f12_frobenius(f12_mul(t3, t0), 2)
    return f12_mul(f12_mul(f12_mul(t3, t1), t6), t4)


# ----------------------------------------------------------------------------------------------
# LITERAL mode: byte-for-byte behaviour of the code as written.
# ----------------------------------------------------------------------------------------------
def literal_multi_miller_loop(pairs):
    """src/miller_loop_native.rs:154-212 as written: the per-term loops are empty, so the driver
    only squares 1 and conjugates -> always Fq12::one()."""
    f = F12_ONE
    found_one = False
    for b in range(63, -1, -1):
        i = (((BLS_X >> 1) >> b) & 1) == 1
        if not found_one:
            found_one = i
            continue
        f = f12_sqr(f)
    return f12_conj(f)


def jac_double(pt):
    """ark-ec 0.4 short_weierstrass Projective::double_in_place (a = 0, dbl-2009-l)."""
    X, Y, Z = pt
    if f2_is_zero(Z):
        return pt
    A = f2_sqr(X)
    B = f2_sqr(Y)
    C = f2_sqr(B)
    D = f2_dbl(f2_sub(f2_sub(f2_sqr(f2_add(X, B)), A), C))
    E = f2_add(f2_dbl(A), A)
    F = f2_sqr(E)
    Z3 = f2_dbl(f2_mul(Y, Z))
    X3 = f2_sub(F, f2_dbl(D))
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), f2_dbl(f2_dbl(f2_dbl(C))))
    return (X3, Y3, Z3)


def jac_add(p1, p2):
    """ark-ec 0.4 Projective += Projective (add-2007-bl); used at
    src/miller_loop_native_optimized.rs:93,98."""
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if f2_is_zero(Z1):
        return p2
    if f2_is_zero(Z2):
        return p1
    z1z1 = f2_sqr(Z1)
    z2z2 = f2_sqr(Z2)
    u1 = f2_mul(X1, z2z2)
    u2 = f2_mul(X2, z1z1)
    s1 = f2_mul(f2_mul(Y1, Z2), z2z2)
    s2 = f2_mul(f2_mul(Y2, Z1), z1z1)
    if u1 == u2 and s1 == s2:
        return jac_double(p1)
    h = f2_sub(u2, u1)
    i = f2_sqr(f2_dbl(h))
    j = f2_mul(h, i)
    rr = f2_dbl(f2_sub(s2, s1))
    v = f2_mul(u1, i)
    X3 = f2_sub(f2_sub(f2_sqr(rr), j), f2_dbl(v))
    Y3 = f2_sub(f2_mul(rr, f2_sub(v, X3)), f2_dbl(f2_mul(s1, j)))
    Z3 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(Z1, Z2)), z1z1), z2z2), h)
    return (X3, Y3, Z3)


def literal_line_function(q1, q2, pp):
    """optimized_line_function, src/miller_loop_native_optimized.rs:8-78.  Returns Fq2 (num, den);
    the reference embeds them in Fq12.c0.c0."""
    x1, y1, z1 = q1
    x2, y2, z2 = q2
    xp, yp, zp = (pp[0], 0), (pp[1], 0), (pp[2], 0)
    num = f2_sub(f2_mul(y2, z1), f2_mul(y1, z2))
    den = f2_sub(f2_mul(x2, z1), f2_mul(x1, z2))
    A = f2_sub(f2_mul(xp, z1), f2_mul(x1, zp))
    B = f2_sub(f2_mul(yp, z1), f2_mul(y1, zp))
    if not f2_is_zero(den):
        return f2_sub(f2_mul(num, A), f2_mul(den, B)), f2_mul(f2_mul(den, zp), z1)
    if f2_is_zero(num):
        num = f2_mul(f2_mul((3, 0), x1), x1)
        den = f2_mul(f2_mul((2, 0), y1), z1)
        return f2_sub(f2_mul(num, A), f2_mul(den, B)), f2_mul(f2_mul(den, zp), z1)
    return A, f2_mul(z1, zp)


def literal_optimized_miller_loop(pp, qq):
    """optimized_miller_loop, src/miller_loop_native_optimized.rs:81-127, exactly as written:
    forward over all 64 PSEUDO_BINARY_ENCODING entries, everything in Fq2 (embedded in c0.c0),
    Jacobian coords used as if homogeneous, 'final exponentiation' = one squaring.
    pp = (x,y,z) ints, qq = (x,y,z) Fq2.  Raises ZeroDivisionError where the reference panics."""
    R = qq
    fnum = F2_ONE
    fden = F2_ONE
    for v in PSEUDO_BINARY_ENCODING:
        n, d = literal_line_function(R, R, pp)
        fnum = f2_mul(f2_mul(fnum, fnum), n)
        fden = f2_mul(f2_mul(fden, fden), d)
        R = jac_add(R, R)
        if v == 1:
            n, d = literal_line_function(R, qq, pp)
            fnum = f2_mul(fnum, n)
            fden = f2_mul(fden, d)
            R = jac_add(R, qq)
    if f2_is_zero(fden):
        raise ZeroDivisionError("f_den == 0 (reference panics in Fq12 Div)")
    f = f2_mul(fnum, f2_inv(fden))
    f = f2_mul(f, f)
    return ((f, F2_ZERO, F2_ZERO), F6_ZERO)


# ----------------------------------------------------------------------------------------------
# textbook cross-check: affine Miller loop on the untwisted curve over Fp12 (independent of the
# projective formulas above).  untwist (x',y') -> (x'/w^2, y'/w^3).
# ----------------------------------------------------------------------------------------------
def _f12_from_f2(c, wpow):
    l = [F2_ZERO] * 6
    l[wpow] = c
    return ((l[0], l[2], l[4]), (l[1], l[3], l[5]))


def f12_add(a, b):
    return (f6_add(a[0], b[0]), f6_add(a[1], b[1]))


def f12_sub(a, b):
    return (f6_sub(a[0], b[0]), f6_sub(a[1], b[1]))


def textbook_miller_loop(p, q):
    if p is None or q is None:
        return F12_ONE
    w = _f12_from_f2(F2_ONE, 1)
    w2i = f12_inv(f12_mul(w, w))
    w3i = f12_inv(f12_mul(f12_mul(w, w), w))
    Qx = f12_mul(_f12_from_f2(q[0], 0), w2i)
    Qy = f12_mul(_f12_from_f2(q[1], 0), w3i)
    Px = _f12_from_f2((p[0], 0), 0)
    Py = _f12_from_f2((p[1], 0), 0)
    three = _f12_from_f2((3, 0), 0)
    two = _f12_from_f2((2, 0), 0)
    Rx, Ry = Qx, Qy
    f = F12_ONE
    for bit in X_BITS_AFTER_LEADING:
        lam = f12_mul(f12_mul(three, f12_mul(Rx, Rx)), f12_inv(f12_mul(two, Ry)))
        line = f12_sub(f12_sub(Py, Ry), f12_mul(lam, f12_sub(Px, Rx)))
        f = f12_mul(f12_sqr(f), line)
        nx = f12_sub(f12_sub(f12_mul(lam, lam), Rx), Rx)
        Ry = f12_sub(f12_mul(lam, f12_sub(Rx, nx)), Ry)
        Rx = nx
        if bit:
            lam = f12_mul(f12_sub(Qy, Ry), f12_inv(f12_sub(Qx, Rx)))
            line = f12_sub(f12_sub(Py, Ry), f12_mul(lam, f12_sub(Px, Rx)))
            f = f12_mul(f, line)
            nx = f12_sub(f12_sub(f12_mul(lam, lam), Rx), Qx)
            Ry = f12_sub(f12_mul(lam, f12_sub(Rx, nx)), Ry)
            Rx = nx
    return f12_conj(f)


# ----------------------------------------------------------------------------------------------
# flat-buffer marshalling used at the C ABI (include/b381.h): Montgomery, 12 x u32 LE per Fp.
# ----------------------------------------------------------------------------------------------
def f12_to_limbs32(a):
    out = []
    for v in f12_flat(a):
        out.extend(fp_to_limbs32(v))
    return out


def f12_from_limbs32(l):
    return f12_unflat([fp_from_limbs32(l[12 * i:12 * i + 12]) for i in range(12)])


def g1_to_limbs32(p):
    if p is None:
        return [0] * 24
    return fp_to_limbs32(p[0]) + fp_to_limbs32(p[1])


def g2_to_limbs32(q):
    if q is None:
        return [0] * 48
    return (fp_to_limbs32(q[0][0]) + fp_to_limbs32(q[0][1]) +
            fp_to_limbs32(q[1][0]) + fp_to_limbs32(q[1][1]))
