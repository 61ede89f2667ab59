"""Mirror of /root/reference/src/fields/helpers.rs: MyFq12 (w-basis Fp12), conversions and small
native helpers.  MyFq12 multiplication runs on the GPU through b381_fp12_mul_wbasis."""
from dataclasses import dataclass
from typing import List

import numpy as np

from .. import _lib
from .types import Fq, Fq2, Fq6, Fq12, MODULUS


@dataclass(frozen=True)
class MyFq12:
    """helpers.rs:8-11.  coeffs = [c000,c100,c010,c110,c020,c120, c001,c101,c011,c111,c021,c121]."""
    coeffs: tuple

    @staticmethod
    def from_fq12(f: Fq12) -> "MyFq12":                 # helpers.rs:14-44
        c0, c1 = f.c0, f.c1
        return MyFq12((c0.c0.c0, c1.c0.c0, c0.c1.c0, c1.c1.c0, c0.c2.c0, c1.c2.c0,
                       c0.c0.c1, c1.c0.c1, c0.c1.c1, c1.c1.c1, c0.c2.c1, c1.c2.c1))

    def to_fq12(self) -> Fq12:                          # helpers.rs:47-76
        m = self.coeffs
        c0 = Fq6(Fq2(m[0], m[6]), Fq2(m[2], m[8]), Fq2(m[4], m[10]))
        c1 = Fq6(Fq2(m[1], m[7]), Fq2(m[3], m[9]), Fq2(m[5], m[11]))
        return Fq12(c0, c1)

    def __add__(self, rhs: "MyFq12") -> "MyFq12":       # helpers.rs:78-88
        return MyFq12(tuple(Fq((a.v + b.v) % MODULUS) for a, b in zip(self.coeffs, rhs.coeffs)))

    def limbs(self):
        out = []
        for c in self.coeffs:
            out += c.limbs()
        return out

    @staticmethod
    def from_limbs(l):
        return MyFq12(tuple(Fq.from_limbs(l[12 * i:12 * i + 12]) for i in range(12)))

    def __mul__(self, rhs: "MyFq12") -> "MyFq12":       # helpers.rs:90-152, on the GPU
        return myfq12_mul_batch([self], [rhs])[0]


def myfq12_mul_batch(a: List[MyFq12], b: List[MyFq12]) -> List[MyFq12]:
    n = len(a)
    if n != len(b) or n == 0:
        raise ValueError("batch sizes")
    lib = _lib.lib()
    xa = np.array([x.limbs() for x in a], dtype=np.uint32).reshape(-1)
    xb = np.array([x.limbs() for x in b], dtype=np.uint32).reshape(-1)
    out = np.zeros(n * 144, dtype=np.uint32)
    _lib.check(lib.b381_fp12_mul_wbasis(_lib.u32(xa)[1], _lib.u32(xb)[1], _lib.u32(out)[1], n))
    return [MyFq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(n)]


def from_biguint_to_fq(x: int) -> Fq:
    """helpers.rs:154-157 (`Fq::from_bigint(x).unwrap()`: panics -> ValueError when x >= p)."""
    return Fq(int(x))


def sgn0_fq(x: Fq) -> bool:
    """helpers.rs:159-167."""
    return (x.v & 1) == 1


def sgn0_fq2(x: Fq2) -> bool:
    """helpers.rs:169-174."""
    return sgn0_fq(x.c0) or (x.c0.is_zero() and sgn0_fq(x.c1))


def get_naf(exp: List[int]) -> List[int]:
    """helpers.rs:197-239: NAF digits (LSB first) of a little-endian u64-limb exponent."""
    exp = list(exp)
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


def pow_fq(a: Fq, exp: List[int]) -> Fq:
    """helpers.rs:176-195 (NAF ladder; host-side scalar helper)."""
    res = a.v
    inv = None
    started = False
    for z in reversed(get_naf(exp)):
        if started:
            res = res * res % MODULUS
        if z != 0:
            if started:
                if z == 1:
                    res = res * a.v % MODULUS
                else:
                    if inv is None:
                        inv = pow(a.v, -1, MODULUS)
                    res = res * inv % MODULUS
            else:
                assert z == 1
                started = True
    return Fq(res)


# ---- batched witness helpers on the GPU: the native computations of the reference's
# ---- circuit generators (fq_target.rs:243-343, fq2_target.rs:320-410, fq6_target.rs:384-418, fq12_target.rs:340-374)
def _flat(xs):
    return np.array([w for x in xs for w in x.limbs()], dtype=np.uint32)


def _u8p(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _unary(fn_name, xs, cls, words):
    n = len(xs)
    if n == 0:
        raise ValueError("empty batch")
    a = _flat(xs)
    out = np.zeros(n * words, dtype=np.uint32)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.u32(a)[1], _lib.u32(out)[1], n))
    return [cls.from_limbs(out[words * i:words * (i + 1)]) for i in range(n)]


def inverse_fq_batch(xs: List[Fq]) -> List[Fq]:
    return _unary("b381_fp_inv", xs, Fq, 12)


def inverse_fq2_batch(xs: List[Fq2]) -> List[Fq2]:
    return _unary("b381_fp2_inv", xs, Fq2, 24)


def inverse_fq6_batch(xs: List[Fq6]) -> List[Fq6]:
    return _unary("b381_fp6_inv", xs, Fq6, 72)


def inverse_fq12_batch(xs: List[Fq12]) -> List[Fq12]:
    return _unary("b381_fp12_inv", xs, Fq12, 144)


def _sqrt(fn_name, xs, sgns, cls, words):
    n = len(xs)
    if n == 0:
        raise ValueError("empty batch")
    a = _flat(xs)
    s = np.array([1 if b else 0 for b in sgns], dtype=np.uint8)
    out = np.zeros(n * words, dtype=np.uint32)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.u32(a)[1], _u8p(s), _lib.u32(out)[1], n))
    return [cls.from_limbs(out[words * i:words * (i + 1)]) for i in range(n)]


def sqrt_with_sgn_fq_batch(xs: List[Fq], sgns: List[bool]) -> List[Fq]:
    """FqSqrtGenerator::run_once, fq_target.rs:316-343 (B381_E_NOT_SQUARE where the reference panics)."""
    return _sqrt("b381_fp_sqrt", xs, sgns, Fq, 12)


def sqrt_with_sgn_fq2_batch(xs: List[Fq2], sgns: List[bool]) -> List[Fq2]:
    return _sqrt("b381_fp2_sqrt", xs, sgns, Fq2, 24)


def _is_square(fn_name, xs):
    n = len(xs)
    if n == 0:
        raise ValueError("empty batch")
    a = _flat(xs)
    out = np.zeros(n, dtype=np.uint8)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.u32(a)[1], _u8p(out), n))
    return [bool(b) for b in out]


def is_square_fq_batch(xs: List[Fq]) -> List[bool]:
    """fq_target.rs:269-280: legendre(x) == 1."""
    return _is_square("b381_fp_is_square", xs)


def is_square_fq2_batch(xs: List[Fq2]) -> List[bool]:
    return _is_square("b381_fp2_is_square", xs)


def pow_fq_batch(xs: List[Fq], exp: List[int]) -> List[Fq]:
    """pow_fq (helpers.rs:176-195) of every element with one exponent (u64 limbs, little-endian)."""
    import ctypes
    n = len(xs)
    if n == 0:
        raise ValueError("empty batch")
    a = _flat(xs)
    e = (ctypes.c_uint64 * len(exp))(*[int(x) for x in exp])
    out = np.zeros(n * 12, dtype=np.uint32)
    _lib.check(_lib.lib().b381_fp_pow(_lib.u32(a)[1], e, len(exp), _lib.u32(out)[1], n))
    return [Fq.from_limbs(out[12 * i:12 * i + 12]) for i in range(n)]
