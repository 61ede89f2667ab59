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
