"""Tower value types and the flat-buffer layout of include/b381.h.

Fq   : canonical integer in [0, p);  wire form = 12 x u32 LE limbs of x * 2^384 mod p (ark Fp384)
Fq2  : (c0, c1);  Fq6 : (c0, c1, c2) of Fq2;  Fq12 : (c0, c1) of Fq6 -- field order as used at
/root/reference/src/fields/helpers.rs:16-37.
"""
from dataclasses import dataclass

import numpy as np

MODULUS = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_R = 1 << 384
_RINV = pow(_R, -1, MODULUS)


def to_limbs32(x):
    """canonical integer -> 12 LE u32 limbs of its Montgomery form."""
    if not (0 <= x < MODULUS):
        raise ValueError("Fq value out of range")
    m = x * _R % MODULUS
    return [(m >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def from_limbs32(limbs):
    m = 0
    for i in range(12):
        m |= int(limbs[i]) << (32 * i)
    if m >= MODULUS:
        raise ValueError("non-canonical Montgomery limbs")
    return m * _RINV % MODULUS


@dataclass(frozen=True)
class Fq:
    v: int

    def __post_init__(self):
        if not (0 <= self.v < MODULUS):
            raise ValueError("Fq value out of range")

    @staticmethod
    def zero():
        return Fq(0)

    @staticmethod
    def one():
        return Fq(1)

    def is_zero(self):
        return self.v == 0

    def limbs(self):
        return to_limbs32(self.v)

    @staticmethod
    def from_limbs(l):
        return Fq(from_limbs32(l))


@dataclass(frozen=True)
class Fq2:
    c0: Fq
    c1: Fq

    @staticmethod
    def new(c0, c1):
        return Fq2(c0, c1)

    @staticmethod
    def zero():
        return Fq2(Fq(0), Fq(0))

    @staticmethod
    def one():
        return Fq2(Fq(1), Fq(0))

    def is_zero(self):
        return self.c0.is_zero() and self.c1.is_zero()

    def limbs(self):
        return self.c0.limbs() + self.c1.limbs()

    @staticmethod
    def from_limbs(l):
        return Fq2(Fq.from_limbs(l[0:12]), Fq.from_limbs(l[12:24]))


@dataclass(frozen=True)
class Fq6:
    c0: Fq2
    c1: Fq2
    c2: Fq2

    @staticmethod
    def new(c0, c1, c2):
        return Fq6(c0, c1, c2)

    @staticmethod
    def zero():
        return Fq6(Fq2.zero(), Fq2.zero(), Fq2.zero())

    def limbs(self):
        return self.c0.limbs() + self.c1.limbs() + self.c2.limbs()

    @staticmethod
    def from_limbs(l):
        return Fq6(Fq2.from_limbs(l[0:24]), Fq2.from_limbs(l[24:48]), Fq2.from_limbs(l[48:72]))


@dataclass(frozen=True)
class Fq12:
    c0: Fq6
    c1: Fq6

    @staticmethod
    def new(c0, c1):
        return Fq12(c0, c1)

    @staticmethod
    def one():
        return Fq12(Fq6(Fq2.one(), Fq2.zero(), Fq2.zero()), Fq6.zero())

    def limbs(self):
        return self.c0.limbs() + self.c1.limbs()

    @staticmethod
    def from_limbs(l):
        return Fq12(Fq6.from_limbs(l[0:72]), Fq6.from_limbs(l[72:144]))

    def flat(self):
        """12 canonical integers in tower order."""
        return [c.v for f6 in (self.c0, self.c1) for f2 in (f6.c0, f6.c1, f6.c2) for c in (f2.c0, f2.c1)]

    @staticmethod
    def from_flat(vals):
        f2 = [Fq2(Fq(vals[2 * i]), Fq(vals[2 * i + 1])) for i in range(6)]
        return Fq12(Fq6(*f2[0:3]), Fq6(*f2[3:6]))

    def to_array(self):
        return np.array(self.limbs(), dtype=np.uint32)
