"""Mirror of the layout facts of /root/reference/src/fields/bls12_381base.rs:21-172:
Bls12_381Base([u64; 6]) holds CANONICAL little-endian limbs (not Montgomery); order() as 12 x u32.
The plonky2 Field-trait implementation of that file is circuit substrate and out of scope."""
from dataclasses import dataclass

from .types import Fq, MODULUS


@dataclass(frozen=True)
class Bls12_381Base:
    limbs: tuple    # 6 x u64, little-endian, canonical

    @staticmethod
    def from_fq(x: Fq) -> "Bls12_381Base":              # bls12_381base.rs:164-172
        return Bls12_381Base(tuple((x.v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)))

    def to_int(self) -> int:
        v = 0
        for i, l in enumerate(self.limbs):
            v |= l << (64 * i)
        return v

    def to_fq(self) -> Fq:
        return Fq(self.to_int() % MODULUS)

    @staticmethod
    def order_u32():                                    # bls12_381base.rs:108-113
        return [(MODULUS >> (32 * i)) & 0xFFFFFFFF for i in range(12)]

    def to_u32_digits(self):
        v = self.to_int()
        return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
