"""Mirror of /root/reference/src/fields/my_fq6.rs:6-56."""
from dataclasses import dataclass

from .types import Fq, Fq2, Fq6, MODULUS


@dataclass(frozen=True)
class MyFq6:
    """coeffs = [c00, c10, c20, c01, c11, c21] (my_fq6.rs:25)."""
    coeffs: tuple

    @staticmethod
    def from_fq6(f: Fq6) -> "MyFq6":                    # my_fq6.rs:12-28
        return MyFq6((f.c0.c0, f.c1.c0, f.c2.c0, f.c0.c1, f.c1.c1, f.c2.c1))

    def to_fq6(self) -> Fq6:                            # my_fq6.rs:31-44
        m = self.coeffs
        return Fq6(Fq2(m[0], m[3]), Fq2(m[1], m[4]), Fq2(m[2], m[5]))

    def __add__(self, rhs: "MyFq6") -> "MyFq6":         # my_fq6.rs:46-56
        return MyFq6(tuple(Fq((a.v + b.v) % MODULUS) for a, b in zip(self.coeffs, rhs.coeffs)))
