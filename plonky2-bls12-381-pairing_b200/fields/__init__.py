"""Host-side value types of the native API surface (mirror of /root/reference/src/fields/mod.rs):
ark-style tower types Fq / Fq2 / Fq6 / Fq12 as thin wrappers over canonical Python integers plus
the reference's own native types MyFq12 / MyFq6 / Bls12_381Base.  These are marshalling types:
all batched arithmetic goes through the C ABI (libb381.so); nothing here is a compute fallback."""
from .types import Fq, Fq2, Fq6, Fq12, MODULUS, to_limbs32, from_limbs32  # noqa: F401
from .helpers import MyFq12, from_biguint_to_fq, sgn0_fq, sgn0_fq2, pow_fq, get_naf  # noqa: F401
from .my_fq6 import MyFq6  # noqa: F401
from .bls12_381base import Bls12_381Base  # noqa: F401
