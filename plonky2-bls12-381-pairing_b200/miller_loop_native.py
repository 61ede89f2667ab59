"""Mirror of /root/reference/src/miller_loop_native.rs: multi_miller_loop / MillerLoopResult,
plus the batched variants and the final exponentiation the north star adds.  All arithmetic runs
on the GPU through libb381.so; there is no CPU path."""
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import MODE_ARK, MODE_ZK, MODE_LITERAL  # noqa: F401
from .curves import G1Affine, G2Affine
from .fields.types import Fq12


@dataclass(frozen=True)
class MillerLoopResult:
    """miller_loop_native.rs:7-14 (Default = Fq12::one())."""
    f: Fq12

    @staticmethod
    def default():
        return MillerLoopResult(Fq12.one())

    def final_exponentiation(self) -> Fq12:
        return final_exponentiation_batch([self.f])[0]


def _marshal(terms: Sequence[Tuple[G1Affine, G2Affine]]):
    n = len(terms)
    if n == 0:
        raise ValueError("empty batch")
    g1 = np.array([p.limbs() for p, _ in terms], dtype=np.uint32).reshape(-1)
    g2 = np.array([q.limbs() for _, q in terms], dtype=np.uint32).reshape(-1)
    inf = np.array([(1 if p.infinity else 0) | (2 if q.infinity else 0) for p, q in terms], dtype=np.uint8)
    return n, g1, g2, inf


def multi_miller_loop(terms: Sequence[Tuple[G1Affine, G2Affine]], mode: int = MODE_ARK) -> MillerLoopResult:
    """miller_loop_native.rs:154-212.  mode=MODE_LITERAL reproduces the code as written (returns 1);
    MODE_ZK completes the zkcrypto-structured loop the file spells out; MODE_ARK (default) is
    ark_bls12_381::Bls12_381::multi_miller_loop, the truth the reference defers to."""
    if len(terms) == 0:
        return MillerLoopResult.default()
    n, g1, g2, inf = _marshal(terms)
    out = np.zeros(144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_multi_miller_loop(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return MillerLoopResult(Fq12.from_limbs(out))


def miller_loop_batch(terms, mode: int = MODE_ARK) -> List[MillerLoopResult]:
    """batched variant: one Miller value per pair."""
    n, g1, g2, inf = _marshal(terms)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_miller_loop(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return [MillerLoopResult(Fq12.from_limbs(out[144 * i:144 * i + 144])) for i in range(n)]


def final_exponentiation_batch(fs: Sequence[Fq12]) -> List[Fq12]:
    n = len(fs)
    if n == 0:
        raise ValueError("empty batch")
    a = np.array([f.limbs() for f in fs], dtype=np.uint32).reshape(-1)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_final_exp(_lib.u32(a)[1], _lib.u32(out)[1], n))
    return [Fq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(n)]


def pairing_batch(terms, mode: int = MODE_ARK) -> List[Fq12]:
    n, g1, g2, inf = _marshal(terms)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_pairing(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return [Fq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(n)]


def multi_pairing(terms, mode: int = MODE_ARK) -> Fq12:
    n, g1, g2, inf = _marshal(terms)
    out = np.zeros(144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_multi_pairing(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return Fq12.from_limbs(out)
