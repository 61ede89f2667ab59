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


# ---- G2Prepared: cached line coefficients (shape of src/miller_loop_target.rs:23-76 / ark-ec G2Prepared) ----
@dataclass(frozen=True)
class G2Prepared:
    """68 coefficient triples (Fq2, Fq2, Fq2) of one Q in the order the Miller loop of `mode`
    consumes them, kept in the C-ABI layout (68 x 72 words); `infinity` as in ark-ec."""
    coeffs: np.ndarray
    infinity: bool
    mode: int

    @staticmethod
    def from_affine(q: G2Affine, mode: int = MODE_ARK) -> "G2Prepared":
        return g2_prepare_batch([q], mode)[0]


def g2_prepare_batch(qs: Sequence[G2Affine], mode: int = MODE_ARK) -> List[G2Prepared]:
    n = len(qs)
    if n == 0:
        raise ValueError("empty batch")
    gen = G2Affine.generator().limbs()
    g2 = np.array([gen if q.infinity else q.limbs() for q in qs], dtype=np.uint32).reshape(-1)
    co = np.zeros(n * _lib.G2PREP_WORDS, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_g2_prepare(_lib.u32(g2)[1], _lib.u32(co)[1], n, mode))
    W = _lib.G2PREP_WORDS
    return [G2Prepared(co[W * i:W * (i + 1)].copy(), bool(q.infinity), mode) for i, q in enumerate(qs)]


def _marshal_prepared(terms: Sequence[Tuple[G1Affine, G2Prepared]]):
    n = len(terms)
    if n == 0:
        raise ValueError("empty batch")
    mode = terms[0][1].mode
    if any(q.mode != mode for _, q in terms):
        raise ValueError("prepared points of different modes in one batch")
    g1 = np.array([p.limbs() for p, _ in terms], dtype=np.uint32).reshape(-1)
    co = np.concatenate([q.coeffs for _, q in terms]).astype(np.uint32)
    inf = np.array([(1 if p.infinity else 0) | (2 if q.infinity else 0) for p, q in terms], dtype=np.uint8)
    return n, g1, co, inf, mode


def miller_loop_prepared_batch(terms: Sequence[Tuple[G1Affine, G2Prepared]]) -> List[MillerLoopResult]:
    n, g1, co, inf, mode = _marshal_prepared(terms)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_miller_loop_prepared(_lib.u32(g1)[1], _lib.u32(co)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return [MillerLoopResult(Fq12.from_limbs(out[144 * i:144 * i + 144])) for i in range(n)]


def pairing_prepared_batch(terms: Sequence[Tuple[G1Affine, G2Prepared]]) -> List[Fq12]:
    n, g1, co, inf, mode = _marshal_prepared(terms)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_pairing_prepared(_lib.u32(g1)[1], _lib.u32(co)[1], _lib.u8(inf)[1], _lib.u32(out)[1], n, mode))
    return [Fq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(n)]


class G2PreparedBatch:
    """A batch of prepared Q's kept ON THE DEVICE in the library's packed layout (include/b381.h
    b381_g2_prepare_packed_dev): the cached stage for callers that pair many P's against the same Q's.
    Opaque: valid for this library build, this mode and this GPU context only."""

    def __init__(self, qs: Sequence[G2Affine], mode: int = MODE_ARK):
        import torch
        n = len(qs)
        if n == 0:
            raise ValueError("empty batch")
        self.n, self.mode = n, mode
        self.q_infinity = np.array([2 if q.infinity else 0 for q in qs], dtype=np.uint8)
        gen = G2Affine.generator().limbs()
        g2 = np.array([gen if q.infinity else q.limbs() for q in qs], dtype=np.uint32).reshape(-1)
        lib = _lib.lib()
        self._dev = torch.device("cuda", torch.cuda.current_device())
        d2 = torch.from_numpy(g2.view(np.int32)).to(self._dev)
        self.packed = torch.empty(lib.b381_g2_packed_words(n), dtype=torch.int32, device=self._dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.b381_g2_prepare_packed_dev(d2.data_ptr(), self.packed.data_ptr(), n, mode, st))
        _lib.check(lib.b381_check_dev(st))

    def _run(self, ps: Sequence[G1Affine], final_exp: int) -> np.ndarray:
        import torch
        if len(ps) != self.n:
            raise ValueError("one P per prepared Q")
        lib = _lib.lib()
        g1 = np.array([p.limbs() for p in ps], dtype=np.uint32).reshape(-1)
        inf = self.q_infinity | np.array([1 if p.infinity else 0 for p in ps], dtype=np.uint8)
        d1 = torch.from_numpy(g1.view(np.int32)).to(self._dev)
        dinf = torch.from_numpy(inf).to(self._dev)
        out = torch.empty(self.n * 144, dtype=torch.int32, device=self._dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.b381_miller_loop_packed_dev(d1.data_ptr(), self.packed.data_ptr(), dinf.data_ptr(), out.data_ptr(), self.n, self.mode, final_exp, st))
        _lib.check(lib.b381_check_dev(st))
        return out.cpu().numpy().view(np.uint32)

    def miller_loop(self, ps: Sequence[G1Affine]) -> List[MillerLoopResult]:
        out = self._run(ps, 0)
        return [MillerLoopResult(Fq12.from_limbs(out[144 * i:144 * i + 144])) for i in range(self.n)]

    def pairing(self, ps: Sequence[G1Affine]) -> List[Fq12]:
        out = self._run(ps, 1)
        return [Fq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(self.n)]
