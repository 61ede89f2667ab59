"""shared helpers for the tests (checker side: may import the oracle)."""
import ctypes
import json
import os
import random

import numpy as np

import b381_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
U32P = ctypes.POINTER(ctypes.c_uint32)
U8P = ctypes.POINTER(ctypes.c_uint8)


def p32(a):
    return a.ctypes.data_as(U32P)


def p8(a):
    return None if a is None else a.ctypes.data_as(U8P)


def ref_vectors():
    return json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))


def pairing_vectors():
    return json.load(open(os.path.join(GOLDEN, "pairing_vectors.json")))


def pairs_256():
    return np.load(os.path.join(GOLDEN, "pairs_256.npz"))


def limbs64_to_int(limbs):
    v = 0
    for i, h in enumerate(limbs):
        v |= int(h, 16) << (64 * i)
    return v


def limbs64_to_words(limbs):
    """6 x u64 hex limbs (as written in the reference) -> 12 x u32 wire words."""
    w = []
    for h in limbs:
        v = int(h, 16)
        w += [v & 0xFFFFFFFF, v >> 32]
    return w


def mont_decode(limbs):
    return limbs64_to_int(limbs) * o.MONT_RINV % o.P


def rng(seed):
    return random.Random(seed)


def rfp(r):
    return r.randrange(o.P)


def rf2(r):
    return (rfp(r), rfp(r))


def rf12(r):
    return o.f12_unflat([rfp(r) for _ in range(12)])


def f2_words(a):
    return o.fp_to_limbs32(a[0]) + o.fp_to_limbs32(a[1])


def f2_from_words(w):
    return (o.fp_from_limbs32(w[:12]), o.fp_from_limbs32(w[12:24]))


def arr(l):
    return np.array(l, dtype=np.uint32)


def load_ref_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libb381_ref.so"))
    for name in ("ref_miller_loop", "ref_pairing", "ref_multi_miller_loop"):
        getattr(lib, name).argtypes = [U32P, U32P, U8P, U32P, ctypes.c_size_t, ctypes.c_int]
    lib.ref_final_exp.argtypes = [U32P, U32P, ctypes.c_size_t, ctypes.c_int]
    lib.ref_fp_mul.argtypes = [U32P, U32P, U32P, ctypes.c_size_t, ctypes.c_int]
    lib.ref_fp12_mul.argtypes = [U32P, U32P, U32P, ctypes.c_size_t, ctypes.c_int]
    return lib


def load_hostsim(track=False):
    name = "libb381_hostsim_track.so" if track else "libb381_hostsim.so"
    over = os.environ.get("B381_HOSTSIM_LIB")       # development aid: try another build of the host simulation
    if over:
        return ctypes.CDLL(over + ("_track.so" if track else ".so"))
    return ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", name))


def random_pairs(seed, n):
    r = rng(seed)
    return [(o.g1_mul(o.G1_GEN, r.randrange(1, o.R_ORDER)), o.g2_mul(o.G2_GEN, r.randrange(1, o.R_ORDER))) for _ in range(n)]


def marshal_pairs(pairs):
    g1 = arr(sum((o.g1_to_limbs32(p) for p, _ in pairs), []))
    g2 = arr(sum((o.g2_to_limbs32(q) for _, q in pairs), []))
    inf = np.array([(1 if p is None else 0) | (2 if q is None else 0) for p, q in pairs], dtype=np.uint8)
    return g1, g2, inf


def f12s(out, n):
    return [o.f12_from_limbs32(out[144 * i:144 * i + 144]) for i in range(n)]
