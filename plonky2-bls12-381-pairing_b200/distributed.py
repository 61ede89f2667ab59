"""Multi-GPU host logic: one process per GPU (torch.distributed), pairings sharded contiguously.

Independent pairings need no collective.  The only exchange step on this path is the
multi-pairing product (BASELINE config #5): every rank reduces its shard to ONE Fq12 (576 B), the
partials are all-gathered (NCCL on GPUs, gloo in the CPU tests), multiplied, and a single final
exponentiation follows.  Fq12 multiplication is commutative, so the result is bit-exact regardless
of rank count.  The compute callbacks default to the CUDA library; tests inject CPU stand-ins to
exercise this logic without a GPU."""
from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [lo, hi) of a batch of n units owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _cuda_partial(g1, g2, inf, mode):
    from . import _lib
    lib = _lib.lib()
    out = np.zeros(144, dtype=np.uint32)
    _lib.check(lib.b381_multi_miller_loop(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u8(inf)[1] if inf is not None else None,
                                          _lib.u32(out)[1], g1.size // 24, mode))
    return out


def _cuda_combine(parts):
    from . import _lib
    lib = _lib.lib()
    prod = np.zeros(144, dtype=np.uint32)
    flat = np.ascontiguousarray(parts, dtype=np.uint32).reshape(-1)
    _lib.check(lib.b381_fp12_product(_lib.u32(flat)[1], _lib.u32(prod)[1], flat.size // 144))
    out = np.zeros(144, dtype=np.uint32)
    _lib.check(lib.b381_final_exp(_lib.u32(prod)[1], _lib.u32(out)[1], 1))
    return out


def multi_pairing_sharded(g1: np.ndarray, g2: np.ndarray, inf: Optional[np.ndarray], mode: int = 0,
                          group=None, partial_fn: Callable = _cuda_partial, combine_fn: Callable = _cuda_combine,
                          device=None) -> np.ndarray:
    """final_exp(prod_i miller(P_i, Q_i)) over a batch sharded across the ranks of `group`.

    g1 / g2 / inf hold THIS rank's shard (flat u32 / u8 arrays, include/b381.h layout).  An empty
    shard contributes 1.  Returns the 144-word result on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if g1.size:
        part = partial_fn(g1, g2, inf, mode)
    else:
        part = one_fq12_words()
    if world == 1:
        return combine_fn(part.reshape(1, 144))
    t = torch.from_numpy(part.astype(np.int64))        # int64 carrier: gloo and nccl both move it
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    parts = np.stack([x.cpu().numpy().astype(np.uint32) for x in gathered])
    return combine_fn(parts)


def one_fq12_words() -> np.ndarray:
    """Fq12::one() in wire form (Montgomery 1 = 2^384 mod p in c0.c0.c0)."""
    from .fields.types import to_limbs32
    w = np.zeros(144, dtype=np.uint32)
    w[:12] = to_limbs32(1)
    return w


# ---- one process, several GPUs: a context per device, a host thread per context --------------------------------
class DeviceSet:
    """One libb381 context per GPU (b381_ctx_create), driven by one host thread each: what a Rust host (SURVEY 8b
    "Threading": one context per device, calls from several host threads) would do with a thread pool.  Independent
    pairings are sharded contiguously over the devices, no data-path collective; the multi-pairing product combines
    the per-device Fq12 partials on the first device."""

    def __init__(self, devices):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.devices = list(devices)
        self.ctxs = []
        for d in self.devices:
            c = ctypes.c_void_p()
            _lib.check(self.lib.b381_ctx_create(int(d), ctypes.byref(c)))
            self.ctxs.append(c)

    def close(self):
        for c in self.ctxs:
            self.lib.b381_ctx_destroy(c)
        self.ctxs = []

    def _run(self, fn):
        """fn(k) on a thread bound to context k, for every k; returns the list of results (first exception re-raised)"""
        import threading
        res, errs = [None] * len(self.ctxs), []

        def work(k):
            try:
                self._lib.check(self.lib.b381_ctx_set_current(self.ctxs[k]))
                res[k] = fn(k)
            except Exception as e:      # noqa: BLE001
                errs.append(e)
            finally:
                self.lib.b381_ctx_set_current(None)

        th = [threading.Thread(target=work, args=(k,)) for k in range(len(self.ctxs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return res

    def pairing(self, g1: np.ndarray, g2: np.ndarray, inf: Optional[np.ndarray], mode: int = 0) -> np.ndarray:
        """out[i] = e(P_i, Q_i) for a host batch, sharded contiguously over the devices"""
        L, lib = self._lib, self.lib
        n = g1.size // 24
        out = np.empty(n * 144, dtype=np.uint32)
        w = len(self.ctxs)

        def shard(k):
            lo, hi = shard_bounds(n, k, w)
            if hi > lo:
                L.check(lib.b381_pairing(L.u32(g1[24 * lo:24 * hi])[1], L.u32(g2[48 * lo:48 * hi])[1],
                                         L.u8(inf[lo:hi])[1] if inf is not None else None, L.u32(out[144 * lo:144 * hi])[1], hi - lo, mode))
        self._run(shard)
        return out

    def multi_pairing(self, g1: np.ndarray, g2: np.ndarray, inf: Optional[np.ndarray], mode: int = 0) -> np.ndarray:
        """final_exp(prod_i miller(P_i, Q_i)): per-device partial products, combined on the first device"""
        L, lib = self._lib, self.lib
        n = g1.size // 24
        w = len(self.ctxs)
        parts = np.tile(one_fq12_words(), (w, 1))

        def shard(k):
            lo, hi = shard_bounds(n, k, w)
            if hi > lo:
                L.check(lib.b381_multi_miller_loop(L.u32(g1[24 * lo:24 * hi])[1], L.u32(g2[48 * lo:48 * hi])[1],
                                                   L.u8(inf[lo:hi])[1] if inf is not None else None, L.u32(parts[k])[1], hi - lo, mode))
        self._run(shard)
        out = np.zeros(144, dtype=np.uint32)

        def combine(k):
            if k == 0:
                prod = np.zeros(144, dtype=np.uint32)
                flat = np.ascontiguousarray(parts).reshape(-1)
                L.check(lib.b381_fp12_product(L.u32(flat)[1], L.u32(prod)[1], w))
                L.check(lib.b381_final_exp(L.u32(prod)[1], L.u32(out)[1], 1))
        self._run(combine)
        return out
