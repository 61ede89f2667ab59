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
