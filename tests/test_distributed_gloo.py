"""world_size-2 (and 3) gloo test of the N>1 host logic: contiguous sharding, all-gather of one
Fq12 per rank, product, single final exponentiation.  The compute callbacks are CPU stand-ins
(the oracle's C port) injected by the test; on GPUs the defaults call libb381.so."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

ROOT = util.ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import b381
    import util as u
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = u.load_ref_lib()
    z = u.pairs_256()
    lo, hi = b381.distributed.shard_bounds(n, rank, world)
    g1 = np.ascontiguousarray(z["g1"][lo:hi]).reshape(-1)
    g2 = np.ascontiguousarray(z["g2"][lo:hi]).reshape(-1)

    def partial(a, b, inf, mode):
        out = np.zeros(144, dtype=np.uint32)
        assert ref.ref_multi_miller_loop(u.p32(a), u.p32(b), None, u.p32(out), a.size // 24, 1) == 0
        return out

    def combine(parts):
        import b381_oracle as o
        pr = o.F12_ONE
        for row in parts:
            pr = o.f12_mul(pr, o.f12_from_limbs32(row))
        fin = np.array(o.f12_to_limbs32(pr), dtype=np.uint32)
        out = np.zeros(144, dtype=np.uint32)
        assert ref.ref_final_exp(u.p32(fin), u.p32(out), 1, 1) == 0
        return out

    res = b381.distributed.multi_pairing_sharded(g1, g2, None, 0, partial_fn=partial, combine_fn=combine)
    ret[rank] = res.tolist()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 9), (3, 2)])
def test_multi_pairing_sharded_gloo(world, n):
    import b381_oracle as o
    z = util.pairs_256()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    pr = o.F12_ONE
    for i in range(n):
        pr = o.f12_mul(pr, o.f12_from_limbs32(z["miller_ark"][i]))
    expect = o.f12_to_limbs32(o.ark_final_exponentiation(pr))
    for rank in range(world):
        assert ret[rank] == expect, "rank %d" % rank
