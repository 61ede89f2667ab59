"""Mirror of /root/reference/src/miller_loop_native_optimized.rs:81-127 (LITERAL semantics: the
function exactly as written), batched on the GPU.  The reference's three banner
println! lines (:123-125) are not reproduced."""
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from .curves import G1Projective, G2Projective
from .fields.types import Fq12


def optimized_miller_loop_batch(pairs: Sequence[Tuple[G1Projective, G2Projective]]) -> List[Fq12]:
    n = len(pairs)
    if n == 0:
        raise ValueError("empty batch")
    g1 = np.array([p.limbs() for p, _ in pairs], dtype=np.uint32).reshape(-1)
    g2 = np.array([q.limbs() for _, q in pairs], dtype=np.uint32).reshape(-1)
    out = np.zeros(n * 144, dtype=np.uint32)
    lib = _lib.lib()
    _lib.check(lib.b381_literal_optimized(_lib.u32(g1)[1], _lib.u32(g2)[1], _lib.u32(out)[1], n))
    return [Fq12.from_limbs(out[144 * i:144 * i + 144]) for i in range(n)]


def optimized_miller_loop(P: G1Projective, Q: G2Projective) -> Fq12:
    return optimized_miller_loop_batch([(P, Q)])[0]
