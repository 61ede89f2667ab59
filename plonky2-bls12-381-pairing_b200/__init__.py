"""B200-native batched BLS12-381 pairing engine -- host-side Python binding.

Mirrors the native API surface of NikolayKostadinov21/plonky2-bls12-381-pairing
(/root/reference/src/lib.rs:1-9 module names): fields, curves, global_constants,
miller_loop_native, miller_loop_native_optimized, utils.  All arithmetic runs in libb381.so
(hand-written CUDA for sm_100a, C ABI in include/b381.h); there is no CPU fallback.

The directory name contains hyphens, so import it with
    importlib.import_module("plonky2-bls12-381-pairing_b200")
or through the repo-root shim `import b381`.
"""
from . import _lib  # noqa: F401
from . import global_constants, fields, curves, utils, distributed  # noqa: F401
from . import miller_loop_native, miller_loop_native_optimized  # noqa: F401
from .miller_loop_native import (MillerLoopResult, multi_miller_loop, miller_loop_batch,  # noqa: F401
                                 final_exponentiation_batch, pairing_batch, multi_pairing)
from .miller_loop_native_optimized import optimized_miller_loop, optimized_miller_loop_batch  # noqa: F401
