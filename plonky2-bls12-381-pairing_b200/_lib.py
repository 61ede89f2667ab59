"""ctypes binding of libb381.so (C ABI: include/b381.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, a RuntimeError is raised.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B381_LIB", os.path.join(_HERE, "libb381.so"))   # B381_LIB: experiment builds only

MODE_ARK, MODE_ZK, MODE_LITERAL = 0, 1, 2

ERRORS = {-1: "B381_E_CUDA", -2: "B381_E_ARG", -3: "B381_E_NOT_CANONICAL", -4: "B381_E_ZERO_DIVISION", -5: "B381_E_NOT_INIT", -6: "B381_E_NOT_SQUARE", -7: "B381_E_BAD_ENCODING"}


class B381Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "B381_E_?"), code, msg))
        self.code = code


_lib = None
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_V = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "b381_init": [ctypes.c_int],
    "b381_shutdown": [],
    "b381_last_error": [],
    "b381_device_info": [ctypes.POINTER(ctypes.c_int)] * 3 + [ctypes.POINTER(ctypes.c_size_t)],
    "b381_kernel_launches": [],
    "b381_fp_mul": [_u32p, _u32p, _u32p, ctypes.c_size_t],
    "b381_fp_mul_chain": [_u32p, _u32p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_fp2_mul": [_u32p, _u32p, _u32p, ctypes.c_size_t],
    "b381_fp12_mul": [_u32p, _u32p, _u32p, ctypes.c_size_t],
    "b381_fp12_mul_wbasis": [_u32p, _u32p, _u32p, ctypes.c_size_t],
    "b381_miller_loop": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_multi_miller_loop": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_final_exp": [_u32p, _u32p, ctypes.c_size_t],
    "b381_pairing": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_multi_pairing": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_fp12_product": [_u32p, _u32p, ctypes.c_size_t],
    "b381_literal_optimized": [_u32p, _u32p, _u32p, ctypes.c_size_t],
    "b381_miller_loop_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_final_exp_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    "b381_pairing_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_multi_miller_loop_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_fp_mul_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    "b381_fp_mul_chain_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_fp2_mul_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    "b381_fp12_mul_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    "b381_fp_inv": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp_sqrt": [_u32p, _u8p, _u32p, ctypes.c_size_t],
    "b381_fp_is_square": [_u32p, _u8p, ctypes.c_size_t],
    "b381_fp_pow": [_u32p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t, _u32p, ctypes.c_size_t],
    "b381_fp2_inv": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp2_sqrt": [_u32p, _u8p, _u32p, ctypes.c_size_t],
    "b381_fp2_is_square": [_u32p, _u8p, ctypes.c_size_t],
    "b381_fp6_inv": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp12_inv": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp_to_u32_digits": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp_from_u32_digits": [_u32p, _u32p, ctypes.c_size_t],
    "b381_fp12_to_witness_limbs": [_u32p, _u32p, ctypes.c_size_t],
    "b381_g1_deserialize": [_u8p, ctypes.c_int, _u32p, _u8p, ctypes.c_size_t],
    "b381_g1_serialize": [_u32p, _u8p, ctypes.c_int, _u8p, ctypes.c_size_t],
    "b381_g2_deserialize": [_u8p, ctypes.c_int, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_serialize": [_u32p, _u8p, ctypes.c_int, _u8p, ctypes.c_size_t],
    "b381_g1_in_subgroup": [_u32p, _u8p, _u8p, ctypes.c_size_t],
    "b381_g2_in_subgroup": [_u32p, _u8p, _u8p, ctypes.c_size_t],
    "b381_g1_scalar_mul": [_u32p, _u8p, _u32p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_scalar_mul": [_u32p, _u8p, _u32p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g1_sum": [_u32p, _u8p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_sum": [_u32p, _u8p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g1_msm": [_u32p, _u8p, _u32p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_msm": [_u32p, _u8p, _u32p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_prepare": [_u32p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_miller_loop_prepared": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_pairing_prepared": [_u32p, _u32p, _u8p, _u32p, ctypes.c_size_t, ctypes.c_int],
    "b381_g2_prepare_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_g2_packed_words": [ctypes.c_size_t],
    "b381_g2_prepare_packed_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_miller_loop_packed_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p],
    "b381_multi_miller_loop_packed_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p],
    "b381_miller_loop_packed_one_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p],
    "b381_miller_loop_prepared_dev": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p],
    "b381_check_dev": [ctypes.c_void_p],
    "b381_ctx_create": [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)],
    "b381_ctx_destroy": [ctypes.c_void_p],
    "b381_ctx_set_current": [ctypes.c_void_p],
    "b381_ctx_get_current": [ctypes.POINTER(ctypes.c_void_p)],
    "b381_g1_clear_cofactor": [_u32p, _u8p, _u32p, _u8p, ctypes.c_size_t],
    "b381_g2_clear_cofactor": [_u32p, _u8p, _u32p, _u8p, ctypes.c_size_t],
    "b381_multi_pairing_dev": [_V, _V, _V, _V, ctypes.c_size_t, ctypes.c_int, _V],
    "b381_fp12_mul_wbasis_dev": [_V, _V, _V, ctypes.c_size_t, _V],
    "b381_fp_inv_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp_sqrt_dev": [_V, _V, _V, ctypes.c_size_t, _V],
    "b381_fp_is_square_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp_pow_dev": [_V, ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t, _V, ctypes.c_size_t, _V],
    "b381_fp2_inv_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp2_sqrt_dev": [_V, _V, _V, ctypes.c_size_t, _V],
    "b381_fp2_is_square_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp6_inv_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp12_inv_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp_to_u32_digits_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp_from_u32_digits_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_fp12_to_witness_limbs_dev": [_V, _V, ctypes.c_size_t, _V],
    "b381_g1_deserialize_dev": [_V, ctypes.c_int, _V, _V, ctypes.c_size_t, _V],
    "b381_g1_serialize_dev": [_V, _V, ctypes.c_int, _V, ctypes.c_size_t, _V],
    "b381_g2_deserialize_dev": [_V, ctypes.c_int, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_serialize_dev": [_V, _V, ctypes.c_int, _V, ctypes.c_size_t, _V],
    "b381_g1_in_subgroup_dev": [_V, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_in_subgroup_dev": [_V, _V, _V, ctypes.c_size_t, _V],
    "b381_g1_clear_cofactor_dev": [_V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_clear_cofactor_dev": [_V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g1_scalar_mul_dev": [_V, _V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_scalar_mul_dev": [_V, _V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g1_sum_dev": [_V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_sum_dev": [_V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g1_msm_dev": [_V, _V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_g2_msm_dev": [_V, _V, _V, _V, _V, ctypes.c_size_t, _V],
    "b381_imad_peak": [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)],
}
G2PREP_WORDS = 68 * 72
_RESTYPES = {"b381_last_error": ctypes.c_char_p, "b381_kernel_launches": ctypes.c_ulonglong, "b381_g2_packed_words": ctypes.c_size_t}


def load():
    """dlopen libb381.so and declare every symbol of include/b381.h.  No compute happens here."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb381.so not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export it
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


_initialised = False


def init(device=None):
    """Bind the process to one GPU (default: LOCAL_RANK or 0)."""
    global _initialised
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    rc = lib.b381_init(int(device))
    if rc != 0:
        raise B381Error(rc, lib.b381_last_error().decode())
    _initialised = True
    return lib


def lib():
    if not _initialised:
        init()
    return _lib


def check(rc):
    if rc != 0:
        raise B381Error(rc, _lib.b381_last_error().decode())


def u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_u32p)


def u8(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)
