//! Raw bindings to `include/b381.h`.  Layouts: Fp = 12 x u32 LE, Montgomery R = 2^384.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const B381_MODE_ARK: c_int = 0;
pub const B381_MODE_ZK: c_int = 1;
pub const B381_MODE_LITERAL: c_int = 2;

pub const B381_OK: c_int = 0;
pub const B381_E_CUDA: c_int = -1;
pub const B381_E_ARG: c_int = -2;
pub const B381_E_NOT_CANONICAL: c_int = -3;
pub const B381_E_ZERO_DIVISION: c_int = -4;
pub const B381_E_NOT_INIT: c_int = -5;
pub const B381_E_NOT_SQUARE: c_int = -6;
pub const B381_E_BAD_ENCODING: c_int = -7;

/// opaque context handle (`b381_ctx_create`): one per GPU, bound to a host thread with `b381_ctx_set_current`
#[repr(C)]
pub struct b381_ctx_s { _private: [u8; 0] }
pub type b381_ctx_t = *mut b381_ctx_s;

extern "C" {
    pub fn b381_init(device: c_int) -> c_int;
    pub fn b381_shutdown() -> c_int;
    pub fn b381_last_error() -> *const c_char;
    pub fn b381_device_info(sm_count: *mut c_int, cc_major: *mut c_int, cc_minor: *mut c_int, scratch_bytes: *mut usize) -> c_int;
    pub fn b381_kernel_launches() -> u64;

    pub fn b381_fp_mul(a: *const u32, b: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp_mul_chain(a: *const u32, b: *const u32, out: *mut u32, n: usize, k: c_int) -> c_int;
    pub fn b381_fp2_mul(a: *const u32, b: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp12_mul(a: *const u32, b: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp12_mul_wbasis(a: *const u32, b: *const u32, out: *mut u32, n: usize) -> c_int;

    pub fn b381_miller_loop(g1: *const u32, g2: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_multi_miller_loop(g1: *const u32, g2: *const u32, inf: *const u8, out144: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_final_exp(f: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_pairing(g1: *const u32, g2: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_multi_pairing(g1: *const u32, g2: *const u32, inf: *const u8, out144: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_fp12_product(input: *const u32, out144: *mut u32, n: usize) -> c_int;
    pub fn b381_literal_optimized(g1proj: *const u32, g2proj: *const u32, out: *mut u32, n: usize) -> c_int;

    pub fn b381_miller_loop_dev(g1: *const u32, g2: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_final_exp_dev(f: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_pairing_dev(g1: *const u32, g2: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_multi_miller_loop_dev(g1: *const u32, g2: *const u32, inf: *const u8, out144: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_fp_mul_dev(a: *const u32, b: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_mul_chain_dev(a: *const u32, b: *const u32, out: *mut u32, n: usize, k: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_fp2_mul_dev(a: *const u32, b: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp12_mul_dev(a: *const u32, b: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_inv(a: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp_sqrt(a: *const u32, sgn: *const u8, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp_is_square(a: *const u32, out: *mut u8, n: usize) -> c_int;
    pub fn b381_fp_pow(a: *const u32, exp: *const u64, exp_limbs: usize, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp2_inv(a: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp2_sqrt(a: *const u32, sgn: *const u8, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp2_is_square(a: *const u32, out: *mut u8, n: usize) -> c_int;
    pub fn b381_fp6_inv(a: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp12_inv(a: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp_to_u32_digits(a: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp_from_u32_digits(digits: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_fp12_to_witness_limbs(f: *const u32, out: *mut u32, n: usize) -> c_int;
    pub fn b381_g1_deserialize(input: *const u8, compressed: c_int, g1: *mut u32, inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g1_serialize(g1: *const u32, inf: *const u8, compressed: c_int, out: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_deserialize(input: *const u8, compressed: c_int, g2: *mut u32, inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_serialize(g2: *const u32, inf: *const u8, compressed: c_int, out: *mut u8, n: usize) -> c_int;
    pub fn b381_g1_in_subgroup(g1: *const u32, inf: *const u8, out: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_in_subgroup(g2: *const u32, inf: *const u8, out: *mut u8, n: usize) -> c_int;
    pub fn b381_g1_scalar_mul(g1: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_scalar_mul(g2: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g1_sum(g1: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_sum(g2: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g1_msm(g1: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_msm(g2: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_prepare(g2: *const u32, coeffs: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_miller_loop_prepared(g1: *const u32, coeffs: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_pairing_prepared(g1: *const u32, coeffs: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int) -> c_int;
    pub fn b381_g2_packed_words(n: usize) -> usize;
    pub fn b381_g2_prepare_packed_dev(g2: *const u32, packed: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_miller_loop_packed_dev(g1: *const u32, packed: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int, final_exp: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_multi_miller_loop_packed_dev(g1: *const u32, packed: *const u32, inf: *const u8, out144: *mut u32, n: usize, final_exp: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_miller_loop_packed_one_dev(g1: *const u32, packed_one: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int, final_exp: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_g2_prepare_dev(g2: *const u32, coeffs: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_miller_loop_prepared_dev(g1: *const u32, coeffs: *const u32, inf: *const u8, out: *mut u32, n: usize, mode: c_int, final_exp: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_check_dev(stream: *mut c_void) -> c_int;

    // ---- round 2: contexts, cofactor clearing, `_dev` twins of every batch entry point ----
    pub fn b381_ctx_create(device: c_int, out: *mut b381_ctx_t) -> c_int;
    pub fn b381_ctx_destroy(ctx: b381_ctx_t) -> c_int;
    pub fn b381_ctx_set_current(ctx: b381_ctx_t) -> c_int;
    pub fn b381_ctx_get_current(out: *mut b381_ctx_t) -> c_int;
    pub fn b381_g1_clear_cofactor(g1: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_g2_clear_cofactor(g2: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize) -> c_int;
    pub fn b381_multi_pairing_dev(g1: *const u32, g2: *const u32, inf: *const u8, out144: *mut u32, n: usize, mode: c_int, stream: *mut c_void) -> c_int;
    pub fn b381_fp12_mul_wbasis_dev(a: *const u32, b: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_inv_dev(a: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_sqrt_dev(a: *const u32, sgn: *const u8, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_is_square_dev(a: *const u32, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_pow_dev(a: *const u32, exp: *const u64, exp_limbs: usize, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp2_inv_dev(a: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp2_sqrt_dev(a: *const u32, sgn: *const u8, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp2_is_square_dev(a: *const u32, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp6_inv_dev(a: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp12_inv_dev(a: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_to_u32_digits_dev(a: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp_from_u32_digits_dev(digits: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_fp12_to_witness_limbs_dev(f: *const u32, out: *mut u32, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_deserialize_dev(in_: *const u8, compressed: c_int, g1: *mut u32, inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_serialize_dev(g1: *const u32, inf: *const u8, compressed: c_int, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_deserialize_dev(in_: *const u8, compressed: c_int, g2: *mut u32, inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_serialize_dev(g2: *const u32, inf: *const u8, compressed: c_int, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_in_subgroup_dev(g1: *const u32, inf: *const u8, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_in_subgroup_dev(g2: *const u32, inf: *const u8, out: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_clear_cofactor_dev(g1: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_clear_cofactor_dev(g2: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_scalar_mul_dev(g1: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_scalar_mul_dev(g2: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_sum_dev(g1: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_sum_dev(g2: *const u32, inf: *const u8, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g1_msm_dev(g1: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_g2_msm_dev(g2: *const u32, inf: *const u8, scalars: *const u32, out: *mut u32, out_inf: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn b381_imad_peak(imad_wide_ginst_per_s: *mut f64, sm_mhz: *mut f64) -> c_int;
}
