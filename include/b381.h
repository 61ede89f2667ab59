/* b381.h -- C ABI of libb381.so: B200-native batched BLS12-381 pairing engine.
 *
 * Drop-in boundary for the native pairing path of NikolayKostadinov21/plonky2-bls12-381-pairing
 * (citations relative to /root/reference).  The reference has no FFI; the boundary is its Rust
 * `pub` surface, and each entry point below names the item it replaces.  A Rust (or ctypes)
 * binding only marshals flat buffers -- see INTEGRATION.md.
 *
 * Data layout (identical to ark-ff 0.4 `Fp384` in memory, which is what the reference's types hold):
 *   Fp    : 12 x uint32 little-endian limbs, Montgomery form with R = 2^384, fully reduced (< p)
 *   Fp2   : c0, c1                                  (24 words)
 *   Fp12  : c0.c0.c0, c0.c0.c1, c0.c1.c0, ... , c1.c2.c1   (144 words; tower order of
 *           src/fields/helpers.rs:16-37)
 *   MyFq12: coeffs[0..12] in the w-basis order of src/fields/helpers.rs:39-41 (144 words)
 *   G1 affine: x, y (24 words);  G2 affine: x.c0, x.c1, y.c0, y.c1 (48 words)
 *   G1/G2 projective (Jacobian, as ark-ec 0.4): x, y, z (36 / 72 words)
 *   inf   : one byte per pair, bit0 = P is the identity, bit1 = Q is the identity (may be NULL)
 *
 * All functions return 0 on success, a negative B381_E_* code otherwise; they never throw or
 * unwind.  Host-pointer functions are synchronous (H2D copy, kernels, D2H copy inside the call);
 * `_dev` functions take device pointers, enqueue on `stream` (a cudaStream_t, may be NULL) and do
 * not synchronise.
 *
 * Contexts and threads (SURVEY 8b "Threading").  A context owns one device's streams, staging buffers and scratch.
 * b381_init(device) creates the process-wide default context; b381_ctx_create makes further ones (one per GPU when
 * one process drives several GPUs) and b381_ctx_set_current binds a context to the calling host thread.  Every
 * function below works on the calling thread's current context (the default one if none was set), takes that
 * context's mutex and makes its device current, so calls may come from any thread; calls on DIFFERENT contexts run
 * concurrently.  All kernels of one context share its scratch: the library orders them on the device (each call's
 * stream waits for the previous call of the context), so `_dev` calls on different streams are safe but serialised.
 * The device error word is per context: b381_check_dev(stream) reports (and clears) errors of every `_dev` call
 * enqueued on the context before it.  b381_last_error() is per host thread.
 *
 * Points are NOT validated beyond canonical limbs (as in arkworks, curve / subgroup membership is the caller's
 * business): use b381_g1/g2_deserialize (curve equation), b381_g1/g2_in_subgroup or b381_g1/g2_clear_cofactor on
 * untrusted input before pairing it.
 * There is NO CPU fallback: without a CUDA device every compute call fails with B381_E_CUDA.
 */
#ifndef B381_H
#define B381_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* parity modes */
#define B381_MODE_ARK 0      /* ark_bls12_381::Bls12_381::multi_miller_loop semantics (the truth the reference
                                defers to: src/miller_loop_native_optimized.rs:131-132,151,163) */
#define B381_MODE_ZK 1       /* the loop spelled out at src/miller_loop_native.rs:27-116 with ell (:139-152) wired in */
#define B381_MODE_LITERAL 2  /* the code exactly as written: multi_miller_loop -> 1 (src/miller_loop_native.rs:163-188) */

/* error codes */
#define B381_OK 0
#define B381_E_CUDA (-1)            /* CUDA runtime failure / no device; see b381_last_error() */
#define B381_E_ARG (-2)             /* null pointer, n == 0, bad mode */
#define B381_E_NOT_CANONICAL (-3)   /* an input limb vector is >= p (reference: Fq::from_bigint(..).unwrap() panics) */
#define B381_E_ZERO_DIVISION (-4)   /* final_exponentiation(0) / LITERAL f_den == 0 (reference panics) */
#define B381_E_NOT_INIT (-5)
#define B381_E_BAD_ENCODING (-7)    /* point encoding with inconsistent flag bits */
#define B381_E_NOT_SQUARE (-6)      /* sqrt of a non-residue, or of zero with sgn0 = 1 (reference: x.sqrt().unwrap() / assert_eq! panic) */

/* lifecycle ------------------------------------------------------------------------------------ */
int b381_init(int device);                  /* create the default context on one GPU, allocate scratch */
int b381_shutdown(void);
const char* b381_last_error(void);          /* last error of the calling thread; valid until that thread's next call */
/* further contexts: one per GPU for a host that shards a batch over several GPUs from one process */
typedef struct b381_ctx_s* b381_ctx_t;
int b381_ctx_create(int device, b381_ctx_t* out);
int b381_ctx_destroy(b381_ctx_t ctx);
int b381_ctx_set_current(b381_ctx_t ctx);   /* per host thread; NULL selects the default context again */
int b381_ctx_get_current(b381_ctx_t* out);
int b381_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* scratch_bytes);
/* number of kernels this library has launched since b381_init (bench.py reports it as gpu_launches) */
unsigned long long b381_kernel_launches(void);

/* field-op microbenchmarks (BASELINE config #2) -------------------------------------------------- */
/* out[i] = a[i] * b[i] in Fq  -- replaces ark `Fq * Fq` as used at src/fields/helpers.rs:102-105 */
int b381_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);
/* out[i] = a[i] * b[i]^k (k dependent multiplications held in registers; pure integer-pipe number) */
int b381_fp_mul_chain(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int k);
/* Fq2 product -- ark `Fq2 * Fq2` at src/miller_loop_native.rs:42,50,53 */
int b381_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);
/* Fq12 product in tower order -- ark `Fq12 * Fq12` at src/miller_loop_native_optimized.rs:91-99 */
int b381_fp12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);
/* MyFq12 product in w-basis order -- `impl Mul for MyFq12`, src/fields/helpers.rs:90-152 */
int b381_fp12_mul_wbasis(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n);

/* Miller loop / final exponentiation / pairing ---------------------------------------------------- */
/* batched variant: out[i] = miller_loop(P_i, Q_i), one Fq12 per pair.  Identity pairs give 1. */
int b381_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode);
/* multi_miller_loop(&[(&G1Affine,&G2Affine)]) -> MillerLoopResult, src/miller_loop_native.rs:154-212:
   ONE Fq12 = product over the n terms (ARK/ZK), or 1 (LITERAL, the code as written). */
int b381_multi_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode);
/* final exponentiation f^(3 (p^12-1)/r) -- ark Bls12::final_exponentiation; algorithm spec
   src/fields_as_trees/miller_loop.rs:128-178; the native stub it completes is
   src/miller_loop_native_optimized.rs:104-121 */
int b381_final_exp(const uint32_t* f, uint32_t* out, size_t n);
/* out[i] = final_exp(miller_loop(P_i, Q_i)) fused in one kernel (no HBM round trip of the Miller value) */
int b381_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode);
/* final_exp(multi_miller_loop(...)): the BLS batch-verify shape, one Fq12 out */
int b381_multi_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode);
/* product of n Fq12 values (used to combine per-GPU partial products after an all-gather) */
int b381_fp12_product(const uint32_t* in, uint32_t* out144, size_t n);
/* optimized_miller_loop(G1Projective, G2Projective) -> Fq12, src/miller_loop_native_optimized.rs:81-127,
   exactly as written (LITERAL): only c0.c0 of the output is non-zero. */
int b381_literal_optimized(const uint32_t* g1proj, const uint32_t* g2proj, uint32_t* out, size_t n);

/* G2Prepared: the line coefficients of Q as a cached stage -------------------------------------------
   ark-ec `G2Prepared { ell_coeffs: Vec<(Fp2, Fp2, Fp2)>, infinity }`, the shape the reference's circuit
   side mirrors at src/miller_loop_target.rs:23-76 (G2PreparedTarget, 68 coefficient triples); the
   native loop that would consume it is src/miller_loop_native.rs:154-212.  coeffs holds, per point,
   B381_G2PREP_WORDS = 68 triples x 3 Fq2 x 24 words in the order the Miller loop of `mode` consumes
   them (63 doublings, 5 additions).  A prepared Q is reused across any number of Miller loops. */
#define B381_G2PREP_TRIPLES 68
#define B381_G2PREP_WORDS (68 * 72)
int b381_g2_prepare(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode);
/* out[i] = miller_loop(P_i, prepared Q_i); identical values to b381_miller_loop on (P_i, Q_i).
   inf[i] bit0 = P_i is the identity, bit1 = Q_i is the identity (its coefficients are ignored). */
int b381_miller_loop_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode);
/* out[i] = final_exp(miller_loop(P_i, prepared Q_i)) */
int b381_pairing_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode);

/* Batched witness-generation helpers: the native computations the reference's circuit generators run
   per witness (src/fields/fq_target.rs:243-280, :316-343; fq2_target.rs:320-352, :373-410;
   fq6_target.rs:384-418; fq12_target.rs:340-374; pow_fq src/fields/helpers.rs:176-195).  Element-wise. */
int b381_fp_inv(const uint32_t* a, uint32_t* out, size_t n);                      /* a = 0 -> B381_E_ZERO_DIVISION */
/* sqrt(a) with sgn0_fq(result) == sgn[i] (sgn = NULL: sgn0 = 0); FqSqrtGenerator, fq_target.rs:316-343 */
int b381_fp_sqrt(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n);
/* out[i] = legendre(a) == 1 (fq_target.rs:269-280): zero is not a square by that definition */
int b381_fp_is_square(const uint32_t* a, uint8_t* out, size_t n);
/* pow_fq(a, exp) with one exponent (u64 limbs, little-endian) for the whole batch; as in the reference
   an all-zero exponent returns a */
int b381_fp_pow(const uint32_t* a, const uint64_t* exp, size_t exp_limbs, uint32_t* out, size_t n);
int b381_fp2_inv(const uint32_t* a, uint32_t* out, size_t n);
/* Fq2 sqrt with sgn0_fq2(result) == sgn[i] (src/fields/helpers.rs:169-174) */
int b381_fp2_sqrt(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n);
int b381_fp2_is_square(const uint32_t* a, uint8_t* out, size_t n);
int b381_fp6_inv(const uint32_t* a, uint32_t* out, size_t n);                     /* 72 words per element */
int b381_fp12_inv(const uint32_t* a, uint32_t* out, size_t n);

/* Wire formats ----------------------------------------------------------------------------------------
   (a) the reference's witness format: canonical (non-Montgomery) integers as 12 x 32-bit little-endian
   digits -- `BigUint::from(fq).to_u32_digits()` padded to 12 (src/fields/fq_target.rs:300-313); the
   inverse is from_biguint_to_fq (src/fields/helpers.rs:154-157); Fq12Target::set_witness
   (src/fields/fq12_target.rs:408-416) writes the twelve MyFq12 coefficients (w-basis order,
   src/fields/helpers.rs:39-41) that way: 144 digits per Fq12. */
int b381_fp_to_u32_digits(const uint32_t* a, uint32_t* out, size_t n);
int b381_fp_from_u32_digits(const uint32_t* digits, uint32_t* out, size_t n);   /* digits >= p -> B381_E_NOT_CANONICAL */
int b381_fp12_to_witness_limbs(const uint32_t* f, uint32_t* out, size_t n);
/* (b) ZCash / IETF point encodings as implemented by ark-bls12-381 0.4 (big-endian; byte 0: 0x80 compressed,
   0x40 infinity, 0x20 y lexicographically largest): G1 48 / 96 bytes, G2 96 / 192 bytes (x.c1 first).
   Deserialisation checks the flags, x < p and the curve equation; it does NOT check subgroup membership. */
int b381_g1_deserialize(const uint8_t* in, int compressed, uint32_t* g1, uint8_t* inf, size_t n);
int b381_g1_serialize(const uint32_t* g1, const uint8_t* inf, int compressed, uint8_t* out, size_t n);
int b381_g2_deserialize(const uint8_t* in, int compressed, uint32_t* g2, uint8_t* inf, size_t n);
int b381_g2_serialize(const uint32_t* g2, const uint8_t* inf, int compressed, uint8_t* out, size_t n);
/* out[i] = 1 iff P_i lies in the prime-order subgroup (order r = x^4 - x^2 + 1, decimal at
   src/miller_loop_native_optimized.rs:110): the check untrusted, deserialised points need before they are paired.
   Endomorphism tests of ark-bls12-381 0.4 (eprint 2021/1130): G1 (BETA x, y) == -[x^2] P, G2 psi(P) == [x] P -- two /
   one 64-bit ladders instead of a 255-bit one; same answers as [r] P == identity.  Points must be on the curve. */
int b381_g1_in_subgroup(const uint32_t* g1, const uint8_t* inf, uint8_t* out, size_t n);
int b381_g2_in_subgroup(const uint32_t* g2, const uint8_t* inf, uint8_t* out, size_t n);
/* out[i] = [h_eff] P_i: maps any curve point into the subgroup (ark-bls12-381 0.4 clear_cofactor; the effective
   cofactors of RFC 9380 section 8.8): G1 [1 - x] P, G2 Budroni-Pintore [x^2 - x - 1] P + [x - 1] psi(P) + psi^2(2P). */
int b381_g1_clear_cofactor(const uint32_t* g1, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n);
int b381_g2_clear_cofactor(const uint32_t* g2, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n);
/* out[i] = [k_i] P_i (affine, + identity flag), k_i = 256-bit scalar as 8 little-endian u32 words; the group
   law is the ark-ec Jacobian add / double the reference's native loop uses (`R + R`, `R + Q`,
   src/miller_loop_native_optimized.rs:93,98).  Generates (a_i G1, b_i G2) test points and aggregates keys
   on the device.  Not constant time. */
int b381_g1_scalar_mul(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n);
int b381_g2_scalar_mul(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n);
/* ONE point out: the sum of the n points (public-key aggregation), and sum_i [k_i] P_i (multi-scalar
   multiplication).  G1: bucket method (Pippenger) over the base field -- window digits, counting sort of the point
   indices by bucket, one thread per bucket, running-sum reduction.  G2: n scalar multiplications + a 16-ary
   reduction tree.  n < 2^32. */
int b381_g1_sum(const uint32_t* g1, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n);
int b381_g2_sum(const uint32_t* g2, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n);
int b381_g1_msm(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n);
int b381_g2_msm(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n);

/* device-pointer variants (inputs already resident in HBM; used for the kernel-only throughput) ---- */
int b381_miller_loop_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, void* stream);
int b381_final_exp_dev(const uint32_t* f, uint32_t* out, size_t n, void* stream);
int b381_pairing_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, void* stream);
int b381_multi_miller_loop_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode, void* stream);
int b381_fp_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream);
int b381_fp_mul_chain_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int k, void* stream);
int b381_fp2_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream);
int b381_fp12_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream);
int b381_g2_prepare_dev(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode, void* stream);
int b381_miller_loop_prepared_dev(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode, int final_exp, void* stream);
/* G2Prepared in the library's own ("packed") layout (the cached stage of src/miller_loop_target.rs:23-76 /
   ark-ec G2Prepared again, consumed by the loop of src/miller_loop_native.rs:154-212): the same 68 triples per Q, kept in the internal number format and
   interleaved over tiles of 256 points so that a warp reads 512 contiguous bytes per access; the Miller loop multiplies
   straight out of the buffer (no conversion per line, no copy).  Device memory only, 16-byte aligned,
   b381_g2_packed_words(n) 32-bit words (n rounded up to a tile); opaque -- valid for this library build and `mode`.
   Values of b381_miller_loop_packed_dev are identical to b381_miller_loop on (P_i, Q_i). */
size_t b381_g2_packed_words(size_t n);
int b381_g2_prepare_packed_dev(const uint32_t* g2, uint32_t* packed, size_t n, int mode, void* stream);
int b381_miller_loop_packed_dev(const uint32_t* g1, const uint32_t* packed, const uint8_t* inf, uint32_t* out, size_t n, int mode, int final_exp, void* stream);
/* ONE cached Q against n points P_i: out[i] = miller_loop(P_i, Q) (final_exp: the pairing), Q = point 0 of `packed_one`
   (b381_g2_prepare_packed_dev with n = 1).  Every thread reads the same lines (broadcast loads): no per-pair coefficient
   memory, any batch size.  inf[i] bit0 = P_i is the identity. */
int b381_miller_loop_packed_one_dev(const uint32_t* g1, const uint32_t* packed_one, const uint8_t* inf, uint32_t* out, size_t n, int mode, int final_exp, void* stream);
/* multi_miller_loop (src/miller_loop_native.rs:154-212; ark Bls12::multi_miller_loop) against cached Q's:
   out144 = product over i of miller_loop(P_i, packed Q_i) (ARK mode: `packed` must come from b381_g2_prepare_packed_dev with
   B381_MODE_ARK), optionally followed by the final exponentiation: the BLS batch-verify shape against cached public keys.
   Four pairs per thread share every squaring of f.  Same value as b381_multi_miller_loop / b381_multi_pairing. */
int b381_multi_miller_loop_packed_dev(const uint32_t* g1, const uint32_t* packed, const uint8_t* inf, uint32_t* out144, size_t n, int final_exp, void* stream);
int b381_multi_pairing_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode, void* stream);
int b381_fp12_mul_wbasis_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream);
int b381_fp_inv_dev(const uint32_t* a, uint32_t* out, size_t n, void* stream);
int b381_fp_sqrt_dev(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n, void* stream);
int b381_fp_is_square_dev(const uint32_t* a, uint8_t* out, size_t n, void* stream);
int b381_fp_pow_dev(const uint32_t* a, const uint64_t* exp /* host */, size_t exp_limbs, uint32_t* out, size_t n, void* stream);
int b381_fp2_inv_dev(const uint32_t* a, uint32_t* out, size_t n, void* stream);
int b381_fp2_sqrt_dev(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n, void* stream);
int b381_fp2_is_square_dev(const uint32_t* a, uint8_t* out, size_t n, void* stream);
int b381_fp6_inv_dev(const uint32_t* a, uint32_t* out, size_t n, void* stream);
int b381_fp12_inv_dev(const uint32_t* a, uint32_t* out, size_t n, void* stream);
int b381_fp_to_u32_digits_dev(const uint32_t* a, uint32_t* out, size_t n, void* stream);
int b381_fp_from_u32_digits_dev(const uint32_t* digits, uint32_t* out, size_t n, void* stream);
int b381_fp12_to_witness_limbs_dev(const uint32_t* f, uint32_t* out, size_t n, void* stream);
int b381_g1_deserialize_dev(const uint8_t* in, int compressed, uint32_t* g1, uint8_t* inf, size_t n, void* stream);
int b381_g1_serialize_dev(const uint32_t* g1, const uint8_t* inf, int compressed, uint8_t* out, size_t n, void* stream);
int b381_g2_deserialize_dev(const uint8_t* in, int compressed, uint32_t* g2, uint8_t* inf, size_t n, void* stream);
int b381_g2_serialize_dev(const uint32_t* g2, const uint8_t* inf, int compressed, uint8_t* out, size_t n, void* stream);
int b381_g1_in_subgroup_dev(const uint32_t* g1, const uint8_t* inf, uint8_t* out, size_t n, void* stream);
int b381_g2_in_subgroup_dev(const uint32_t* g2, const uint8_t* inf, uint8_t* out, size_t n, void* stream);
int b381_g1_clear_cofactor_dev(const uint32_t* g1, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g2_clear_cofactor_dev(const uint32_t* g2, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g1_scalar_mul_dev(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g2_scalar_mul_dev(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g1_sum_dev(const uint32_t* g1, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g2_sum_dev(const uint32_t* g2, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g1_msm_dev(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
int b381_g2_msm_dev(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n, void* stream);
/* fetch-and-clear the context's device error word after synchronising `stream`; returns 0 or a B381_E_* code */
int b381_check_dev(void* stream);

/* integer-multiply roofline probe: sustained IMAD.WIDE issue rate of this GPU, in 1e9 thread-instructions/s,
   and the SM clock (MHz) seen while it ran (the denominator of roofline.frac) */
int b381_imad_peak(double* imad_wide_ginst_per_s, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* B381_H */
