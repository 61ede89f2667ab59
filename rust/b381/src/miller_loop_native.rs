//! Drop-in for the reference's src/miller_loop_native.rs: same names, GPU-backed.
use crate::ffi::{check, push_g1, push_g2, read_fq12, B381Error};
use ark_bls12_381::{Fq12, G1Affine, G2Affine};
use ark_std::One;

#[derive(Copy, Clone, Debug)]
pub struct MillerLoopResult(pub Fq12);            // src/miller_loop_native.rs:7-8

impl Default for MillerLoopResult {
    fn default() -> Self { MillerLoopResult(Fq12::one()) }     // :10-14
}

#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Mode { Ark = 0, Zk = 1, Literal = 2 }

fn marshal(terms: &[(&G1Affine, &G2Affine)]) -> (Vec<u32>, Vec<u32>, Vec<u8>) {
    let (mut g1, mut g2, mut inf) = (Vec::with_capacity(terms.len() * 24), Vec::with_capacity(terms.len() * 48), Vec::with_capacity(terms.len()));
    for (p, q) in terms {
        push_g1(&mut g1, p);
        push_g2(&mut g2, q);
        inf.push(p.infinity as u8 | ((q.infinity as u8) << 1));
    }
    (g1, g2, inf)
}

/// `multi_miller_loop(&[(&G1Affine,&G2Affine)]) -> MillerLoopResult` (src/miller_loop_native.rs:154).
/// `Mode::Literal` reproduces the file as written (returns 1); `Mode::Ark` is what the reference
/// treats as truth.
pub fn multi_miller_loop_mode(terms: &[(&G1Affine, &G2Affine)], mode: Mode) -> Result<MillerLoopResult, B381Error> {
    if terms.is_empty() { return Ok(MillerLoopResult::default()); }
    let (g1, g2, inf) = marshal(terms);
    let mut out = [0u32; 144];
    check(unsafe { b381_sys::b381_multi_miller_loop(g1.as_ptr(), g2.as_ptr(), inf.as_ptr(), out.as_mut_ptr(), terms.len(), mode as i32) })?;
    Ok(MillerLoopResult(read_fq12(&out)))
}

pub fn multi_miller_loop(terms: &[(&G1Affine, &G2Affine)]) -> MillerLoopResult {
    multi_miller_loop_mode(terms, Mode::Ark).expect("b381_multi_miller_loop")
}

/// batched variant: one Miller value per pair
pub fn miller_loop_batch(terms: &[(&G1Affine, &G2Affine)], mode: Mode) -> Result<Vec<MillerLoopResult>, B381Error> {
    let (g1, g2, inf) = marshal(terms);
    let mut out = vec![0u32; 144 * terms.len()];
    check(unsafe { b381_sys::b381_miller_loop(g1.as_ptr(), g2.as_ptr(), inf.as_ptr(), out.as_mut_ptr(), terms.len(), mode as i32) })?;
    Ok(out.chunks_exact(144).map(|w| MillerLoopResult(read_fq12(w))).collect())
}

/// full pairings e(P_i, Q_i), Miller loop and final exponentiation fused on the GPU
pub fn pairing_batch(terms: &[(&G1Affine, &G2Affine)]) -> Result<Vec<Fq12>, B381Error> {
    let (g1, g2, inf) = marshal(terms);
    let mut out = vec![0u32; 144 * terms.len()];
    check(unsafe { b381_sys::b381_pairing(g1.as_ptr(), g2.as_ptr(), inf.as_ptr(), out.as_mut_ptr(), terms.len(), Mode::Ark as i32) })?;
    Ok(out.chunks_exact(144).map(read_fq12).collect())
}

impl MillerLoopResult {
    /// f^(3 (p^12-1)/r); algorithm spec: src/fields_as_trees/miller_loop.rs:128-178
    pub fn final_exponentiation(&self) -> Result<Fq12, B381Error> {
        let mut a = Vec::with_capacity(144);
        crate::ffi::push_fq12(&mut a, &self.0);
        let mut out = [0u32; 144];
        check(unsafe { b381_sys::b381_final_exp(a.as_ptr(), out.as_mut_ptr(), 1) })?;
        Ok(read_fq12(&out))
    }
}

/// `G2Prepared` as a cached stage (ark-ec `G2Prepared`; the shape of `G2PreparedTarget`,
/// src/miller_loop_target.rs:23-76): 68 coefficient triples per point in the C-ABI layout.
pub struct G2Prepared { pub coeffs: Vec<u32>, pub infinity: bool, pub mode: Mode }

impl G2Prepared {
    pub fn from_affine(q: &G2Affine, mode: Mode) -> Result<Self, B381Error> {
        let mut g2 = Vec::with_capacity(48);
        push_g2(&mut g2, q);
        let mut coeffs = vec![0u32; 68 * 72];
        if !q.infinity {
            check(unsafe { b381_sys::b381_g2_prepare(g2.as_ptr(), coeffs.as_mut_ptr(), 1, mode as i32) })?;
        }
        Ok(G2Prepared { coeffs, infinity: q.infinity, mode })
    }
}

/// one Miller value per (P, prepared Q); identical values to `miller_loop_batch` on (P, Q)
pub fn miller_loop_prepared_batch(terms: &[(&G1Affine, &G2Prepared)]) -> Result<Vec<MillerLoopResult>, B381Error> {
    let mode = terms.first().map(|t| t.1.mode).unwrap_or(Mode::Ark);
    let (mut g1, mut co, mut inf) = (Vec::new(), Vec::new(), Vec::new());
    for (p, q) in terms {
        push_g1(&mut g1, p);
        co.extend_from_slice(&q.coeffs);
        inf.push(p.infinity as u8 | ((q.infinity as u8) << 1));
    }
    let mut out = vec![0u32; 144 * terms.len()];
    check(unsafe { b381_sys::b381_miller_loop_prepared(g1.as_ptr(), co.as_ptr(), inf.as_ptr(), out.as_mut_ptr(), terms.len(), mode as i32) })?;
    Ok(out.chunks_exact(144).map(|w| MillerLoopResult(read_fq12(w))).collect())
}
