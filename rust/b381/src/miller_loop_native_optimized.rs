//! Drop-in for src/miller_loop_native_optimized.rs:81-127 (the function exactly as written, batched).
use crate::ffi::{check, push_fq, push_fq2, read_fq12, B381Error};
use ark_bls12_381::{Fq12, G1Projective, G2Projective};

pub fn optimized_miller_loop_batch(pairs: &[(G1Projective, G2Projective)]) -> Result<Vec<Fq12>, B381Error> {
    let (mut g1, mut g2) = (Vec::new(), Vec::new());
    for (p, q) in pairs {
        push_fq(&mut g1, &p.x); push_fq(&mut g1, &p.y); push_fq(&mut g1, &p.z);
        push_fq2(&mut g2, &q.x); push_fq2(&mut g2, &q.y); push_fq2(&mut g2, &q.z);
    }
    let mut out = vec![0u32; 144 * pairs.len()];
    check(unsafe { b381_sys::b381_literal_optimized(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), pairs.len()) })?;
    Ok(out.chunks_exact(144).map(read_fq12).collect())
}

#[allow(non_snake_case)]
pub fn optimized_miller_loop(P: G1Projective, Q: G2Projective) -> Fq12 {
    optimized_miller_loop_batch(&[(P, Q)]).expect("b381_literal_optimized (the reference panics when f_den == 0)")[0]
}
