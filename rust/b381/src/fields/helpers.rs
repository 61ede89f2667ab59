// mirror of the reference's src/fields/helpers.rs: MyFq12 with GPU-backed multiplication
use crate::ffi::{check, push_fq, read_fq, B381Error};
use ark_bls12_381::{Fq, Fq12, Fq2, Fq6};
use std::ops::{Add, Mul};

#[derive(Debug, Clone, Copy, PartialEq)]
pub struct MyFq12 { pub coeffs: [Fq; 12] }

impl From<Fq12> for MyFq12 {              // helpers.rs:14-44
    fn from(f: Fq12) -> Self {
        let (a, b) = (f.c0, f.c1);
        Self { coeffs: [a.c0.c0, b.c0.c0, a.c1.c0, b.c1.c0, a.c2.c0, b.c2.c0, a.c0.c1, b.c0.c1, a.c1.c1, b.c1.c1, a.c2.c1, b.c2.c1] }
    }
}
impl From<MyFq12> for Fq12 {              // helpers.rs:47-76
    fn from(m: MyFq12) -> Self {
        let c = m.coeffs;
        Fq12::new(Fq6::new(Fq2::new(c[0], c[6]), Fq2::new(c[2], c[8]), Fq2::new(c[4], c[10])),
                  Fq6::new(Fq2::new(c[1], c[7]), Fq2::new(c[3], c[9]), Fq2::new(c[5], c[11])))
    }
}
impl Add for MyFq12 {
    type Output = Self;
    fn add(self, rhs: Self) -> Self { let mut c = self.coeffs; for i in 0..12 { c[i] += rhs.coeffs[i]; } Self { coeffs: c } }
}

/// batched w-basis products on the GPU (helpers.rs:90-152)
pub fn myfq12_mul_batch(a: &[MyFq12], b: &[MyFq12]) -> Result<Vec<MyFq12>, B381Error> {
    assert_eq!(a.len(), b.len());
    let (mut wa, mut wb) = (Vec::with_capacity(a.len() * 144), Vec::with_capacity(a.len() * 144));
    for (x, y) in a.iter().zip(b) { for c in &x.coeffs { push_fq(&mut wa, c); } for c in &y.coeffs { push_fq(&mut wb, c); } }
    let mut out = vec![0u32; a.len() * 144];
    check(unsafe { b381_sys::b381_fp12_mul_wbasis(wa.as_ptr(), wb.as_ptr(), out.as_mut_ptr(), a.len()) })?;
    Ok(out.chunks_exact(144).map(|w| { let mut c = [Fq::from(0u64); 12]; for i in 0..12 { c[i] = read_fq(&w[12 * i..12 * i + 12]); } MyFq12 { coeffs: c } }).collect())
}
impl Mul for MyFq12 {
    type Output = Self;
    fn mul(self, rhs: Self) -> Self { myfq12_mul_batch(&[self], &[rhs]).expect("b381_fp12_mul_wbasis")[0] }
}
