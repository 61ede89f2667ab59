// mirror of the reference's src/fields/my_fq6.rs:6-56 (host-side permutation type)
use ark_bls12_381::{Fq, Fq2, Fq6};
use std::ops::Add;

#[derive(Debug, Clone, Copy, PartialEq)]
pub struct MyFq6 { pub coeffs: [Fq; 6] }

impl From<Fq6> for MyFq6 {
    fn from(f: Fq6) -> Self { Self { coeffs: [f.c0.c0, f.c1.c0, f.c2.c0, f.c0.c1, f.c1.c1, f.c2.c1] } }
}
impl From<MyFq6> for Fq6 {
    fn from(m: MyFq6) -> Self {
        let c = m.coeffs;
        Fq6::new(Fq2::new(c[0], c[3]), Fq2::new(c[1], c[4]), Fq2::new(c[2], c[5]))
    }
}
impl Add for MyFq6 {
    type Output = Self;
    fn add(self, rhs: Self) -> Self { let mut c = self.coeffs; for i in 0..6 { c[i] += rhs.coeffs[i]; } Self { coeffs: c } }
}
