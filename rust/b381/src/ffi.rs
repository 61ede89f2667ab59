//! Marshalling between arkworks values and the flat buffers of include/b381.h.
//! ark-ff 0.4 `Fp384` stores `BigInt<6>` = 6 LE u64 limbs in Montgomery form (R = 2^384): that is
//! exactly 12 LE u32 words on a little-endian host, i.e. the wire format.
use ark_bls12_381::{Fq, Fq12, Fq2, Fq6, G1Affine, G2Affine};
use std::ffi::CStr;

#[derive(Debug)]
pub struct B381Error { pub code: i32, pub message: String }

pub fn check(rc: i32) -> Result<(), B381Error> {
    if rc == 0 { return Ok(()); }
    let message = unsafe { CStr::from_ptr(b381_sys::b381_last_error()) }.to_string_lossy().into_owned();
    Err(B381Error { code: rc, message })
}

pub fn init(device: i32) -> Result<(), B381Error> { check(unsafe { b381_sys::b381_init(device) }) }

#[inline]
pub fn push_fq(out: &mut Vec<u32>, x: &Fq) {
    for limb in x.0 .0.iter() {            // Montgomery-form limbs, no conversion
        out.push(*limb as u32);
        out.push((*limb >> 32) as u32);
    }
}
pub fn push_fq2(out: &mut Vec<u32>, x: &Fq2) { push_fq(out, &x.c0); push_fq(out, &x.c1); }
pub fn push_fq12(out: &mut Vec<u32>, x: &Fq12) {
    for f6 in [&x.c0, &x.c1] { for f2 in [&f6.c0, &f6.c1, &f6.c2] { push_fq2(out, f2); } }
}
pub fn push_g1(out: &mut Vec<u32>, p: &G1Affine) { push_fq(out, &p.x); push_fq(out, &p.y); }
pub fn push_g2(out: &mut Vec<u32>, q: &G2Affine) { push_fq2(out, &q.x); push_fq2(out, &q.y); }

#[inline]
pub fn read_fq(w: &[u32]) -> Fq {
    let mut l = [0u64; 6];
    for i in 0..6 { l[i] = w[2 * i] as u64 | ((w[2 * i + 1] as u64) << 32); }
    ark_ff::Fp(ark_ff::BigInt(l), core::marker::PhantomData)     // limbs are already Montgomery form
}
pub fn read_fq2(w: &[u32]) -> Fq2 { Fq2::new(read_fq(&w[0..12]), read_fq(&w[12..24])) }
pub fn read_fq12(w: &[u32]) -> Fq12 {
    let f2 = |i: usize| read_fq2(&w[24 * i..24 * i + 24]);
    Fq12::new(Fq6::new(f2(0), f2(1), f2(2)), Fq6::new(f2(3), f2(4), f2(5)))
}
