//! Same module tree as the reference's `src/lib.rs:1-9` for the native path.
pub mod ffi;
pub mod fields;
pub mod global_constants;
pub mod miller_loop_native;
pub mod miller_loop_native_optimized;
pub mod utils;
