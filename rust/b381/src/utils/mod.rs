pub mod constants;
pub use crate::fields::helpers;
