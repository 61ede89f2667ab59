pub mod helpers;
pub mod my_fq6;
