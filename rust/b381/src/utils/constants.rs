// mirror of the reference's src/utils/constants.rs:1-2
pub const BLS_X: u64 = 0xd201_0000_0001_0000;
pub const BLS_X_IS_NEGATIVE: bool = true;
