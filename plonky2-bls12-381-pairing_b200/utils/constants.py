"""Mirror of /root/reference/src/utils/constants.rs:1-2."""
BLS_X = 0xD201_0000_0001_0000
BLS_X_IS_NEGATIVE = True
