"""Mirror of /root/reference/src/global_constants.rs:1-8."""
LOG_ATE_LOOP_COUNT = 62
ATE_LOOP_COUNT = 15132376222941642752
PSEUDO_BINARY_ENCODING = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1,
]
BLS_X = 0xD201_0000_0001_0000
BLS_X_IS_NEGATIVE = True
