"""Golden values that live in the reference tree (data, not code).

ERASURE_SELFTEST: cmd/erasure-coding.go:160 — xxhash64(seed 0) over byte(i)||shard_i for the
encoding of bytes 0..255 with (data, parity), for total=4..15, data=total/2..total-1.
BITROT_SELFTEST: cmd/bitrot.go:225-229.
"""
ERASURE_SELFTEST = {
    (2, 2): 0x23fb21be2496f5d3, (2, 3): 0xa5cd5600ba0d8e7c, (3, 1): 0x60ab052148b010b4, (3, 2): 0xe64927daef76435a,
    (3, 3): 0x672f6f242b227b21, (3, 4): 0x571e41ba23a6dc6, (4, 1): 0x524eaa814d5d86e2, (4, 2): 0x62b9552945504fef,
    (4, 3): 0xcbf9065ee053e518, (4, 4): 0x9a07581dcd03da8, (4, 5): 0xbf2d27b55370113f, (5, 1): 0xf71031a01d70daf,
    (5, 2): 0x8e5845859939d0f4, (5, 3): 0x7ad9161acbb4c325, (5, 4): 0xc446b88830b4f800, (5, 5): 0xabf1573cc6f76165,
    (5, 6): 0x7b5598a85045bfb8, (6, 1): 0xe2fc1e677cc7d872, (6, 2): 0x7ed133de5ca6a58e, (6, 3): 0x39ef92d0a74cc3c0,
    (6, 4): 0xcfc90052bc25d20, (6, 5): 0x71c96f6baeef9c58, (6, 6): 0x4b79056484883e4c, (6, 7): 0xb1a0e2427ac2dc1a,
    (7, 1): 0x937ba2b7af467a22, (7, 2): 0x5fd13a734d27d37a, (7, 3): 0x3be2722d9b66912f, (7, 4): 0x14c628e59011be3d,
    (7, 5): 0xcc3b39ad4c083b9f, (7, 6): 0x45af361b7de7a4ff, (7, 7): 0x456cc320cec8a6e6, (7, 8): 0x1867a9f4db315b5c,
    (8, 1): 0xbc5756b9a9ade030, (8, 2): 0xdfd7d9d0b3e36503, (8, 3): 0x72bb72c2cdbcf99d, (8, 4): 0x3ba5e9b41bf07f0,
    (8, 5): 0xd7dabc15800f9d41, (8, 6): 0xb482a6169fd270f, (8, 7): 0x50748e0099d657e8, (9, 1): 0xc77ae0144fcaeb6e,
    (9, 2): 0x8a86c7dbebf27b68, (9, 3): 0xa64e3be6d6fe7e92, (9, 4): 0x239b71c41745d207, (9, 5): 0x2d0803094c5a86ce,
    (9, 6): 0xa3c2539b3af84874, (10, 1): 0x7d30d91b89fcec21, (10, 2): 0xfa5af9aa9f1857a3, (10, 3): 0x84bc4bda8af81f90,
    (10, 4): 0x6c1cba8631de994a, (10, 5): 0x4383e58a086cc1ac, (11, 1): 0x4ed2929a2df690b, (11, 2): 0xecd6f1b1399775c0,
    (11, 3): 0xc78cfbfc0dc64d01, (11, 4): 0xb2643390973702d6, (12, 1): 0x3b2a88686122d082, (12, 2): 0xfd2f30a48a8e2e9,
    (12, 3): 0xd5ce58368ae90b13, (13, 1): 0x9c88e2a9d1b8fff8, (13, 2): 0xcb8460aa4cf6613, (14, 1): 0x78a28bbaec57996e,
}
assert len(ERASURE_SELFTEST) == 60

BITROT_SELFTEST = {  # algo id -> hex digest (cmd/bitrot.go:225-229)
    1: "a7677ff19e0182e4d52e3a3db727804abc82a5818749336369552e54b838b004",
    4: "e519b7d84b1c3c917985f544773a35cf265dcab10948be3550320d156bab612124a5ae2ae5a8c73c0eea360f68b0e28136f26e858756dbfe7375a7389f26c669",
    2: "39c0407ed3f01b18d22c85db4aeff11e060ca5f43131b0126731ca197cd42313",
    3: "39c0407ed3f01b18d22c85db4aeff11e060ca5f43131b0126731ca197cd42313",
}
MAGIC_KEY_HEX = "4be734fa8e238acd263e83e6bb968552040f935da39f441497e09d1322de36a0"  # cmd/bitrot.go:37
PI_100 = "1415926535897932384626433832795028841971693993751058209749445923078164062862089986280348253421170679"
INLINE_NOTINLINE_MD5 = "fffb6377948ebea75ad2b8058e849ef5"  # cmd/erasure-object_test.go:1177

# DERIVED (not in the reference tree; produced by the oracle after it passed every golden above;
# SURVEY.md §8c "derived KATs") — self-test-style xxh64 for configs MinIO's self-test does not reach.
DERIVED_SELFTEST = {(12, 4): 0x8c7623c5dc637f21, (16, 4): 0xcd7ece64401dc841, (8, 8): 0xf7c3766772a1198a}
DERIVED_PARITY_ROWS = {
    (4, 2): ["1b1c1214", "1c1b1412"],
    (12, 4): ["afb4968cf5e8c4d81b1c1214", "b4af8c96e8f5d8c41c1b1412", "968cafb4c4d8f5e812141b1c", "8c96b4afd8c4e8f514121c1b"],
    (16, 4): ["21b5f685df02b7873edd4aa48dda6130", "b52185f602df87b7dd3ea44ada8d3061",
              "f68521b5b787df024aa43edd61308dda", "85f6b52187b702dfa44add3e3061da8d"],
}
DERIVED_HH = {  # HH256(magic key, pat(n)), pat(n)[i] = (7*i+3) & 0xff
    0: "5e76d207cf4ab20866fdc03c83e8a0f4e8f458e880777956ec0bae4e9f23f6c5",
    22: "1a0d02c78ed9dbfdc8a3e6ff822c3f9552604c3121c8f61726dcc19b5858de85",
    87382: "1cbbd6e468d701a514c90e7989c9742a5c561632ee44cfc8f1a2a9da5efa07ee",
    65536: "c9c9b4cf2349f3dc0866db98ff703ab3c9409b851bfd37bde5d41dc1591c3f7f",
    262144: "17a1aa676b0026c123d8ec63157bee5586fa3b4c495794c3f20da05f06395087",
}
