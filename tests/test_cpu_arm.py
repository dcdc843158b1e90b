"""The CPU arm of bench.py (oracle/simd.c persistent pool) must produce exactly the oracle's bytes: it is the baseline the
GPU path is compared with, so it has to be doing the same work."""
import numpy as np


def test_pool_matches_scalar_oracle(oracle):
    L = oracle.lib()
    k, m, bs, nb = 12, 4, 1 << 20, 12
    S = -(-bs // k)
    src = np.empty(nb * bs, dtype=np.uint8)
    par = np.empty(nb * m * S, dtype=np.uint8)
    dig = np.empty(nb * (k + m) * 32, dtype=np.uint8)
    pool = L.orc_pool_new(5)
    L.orc_pool_fill(pool, k, m, bs, src.ctypes.data, nb, par.ctypes.data, dig.ctypes.data, 1234)
    assert not par.any() and len(np.unique(src[:4096])) > 200          # outputs cleared, source is not constant
    sec = L.orc_pool_encode_hash(pool, k, m, bs, src.ctypes.data, nb, par.ctypes.data, dig.ctypes.data, 2)
    assert sec > 0
    for b in (0, 5, nb - 1):
        sh = oracle.encode_data(k, m, src[b * bs:(b + 1) * bs])        # scalar restatement
        for j in range(m):
            assert np.array_equal(par[(b * m + j) * S:(b * m + j + 1) * S], sh[k + j])
        for i in range(k + m):
            assert dig[(b * (k + m) + i) * 32:(b * (k + m) + i + 1) * 32].tobytes() == oracle.hh256(sh[i])
    # a second run on the same pool (threads persist between runs) gives the same bytes
    ref = par.copy()
    L.orc_pool_encode_hash(pool, k, m, bs, src.ctypes.data, nb, par.ctypes.data, dig.ctypes.data, 1)
    assert np.array_equal(ref, par)
    L.orc_pool_free(pool)
