"""Hypothesis-driven random (k, m, block size, length, erasure pattern) round trips: GPU vs oracle, bit-exact."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import minio_b200
    return minio_b200


@st.composite
def cases(draw):
    k = draw(st.integers(1, 16))
    m = draw(st.integers(0, min(8, 16 - k) if k < 16 else 0))
    bs = draw(st.sampled_from([64, 1000, 4096, 65536, 1 << 20, (1 << 20) - 1, 12345]))
    nblk = draw(st.integers(0, 5))
    tail = draw(st.integers(0, bs - 1))
    length = min(nblk * bs + tail, 6 << 20)
    seed = draw(st.integers(0, 2**31))
    nerase = draw(st.integers(0, m))
    erase = draw(st.permutations(list(range(k + m)))) [:nerase]
    return k, m, bs, length, seed, sorted(erase)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(cases())
def test_random_roundtrip(mb, oracle, case):
    k, m, bs, length, seed, erase = case
    data = np.random.default_rng(seed).integers(0, 256, length, dtype=np.uint8)
    c = mb.Codec(k, m, bs)
    files = c.encode(data)
    want, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    for i in range(k + m):
        assert np.array_equal(files[i], want[i]), (case, i)
    if length:
        offl = [None if i in erase else files[i] for i in range(k + m)]
        out, hint = c.decode(offl, 0, length, length)
        assert np.array_equal(out, data) and hint == 0
        if erase:
            healed = c.heal(offl, [i in erase for i in range(k + m)], length)
            for i in erase:
                assert np.array_equal(healed[i], files[i]), (case, i)
        assert c.bitrot_verify(files[0], c.shard_file_size(length)) == 0
    c.close()
