"""N>1 host-side logic on CPU: world-size-2 gloo run of the partitioning used by `bench.py --gpus N`
(one contiguous block range per rank, no data-path collective; results combined by all-gather of digests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minio_b200.partition import partition_blocks, partition_objects


def test_partition_blocks_properties():
    for n in [0, 1, 7, 8, 10240, 12345]:
        for world in [1, 2, 3, 4, 8]:
            seen = []
            for r in range(world):
                s, c = partition_blocks(n, world, r)
                seen.extend(range(s, s + c))
                assert c in (n // world, n // world + 1)
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        partition_blocks(10, 2, 2)


def test_partition_objects_balanced():
    sizes = [64] * 1024                      # BASELINE config 4: 1024 objects x 64 MiB over 8 GPUs
    parts = partition_objects(sizes, 8)
    assert sorted(sum(parts, [])) == list(range(1024)) and all(len(p) == 128 for p in parts)
    rng = np.random.default_rng(0)
    sizes = [int(x) for x in rng.integers(1, 1000, 200)]
    parts = partition_objects(sizes, 4)
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, m, bs, nblocks = 4, 2, 4096, 13
    data = np.random.default_rng(5).integers(0, 256, nblocks * bs, dtype=np.uint8)   # same stream on every rank
    start, count = partition_blocks(nblocks, world, rank)
    dig = torch.zeros((nblocks, k + m, 32), dtype=torch.uint8)
    for b in range(start, start + count):                    # each rank encodes+hashes only its own blocks
        sh = o.encode_data(k, m, data[b * bs:(b + 1) * bs])
        for i in range(k + m):
            dig[b, i] = torch.frombuffer(bytearray(o.hh256(sh[i])), dtype=torch.uint8)
    dist.all_reduce(dig, op=dist.ReduceOp.SUM)               # ranges are disjoint: sum == concatenation
    t = torch.tensor([float(count)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # the max-over-ranks reduction bench.py uses for timing
    if rank == 0:
        q.put((dig.numpy().tobytes(), float(t.item())))
    dist.destroy_process_group()


def test_world2_gloo_partitioned_encode(oracle):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, m, bs, nblocks = 4, 2, 4096, 13
    data = np.random.default_rng(5).integers(0, 256, nblocks * bs, dtype=np.uint8)
    want = b""
    for b in range(nblocks):
        sh = oracle.encode_data(k, m, data[b * bs:(b + 1) * bs])
        want += b"".join(oracle.hh256(x) for x in sh)
    assert got == want and mx == 7.0
