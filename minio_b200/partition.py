"""Multi-GPU partitioning of the erasure hot path.

Erasure blocks (and whole objects / erasure sets) are independent units: there is no exchange step, so
each rank owns a contiguous range and no data-path collective is needed.  Whole-file bitrot algorithms
chain a hash along one shard file, so the unit handed to a rank is never smaller than one object.
"""


def partition_blocks(nblocks, world, rank):
    """Contiguous, balanced [start, start+count) block range of `rank` (ranks differ by at most one block)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(nblocks, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def partition_objects(sizes, world):
    """Greedy longest-first assignment of whole objects to ranks (BASELINE config 4: one erasure set per GPU).
    Returns a list of index lists, one per rank."""
    order = sorted(range(len(sizes)), key=lambda i: -sizes[i])
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: load[j])
        out[r].append(i)
        load[r] += sizes[i]
    for lst in out:
        lst.sort()
    return out
