"""world_size-2 gloo worker for tests/test_multigpu_cpu.py: shards windows, fabricates each rank's records as a
pure function of the GLOBAL window index, all-gathers them and checks the assembled table."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch.distributed as dist  # noqa: E402

from genomics_general_b200 import multigpu  # noqa: E402


def records(idx, C):
    return np.stack([np.sin(idx * (c + 1.0)) + idx for c in range(C)], axis=1) if len(idx) else np.zeros((0, C))


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(5)
    for W in (0, 1, 2, 7, 101):
        lo = np.sort(rng.integers(0, 100000, W)).astype(np.int64)
        hi = lo + rng.integers(0, 5000, W)
        shards = multigpu.shard_windows(lo, hi, world)
        assert shards[0][0] == 0 and shards[-1][1] == W
        assert all(shards[r][1] == shards[r + 1][0] for r in range(world - 1))
        b, e = shards[rank]
        C = 11
        full = multigpu.all_gather_rows(records(np.arange(b, e, dtype=np.float64), C), [s[1] - s[0] for s in shards])
        assert full.shape == (W, C), full.shape
        assert np.array_equal(full, records(np.arange(W, dtype=np.float64), C))
        if e > b:
            s0, s1 = multigpu.shard_site_range(lo, hi, b, e)
            assert s0 == lo[b:e].min() and s1 == hi[b:e].max()
    # pack / unpack of popgen records
    P = 3
    res = dict(sites=np.arange(4), pos_sum=np.arange(4) * 10 ** 12, path=np.array([0, 1, 2, 1], dtype=np.int32),
               pi=np.random.rand(4, 3), dxy=np.random.rand(4, 3), fst=np.full((4, 3), np.nan))
    back = multigpu.unpack_popgen_records(multigpu.popgen_records(res), P)
    for k in res:
        assert np.array_equal(np.asarray(res[k], dtype=np.float64), np.asarray(back[k], dtype=np.float64), equal_nan=True), k
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK")


if __name__ == "__main__":
    main()
