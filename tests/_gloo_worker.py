"""world_size-2 gloo worker for tests/test_multigpu_cpu.py: shards windows, fabricates each rank's records as a
pure function of the GLOBAL window index, all-gathers them and checks the assembled table."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch.distributed as dist  # noqa: E402

from genomics_general_b200 import multigpu  # noqa: E402


def all_gather_rows(local_rows, counts, device=None):
    """All-gather per-window records (float64 [W_local, C]) from every rank into [W_total, C], in rank order, with ONE
    torch.distributed.all_gather on a padded buffer (gloo here; the product issues the same exchange through its own
    native NCCL call, pg_*_allgather)."""
    import torch
    world = dist.get_world_size()
    C = local_rows.shape[1]
    wmax = max(int(c) for c in counts) if len(counts) else 0
    buf = torch.zeros((max(wmax, 1), C), dtype=torch.float64, device=device)
    if local_rows.shape[0]:
        buf[: local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return multigpu.gathered_rows(torch.stack(out).reshape(world * max(wmax, 1), C).numpy(), counts, max(wmax, 1)) \
        if sum(int(c) for c in counts) else np.zeros((0, C))


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
        full = all_gather_rows(records(np.arange(b, e, dtype=np.float64), C), [s[1] - s[0] for s in shards])
        assert full.shape == (W, C), full.shape
        assert np.array_equal(full, records(np.arange(W, dtype=np.float64), C))
        if e > b:
            s0, s1 = multigpu.shard_site_range(lo, hi, b, e)
            assert s0 == lo[b:e].min() and s1 == hi[b:e].max()
    # layout of the device records (pg_popgen_device): three int64 bit patterns, then pi / dxy / fst
    P = 3
    rec = np.zeros((4, 4 + 5 * P + 2 * 3), dtype=np.float64)
    rec[:, :3] = np.array([[5, 10 ** 12, 2]] * 4, dtype=np.int64).view(np.float64)
    rec[:, 3:3 + P] = 0.25
    back = multigpu.unpack_device_records(rec, P)
    assert back["sites"].tolist() == [5] * 4 and back["pos_sum"].tolist() == [10 ** 12] * 4 and back["path"].tolist() == [2] * 4
    assert np.all(back["pi"] == 0.25) and back["dxy"].shape == (4, 3) and back["fst"].shape == (4, 3)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK")


if __name__ == "__main__":
    main()
