"""Window sharding across the GPUs of one box: one process per GPU (torchrun), contiguous window
shards balanced by site count, no data-path collective, and ONE all-gather of the fixed-width per-window
records at the end.  torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU
tests); the statistics come from libpgwin.so.
"""
from __future__ import annotations

import numpy as np


def shard_windows(lo, hi, world: int):
    """Split windows (in order) into `world` contiguous shards with roughly equal total sites.
    Returns a list of (w_begin, w_end) per rank."""
    lo = np.asarray(lo, dtype=np.int64)
    hi = np.asarray(hi, dtype=np.int64)
    W = len(lo)
    if W == 0:
        return [(0, 0)] * world
    csum = np.concatenate([[0], np.cumsum(np.maximum(hi - lo, 1))])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        b = max(bounds[-1], min(b, W))
        bounds.append(b)
    bounds.append(W)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_site_range(lo, hi, w_begin: int, w_end: int):
    """Site range [s0, s1) a rank must hold to compute windows [w_begin, w_end) — overlapping windows
    simply replicate their halo sites on the neighbouring shard."""
    if w_end <= w_begin:
        return 0, 0
    lo = np.asarray(lo, dtype=np.int64)[w_begin:w_end]
    hi = np.asarray(hi, dtype=np.int64)[w_begin:w_end]
    return int(lo.min()), int(hi.max())


def all_gather_rows(local_rows: np.ndarray, counts, device=None):
    """All-gather per-window records (float64 [W_local, C]) from every rank into [W_total, C], in rank order.

    counts: number of windows per rank (known to every rank from shard_windows).  Uses one
    torch.distributed.all_gather on a padded buffer (NCCL when `device` is a CUDA device, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    C = local_rows.shape[1]
    wmax = max(int(c) for c in counts) if len(counts) else 0
    buf = torch.zeros((max(wmax, 1), C), dtype=torch.float64, device=device)
    if local_rows.shape[0]:
        buf[: local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [out[r][: int(counts[r])].cpu().numpy() for r in range(world)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, C))


def popgen_records(res: dict) -> np.ndarray:
    """Pack a popgen result dict into fixed-width float64 records [W, 3 + P + 2*npairs]
    (sites, pos_sum and path are exactly representable in float64)."""
    return np.concatenate([res["sites"][:, None].astype(np.float64), res["pos_sum"][:, None].astype(np.float64),
                           res["path"][:, None].astype(np.float64), res["pi"], res["dxy"], res["fst"]], axis=1)


def unpack_popgen_records(rec: np.ndarray, P: int) -> dict:
    npairs = P * (P - 1) // 2
    return dict(sites=rec[:, 0].astype(np.int64), pos_sum=rec[:, 1].astype(np.int64), path=rec[:, 2].astype(np.int32),
                pi=rec[:, 3:3 + P], dxy=rec[:, 3 + P:3 + P + npairs], fst=rec[:, 3 + P + npairs:3 + P + 2 * npairs])


class DeviceGather:
    """Preallocated buffers for the per-step all-gather of device-resident popgen records (NCCL):
    each rank's engine writes its records straight into `local` (pg_popgen_device), one
    all_gather_into_tensor moves them, one D2H brings the table to the host."""

    def __init__(self, counts, width, device):
        import torch
        self.counts = [int(c) for c in counts]
        self.width = int(width)
        self.wmax = max(max(self.counts), 1)
        self.world = len(self.counts)
        self.local = torch.zeros((self.wmax, self.width), dtype=torch.float64, device=device)
        self.all = torch.zeros((self.world * self.wmax, self.width), dtype=torch.float64, device=device)
        self.host = torch.zeros((self.world * self.wmax, self.width), dtype=torch.float64).pin_memory()

    def gather(self) -> np.ndarray:
        import torch
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.all, self.local)
        self.host.copy_(self.all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = self.host.numpy()
        return np.concatenate([h[r * self.wmax: r * self.wmax + self.counts[r]] for r in range(self.world)], axis=0)


def unpack_device_records(rec: np.ndarray, P: int) -> dict:
    """records as written by pg_popgen_device: the first three words are int64 bit patterns"""
    npairs = P * (P - 1) // 2
    ints = np.ascontiguousarray(rec[:, :3]).view(np.int64)
    return dict(sites=ints[:, 0].copy(), pos_sum=ints[:, 1].copy(), path=ints[:, 2].astype(np.int32),
                pi=rec[:, 3:3 + P], dxy=rec[:, 3 + P:3 + P + npairs], fst=rec[:, 3 + P + npairs:3 + P + 2 * npairs],
                popfreq=rec[:, 3 + P + 2 * npairs:])


def gathered_rows(table: np.ndarray, counts, w_max: int) -> np.ndarray:
    """rows of every rank (rank order) out of a gathered [world * w_max, C] table"""
    return np.concatenate([table[r * w_max: r * w_max + int(counts[r])] for r in range(len(counts))], axis=0)


def unpack_abba_records(rec: np.ndarray) -> dict:
    """records of pg_abbababa_allgather: [sites, pos_sum (int64 bit patterns), ABBA, BABA, D, fd, fdM, sitesUsed]"""
    ints = np.ascontiguousarray(rec[:, :2]).view(np.int64)
    return dict(sites=ints[:, 0].copy(), pos_sum=ints[:, 1].copy(), ABBA=rec[:, 2], BABA=rec[:, 3], D=rec[:, 4], fd=rec[:, 5],
                fdM=rec[:, 6], sitesUsed=rec[:, 7])


FOURPOP_KEYS = ('fhom', "fhom'", 'D', 'fd', "fd'", 'fdm', "fdm'", 'fdh', 'fdh2', 'fh', "ABBA", "BABA", "ABAA", "BAAA")


def unpack_fourpop_records(rec: np.ndarray) -> dict:
    """records of pg_fourpop_allgather: [sites, pos_sum, 14 statistics, sitesUsed]"""
    ints = np.ascontiguousarray(rec[:, :2]).view(np.int64)
    out = {k: rec[:, 2 + i] for i, k in enumerate(FOURPOP_KEYS)}
    out.update(sites=ints[:, 0].copy(), pos_sum=ints[:, 1].copy(), sitesUsed=rec[:, 16])
    return out
