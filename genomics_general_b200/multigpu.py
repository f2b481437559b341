"""Window sharding across the GPUs of one box: one process per GPU, contiguous window shards balanced by site
count, no data-path collective, and ONE all-gather of the fixed-width per-window records at the end —
issued by the C-ABI itself (pg_*_allgather, native NCCL).  This module is host arithmetic only (no torch): shard
boundaries and the layouts of the gathered record tables.  The command lines' multi-GPU path is in mgpu.py.
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
