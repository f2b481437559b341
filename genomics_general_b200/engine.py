"""Thin object wrapper over the C-ABI: one ``Engine`` = one ``pg_ctx`` = one GPU.

All numerics run in libpgwin.so (hand-written CUDA, sm_100a).  numpy is used only for host
buffers.  Nothing here computes statistics on the CPU.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._lib import PgError, check


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PinnedArray:
    """numpy view over cudaHostAlloc'd memory (freed on close/GC)."""

    def __init__(self, shape, dtype):
        self._lib = _lib.lib()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(self._lib.pg_host_alloc(C.byref(p), max(nbytes, 1)), "pg_host_alloc")
        self._p = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if self._p is not None:
            self.array = None
            self._lib.pg_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SFS_MAX_CELLS = 1 << 28      # dense histograms of pg_sfs / pg_sfs_tables (k1.cu); the reference's sparse dicts have no such limit


def _check_sfs_cells(cells):
    """refuse before 16 bytes per cell are allocated on the host: e.g. a 4-D spectrum of > 127 haplotypes per population"""
    if sum(cells) > SFS_MAX_CELLS:
        raise PgError("sfs: %d histogram cells requested, the dense spectra are limited to %d in total (fewer dimensions, "
                      "fewer joint spectra per run, or --subsample smaller populations)" % (sum(cells), SFS_MAX_CELLS))


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.lib()
        ctx = C.c_void_p()
        check(self._lib.pg_ctx_create(int(device), C.byref(ctx)), "pg_ctx_create")
        self._ctx = ctx
        self.device = int(device)
        self.S = 0
        self.H = 0
        self.P = 0
        self.W = 0

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_ctx", None) is not None:
            self._lib.pg_ctx_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ----
    def upload(self, geno: np.ndarray, pos=None):
        """geno: int8 [S, H] site-major (A0 C1 G2 T3, negative = missing); pos: int32 [S]."""
        geno = np.ascontiguousarray(geno, dtype=np.int8)
        assert geno.ndim == 2
        S, H = geno.shape
        if pos is not None:
            pos = np.ascontiguousarray(pos, dtype=np.int32)
            assert pos.shape == (S,)
        check(self._lib.pg_upload(self._ctx, _ptr(geno), S, H, _ptr(pos)), "pg_upload")
        self.S, self.H = S, H
        self.W = 0

    def synth_fill(self, spec, n_sites: int, spacing: int = 10):
        tv, to, tt, tm = spec.thresholds()
        check(self._lib.pg_synth_fill(self._ctx, int(n_sites), spec.n_pops, spec.samples_per_pop, spec.ploidy,
                                      spec.seed, tv, to, tt, tm, int(spacing)), "pg_synth_fill")
        self.S, self.H = int(n_sites), spec.n_haps
        self.W = 0

    def download(self, site0: int, n: int, into_geno=None, into_pos=None, want_geno=True, want_pos=True):
        """Device -> host copy (decoded to A0 C1 G2 T3 / -1).  Destination arrays must be C-contiguous."""
        g = p = None
        if want_geno:
            g = into_geno if into_geno is not None else np.empty((n, self.H), dtype=np.int8)
            assert g.flags.c_contiguous and g.dtype == np.int8 and g.shape == (n, self.H)
        if want_pos:
            p = into_pos if into_pos is not None else np.empty(n, dtype=np.int32)
            assert p.flags.c_contiguous and p.dtype == np.int32 and p.shape == (n,)
        check(self._lib.pg_download(self._ctx, int(site0), int(n), _ptr(g), _ptr(p)), "pg_download")
        return g, p

    def set_pops(self, hap_pop, n_pops=None):
        hap_pop = np.ascontiguousarray(hap_pop, dtype=np.int32)
        assert hap_pop.shape == (self.H,), (hap_pop.shape, self.H)
        P = int(n_pops) if n_pops is not None else int(hap_pop.max()) + 1
        check(self._lib.pg_set_pops(self._ctx, P, _ptr(hap_pop)), "pg_set_pops")
        self.P = P

    def set_windows(self, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.int64)
        hi = np.ascontiguousarray(hi, dtype=np.int64)
        assert lo.shape == hi.shape and lo.ndim == 1
        check(self._lib.pg_set_windows(self._ctx, len(lo), _ptr(lo), _ptr(hi)), "pg_set_windows")
        self.W = len(lo)

    # ---- statistics ----
    def set_freqstats(self, enable: bool = True):
        """Carry the popFreq counters (groupFreqStats) in the popgen site pass (opt-in: ~4 % of the pass)."""
        check(self._lib.pg_set_freqstats(self._ctx, 1 if enable else 0), "pg_set_freqstats")

    def popgen(self, min_sites: int = 1, min_data: float = 0.01, force_pairwise: bool = False):
        """-> dict(pi [W,P], dxy [W,npairs], fst [W,npairs], sites [W], pos_sum [W], path [W])."""
        W, P = self.W, self.P
        npairs = P * (P - 1) // 2
        pi = np.empty((W, P), dtype=np.float64)
        dxy = np.empty((W, npairs), dtype=np.float64)
        fst = np.empty((W, npairs), dtype=np.float64)
        sites = np.empty(W, dtype=np.int64)
        pos_sum = np.empty(W, dtype=np.int64)
        path = np.empty(W, dtype=np.int32)
        check(self._lib.pg_popgen(self._ctx, int(min_sites) if min_sites else 0, float(min_data),
                                  2 if force_pairwise else 0, _ptr(pi), _ptr(dxy), _ptr(fst), _ptr(sites),
                                  _ptr(pos_sum), _ptr(path)), "pg_popgen")
        return dict(pi=pi, dxy=dxy, fst=fst, sites=sites, pos_sum=pos_sum, path=path,
                    pairs=list(itertools.combinations(range(P), 2)))

    def popgen_record_width(self) -> int:
        return 4 + 5 * self.P + 2 * (self.P * (self.P - 1) // 2)

    def popgen_freqstats(self):
        """popFreq columns (Alignment.groupFreqStats) of the most recent popgen() call:
        dict(l [W], S, thetaPi, thetaW, TajD [W,P])."""
        W, P = self.W, self.P
        l = np.empty(W, dtype=np.float64)
        out = {k: np.empty((W, P), dtype=np.float64) for k in ("S", "thetaPi", "thetaW", "TajD")}
        check(self._lib.pg_popgen_freqstats(self._ctx, _ptr(l), _ptr(out["S"]), _ptr(out["thetaPi"]), _ptr(out["thetaW"]),
                                            _ptr(out["TajD"])), "pg_popgen_freqstats")
        out["l"] = l
        return out

    def popgen_device(self, d_rec_ptr: int, min_sites: int = 1, min_data: float = 0.01, force_pairwise: bool = False) -> int:
        """Statistics left on the device as fixed-width records (see pg_popgen_device); `d_rec_ptr` is a device
        pointer to W * popgen_record_width() 8-byte words.  Returns the number of windows that took the pairwise path."""
        n = C.c_int64(0)
        check(self._lib.pg_popgen_device(self._ctx, int(min_sites) if min_sites else 0, float(min_data),
                                         2 if force_pairwise else 0, C.c_void_p(int(d_rec_ptr)), C.byref(n)),
              "pg_popgen_device")
        return int(n.value)

    # ---- multi-GPU: native NCCL gather (one process per GPU) ----
    def nccl_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        check(self._lib.pg_nccl_unique_id(buf), "pg_nccl_unique_id")
        return bytes(buf)

    def nccl_init(self, world: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self._lib.pg_nccl_init(self._ctx, int(world), int(rank), buf), "pg_nccl_init")
        self._world, self._rank = int(world), int(rank)

    def nccl_finalize(self):
        check(self._lib.pg_nccl_finalize(self._ctx), "pg_nccl_finalize")
        self._world, self._rank = 1, 0

    def popgen_allgather(self, w_max: int, table: np.ndarray, min_sites: int = 1, min_data: float = 0.01,
                         force_pairwise: bool = False) -> int:
        """Statistics of this rank's windows, then ONE ncclAllGather of every rank's records into `table`
        (float64 [world * w_max, popgen_record_width()], ideally pinned).  Returns this rank's pairwise-window count."""
        assert table.dtype == np.float64 and table.flags.c_contiguous
        assert table.shape == (self._world * int(w_max), self.popgen_record_width())
        n = C.c_int64(0)
        check(self._lib.pg_popgen_allgather(self._ctx, int(min_sites) if min_sites else 0, float(min_data),
                                            2 if force_pairwise else 0, int(w_max), _ptr(table), C.byref(n)),
              "pg_popgen_allgather")
        return int(n.value)

    def popgen_gather_begin(self, w_max: int, slot: int, min_sites: int = 1, min_data: float = 0.01):
        """Pipelined popgen_allgather: enqueue one batch (statistics on the main stream, all-gather + D2H on a side stream)."""
        check(self._lib.pg_popgen_gather_begin(self._ctx, int(min_sites) if min_sites else 0, float(min_data), int(w_max),
                                               int(slot)), "pg_popgen_gather_begin")

    def popgen_gather_end(self, w_max: int, slot: int) -> np.ndarray:
        """-> float64 [world * w_max, popgen_record_width()] view of the slot's pinned table (valid until its next begin)."""
        p = C.c_void_p()
        n = C.c_int64(0)
        check(self._lib.pg_popgen_gather_end(self._ctx, int(slot), C.byref(p), C.byref(n)), "pg_popgen_gather_end")
        world = getattr(self, "_world", 1) or 1
        shape = (world * int(w_max), self.popgen_record_width())
        buf = (C.c_double * (shape[0] * shape[1])).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def abbababa_allgather(self, p1: int, p2: int, p3: int, o: int, min_data: float, w_max: int, table: np.ndarray):
        """ABBA-BABA statistics of this rank's windows + ONE ncclAllGather: `table` float64 [world * w_max, 8] receives
        [sites, pos_sum (int64 bit patterns), ABBA, BABA, D, fd, fdM, sitesUsed] per window (multigpu.unpack_abba_records)."""
        assert table.dtype == np.float64 and table.flags.c_contiguous and table.shape == (self._world * int(w_max), 8)
        check(self._lib.pg_abbababa_allgather(self._ctx, p1, p2, p3, o, float(min_data), int(w_max), _ptr(table)),
              "pg_abbababa_allgather")

    def fourpop_allgather(self, p1: int, p2: int, p3: int, p4: int, min_data: float, w_max: int, table: np.ndarray,
                          polarize: bool = False, fixed: bool = False):
        """genomics.fourPop of this rank's windows + ONE ncclAllGather: `table` float64 [world * w_max, 17]."""
        assert table.dtype == np.float64 and table.flags.c_contiguous and table.shape == (self._world * int(w_max), 17)
        mode = 1 if polarize else (2 if fixed else 0)
        check(self._lib.pg_fourpop_allgather(self._ctx, p1, p2, p3, p4, float(min_data), mode, int(w_max), _ptr(table)),
              "pg_fourpop_allgather")

    def abbababa(self, p1: int, p2: int, p3: int, o: int, min_data: float = 0.01):
        """-> dict(ABBA,BABA,D,fd,fdM [W], sitesUsed [W] (nan = no good site), sites, pos_sum)."""
        W = self.W
        out = np.empty((W, 5), dtype=np.float64)
        used = np.empty(W, dtype=np.float64)
        sites = np.empty(W, dtype=np.int64)
        pos_sum = np.empty(W, dtype=np.int64)
        check(self._lib.pg_abbababa(self._ctx, p1, p2, p3, o, float(min_data), _ptr(out), _ptr(used), _ptr(sites),
                                    _ptr(pos_sum)), "pg_abbababa")
        return dict(ABBA=out[:, 0], BABA=out[:, 1], D=out[:, 2], fd=out[:, 3], fdM=out[:, 4], sitesUsed=used,
                    sites=sites, pos_sum=pos_sum)

    FOURPOP_KEYS = ('fhom', "fhom'", 'D', 'fd', "fd'", 'fdm', "fdm'", 'fdh', 'fdh2', 'fh', "ABBA", "BABA", "ABAA", "BAAA")

    def fourpop(self, p1: int, p2: int, p3: int, p4: int, min_data: float = 0.01, polarize: bool = False,
                fixed: bool = False):
        """genomics.fourPop per window -> dict(<14 statistics> [W], sitesUsed [W], sites, pos_sum)."""
        W = self.W
        out = np.empty((W, 14), dtype=np.float64)
        used = np.empty(W, dtype=np.float64)
        sites = np.empty(W, dtype=np.int64)
        pos_sum = np.empty(W, dtype=np.int64)
        mode = 1 if polarize else (2 if fixed else 0)          # genomics.py:1610-1615: polarize wins over fixed
        check(self._lib.pg_fourpop(self._ctx, p1, p2, p3, p4, float(min_data), mode, _ptr(out), _ptr(used), _ptr(sites),
                                   _ptr(pos_sum)), "pg_fourpop")
        r = {k: out[:, i] for i, k in enumerate(self.FOURPOP_KEYS)}
        r.update(sitesUsed=used, sites=sites, pos_sum=pos_sum)
        return r

    def ingest_text(self, data: bytes, fmt: int, col_hap, col_ploidy, H: int, offset: int = 0) -> int:
        """Device-side .geno tokenizer (pg_ingest_text): data[offset:] = the file's data lines (no copy is made).
        Returns the number of sites now resident."""
        col_hap = np.ascontiguousarray(col_hap, dtype=np.int32)
        col_ploidy = np.ascontiguousarray(col_ploidy, dtype=np.int8)
        assert col_hap.shape == col_ploidy.shape
        n = C.c_int64(0)
        addr = C.cast(C.c_char_p(data), C.c_void_p).value or 0        # `data` stays referenced by the caller
        check(self._lib.pg_ingest_text(self._ctx, C.c_void_p(addr + offset), len(data) - offset, int(fmt), len(col_hap),
                                       _ptr(col_hap), _ptr(col_ploidy), int(H), C.byref(n)), "pg_ingest_text")
        self.S, self.H = int(n.value), int(H)
        return self.S

    def ingest_file(self, path: str, body_offset: int, fmt: int, col_hap, col_ploidy, H: int) -> int:
        """The same, reading the file straight into the pinned staging buffers (pg_ingest_file)."""
        col_hap = np.ascontiguousarray(col_hap, dtype=np.int32)
        col_ploidy = np.ascontiguousarray(col_ploidy, dtype=np.int8)
        n = C.c_int64(0)
        check(self._lib.pg_ingest_file(self._ctx, path.encode(), int(body_offset), int(fmt), len(col_hap), _ptr(col_hap),
                                       _ptr(col_ploidy), int(H), C.byref(n)), "pg_ingest_file")
        self.S, self.H = int(n.value), int(H)
        return self.S

    def ingest_file_range(self, path: str, byte_lo: int, byte_hi: int, fmt: int, col_hap, col_ploidy, H: int) -> int:
        """One rank's share of the file: bytes [byte_lo, byte_hi) (line starts; byte_hi < 0 = end of file)."""
        col_hap = np.ascontiguousarray(col_hap, dtype=np.int32)
        col_ploidy = np.ascontiguousarray(col_ploidy, dtype=np.int8)
        n = C.c_int64(0)
        check(self._lib.pg_ingest_file_range(self._ctx, path.encode(), int(byte_lo), int(byte_hi), int(fmt), len(col_hap),
                                             _ptr(col_hap), _ptr(col_ploidy), int(H), C.byref(n)), "pg_ingest_file_range")
        self.S, self.H = int(n.value), int(H)
        return self.S

    def append_sites(self, geno: np.ndarray, pos=None):
        """Append sites (int8 [n, H]) after the resident ones (halo of the next rank's first sites)."""
        geno = np.ascontiguousarray(geno, dtype=np.int8)
        assert geno.ndim == 2 and geno.shape[1] == self.H
        if pos is not None:
            pos = np.ascontiguousarray(pos, dtype=np.int32)
        check(self._lib.pg_append_sites(self._ctx, geno.shape[0], _ptr(geno), _ptr(pos)), "pg_append_sites")
        self.S += geno.shape[0]

    def ingest_meta(self, S: int):
        """(pos int32 [S], new_scaffold int8 [S], line_off int64 [S]) of the last ingest_text."""
        pos = np.empty(S, dtype=np.int32)
        newsc = np.empty(S, dtype=np.int8)
        off = np.empty(S, dtype=np.int64)
        check(self._lib.pg_ingest_meta(self._ctx, _ptr(pos), _ptr(newsc), _ptr(off)), "pg_ingest_meta")
        check(self._lib.pg_ingest_release(self._ctx), "pg_ingest_release")
        return pos, newsc, off

    def site_counts(self, site0: int = 0, n: int = None, out=None):
        """uint16 [n, P, 4] A,C,G,T counts per population (`out`: a caller-owned array to fill, e.g. one whose pages are
        already resident — a fresh 100 MB array costs more in page faults than the kernel and the copy together)."""
        n = self.S - site0 if n is None else int(n)
        if out is None:
            out = np.empty((n, self.P, 4), dtype=np.uint16)
        else:
            out = out[:n]
            assert out.dtype == np.uint16 and out.flags.c_contiguous and out.shape == (n, self.P, 4)
        check(self._lib.pg_site_counts(self._ctx, int(site0), n, _ptr(out)), "pg_site_counts")
        return out

    def site_target_freqs(self, target: str, site0: int = 0, n: int | None = None, min_data: float = 0.0,
                          as_counts: bool = False):
        """freq.py --target derived|minor -> (values float64 [n,P], tie bool [n])."""
        n = self.S - site0 if n is None else n
        out = np.empty((n, self.P), dtype=np.float64)
        tie = np.zeros(n, dtype=np.uint8)
        code = {"derived": 1, "minor": 2}[target]
        check(self._lib.pg_site_target_freqs(self._ctx, int(site0), int(n), code, float(min_data), 1 if as_counts else 0,
                                             _ptr(out), _ptr(tie)), "pg_site_target_freqs")
        return out, tie.astype(bool)

    def sfs(self, n_in: int, groups, pop_sizes, outgroup: int = -1, site_mask=None):
        """sfs.py for genotype input: `groups` = list of tuples of in-group population indices; pop_sizes[X] = haplotypes of
        population X.  Returns (list of dense int64 spectra shaped (N_k+1, ...), list of first-site arrays, sites counted)."""
        goff = np.zeros(len(groups) + 1, dtype=np.int32)
        for k, grp in enumerate(groups):
            goff[k + 1] = goff[k] + len(grp)
        gp = np.array([x for grp in groups for x in grp], dtype=np.int32)
        shapes = [tuple(int(pop_sizes[x]) + 1 for x in grp) for grp in groups]
        cells = [int(np.prod([int(d) for d in sh], dtype=object)) for sh in shapes]
        _check_sfs_cells(cells)
        hist = np.zeros(sum(cells), dtype=np.int64)
        first = np.zeros(sum(cells), dtype=np.int64)
        mask = None if site_mask is None else np.ascontiguousarray(site_mask, dtype=np.uint8)
        if mask is not None:
            assert mask.shape == (self.S,)
        n = C.c_int64(0)
        check(self._lib.pg_sfs(self._ctx, int(n_in), int(outgroup), len(groups), _ptr(goff), _ptr(gp), _ptr(mask), _ptr(hist),
                               _ptr(first), C.byref(n)), "pg_sfs")
        offs = np.concatenate([[0], np.cumsum(cells)])
        return ([hist[offs[k]:offs[k + 1]].reshape(shapes[k]) for k in range(len(groups))],
                [first[offs[k]:offs[k + 1]].reshape(shapes[k]) for k in range(len(groups))], int(n.value))

    def sfs_tables(self, kind: str, table, n_in: int, groups, outgroup: int = -1, site_mask=None):
        """sfs.py on count tables: kind "base" = uint16 [n,P,4] base counts per population, "target" = int32 [n,P] counts of
        the target allele.  Returns (dense spectra, first-site arrays, sites counted) like sfs()."""
        if kind == "base":
            table = np.ascontiguousarray(table, dtype=np.uint16)
            n, P = table.shape[0], table.shape[1]
            dims = (table.sum(axis=2, dtype=np.int64).max(axis=0) + 1 if n else np.ones(P, np.int64)).astype(np.int32)
        else:
            table = np.ascontiguousarray(table, dtype=np.int32)
            n, P = table.shape
            dims = (table.max(axis=0) + 1 if n else np.ones(P, np.int64)).astype(np.int32)
            assert n == 0 or table.min() >= 0
        goff = np.zeros(len(groups) + 1, dtype=np.int32)
        for k, grp in enumerate(groups):
            goff[k + 1] = goff[k] + len(grp)
        gp = np.array([x for grp in groups for x in grp], dtype=np.int32)
        shapes = [tuple(int(dims[x]) for x in grp) for grp in groups]
        cells = [int(np.prod([int(d) for d in sh], dtype=object)) for sh in shapes]
        _check_sfs_cells(cells)
        hist = np.zeros(sum(cells), dtype=np.int64)
        first = np.zeros(sum(cells), dtype=np.int64)
        mask = None if site_mask is None else np.ascontiguousarray(site_mask, dtype=np.uint8)
        cnt = C.c_int64(0)
        check(self._lib.pg_sfs_tables(self._ctx, 0 if kind == "base" else 1, _ptr(table), int(n), int(P), _ptr(dims), int(n_in),
                                      int(outgroup), len(groups), _ptr(goff), _ptr(gp), _ptr(mask), _ptr(hist), _ptr(first),
                                      C.byref(cnt)), "pg_sfs_tables")
        offs = np.concatenate([[0], np.cumsum(cells)])
        return ([hist[offs[k]:offs[k + 1]].reshape(shapes[k]) for k in range(len(groups))],
                [first[offs[k]:offs[k + 1]].reshape(shapes[k]) for k in range(len(groups))], int(cnt.value))

    def pairdist(self, hap_ind, n_ind: int, include_same_with_same: bool = False, min_sites: int = 0, out=None):
        """-> dict(dist [W,n_ind,n_ind], sites [W], pos_sum [W]).  min_sites > 0 masks haplotype pairs with fewer
        jointly non-missing sites (what an earlier groupDistStats does to the reference's cached matrix).
        `out`: a caller-owned float64 [W, n_ind, n_ind] array; a PinnedArray's `.array` is written by the copy engine
        directly (hundreds of MB of matrices otherwise spend most of their time in page faults of a fresh array)."""
        hap_ind = np.ascontiguousarray(hap_ind, dtype=np.int32)
        assert hap_ind.shape == (self.H,)
        W = self.W
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (W, n_ind, n_ind)
        dist = out if out is not None else np.empty((W, n_ind, n_ind), dtype=np.float64)
        sites = np.empty(W, dtype=np.int64)
        pos_sum = np.empty(W, dtype=np.int64)
        check(self._lib.pg_pairdist(self._ctx, int(n_ind), _ptr(hap_ind), 1 if include_same_with_same else 0,
                                    int(min_sites or 0), _ptr(dist), _ptr(sites), _ptr(pos_sum)), "pg_pairdist")
        return dict(dist=dist, sites=sites, pos_sum=pos_sum)

    def pairdist_cat(self, hap_ind, n_ind: int, include_same_with_same: bool = False):
        """distMat.py --windType cat: one matrix over every uploaded site (summed over the ranks of the NCCL
        communicator when one is set) -> (dist [n_ind,n_ind], total_sites)."""
        hap_ind = np.ascontiguousarray(hap_ind, dtype=np.int32)
        assert hap_ind.shape == (self.H,)
        dist = np.empty((n_ind, n_ind), dtype=np.float64)
        tot = C.c_int64(0)
        check(self._lib.pg_pairdist_cat(self._ctx, int(n_ind), _ptr(hap_ind), 1 if include_same_with_same else 0,
                                        _ptr(dist), C.byref(tot)), "pg_pairdist_cat")
        return dist, int(tot.value)

    def seq_nonnan(self):
        """Alignment.seqNonNan() per window -> int64 [W, H]."""
        out = np.empty((self.W, self.H), dtype=np.int64)
        check(self._lib.pg_seq_nonnan(self._ctx, _ptr(out)), "pg_seq_nonnan")
        return out

    def ind_het(self, hap_ind, n_ind: int, min_sites: int = 0):
        """Alignment.sampleHet() per window -> [W, n_ind]."""
        hap_ind = np.ascontiguousarray(hap_ind, dtype=np.int32)
        assert hap_ind.shape == (self.H,)
        het = np.empty((self.W, n_ind), dtype=np.float64)
        check(self._lib.pg_ind_het(self._ctx, int(n_ind), _ptr(hap_ind), int(min_sites or 0), _ptr(het)), "pg_ind_het")
        return het

    def hapstats(self, max_dist: float = 0.0, min_sites: int = 0, diag_nan: bool = False):
        """Alignment.H12stats(maxDist) per window -> [W, P, 3] = H1, H12, H2."""
        out = np.empty((self.W, self.P, 3), dtype=np.float64)
        check(self._lib.pg_hapstats(self._ctx, float(max_dist), int(min_sites or 0), 1 if diag_nan else 0, _ptr(out)),
              "pg_hapstats")
        return out

    def pair_counts(self, window: int):
        """(diff, n) int32 [H,H] of one window (Alignment.distMatrix / pairNonNan numerators)."""
        diff = np.empty((self.H, self.H), dtype=np.int32)
        n = np.empty((self.H, self.H), dtype=np.int32)
        check(self._lib.pg_pair_counts(self._ctx, int(window), _ptr(diff), _ptr(n)), "pg_pair_counts")
        return diff, n

    # ---- introspection ----
    def last_timings(self):
        cap = 32
        names = (C.c_char * 32 * cap)()
        ms = (C.c_float * cap)()
        launches = (C.c_int32 * cap)()
        cnt = C.c_int32(0)
        check(self._lib.pg_last_timings(self._ctx, cap, names, ms, launches, C.byref(cnt)), "pg_last_timings")
        return {names[k].value.decode(): dict(ms=float(ms[k]), launches=int(launches[k])) for k in range(cnt.value)}

    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(self._lib.pg_launch_count(self._ctx, C.byref(n)), "pg_launch_count")
        return int(n.value)


def k1_plan(S: int, H: int):
    """Host-only: the site-pass launch geometry for a shape (works without a GPU)."""
    L = _lib.lib()
    v = [C.c_int32(0) for _ in range(5)]
    check(L.pg_debug_k1_plan(int(S), int(H), *[C.byref(x) for x in v]), "pg_debug_k1_plan")
    return dict(pitch=v[0].value, lanes_per_site=v[1].value, tile_sites=v[2].value, stages=v[3].value,
                smem_bytes=v[4].value)
