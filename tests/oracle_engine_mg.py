"""TEST INFRASTRUCTURE — the oracle-backed Engine look-alike for the multi-GPU command lines (`--devices N`) on the CPU.

Stands in for one rank's engine: the byte range of the file is tokenised by the native HOST tokenizer (pg_geno_parse),
the statistics come from the CPU oracle, and the one all-gather of the fixed-width records goes through files in a
directory whose path travels where the NCCL id travels (the 128 bytes rank 0 publishes).  Everything else of the N > 1
path — byte ranges cut at line starts, the global picture of positions / scaffold runs, window ownership, the halo sites
re-read from the next range, the order of the gathered table, rank 0 writing the rows — is the product's own code
(genomics_general_b200/mgpu.py and the command lines).  Never imported by the product."""
import ctypes as C
import os
import tempfile
import time

import numpy as np

from genomics_general_b200 import _lib
from oracle_engine import OracleEngine


class OracleEngineMG(OracleEngine):
    def __init__(self, device=0):
        super().__init__(device)
        self._world, self._rank, self._xdir, self._round = 1, 0, None, 0

    # ---- ingest of a byte range (what pg_ingest_file_range + pg_ingest_meta give) ----
    def ingest_file_range(self, path, byte_lo, byte_hi, fmt, col_hap, col_ploidy, H):
        with open(path, "rb") as f:
            f.seek(byte_lo)
            body = f.read((os.path.getsize(path) if byte_hi < 0 else byte_hi) - byte_lo)
        col_hap = np.asarray(col_hap, dtype=np.int64)
        wanted = np.flatnonzero(col_hap >= 0)
        order = wanted[np.argsort(col_hap[wanted], kind="stable")]          # output sample k <- file column order[k]
        col_take = np.ascontiguousarray(order, dtype=np.int32)
        pl = np.ascontiguousarray(np.asarray(col_ploidy)[order], dtype=np.int8)
        L = _lib.lib()
        n = C.c_int64(0)
        assert L.pg_geno_count_lines(body, len(body), C.byref(n)) == 0
        S = int(n.value)
        geno = np.empty((S, int(H)), dtype=np.int8)
        pos = np.empty(S, dtype=np.int32)
        newsc = np.empty(S, dtype=np.int8)
        off = np.empty(S, dtype=np.int64)
        rc = L.pg_geno_parse(body, len(body), int(fmt), len(col_take), col_take.ctypes.data_as(C.c_void_p),
                             pl.ctypes.data_as(C.c_void_p), int(H), S, geno.ctypes.data_as(C.c_void_p),
                             pos.ctypes.data_as(C.c_void_p), newsc.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), 1)
        assert rc == 0, L.pg_last_error()
        self.upload(geno, pos)
        self._meta = (pos, newsc, off)
        return S

    def ingest_meta(self, S):
        assert S == len(self._meta[0])
        return self._meta

    def append_sites(self, geno, pos=None):
        geno = np.asarray(geno, dtype=np.int8)
        assert geno.ndim == 2 and geno.shape[1] == self.H
        self.g = np.concatenate([self.g, geno], axis=0)
        self.pos = np.concatenate([self.pos, np.zeros(len(geno), np.int64) if pos is None else np.asarray(pos, np.int64)])
        self.S = len(self.g)

    def last_timings(self):
        return {}

    def set_freqstats(self, enable=True):
        self._want_freq = bool(enable)

    # ---- the exchange: the "NCCL id" is the path of a directory ----
    def nccl_unique_id(self):
        d = tempfile.mkdtemp(prefix="pg_fake_nccl_").encode()
        assert len(d) < 128
        return d + b"\0" * (128 - len(d))

    def nccl_init(self, world, rank, unique_id):
        assert len(unique_id) == 128
        self._world, self._rank = int(world), int(rank)
        self._xdir = unique_id.rstrip(b"\0").decode()

    def nccl_finalize(self):
        self._world, self._rank = 1, 0

    def _allgather(self, mine, w_max, table):
        """every rank's [w_max, C] block into rows rank * w_max .. of `table` (what ncclAllGather does in place)"""
        self._round += 1
        block = np.zeros((w_max, table.shape[1]), dtype=np.float64)
        block[:len(mine)] = mine
        p = os.path.join(self._xdir, "g%d.r%d.npy" % (self._round, self._rank))
        with open(p + ".tmp", "wb") as f:
            np.save(f, block)
        os.rename(p + ".tmp", p)
        for r in range(self._world):
            q = os.path.join(self._xdir, "g%d.r%d.npy" % (self._round, r))
            t0 = time.time()
            while not os.path.exists(q):
                assert time.time() - t0 < 300, "rank %d never published its records" % r
                time.sleep(0.005)
            table[r * w_max:(r + 1) * w_max] = np.load(q)

    def popgen_record_width(self):
        return 4 + 5 * self.P + 2 * (self.P * (self.P - 1) // 2)

    def popgen_allgather(self, w_max, table, min_sites=1, min_data=0.01, force_pairwise=False):
        assert table.shape == (self._world * int(w_max), self.popgen_record_width())
        r = self.popgen(min_sites, min_data)
        rec = np.zeros((self.W, table.shape[1]), dtype=np.float64)
        ints = np.stack([r["sites"].astype(np.int64), r["pos_sum"].astype(np.int64), r["path"].astype(np.int64)], axis=1)
        rec[:, :3] = ints.view(np.float64)                                  # int64 bit patterns, as the device writes them
        P = self.P
        npairs = P * (P - 1) // 2
        rec[:, 3:3 + P] = r["pi"]
        rec[:, 3 + P:3 + P + npairs] = r["dxy"]
        rec[:, 3 + P + npairs:3 + P + 2 * npairs] = r["fst"]
        if getattr(self, "_want_freq", False):       # [l, S[P], thetaPi[P], thetaW[P], TajD[P]] (pg_popgen_freqstats decodes the same)
            f = self.popgen_freqstats()
            b = 3 + P + 2 * npairs
            rec[:, b] = f["l"]
            for k, key in enumerate(("S", "thetaPi", "thetaW", "TajD")):
                rec[:, b + 1 + k * P:b + 1 + (k + 1) * P] = f[key]
        self._allgather(rec, int(w_max), table)
        return int((r["path"] == 2).sum())

    def abbababa_allgather(self, p1, p2, p3, o, min_data, w_max, table):
        assert table.shape == (self._world * int(w_max), 8)
        r = self.abbababa(p1, p2, p3, o, min_data)
        rec = np.zeros((self.W, 8), dtype=np.float64)
        rec[:, :2] = np.stack([r["sites"].astype(np.int64), r["pos_sum"].astype(np.int64)], axis=1).view(np.float64)
        for k, key in enumerate(("ABBA", "BABA", "D", "fd", "fdM", "sitesUsed")):
            rec[:, 2 + k] = r[key]
        self._allgather(rec, int(w_max), table)

    def fourpop_allgather(self, p1, p2, p3, p4, min_data, w_max, table, polarize=False, fixed=False):
        assert table.shape == (self._world * int(w_max), 17)
        r = self.fourpop(p1, p2, p3, p4, min_data, polarize=polarize, fixed=fixed)
        rec = np.zeros((self.W, 17), dtype=np.float64)
        rec[:, :2] = np.stack([r["sites"].astype(np.int64), r["pos_sum"].astype(np.int64)], axis=1).view(np.float64)
        for k, key in enumerate(self.FOURPOP_KEYS):
            rec[:, 2 + k] = r[key]
        rec[:, 16] = r["sitesUsed"]
        self._allgather(rec, int(w_max), table)

    def pairdist_cat(self, hap_ind, n_ind, include_same_with_same=False):
        """with a communicator: the integer pair matrices of the ranks' site shards are added before the division — the
        same numbers as one pass over the concatenated shards, which is what this computes"""
        if self._world <= 1:
            return super().pairdist_cat(hap_ind, n_ind, include_same_with_same)
        self._round += 1
        p = os.path.join(self._xdir, "cat%d.r%d.npy" % (self._round, self._rank))
        with open(p + ".tmp", "wb") as f:
            np.save(f, self.g)
        os.rename(p + ".tmp", p)
        parts = []
        for r in range(self._world):
            q = os.path.join(self._xdir, "cat%d.r%d.npy" % (self._round, r))
            t0 = time.time()
            while not os.path.exists(q):
                assert time.time() - t0 < 300
                time.sleep(0.005)
            parts.append(np.load(q))
        whole = OracleEngine()
        whole.upload(np.concatenate(parts, axis=0))
        return whole.pairdist_cat(hap_ind, n_ind, include_same_with_same)
