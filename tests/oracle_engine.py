"""TEST INFRASTRUCTURE — an Engine look-alike backed by the CPU oracle.

The CPU tests (`-m "not gpu"`) use it to run the command lines' HOST logic (argument handling, populations and ploidy,
window generation, row assembly and number formatting) against the reference scripts' own output without a GPU.
It is never imported by the product; the GPU tests run the same command lines on the real engine."""
import warnings

import numpy as np

from oracle import dense_oracle as do


class OracleEngine:
    FOURPOP_KEYS = do.FOURPOP_KEYS[:-1]

    def __init__(self, device=0):
        self.g = None
        self.pos = None
        self.hap_pop = None
        self.P = 0
        self.lo = self.hi = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    # ---- data ----
    def upload(self, geno, pos=None):
        self.g = np.asarray(geno, dtype=np.int8)
        self.pos = np.zeros(len(self.g), dtype=np.int64) if pos is None else np.asarray(pos, dtype=np.int64)
        self.S, self.H = self.g.shape

    def set_pops(self, hap_pop, P):
        self.hap_pop = np.asarray(hap_pop, dtype=np.int32)
        self.P = int(P)

    def set_windows(self, lo, hi):
        self.lo = np.asarray(lo, dtype=np.int64)
        self.hi = np.asarray(hi, dtype=np.int64)
        self.W = len(self.lo)

    def set_freqstats(self, enable=True):
        pass

    def _win(self, w):
        return self.g[self.lo[w]:self.hi[w]]

    def _book(self):
        sites = (self.hi - self.lo).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(self.pos)])
        return sites, (csum[self.hi] - csum[self.lo]).astype(np.int64)

    # ---- statistics ----
    def popgen(self, min_sites=1, min_data=0.01, force_pairwise=False):
        P = self.P
        npairs = P * (P - 1) // 2
        sites, pos_sum = self._book()
        pi = np.full((self.W, P), np.nan)
        dxy = np.full((self.W, npairs), np.nan)
        fst = np.full((self.W, npairs), np.nan)
        path = np.zeros(self.W, dtype=np.int32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for w in range(self.W):
                if sites[w] < (min_sites or 0) or sites[w] == 0:
                    continue
                path[w] = 2
                pi[w], dxy[w], fst[w] = do.group_dist_stats(self._win(w), self.hap_pop, P, min_sites or None, min_data)
        return dict(pi=pi, dxy=dxy, fst=fst, sites=sites, pos_sum=pos_sum, path=path)

    def popgen_freqstats(self):
        out = {k: np.full((self.W, self.P), np.nan) for k in ("S", "thetaPi", "thetaW", "TajD")}
        out["l"] = np.full(self.W, np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for w in range(self.W):
                if self.hi[w] > self.lo[w]:
                    f = do.group_freq_stats(self._win(w)[:, self.hap_pop >= 0], self.hap_pop[self.hap_pop >= 0], self.P)
                    for k in ("S", "thetaPi", "thetaW", "TajD"):
                        out[k][w] = f[k]
                    out["l"][w] = f["l"][0]
        return out

    def abbababa(self, p1, p2, p3, o, min_data=0.01):
        sites, pos_sum = self._book()
        keys = ("ABBA", "BABA", "D", "fd", "fdM", "sitesUsed")
        out = {k: np.full(self.W, np.nan) for k in keys}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for w in range(self.W):
                if self.hi[w] > self.lo[w]:
                    r = do.abbababa(self._win(w), self.hap_pop, p1, p2, p3, o, min_data)
                    for k in keys:
                        out[k][w] = r[k]
        out.update(sites=sites, pos_sum=pos_sum)
        return out

    def fourpop(self, p1, p2, p3, p4, min_data=0.01, polarize=False, fixed=False):
        sites, pos_sum = self._book()
        out = {k: np.full(self.W, np.nan) for k in self.FOURPOP_KEYS}
        out["sitesUsed"] = np.zeros(self.W)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for w in range(self.W):
                if self.hi[w] > self.lo[w]:
                    r = do.four_pop(self._win(w), self.hap_pop, p1, p2, p3, p4, min_data, polarize=polarize,
                                    fixed=fixed and not polarize)
                    for k in self.FOURPOP_KEYS + ("sitesUsed",):
                        out[k][w] = r[k]
        out.update(sites=sites, pos_sum=pos_sum)
        return out

    def site_counts(self, site0=0, n=None):
        n = self.S - site0 if n is None else n
        return do.site_counts(self.g[site0:site0 + n], self.hap_pop, self.P).astype(np.uint16)

    def site_target_freqs(self, target, site0=0, n=None, min_data=0.0, as_counts=False):
        n = self.S - site0 if n is None else n
        v, tie = do.target_freqs(self.g[site0:site0 + n], self.hap_pop, self.P, target, min_data=min_data, as_counts=as_counts)
        return v, tie

    def pairdist(self, hap_ind, n_ind, include_same_with_same=False, min_sites=0):
        sites, pos_sum = self._book()
        dist = np.full((self.W, n_ind, n_ind), np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for w in range(self.W):
                if self.hi[w] > self.lo[w]:
                    dist[w] = do.ind_pair_dists(self._win(w), hap_ind, n_ind, include_same_with_same, min_sites or None)
        return dict(dist=dist, sites=sites, pos_sum=pos_sum)

    def pairdist_cat(self, hap_ind, n_ind, include_same_with_same=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return do.ind_pair_dists(self.g, hap_ind, n_ind, include_same_with_same), int(self.S)

    def seq_nonnan(self):
        return np.stack([(self._win(w) >= 0).sum(axis=0) for w in range(self.W)]).astype(np.int64)

    def ind_het(self, hap_ind, n_ind, min_sites=0):
        out = np.full((self.W, n_ind), np.nan)
        for w in range(self.W):
            if self.hi[w] > self.lo[w]:
                out[w] = do.sample_het(self._win(w), hap_ind, n_ind, min_sites or None)
        return out

    def hapstats(self, max_dist=0.0, min_sites=0, diag_nan=False):
        out = np.full((self.W, self.P, 3), np.nan)
        for w in range(self.W):
            if self.hi[w] > self.lo[w]:
                out[w] = do.h12_stats(self._win(w), self.hap_pop, self.P, max_dist, min_sites or None, diag_nan or bool(min_sites))
        return out

    def sfs(self, n_in, groups, pop_sizes, outgroup=-1, site_mask=None):
        tc, used = do.sfs_target_counts(self.g, self.hap_pop, n_in, outgroup)
        if site_mask is not None:
            used = used & np.asarray(site_mask, dtype=bool)
        hists, firsts = [], []
        for grp in groups:
            shape = tuple(int(pop_sizes[x]) + 1 for x in grp)
            h = np.zeros(shape, dtype=np.int64)
            f = np.full(shape, -1, dtype=np.int64)
            for s in np.where(used)[0]:
                cell = tuple(int(tc[s, x]) for x in grp)
                h[cell] += 1
                if f[cell] < 0:
                    f[cell] = s
            hists.append(h)
            firsts.append(f)
        return hists, firsts, int(used.sum())

    def sfs_tables(self, kind, table, n_in, groups, outgroup=-1, site_mask=None):
        """same contract as Engine.sfs_tables: target allele of every row with the oracle, dense histograms + first rows"""
        table = np.asarray(table)
        n = table.shape[0]
        if kind == "base":
            tc, used = do.sfs_target_counts_from_counts(table, n_in, outgroup)
            dims = table.sum(axis=2).max(axis=0) + 1 if n else np.ones(table.shape[1], dtype=np.int64)
        else:
            tc, used = table, np.ones(n, dtype=bool)
            dims = table.max(axis=0) + 1 if n else np.ones(table.shape[1], dtype=np.int64)
        if site_mask is not None:
            used = used & np.asarray(site_mask, dtype=bool)
        hists, firsts = [], []
        for grp in groups:
            shape = tuple(int(dims[x]) for x in grp)
            h = np.zeros(shape, dtype=np.int64)
            f = np.full(shape, -1, dtype=np.int64)
            for s in np.where(used)[0]:
                cell = tuple(int(tc[s, x]) for x in grp)
                h[cell] += 1
                if f[cell] < 0:
                    f[cell] = s
            hists.append(h)
            firsts.append(f)
        return hists, firsts, int(used.sum())
