"""Seeded differential fuzz of the C-ABI against the oracle: random shapes (populations of 1..300 haplotypes, interleaved
columns, unused haplotypes, ploidy 1/2), random windows (empty, single-site, overlapping), random missingness."""
import warnings

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-9, atol=1e-12)
FP_KEYS = ('fhom', "fhom'", 'D', 'fd', "fd'", 'fdm', "fdm'", 'fdh', 'fdh2', 'fh', "ABBA", "BABA", "ABAA", "BAAA")


@pytest.fixture(scope="module")
def eng():
    from genomics_general_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _case(rng):
    P = int(rng.integers(1, 7))
    sizes = [int(rng.choice([1, 2, 3, 5, 8, 17, 40, 130, 300], p=[.1, .1, .1, .15, .15, .15, .15, .05, .05])) for _ in range(P)]
    unused = int(rng.integers(0, 4))
    hap_pop = np.concatenate([np.full(n, x) for x, n in enumerate(sizes)] + [np.full(unused, -1)]).astype(np.int32)
    if rng.random() < 0.5:
        hap_pop = rng.permutation(hap_pop)
    H = len(hap_pop)
    L = int(rng.integers(1, 700))
    ref = rng.integers(0, 4, L)
    alt = (ref + rng.integers(1, 4, L)) % 4
    var = rng.random(L) < rng.choice([0.05, 0.3, 0.9])
    freq = rng.random((L, P + 1)) * var[:, None]
    g = np.where(rng.random((L, H)) < freq[:, hap_pop], alt[:, None], ref[:, None]).astype(np.int8)
    third = (rng.random((L, H)) < 0.2) & (rng.random(L) < 0.05)[:, None]
    g[third] = (g[third] + 1) % 4
    miss = rng.choice([0.0, 0.0, 0.02, 0.3])
    if miss:
        g[rng.random((L, H)) < miss] = -1
    if rng.random() < 0.3:
        g[rng.random(L) < 0.2] = -1
    nw = int(rng.integers(1, 6))
    lo = np.sort(rng.integers(0, L + 1, nw)).astype(np.int64)
    hi = np.minimum(L, lo + rng.integers(0, L + 1, nw)).astype(np.int64)
    return P, hap_pop, g, lo, hi


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_against_oracle(eng, seed):
    from oracle import dense_oracle as do
    rng = np.random.default_rng(1000 + seed)
    P, hap_pop, g, lo, hi = _case(rng)
    L, H = g.shape
    pos = np.cumsum(rng.integers(1, 50, L)).astype(np.int32)
    eng.upload(g, pos)
    eng.set_pops(hap_pop, P)
    eng.set_windows(lo, hi)
    min_sites = int(rng.choice([0, 1, 5, 50]))
    min_data = float(rng.choice([0.0, 0.01, 0.5, 0.9]))
    r = eng.popgen(min_sites, min_data)
    hap_ind = (np.arange(H) // 2).astype(np.int32)
    n_ind = (H + 1) // 2
    dm = eng.pairdist(hap_ind, n_ind, False, min_sites=min_sites)["dist"]
    het = eng.ind_het(hap_ind, n_ind, min_sites=min_sites)
    md = float(rng.choice([0.0, 0.01, 0.2]))
    dn = bool(rng.integers(0, 2))
    hs = eng.hapstats(md, min_sites=min_sites, diag_nan=dn)
    cnt = eng.site_counts()
    assert np.array_equal(cnt.astype(np.int64), do.site_counts(g, hap_pop, P))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for w in range(len(lo)):
            gw = g[lo[w]:hi[w]]
            assert r["sites"][w] == hi[w] - lo[w]
            assert r["pos_sum"][w] == int(pos[lo[w]:hi[w]].astype(np.int64).sum())
            if hi[w] - lo[w] < max(min_sites, 0) or hi[w] == lo[w]:
                continue
            if r["path"][w] != 0:
                pi, dxy, fst = do.group_dist_stats(gw, hap_pop, P, min_sites or None, min_data)
                assert_close(r["pi"][w], pi, "pi seed %d w %d" % (seed, w), **TOL)
                assert_close(r["dxy"][w], dxy, "dxy", **TOL)
                assert_close(r["fst"][w], fst, "fst", rtol=1e-7, atol=1e-11)
            assert_close(dm[w], do.ind_pair_dists(gw, hap_ind, n_ind, False, min_sites or None), "indpair", **TOL)
            assert_close(het[w], do.sample_het(gw, hap_ind, n_ind, min_sites or None), "het", **TOL)
            assert_close(hs[w], do.h12_stats(gw, hap_pop, P, md, min_sites or None, dn), "h12", **TOL)
    if P >= 4:
        sel = [int(x) for x in rng.permutation(P)[:4]]
        tot = np.stack([((g == a) & (np.isin(hap_pop, sel))[None, :]).sum(axis=1) for a in range(4)], axis=1)
        srt = np.sort(tot, axis=1)
        tied = (srt[:, 2] > 0) & (srt[:, 2] == srt[:, 3])
        for kw in ({}, dict(polarize=True), dict(fixed=True)):
            g2 = g.copy()
            if not kw:
                g2[tied] = -1            # default mode: exact ties are implementation-defined in the reference
                eng.upload(g2, pos)
                eng.set_pops(hap_pop, P)
                eng.set_windows(lo, hi)
            mdv = float(rng.choice([0.0, 0.3, 1.0]))
            fp = eng.fourpop(*sel, mdv, **kw)
            ab = eng.abbababa(*sel, mdv) if kw == dict(polarize=True) else None
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for w in range(len(lo)):
                    want = do.four_pop(g2[lo[w]:hi[w]], hap_pop, *sel, mdv, **kw)
                    assert fp["sitesUsed"][w] == want["sitesUsed"], (seed, w, kw)
                    assert_close([fp[k][w] for k in FP_KEYS], [want[k] for k in FP_KEYS], "fourpop %s seed %d w %d" % (kw, seed, w),
                                 rtol=1e-8, atol=1e-11)
                    if ab is not None:
                        wa = do.abbababa(g2[lo[w]:hi[w]], hap_pop, *sel, mdv)
                        assert_close([ab[k][w] for k in ("ABBA", "BABA", "D", "fd", "fdM")],
                                     [wa[k] for k in ("ABBA", "BABA", "D", "fd", "fdM")], "abba", rtol=1e-8, atol=1e-11)
            if not kw:
                eng.upload(g, pos)
                eng.set_pops(hap_pop, P)
                eng.set_windows(lo, hi)
    if P >= 2:
        for target in ("derived", "minor"):
            v, tie = eng.site_target_freqs(target, min_data=float(rng.choice([0, 2])), as_counts=bool(rng.integers(0, 2)))
        v, tie = eng.site_target_freqs("minor")
        want, wtie = do.target_freqs(g, hap_pop, P, "minor")
        assert np.array_equal(tie, wtie)
        assert np.array_equal(v[~tie], want[~tie], equal_nan=True)
        v, _ = eng.site_target_freqs("derived", min_data=1.0)
        want, _ = do.target_freqs(g, hap_pop, P, "derived", min_data=1.0)
        assert np.array_equal(v, want, equal_nan=True)
