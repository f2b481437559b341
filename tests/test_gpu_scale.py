"""GPU tests at BASELINE-like sizes through size-independent properties (the oracle is too slow there):
permutation invariance, K1 == K2 on complete data, window additivity, bookkeeping identities."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

S_BIG = 2_000_000          # x 400 haplotypes = 0.8 GB on the device


@pytest.fixture(scope="module")
def eng():
    from genomics_general_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_bookkeeping_identities_at_scale(eng):
    from genomics_general_b200 import synth, windows
    spec = synth.SynthSpec(4, 50, miss=0.0, seed=42)
    eng.synth_fill(spec, S_BIG)
    eng.set_pops(spec.hap_pop(), 4)
    pos = synth.synth_positions(S_BIG, seed=42)
    ws = windows.sliding_coord_windows(np.zeros(S_BIG, dtype=np.int32), ["chr1"], pos, 50000)
    lo, hi = ws.ranges()
    eng.set_windows(lo, hi)
    r = eng.popgen(100, 0.01)
    assert r["sites"].sum() == S_BIG and np.array_equal(r["sites"], hi - lo)
    csum = np.concatenate([[0], np.cumsum(pos.astype(np.int64))])
    assert np.array_equal(r["pos_sum"], csum[hi] - csum[lo])
    assert np.all(r["path"] == 1)
    assert np.all((r["pi"] >= 0) & (r["pi"] <= 1)) and np.all((r["dxy"] >= 0) & (r["dxy"] <= 1))
    # one window over everything == site-weighted combination of the per-window sums (integer sums are additive)
    eng.set_windows([0], [S_BIG])
    whole = eng.popgen(100, 0.01)
    wgt = (hi - lo) / float(S_BIG)
    assert_close(whole["pi"][0], (r["pi"] * wgt[:, None]).sum(axis=0), "pi additivity", rtol=1e-9, atol=0)
    assert_close(whole["dxy"][0], (r["dxy"] * wgt[:, None]).sum(axis=0), "dxy additivity", rtol=1e-9, atol=0)


def test_k2_equals_k1_on_complete_data_at_scale(eng):
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(4, 50, miss=0.0, seed=43)
    S = 400_000
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, lo + 5000)
    a = eng.popgen(100, 0.01)
    b = eng.popgen(100, 0.01, force_pairwise=True)
    assert np.all(a["path"] == 1) and np.all(b["path"] == 2)
    for k in ("pi", "dxy"):
        assert_close(b[k], a[k], k, rtol=1e-12, atol=0)
    assert_close(b["fst"], a["fst"], "fst", rtol=1e-9, atol=1e-12)


def test_haplotype_permutation_invariance(eng):
    """Shuffling haplotype columns (and the population map with them) must not change any statistic:
    K1 integer sums are order-free (bit-identical); the pairwise block sums change summation order only."""
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(4, 25, miss=0.02, seed=44)
    S = 300_000
    g = synth.synth_genotypes(spec, 0, S)
    pos = synth.synth_positions(S)
    hp = spec.hap_pop()
    lo = np.arange(0, S, 10000, dtype=np.int64)
    perm = np.random.default_rng(0).permutation(spec.n_haps)
    res = []
    for cols in (np.arange(spec.n_haps), perm):
        eng.upload(np.ascontiguousarray(g[:, cols]), pos)
        eng.set_pops(hp[cols], 4)
        eng.set_windows(lo, lo + 10000)
        res.append((eng.popgen(100, 0.01), eng.abbababa(0, 1, 2, 3, 0.5), eng.site_counts(0, 50000)))
    (p0, a0, c0), (p1, a1, c1) = res
    assert np.array_equal(c0, c1)                                  # allele counts: bit-exact
    for k in ("ABBA", "BABA", "D", "fd", "fdM", "sitesUsed"):
        assert_close(a1[k], a0[k], k, rtol=1e-12, atol=1e-12)
    assert np.array_equal(p0["path"], p1["path"])
    for k in ("pi", "dxy"):
        assert_close(p1[k], p0[k], k, rtol=1e-12, atol=0)
    assert_close(p1["fst"], p0["fst"], "fst", rtol=1e-9, atol=1e-12)


def test_abba_counts_are_consistent_with_site_counts(eng):
    """sitesUsed recomputed on the host from pg_site_counts (integers) equals the kernel's own classification."""
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(4, 50, miss=0.05, seed=45)
    S = 500_000
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, lo + 5000)
    r = eng.abbababa(0, 1, 2, 3, 0.5)
    c = eng.site_counts().astype(np.int64)                         # [S,4,4]
    n = c.sum(axis=2)
    tot = c.sum(axis=1)
    good = ((tot > 0).sum(axis=1) == 2) & np.all(n / 100.0 >= 0.5, axis=1)
    derived = (tot > 0) & (c[:, 3, :] == 0) & (n[:, 3] > 0)[:, None]
    used = (good[:, None] & derived).sum(axis=1)
    exp = np.add.reduceat(used, lo)
    ngood = np.add.reduceat(good.astype(np.int64), lo)
    got = r["sitesUsed"]
    assert np.array_equal(np.isnan(got), ngood == 0)
    assert np.array_equal(got[~np.isnan(got)].astype(np.int64), exp[ngood > 0])


def test_sites_windows_shape_c5(eng):
    """Config-5 row length (8 pops x 100 diploid = 1600 haplotypes, 2 lanes per site) against the oracle on a
    few windows, the rest through determinism + K2 agreement."""
    import warnings
    from genomics_general_b200 import synth
    from oracle import dense_oracle as do
    spec = synth.SynthSpec(8, 100, miss=0.0, seed=46)
    S = 60_000
    eng.synth_fill(spec, S)
    hp = spec.hap_pop()
    eng.set_pops(hp, 8)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, lo + 5000)
    r = eng.popgen(5000, 0.01)
    g, _ = eng.download(0, 10000)
    for w in (0, 1):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pi, dxy, fst = do.group_dist_stats(g[w * 5000:(w + 1) * 5000], hp, 8, 5000, 0.01)
        assert_close(r["pi"][w], pi, "pi", rtol=1e-9, atol=0)
        assert_close(r["dxy"][w], dxy, "dxy", rtol=1e-9, atol=0)
        assert_close(r["fst"][w], fst, "fst", rtol=1e-8, atol=1e-12)
    eng.set_windows(lo[:3], lo[:3] + 5000)
    a = eng.popgen(5000, 0.01)
    b = eng.popgen(5000, 0.01, force_pairwise=True)
    assert_close(b["pi"], a["pi"], "pi K2", rtol=1e-12, atol=0)
    assert_close(b["dxy"], a["dxy"], "dxy K2", rtol=1e-12, atol=0)
