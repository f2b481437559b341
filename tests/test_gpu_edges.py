"""GPU edge cases: empty inputs, tiny shapes, very many / heavily overlapping windows, long rows."""
import warnings

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-9, atol=1e-12)


@pytest.fixture(scope="module")
def eng():
    from genomics_general_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _check_popgen(eng, g, hap_pop, P, lo, hi, min_sites=1, min_data=0.01, every=1):
    from oracle import dense_oracle as do
    r = eng.popgen(min_sites, min_data)
    assert np.array_equal(r["sites"], np.asarray(hi) - np.asarray(lo))
    for w in range(0, len(lo), every):
        if hi[w] - lo[w] < min_sites:
            assert r["path"][w] == 0
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pi, dxy, fst = do.group_dist_stats(g[lo[w]:hi[w]], hap_pop, P, min_sites, min_data)
        assert_close(r["pi"][w], pi, "pi w%d" % w, **TOL)
        assert_close(r["dxy"][w], dxy, "dxy w%d" % w, **TOL)
        assert_close(r["fst"][w], fst, "fst w%d" % w, rtol=1e-8, atol=1e-10)
    return r


def test_no_sites_and_no_windows(eng):
    eng.upload(np.zeros((0, 6), dtype=np.int8), np.zeros(0, dtype=np.int32))
    eng.set_pops(np.array([0, 0, 0, 1, 1, 1], dtype=np.int32), 2)
    eng.set_windows([0, 0], [0, 0])
    r = eng.popgen(1, 0.01)
    assert r["sites"].tolist() == [0, 0] and np.all(np.isnan(r["pi"])) and np.all(r["path"] == 0)
    eng.upload(np.zeros((5, 6), dtype=np.int8))
    eng.set_pops(np.array([0, 0, 0, 1, 1, 1], dtype=np.int32), 2)
    eng.set_windows([], [])
    r = eng.popgen(1, 0.01)
    assert r["pi"].shape == (0, 2)


def test_single_site_single_haplotype_rows(eng):
    from genomics_general_b200 import synth
    g = np.array([[0], [1], [-1], [3]], dtype=np.int8)
    eng.upload(g, np.array([3, 5, 9, 11], dtype=np.int32))
    eng.set_pops(np.array([0], dtype=np.int32), 1)
    eng.set_windows([0, 1, 0], [1, 2, 4])
    r = eng.popgen(1, 0.0)
    assert np.all(np.isnan(r["pi"]))                      # one haplotype: the only entry is the nan diagonal
    assert r["pos_sum"].tolist() == [3, 5, 28]
    assert np.array_equal(eng.site_counts()[:, 0, :], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 1]])


@pytest.mark.parametrize("miss", [0.0, 0.05])
def test_thousands_of_tiny_windows(eng, miss):
    """3-site windows: every tile holds dozens of segments (exercises the per-warp segment flush)."""
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(2, 6, miss=miss, seed=12)
    S = 20000
    g = synth.synth_genotypes(spec, 0, S)
    eng.upload(g, synth.synth_positions(S))
    eng.set_pops(spec.hap_pop(), 2)
    lo = np.arange(0, S - 3, 3, dtype=np.int64)
    hi = lo + 3
    eng.set_windows(lo, hi)
    _check_popgen(eng, g, spec.hap_pop(), 2, lo, hi, every=97)


def test_more_than_65535_windows(eng):
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(2, 4, miss=0.0, seed=13)
    S = 140000
    g = synth.synth_genotypes(spec, 0, S)
    eng.upload(g, synth.synth_positions(S))
    eng.set_pops(spec.hap_pop(), 2)
    lo = np.arange(0, S, 2, dtype=np.int64)                # 70000 windows
    hi = lo + 2
    eng.set_windows(lo, hi)
    _check_popgen(eng, g, spec.hap_pop(), 2, lo, hi, every=4999)


def test_heavily_overlapping_windows(eng):
    """-w 400 -s 7 style windows: each site belongs to ~57 windows; one pass still serves all of them."""
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(3, 5, miss=0.02, seed=14)
    S = 6000
    g = synth.synth_genotypes(spec, 0, S)
    eng.upload(g, synth.synth_positions(S))
    eng.set_pops(spec.hap_pop(), 3)
    lo = np.arange(0, S - 400, 7, dtype=np.int64)
    hi = lo + 400
    eng.set_windows(lo, hi)
    _check_popgen(eng, g, spec.hap_pop(), 3, lo, hi, min_sites=50, every=41)


def test_long_rows_four_lanes_per_site(eng):
    """3000 haplotypes per site: 4 lanes share one row (G = 4)."""
    from genomics_general_b200 import synth
    from genomics_general_b200.engine import k1_plan
    assert k1_plan(1000, 3000)["lanes_per_site"] >= 2
    spec = synth.SynthSpec(3, 500, miss=0.0, seed=15)
    S = 700
    g = synth.synth_genotypes(spec, 0, S)
    g[::50] = -1                                            # some all-missing sites
    eng.upload(g, synth.synth_positions(S))
    eng.set_pops(spec.hap_pop(), 3)
    lo = np.array([0, 350, 100], dtype=np.int64)
    hi = np.array([350, 700, 600], dtype=np.int64)
    eng.set_windows(lo, hi)
    r = _check_popgen(eng, g, spec.hap_pop(), 3, lo, hi, min_sites=10)
    assert np.all(r["path"] == 1)
    from oracle import dense_oracle as do
    assert np.array_equal(eng.site_counts().astype(np.int64), do.site_counts(g, spec.hap_pop(), 3))


def test_windows_that_skip_sites_and_repeat(eng):
    """Predefined-style windows: gaps between windows, identical windows listed twice, nested windows."""
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(2, 8, miss=0.01, seed=16)
    S = 5000
    g = synth.synth_genotypes(spec, 0, S)
    eng.upload(g, synth.synth_positions(S))
    eng.set_pops(spec.hap_pop(), 2)
    lo = np.array([100, 100, 900, 1000, 1200, 4000, 4990], dtype=np.int64)
    hi = np.array([600, 600, 2500, 1100, 1201, 4999, 5000], dtype=np.int64)
    eng.set_windows(lo, hi)
    r = _check_popgen(eng, g, spec.hap_pop(), 2, lo, hi)
    assert np.array_equal(r["pi"][0], r["pi"][1], equal_nan=True)


def test_more_than_eight_populations(eng):
    """P = 11: the site pass only does the bookkeeping, every window's statistics come from the pairwise path."""
    from genomics_general_b200 import synth
    for miss in (0.0, 0.04):
        spec = synth.SynthSpec(11, 3, miss=miss, seed=18)
        S = 2400
        g = synth.synth_genotypes(spec, 0, S)
        eng.upload(g, synth.synth_positions(S))
        eng.set_pops(spec.hap_pop(), 11)
        lo = np.array([0, 800, 1600, 100, 2390], dtype=np.int64)
        hi = np.array([800, 1600, 2400, 900, 2400], dtype=np.int64)
        eng.set_windows(lo, hi)
        r = _check_popgen(eng, g, spec.hap_pop(), 11, lo, hi, min_sites=20)
        assert r["path"].tolist() == [2, 2, 2, 2, 0]
