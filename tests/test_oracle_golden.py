"""CPU tests: the oracle (oracle/dense_oracle.py, oracle/ref_port.py) against fixtures produced by
running the unmodified reference (oracle/make_golden.py).  This is what 'pins' the oracle."""
import itertools
import json
import os
import warnings

import numpy as np
import pytest

from oracle import dense_oracle as do
from oracle import ref_port as rp
from helpers import GOLDEN, assert_close, gds_to_arrays, load_window_cases

ARR, META = load_window_cases()
IDS = [m["name"] for m in META]


@pytest.mark.parametrize("case", META, ids=IDS)
def test_pair_counts_match_reference(case):
    g = ARR[case["name"] + "__g_aln"]
    diff, n = do.pair_counts(g)
    d = do.dist_matrix(diff, n)
    assert_close(d, ARR[case["name"] + "__distMatrix"], "distMatrix", rtol=1e-15, atol=0)
    nn = n.copy()
    np.fill_diagonal(nn, 0)
    assert np.array_equal(nn, ARR[case["name"] + "__pairNonNan"])           # bit-exact integers


@pytest.mark.parametrize("case", META, ids=IDS)
def test_group_dist_stats(case):
    g = ARR[case["name"] + "__g_aln"]
    hap_pop = ARR[case["name"] + "__hap_pop"]
    P = len(case["pop_names"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pi, dxy, fst = do.group_dist_stats(g, hap_pop, P, case["minSites"], case["minData"])
    rpi, rdxy, rfst = gds_to_arrays(case["groupDistStats"], case["pop_names"])
    assert_close(pi, rpi, "pi", rtol=1e-13)
    assert_close(dxy, rdxy, "dxy", rtol=1e-13)
    assert_close(fst, rfst, "fst", rtol=1e-10)


@pytest.mark.parametrize("case", META, ids=IDS)
def test_closed_form_is_exact_when_not_ragged(case):
    g = ARR[case["name"] + "__g_aln"]
    hap_pop = ARR[case["name"] + "__hap_pop"]
    P = len(case["pop_names"])
    ok, pi, dxy, fst = do.group_dist_stats_closed_form(g, hap_pop, P, case["minSites"], case["minData"])
    used = hap_pop >= 0
    nv = (g[:, used] >= 0).sum(axis=1)
    assert ok == (not np.any((nv > 0) & (nv < used.sum())))
    if ok:
        rpi, rdxy, rfst = gds_to_arrays(case["groupDistStats"], case["pop_names"])
        assert_close(pi, rpi, "pi closed form", rtol=1e-12)
        assert_close(dxy, rdxy, "dxy closed form", rtol=1e-12)
        assert_close(fst, rfst, "fst closed form", rtol=1e-9)


def test_closed_form_covers_some_cases():
    n_ok = 0
    for case in META:
        ok, *_ = do.group_dist_stats_closed_form(ARR[case["name"] + "__g_aln"], ARR[case["name"] + "__hap_pop"],
                                                 len(case["pop_names"]), case["minSites"], case["minData"])
        n_ok += bool(ok)
    assert n_ok >= 3


@pytest.mark.parametrize("case", [m for m in META if m.get("groupFreqStats")], ids=[m["name"] for m in META if m.get("groupFreqStats")])
def test_group_freq_stats(case):
    g = ARR[case["name"] + "__g_aln"]
    r = do.group_freq_stats(g, ARR[case["name"] + "__hap_pop"], len(case["pop_names"]))
    for x, p in enumerate(case["pop_names"]):
        for k in ("l", "S", "thetaPi", "thetaW", "TajD"):
            assert_close(r[k][x], case["groupFreqStats"]["%s_%s" % (k, p)], "%s_%s" % (k, p), rtol=1e-12)


@pytest.mark.parametrize("case", META, ids=IDS)
def test_site_counts_bit_exact(case):
    g = ARR[case["name"] + "__g_aln"]
    sc = do.site_counts(g, ARR[case["name"] + "__hap_pop"], len(case["pop_names"]))
    assert np.array_equal(sc, ARR[case["name"] + "__site_counts"])


@pytest.mark.parametrize("case", [m for m in META if "ABBABABA" in m], ids=[m["name"] for m in META if "ABBABABA" in m])
def test_abbababa(case):
    g = ARR[case["name"] + "__g_aln"]
    hap_pop = ARR[case["name"] + "__hap_pop"]
    for md, refd in case["ABBABABA"].items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = do.abbababa(g, hap_pop, 0, 1, 2, 3, float(md))
        for k in ("D", "fd", "fdM", "ABBA", "BABA", "sitesUsed"):
            assert_close(r[k], refd[k], "%s minData=%s" % (k, md), rtol=1e-10)


def _hap_ind(case):
    # individual index (file order of sample names) of each alignment-order haplotype
    return np.array([case["sample_names"].index(s) for s in case["hap_samples"]], dtype=np.int32)


@pytest.mark.parametrize("case", META, ids=IDS)
@pytest.mark.parametrize("inc", [0, 1])
def test_ind_pair_dists(case, inc):
    g = ARR[case["name"] + "__g_aln"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = do.ind_pair_dists(g, _hap_ind(case), len(case["sample_names"]), bool(inc))
    assert_close(m, ARR[case["name"] + "__indPairDists_%d" % inc], "indPairDists", rtol=1e-13)


# ---------------- the loop-faithful port (what bench.py times as the CPU baseline) ----------------
SMALL = [m for m in META if ARR[m["name"] + "__g_aln"].shape[1] <= 30]


@pytest.mark.parametrize("case", SMALL, ids=[m["name"] for m in SMALL])
def test_port_group_dist_stats(case):
    g = ARR[case["name"] + "__g_aln"]
    hap_pop = ARR[case["name"] + "__hap_pop"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pi, dxy, fst = rp.popgen_window_port(g, hap_pop, len(case["pop_names"]), case["minSites"], case["minData"])
    rpi, rdxy, rfst = gds_to_arrays(case["groupDistStats"], case["pop_names"])
    assert_close(pi, rpi, "pi", rtol=1e-13)
    assert_close(dxy, rdxy, "dxy", rtol=1e-13)
    assert_close(fst, rfst, "fst", rtol=1e-10)


@pytest.mark.parametrize("case", [m for m in SMALL if "ABBABABA" in m], ids=[m["name"] for m in SMALL if "ABBABABA" in m])
def test_port_abbababa(case):
    g = ARR[case["name"] + "__g_aln"]
    aln = rp.PortAlignment(g, ARR[case["name"] + "__hap_pop"])
    for md, refd in case["ABBABABA"].items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = rp.abbababa_port(aln, 0, 1, 2, 3, float(md))
        for k in ("D", "fd", "fdM", "ABBA", "BABA", "sitesUsed"):
            assert_close(r[k], refd[k], "%s minData=%s" % (k, md), rtol=1e-12)


# ---------------- window generators ----------------
GEN = json.load(open(os.path.join(GOLDEN, "generator_cases.json")))


@pytest.mark.parametrize("idx", range(len(GEN["cases"])))
def test_window_generators(idx):
    case = GEN["cases"][idx]
    scaf, pos = GEN["scaffolds"], GEN["positions"]
    p = case["params"]
    if case["kind"] == "coordinate":
        wins = do.sliding_coord_windows(scaf, pos, p["windSize"], p["stepSize"], exclude=p.get("exclude"))
    elif case["kind"] == "sites":
        wins = do.sliding_sites_windows(scaf, pos, p["windSites"], p["overlap"],
                                        p["maxDist"] if p["maxDist"] else float("inf"), p["minSites"],
                                        exclude=p.get("exclude"))
    else:
        wins = do.predefined_coord_windows(scaf, pos, [tuple(c) for c in p["windCoords"]])
    assert len(wins) == len(case["windows"]), (len(wins), len(case["windows"]))
    for w, r in zip(wins, case["windows"]):
        assert w["scaffold"] == r["scaffold"]
        assert [pos[k] for k in w["sites"]] == r["positions"]
        if case["kind"] != "sites":
            assert [w["start"], w["end"]] == r["limits"]


@pytest.mark.parametrize("case", [m for m in META if m.get("groupFreqStats")], ids=[m["name"] for m in META if m.get("groupFreqStats")])
def test_alignment_mirror_group_freq_stats(case):
    """Alignment.groupFreqStats of the drop-in API (host logic on the oracle-backed engine) against the reference's dict"""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from genomics_general_b200 import genomics as G
    hp = ARR[case["name"] + "__hap_pop"]
    if np.any(hp < 0):
        pytest.skip("sequences outside every group")
    groups = [case["pop_names"][x] for x in hp]
    a = G.Alignment(ARR[case["name"] + "__g_aln"], groups=groups, engine=OracleEngine())
    got = a.groupFreqStats()
    assert set(got) == set(case["groupFreqStats"])
    for k, want in case["groupFreqStats"].items():
        assert_close(got[k], want, k, rtol=1e-12)
    assert all(isinstance(got["l_" + p], int) for p in case["pop_names"])
