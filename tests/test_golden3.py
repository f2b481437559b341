"""Round-2 parity corners against fixtures produced by the unmodified reference (oracle/make_golden3.py):
ABBABABA(polarize=False | fixed=True), sampleHet(minSites=k), the haplotype order of genoToAlignment with names that
sort differently as samples and as haplotypes (s1 / s10) under missing data and maxDist > 0, freq.py --ploidy with
--haploid, and genotypes piped on stdin.

Every test runs twice: on the CPU with the oracle-backed engine (host logic of the mirror / command lines) and, marked
gpu, on the real engine through the C-ABI."""
import hashlib
import io
import json
import os
import sys
import warnings

import numpy as np
import pytest

from helpers import GOLDEN, assert_close
from oracle_engine import OracleEngine

C3 = json.load(open(os.path.join(GOLDEN, "cases3.json")))
META2 = {m["name"]: m for m in json.load(open(os.path.join(GOLDEN, "window_cases2.json")))}
ARR2 = np.load(os.path.join(GOLDEN, "window_cases2.npz"))
TOL = dict(rtol=1e-9, atol=1e-12)
AB_KEYS = ["D", "fd", "fdM", "ABBA", "BABA", "sitesUsed"]


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def engine(request, monkeypatch):
    """an engine object for the Alignment mirror + the command lines patched to the same kind"""
    from genomics_general_b200.cli import _common, freq, popgenWindows
    if request.param == "oracle":
        for mod in (freq, popgenWindows):
            monkeypatch.setattr(mod, "Engine", OracleEngine)
        real = _common.load_geno
        monkeypatch.setattr(_common, "load_geno", lambda args, samples, pl, header=None, engine=None: real(args, samples, pl, header, None))
        yield OracleEngine()
    else:
        from genomics_general_b200.engine import Engine
        e = Engine(0)
        yield e
        e.close()


def _aln(engine, name, key="__g_aln"):
    from genomics_general_b200 import genomics as G
    m = META2[name]
    hp = ARR2[name + "__hap_pop"]
    groups = [m["pop_names"][x] if x >= 0 else None for x in hp]
    return G.Alignment(ARR2[name + key], names=m["hap_names"], groups=groups, sampleNames=m["hap_samples"], engine=engine)


@pytest.mark.parametrize("name", list(C3["abba_modes"]))
def test_abbababa_unpolarized_modes(engine, name):
    """genomics.py:1672-1677: fixed=True and the minor-allele mode (tie-free input, see make_golden2.py)"""
    from genomics_general_b200 import genomics as G
    for mode, key, kw in (("fixed", "__g_aln", dict(polarize=False, fixed=True)),
                          ("minor", "__g_aln_notie", dict(polarize=False, fixed=False))):
        for md in (0.0, 0.5, 1.0):
            want = C3["abba_modes"][name]["%s_%g" % (mode, md)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = G.ABBABABA(_aln(engine, name, key), "pop0", "pop1", "pop2", "pop3", md, **kw)
            assert_close([got[k] for k in AB_KEYS], [want[k] for k in AB_KEYS], "%s %s md=%g" % (name, mode, md), **TOL)


@pytest.mark.parametrize("name", list(C3["het_minsites"]))
def test_sample_het_min_sites(engine, name):
    """genomics.py:924: `len(x)==2 & n >= minSites` is a chained comparison — minSites > 2 leaves no value at all"""
    for k, want in C3["het_minsites"][name].items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = _aln(engine, name).sampleHet(minSites=int(k))
        assert list(got) == list(want)
        assert_close([got[s] for s in want], [want[s] for s in want], "%s minSites=%s" % (name, k), **TOL)


def test_haplotype_order_and_h12_with_ties(engine):
    """genoToAlignment sorts SEQUENCE names (s10_A < s1_A); H12's greedy clustering is order dependent"""
    from genomics_general_b200 import genomics as G
    d = C3["hap_order"]
    g = np.array(d["g"], dtype=np.int8)
    names = d["names"]
    lut = np.array(list("ACGTN"))
    ch = lut[np.where(g < 0, 4, g)]
    seqDict = {n: [ch[s, 2 * k] + "/" + ch[s, 2 * k + 1] for s in range(len(g))] for k, n in enumerate(names)}
    sd = G.SampleData(indNames=sorted(names), popNames=list(d["pops"]), popInds=[d["pops"][p] for p in d["pops"]])
    for key, want in d["H12stats"].items():
        parts = key.split("_")
        lo, hi, md = int(parts[0]), int(parts[1]), float(parts[2])
        sub = {n: v[lo:hi] for n, v in seqDict.items()}
        a = G.genoToAlignment(sub, sd, genoFormat="phased")
        a._eng = engine
        assert [str(x) for x in a.names] == d["aln_names"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if key.endswith("masked350"):
                a.groupDistStats(doPairs=True, minSites=350, minData=0.01)
            got = a.H12stats(maxDist=md)
        assert_close([got[k] for k in sorted(want)], [want[k] for k in sorted(want)], key, **TOL)


def _write_hap_inputs(tmp_path):
    from genomics_general_b200 import synth
    d = C3["hap_order"]
    path = str(tmp_path / "hap.geno")
    synth.write_geno(path, np.array(d["g"], dtype=np.int8), d["pos"], ["chr1"] * d["S"], d["names"])
    pops = str(tmp_path / "hap.pops")
    with open(pops, "wt") as f:
        for p, inds in d["pops"].items():
            for n in inds:
                f.write("%s %s\n" % (n, p))
    return path, pops


@pytest.mark.parametrize("key", ["cli_hapStats", "cli_popDist_hapStats", "cli_indPairDist_hapStats"])
def test_popgenWindows_hapstats_unsorted_header(engine, key, tmp_path):
    from genomics_general_b200.cli import popgenWindows
    path, pops = _write_hap_inputs(tmp_path)
    o = str(tmp_path / "o.csv")
    popgenWindows.main(["-g", path, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops, "-p", "popA", "-p", "popB",
                        "--windType", "sites", "-w", "400", "-m", "380", "--roundTo", "10"] + C3["hap_order"][key + "_args"])
    ours, ref = open(o).read().strip().split("\n"), C3["hap_order"][key].strip().split("\n")
    assert ours[0] == ref[0] and len(ours) == len(ref)
    for a, b in zip(ours[1:], ref[1:]):
        a, b = a.split(","), b.split(",")
        assert a[:5] == b[:5]
        assert_close([float(x) for x in a[5:]], [float(x) for x in b[5:]], "row " + a[1], rtol=1e-6, atol=2e-10)


def _freq_inputs(tmp_path, mixed):
    from genomics_general_b200 import synth
    c = C3["freq_cli"]["cfg"]
    spec = synth.SynthSpec(c["n_pops"], c["spp"], miss=c["miss"], seed=c["seed"])
    g = synth.synth_genotypes(spec, 0, c["S"])
    pos = synth.synth_positions(c["S"], seed=c["pos_seed"])
    names = spec.sample_names()
    path = str(tmp_path / "f.geno")
    if not mixed:
        synth.write_geno(path, g, pos, ["chr1"] * c["S"], names)
    else:       # --haploid samples carry one-letter tokens
        ch = np.array(list("ACGTN"))[np.where(g < 0, 4, g)]
        hap_idx = {names.index(n) for n in C3["freq_cli"]["haploid_samples"]}
        with open(path, "wt") as f:
            f.write("#CHROM\tPOS\t" + "\t".join(names) + "\n")
            for s in range(c["S"]):
                toks = [ch[s, 2 * k] if k in hap_idx else ch[s, 2 * k] + "/" + ch[s, 2 * k + 1] for k in range(len(names))]
                f.write("chr1\t%d\t%s\n" % (pos[s], "\t".join(toks)))
    pops = str(tmp_path / "f.pops")
    with open(pops, "wt") as f:
        for i, n in enumerate(names):
            f.write("%s pop%d\n" % (n, i // c["spp"]))
    return path, pops


def _same_text(path, want):
    txt = open(path).read().splitlines()
    assert txt[:200] == want["head"] and len(txt) == want["nlines"]
    assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == want["sha256"]


@pytest.mark.parametrize("key", ["ploidy_haploid", "ploidy_haploid_derived"])
def test_freq_ploidy_with_haploid_override(engine, key, tmp_path):
    """freq.py:283-284: --haploid applies after --ploidy as well"""
    from genomics_general_b200.cli import freq
    path, pops = _freq_inputs(tmp_path, mixed=True)
    o = str(tmp_path / "f.tsv")
    freq.main(["-g", path, "-o", o, "-f", "phased", "-t", "1", "--popsFile", pops, "-p", "pop0", "-p", "pop1", "-p", "pop2"]
              + C3["freq_cli"][key]["args"])
    _same_text(o, C3["freq_cli"][key])


def test_freq_reads_piped_genotypes(engine, tmp_path, monkeypatch):
    """no -g: header and genotypes come from stdin (freq.py:228-233)"""
    from genomics_general_b200.cli import _common, freq
    path, pops = _freq_inputs(tmp_path, mixed=False)
    o = str(tmp_path / "f.tsv")

    class _Stdin:
        buffer = io.BytesIO(open(path, "rb").read())
    monkeypatch.setattr(sys, "stdin", _Stdin())
    _common._STDIN.clear()
    freq.main(["-o", o, "-f", "phased", "-t", "1", "--popsFile", pops, "-p", "pop0", "-p", "pop1", "-p", "pop2"])
    _common._STDIN.clear()
    _same_text(o, C3["freq_cli"]["stdin"])


@pytest.mark.gpu
def test_gbin_cache_second_run_skips_tokenisation(tmp_path):
    """--cache: the second run of the command line loads <geno>.gbin (no text ingest) and writes the same rows"""
    from genomics_general_b200.cli import _common, popgenWindows
    path, pops = _write_hap_inputs(tmp_path)
    outs = []
    real = _common._load_geno_uncached
    calls = []

    def counting(*a, **k):
        calls.append(1)
        return real(*a, **k)
    _common._load_geno_uncached = counting
    try:
        for k in range(2):
            o = str(tmp_path / ("o%d.csv" % k))
            popgenWindows.main(["-g", path, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops, "-p", "popA", "-p", "popB",
                                "--windType", "sites", "-w", "400", "-m", "380", "--roundTo", "10", "--cache",
                                "--analysis", "popDist", "popPairDist", "hapStats", "--hapDist", "0.01"])
            outs.append(open(o).read())
    finally:
        _common._load_geno_uncached = real
    assert os.path.exists(path + ".gbin") and len(calls) == 1          # tokenised once
    assert outs[0] == outs[1] and outs[0].count("\n") > 2
