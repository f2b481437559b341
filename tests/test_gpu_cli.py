"""GPU tests of the drop-in layer: the four command lines and the genomics-compatible API, against the
outputs of the UNMODIFIED reference scripts / module committed in tests/golden (oracle/make_golden.py)."""
import hashlib
import io
import json
import os
import warnings

import numpy as np
import pytest

from helpers import GOLDEN, assert_close, load_window_cases

pytestmark = pytest.mark.gpu

CLI = json.load(open(os.path.join(GOLDEN, "cli_cases.json")))
ARR, META = load_window_cases()
BASES = np.array(list("ACGTN"))


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    """Re-create the .geno / .pops files the golden outputs were produced from (same spec + seed)."""
    from genomics_general_b200 import synth
    d = tmp_path_factory.mktemp("cli")
    out = {}
    for name, res in CLI.items():
        c = res["cfg"]
        spec = synth.SynthSpec(c["n_pops"], c["spp"], seed=c["seed"], miss=c["miss"])
        g = synth.synth_genotypes(spec, 0, c["S"])
        nsc = c.get("scaffolds", 1)
        per = c["S"] // nsc
        scafs, pos = [], []
        for k in range(nsc):
            n = per if k < nsc - 1 else c["S"] - per * (nsc - 1)
            scafs += ["chr%d" % (k + 1)] * n
            pos.append(synth.synth_positions(n, seed=c["seed"] + k))
        path = str(d / (name + ".geno"))
        synth.write_geno(path, g, np.concatenate(pos), scafs, spec.sample_names())
        pops = str(d / (name + ".pops"))
        with open(pops, "wt") as f:
            for i, n in enumerate(spec.sample_names()):
                f.write("%s pop%d\n" % (n, i // c["spp"]))
        out[name] = dict(geno=path, pops=pops, spec=spec, cfg=c, dir=str(d))
    return out


def _rows(text):
    lines = text.strip("\n").split("\n")
    return lines[0], [l.split(",") for l in lines[1:]]


def _compare_csv(ours, ref, n_prefix, rtol, atol):
    h1, r1 = _rows(ours)
    h2, r2 = _rows(ref)
    assert h1 == h2
    assert len(r1) == len(r2), (len(r1), len(r2))
    for a, b in zip(r1, r2):
        assert a[:n_prefix] == b[:n_prefix], (a[:n_prefix], b[:n_prefix])
        assert_close(np.array([float(x) for x in a[n_prefix:]]), np.array([float(x) for x in b[n_prefix:]]),
                     "row %s" % a[:3], rtol=rtol, atol=atol)


def _popargs(spec):
    a = []
    for p in spec.pop_names():
        a += ["-p", p]
    return a


@pytest.mark.parametrize("name", list(CLI))
def test_popgenWindows_cli(inputs, name):
    from genomics_general_b200.cli import popgenWindows
    i = inputs[name]
    c = i["cfg"]
    o = os.path.join(i["dir"], "o.csv")
    base = ["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1", "--popsFile", i["pops"]]
    popgenWindows.main(base + ["--roundTo", "12"] + _popargs(i["spec"]))
    _compare_csv(open(o).read(), CLI[name]["popgenWindows_roundTo12"], 5, 1e-6, 2e-12)
    popgenWindows.main(base + ["--writeFailedWindows"] + _popargs(i["spec"]))
    ours = open(o).read()
    _compare_csv(ours, CLI[name]["popgenWindows_default"], 5, 0, 1.0001e-4)
    same = sum(a == b for a, b in zip(ours.split("\n"), CLI[name]["popgenWindows_default"].split("\n")))
    assert same >= 0.98 * len(ours.split("\n"))           # 4-decimal strings are identical up to rare rounding ties


def test_popgenWindows_sites_windows_cli(inputs):
    from genomics_general_b200.cli import popgenWindows
    i = inputs["four_pops"]
    o = os.path.join(i["dir"], "o2.csv")
    popgenWindows.main(["--windType", "sites", "-w", "500", "-O", "100", "-m", "200", "-g", i["geno"], "-o", o, "-f", "phased",
                        "--popsFile", i["pops"], "--roundTo", "10"] + _popargs(i["spec"]))
    _compare_csv(open(o).read(), CLI["four_pops"]["popgenWindows_sites"], 5, 1e-6, 2e-10)


def test_ABBABABAwindows_cli(inputs):
    from genomics_general_b200.cli import ABBABABAwindows
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "ab.csv")
    ABBABABAwindows.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1",
                          "--popsFile", i["pops"], "--minData", "0.5", "-P1", "pop0", "-P2", "pop1", "-P3", "pop2", "-O", "pop3"])
    _compare_csv(open(o).read(), CLI["four_pops"]["ABBABABAwindows"], 6, 0, 1.0001e-4)


def test_freq_cli_bit_exact(inputs):
    from genomics_general_b200.cli import freq
    i = inputs["four_pops"]
    o = os.path.join(i["dir"], "f.tsv")
    freq.main(["-g", i["geno"], "-o", o, "-f", "phased", "-t", "1", "--popsFile", i["pops"]] + _popargs(i["spec"]))
    txt = open(o).read().splitlines()
    assert txt[:400] == CLI["four_pops"]["freq_head"]
    assert len(txt) == CLI["four_pops"]["freq_nlines"]
    assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == CLI["four_pops"]["freq_sha256"]


def _floats(text):
    return np.array([float(x) for x in text.split() if x.replace(".", "", 1).replace("e-", "", 1).replace("nan", "0").isdigit()
                     or x == "nan"])


def test_distMat_cli(inputs):
    from genomics_general_b200.cli import distMat
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "d.txt")
    distMat.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1",
                  "--outFormat", "raw", "--roundTo", "10"])
    ours, ref = open(o).read(), CLI["four_pops"]["distMat_raw"]
    assert ours.count("\n") == ref.count("\n")
    assert_close(_floats(ours), _floats(ref), "raw", rtol=1e-6, atol=2e-10)
    distMat.main(["--windType", "cat", "-g", i["geno"], "-o", o, "-f", "phased", "--outFormat", "phylip", "--roundTo", "8"])
    ours, ref = open(o).read(), CLI["four_pops"]["distMat_cat_phylip"]
    assert [l.split()[0] for l in ours.splitlines()] == [l.split()[0] for l in ref.splitlines()]
    a = np.array([[float(x) for x in l.split()[1:]] for l in ours.splitlines()[1:]])
    b = np.array([[float(x) for x in l.split()[1:]] for l in ref.splitlines()[1:]])
    assert_close(a, b, "phylip", rtol=1e-6, atol=2e-8)


# ------------------------------------------------------------------------------------------------
# genomics-compatible API, used the way the reference's workers use it (popgenWindows.py:44-52)
# ------------------------------------------------------------------------------------------------
def _tokens(g, ploidies):
    ch = BASES[np.where(g < 0, 4, g)]
    toks, h = [], 0
    cols = []
    for p in ploidies:
        cols.append(["/".join(ch[s, h:h + p]) for s in range(g.shape[0])])
        h += p
    return cols


@pytest.mark.parametrize("case", META, ids=[m["name"] for m in META])
def test_genomics_api_like_the_reference_worker(case):
    from genomics_general_b200 import genomics
    g = ARR[case["name"] + "__g_file"]
    names, ploidies = case["sample_names"], case["ploidies"]
    sd = genomics.SampleData(indNames=list(names), popNames=list(case["pop_names"]),
                             popInds=[list(p) for p in case["pop_inds"]], ploidyDict=dict(zip(names, ploidies)))
    seqDict = dict(zip(names, _tokens(g, ploidies)))
    aln = genomics.genoToAlignment(seqDict, sd, genoFormat="phased")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gds = aln.groupDistStats(doPairs=True, minSites=case["minSites"], minData=case["minData"])
    assert set(gds) == set(case["groupDistStats"])
    for k, v in case["groupDistStats"].items():
        assert_close(gds[k], v, k, rtol=1e-9, atol=1e-12)
    # a fresh alignment, as in make_golden.py: after groupDistStats the reference's cached matrix stays masked in place
    pdd = genomics.genoToAlignment(seqDict, sd, genoFormat="phased").indPairDists(includeSameWithSame=False)
    m = np.array([[pdd[a][b] for b in names] for a in names])
    assert_close(m, ARR[case["name"] + "__indPairDists_0"], "indPairDists", rtol=1e-9, atol=1e-12)
    # haplotype-level matrices: our row order is file order, the reference sorts by name
    order = [list(aln.names).index(n) for n in case["hap_names"]]
    dm = aln.distMatrix()[np.ix_(order, order)]
    assert_close(dm, ARR[case["name"] + "__distMatrix"], "distMatrix", rtol=1e-15, atol=0)
    assert np.array_equal(aln.pairNonNan()[np.ix_(order, order)], ARR[case["name"] + "__pairNonNan"])
    for x, p in enumerate(case["pop_names"]):
        assert np.array_equal(aln.subset(groups=[p]).siteFreqs(asCounts=True), ARR[case["name"] + "__site_counts"][:, x, :])
    if "ABBABABA" in case:
        for md, refd in case["ABBABABA"].items():
            r = genomics.ABBABABA(aln, "pop0", "pop1", "pop2", "pop3", float(md))
            for k in ("D", "fd", "fdM", "ABBA", "BABA", "sitesUsed"):
                assert_close(r[k], refd[k], "%s minData=%s" % (k, md), rtol=1e-9, atol=1e-12)


def test_window_generators_yield_reference_windows(inputs):
    """slidingCoordWindows on a real file: same windows (scaffold, limits, positions) as the reference CLI rows imply."""
    from genomics_general_b200 import genomics
    i = inputs["four_pops"]
    c = i["cfg"]
    wins = list(genomics.slidingCoordWindows(i["geno"], c["w"], c["w"], names=i["spec"].sample_names()))
    _, rows = _rows(CLI["four_pops"]["popgenWindows_default"])
    assert len(wins) == len(rows)
    for w, r in zip(wins, rows):
        assert [w.scaffold, str(w.limits[0]), str(w.limits[1]), str(w.midPos()), str(w.seqLen())] == r[:5]
    sd = genomics.SampleData(indNames=i["spec"].sample_names(), popNames=i["spec"].pop_names(),
                             popInds=[[n for n in i["spec"].sample_names() if n.startswith("p%d_" % k)] for k in range(4)])
    aln = genomics.genoToAlignment(wins[1].seqDict(), sd, genoFormat="phased")
    gds = aln.groupDistStats(minSites=c["m"], minData=0.01)
    h, _ = _rows(CLI["four_pops"]["popgenWindows_roundTo12"])
    _, rows12 = _rows(CLI["four_pops"]["popgenWindows_roundTo12"])
    for name, val in zip(h.split(",")[5:], rows12[1][5:]):
        assert_close(gds[name], float(val), name, rtol=1e-6, atol=2e-12)


def test_popgenWindows_popFreq_and_indPairDist_cli(inputs):
    """--analysis popFreq popDist popPairDist indPairDist: same header (column order) and rows as the reference."""
    from genomics_general_b200.cli import popgenWindows
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "pf.csv")
    popgenWindows.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1",
                        "--popsFile", i["pops"], "--roundTo", "8", "--analysis", "popFreq", "popDist", "popPairDist",
                        "indPairDist"] + _popargs(i["spec"]))
    ours, ref = open(o).read(), CLI["four_pops"]["popgenWindows_popFreq_indPairDist"]
    _compare_csv(ours, ref, 5, 1e-6, 2e-8)
    # integer columns (l_, S_) are printed without a decimal point, exactly like the reference
    h, r1 = _rows(ours)
    _, r2 = _rows(ref)
    cols = [k for k, n in enumerate(h.split(",")) if n.startswith(("l_", "S_"))]
    assert cols and all(a[k] == b[k] for a, b in zip(r1, r2) for k in cols)


# ------------------------------------------------------------------------------------------------
# more flags, all against outputs of the unmodified reference scripts
# ------------------------------------------------------------------------------------------------
def test_popgenWindows_predefined_gz_windowID(inputs):
    import gzip
    from genomics_general_b200.cli import popgenWindows
    i = inputs["four_pops"]
    gz = i["geno"] + ".gz"
    with open(i["geno"], "rb") as fi, gzip.open(gz, "wb") as fo:
        fo.write(fi.read())
    coords = os.path.join(i["dir"], "coords.txt")
    open(coords, "wt").write(CLI["four_pops"]["coords_file"])
    o = os.path.join(i["dir"], "pre.csv")
    popgenWindows.main(["--windType", "predefined", "--windCoords", coords, "-m", "10", "-g", gz, "-o", o, "-f", "phased",
                        "--popsFile", i["pops"], "--roundTo", "9", "--addWindowID", "--writeFailedWindows"] + _popargs(i["spec"]))
    _compare_csv(open(o).read(), CLI["four_pops"]["popgenWindows_predefined_gz_id"], 6, 1e-6, 2e-9)


def test_popgenWindows_diplo_format_and_step(inputs):
    from genomics_general_b200 import synth
    from genomics_general_b200.cli import popgenWindows
    i = inputs["four_pops"]
    c = i["cfg"]
    spec = i["spec"]
    g = synth.synth_genotypes(spec, 0, c["S"])
    per = c["S"] // 3
    scafs, pos = [], []
    for k in range(3):
        n = per if k < 2 else c["S"] - 2 * per
        scafs += ["chr%d" % (k + 1)] * n
        pos.append(synth.synth_positions(n, seed=c["seed"] + k))
    dpath = os.path.join(i["dir"], "diplo.geno")
    synth.write_geno(dpath, g, np.concatenate(pos), scafs, spec.sample_names(), fmt="diplo")
    o = os.path.join(i["dir"], "dip.csv")
    popgenWindows.main(["-w", str(c["w"]), "-s", "10000", "-m", str(c["m"]), "-g", dpath, "-o", o, "-f", "diplo",
                        "--popsFile", i["pops"], "--roundTo", "9"] + _popargs(spec))
    _compare_csv(open(o).read(), CLI["four_pops"]["popgenWindows_diplo_step"], 5, 1e-6, 2e-9)


def test_ABBABABAwindows_sites_overlap_failed_windows(inputs):
    from genomics_general_b200.cli import ABBABABAwindows
    i = inputs["four_pops"]
    o = os.path.join(i["dir"], "ab2.csv")
    ABBABABAwindows.main(["--windType", "sites", "-w", "1000", "--overlap", "250", "-m", "100", "-g", i["geno"], "-o", o,
                          "-f", "phased", "--popsFile", i["pops"], "--minData", "0.9", "-P1", "pop1", "-P2", "pop0",
                          "-P3", "pop2", "-O", "pop3", "--writeFailedWindows", "--addWindowID"])
    _compare_csv(open(o).read(), CLI["four_pops"]["ABBABABAwindows_sites_overlap"], 7, 0, 1.0001e-4)


def test_distMat_nexus_subset_windowData(inputs):
    from genomics_general_b200.cli import distMat
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "d.nex")
    wd = os.path.join(i["dir"], "wd.txt")
    distMat.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "--outFormat", "nexus",
                  "--roundTo", "7", "--includeSameWithSame", "--windowDataOutFile", wd, "--samples"]
                 + CLI["four_pops"]["distMat_subset_samples"])
    assert open(wd).read() == CLI["four_pops"]["distMat_windowData"]
    ours, ref = open(o).read().splitlines(), CLI["four_pops"]["distMat_nexus_subset"].splitlines()
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        if a.startswith("[") and "    " in a:
            la, va = a.split("    ", 1)
            lb, vb = b.split("    ", 1)
            assert la == lb
            assert_close(np.array([float(x) for x in va.split()]), np.array([float(x) for x in vb.split()]), la,
                         rtol=1e-6, atol=2e-7)
        else:
            assert a == b


def test_freq_indFreqs_bit_exact(inputs):
    from genomics_general_b200.cli import freq
    i = inputs["four_pops"]
    o = os.path.join(i["dir"], "fi.tsv")
    freq.main(["-g", i["geno"], "-o", o, "-f", "phased", "-t", "1", "--indFreqs"])
    txt = open(o).read().splitlines()
    assert txt[:50] == CLI["four_pops"]["freq_indFreqs_head"]
    assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == CLI["four_pops"]["freq_indFreqs_sha256"]
