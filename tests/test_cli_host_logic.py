"""Host logic of the drop-in command lines on the CPU: every script runs with the engine replaced by the oracle
(tests/oracle_engine.py) and its rows are compared with the reference scripts' own output (tests/golden).  What this
covers is everything AROUND the kernels — flags, populations / ploidy, the native text tokenizer, window generation,
row prefixes, rounding and number formatting, failed-window handling; the GPU tests run the same commands on the real
engine."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_close
from oracle_engine import OracleEngine

CLI = json.load(open(os.path.join(GOLDEN, "cli_cases.json")))
CLI2 = json.load(open(os.path.join(GOLDEN, "cli_cases2.json")))["four_pops"]


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    from genomics_general_b200 import synth
    d = tmp_path_factory.mktemp("cli_cpu")
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
        popargs = []
        for p in spec.pop_names():
            popargs += ["-p", p]
        out[name] = dict(geno=path, pops=pops, spec=spec, cfg=c, dir=str(d), popargs=popargs)
    return out


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    """every command line of this module talks to the oracle instead of the GPU and tokenises on the host"""
    from genomics_general_b200.cli import ABBABABAwindows, _common, distMat, fourPopWindows, freq, popgenWindows, sfs
    for mod in (ABBABABAwindows, distMat, fourPopWindows, freq, popgenWindows, sfs):
        monkeypatch.setattr(mod, "Engine", OracleEngine)
    real = _common.load_geno
    monkeypatch.setattr(_common, "load_geno", lambda args, samples, pl, header=None, engine=None: real(args, samples, pl, header, None))


def _table(text):
    lines = text.strip("\n").split("\n")
    hdr = lines[0].split(",")
    return hdr, [dict(zip(hdr, l.split(","))) for l in lines[1:]]


def _same_rows(ours, ref, n_prefix, atol, rtol=1e-6):
    h1, r1 = _table(ours)
    h2, r2 = _table(ref)
    assert sorted(h1) == sorted(h2)
    assert len(r1) == len(r2), (len(r1), len(r2))
    for a, b in zip(r1, r2):
        for k in h2[:n_prefix]:
            assert a[k] == b[k], (k, a[k], b[k])
        keys = h2[n_prefix:]
        assert_close([float(a[k]) for k in keys], [float(b[k]) for k in keys], "row " + a["start"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", list(CLI))
def test_popgenWindows_default_rows(inputs, name):
    from genomics_general_b200.cli import popgenWindows
    i = inputs[name]
    c = i["cfg"]
    o = os.path.join(i["dir"], "o.csv")
    base = ["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1", "--popsFile", i["pops"]]
    popgenWindows.main(base + ["--roundTo", "12"] + i["popargs"])
    _same_rows(open(o).read(), CLI[name]["popgenWindows_roundTo12"], 5, 2e-12)
    popgenWindows.main(base + ["--writeFailedWindows"] + i["popargs"])
    ours, ref = open(o).read(), CLI[name]["popgenWindows_default"]
    _same_rows(ours, ref, 5, 1.0001e-4, rtol=0)
    assert sum(a == b for a, b in zip(ours.split("\n"), ref.split("\n"))) >= 0.98 * len(ref.split("\n"))


def test_popgenWindows_window_types_and_analyses(inputs):
    import gzip
    from genomics_general_b200.cli import popgenWindows
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "o2.csv")
    common = ["-g", i["geno"], "-o", o, "-f", "phased", "--popsFile", i["pops"]] + i["popargs"]
    popgenWindows.main(["--windType", "sites", "-w", "500", "-O", "100", "-m", "200", "--roundTo", "10"] + common)
    _same_rows(open(o).read(), CLI["four_pops"]["popgenWindows_sites"], 5, 2e-10)
    popgenWindows.main(["-w", str(c["w"]), "-m", str(c["m"]), "--roundTo", "8", "--analysis", "popFreq", "popDist", "popPairDist",
                        "indPairDist"] + common)
    ours, ref = open(o).read(), CLI["four_pops"]["popgenWindows_popFreq_indPairDist"]
    _same_rows(ours, ref, 5, 2e-8)
    h, r1 = _table(ours)
    _, r2 = _table(ref)
    ints = [n for n in h if n.startswith(("l_", "S_"))]
    assert ints and all(a[k] == b[k] for a, b in zip(r1, r2) for k in ints)        # printed without a decimal point
    gz = i["geno"] + ".gz"
    with open(i["geno"], "rb") as fi, gzip.open(gz, "wb") as fo:
        fo.write(fi.read())
    coords = os.path.join(i["dir"], "coords.txt")
    open(coords, "wt").write(CLI["four_pops"]["coords_file"])
    popgenWindows.main(["--windType", "predefined", "--windCoords", coords, "-m", "10", "-g", gz, "-o", o, "-f", "phased",
                        "--popsFile", i["pops"], "--roundTo", "9", "--addWindowID", "--writeFailedWindows"] + i["popargs"])
    _same_rows(open(o).read(), CLI["four_pops"]["popgenWindows_predefined_gz_id"], 6, 2e-9)
    for key, extra in (("popgen_indHet_alone", ["--windType", "sites", "-w", "300", "-m", "290", "--analysis", "indHet"]),
                       ("popgen_popDist_indHet_hapStats", ["--windType", "sites", "-w", "300", "-m", "290", "--analysis", "popDist",
                                                           "indHet", "hapStats", "--hapDist", "0.05"]),
                       ("popgen_indPairDist_hapStats_indHet", ["-w", "20000", "-m", "50", "--analysis", "indPairDist", "hapStats",
                                                               "indHet", "--hapDist", "0.08"])):
        popgenWindows.main(["--roundTo", "8"] + extra + common)
        _same_rows(open(o).read(), CLI2[key], 5, 2e-8)


def test_abba_and_fourpop_rows(inputs):
    from genomics_general_b200.cli import ABBABABAwindows, fourPopWindows
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "ab.csv")
    pops4 = ["-P1", "pop0", "-P2", "pop1", "-P3", "pop2", "-O", "pop3"]
    ABBABABAwindows.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1",
                          "--popsFile", i["pops"], "--minData", "0.5"] + pops4)
    _same_rows(open(o).read(), CLI["four_pops"]["ABBABABAwindows"], 6, 1.0001e-4, rtol=0)
    ABBABABAwindows.main(["--windType", "sites", "-w", "1000", "--overlap", "250", "-m", "100", "-g", i["geno"], "-o", o,
                          "-f", "phased", "--popsFile", i["pops"], "--minData", "0.9", "-P1", "pop1", "-P2", "pop0",
                          "-P3", "pop2", "-O", "pop3", "--writeFailedWindows", "--addWindowID"])
    _same_rows(open(o).read(), CLI["four_pops"]["ABBABABAwindows_sites_overlap"], 7, 1.0001e-4, rtol=0)
    fourPopWindows.main(["-g", i["geno"], "-o", o, "-f", "phased", "-T", "1", "--popsFile", i["pops"], "-w", "20000", "-m", "50",
                         "--minData", "0.5", "--polarize"] + pops4)
    _same_rows(open(o).read(), CLI2["fourPopWindows_polarize"], 6, 1.0001e-4, rtol=0)
    fourPopWindows.main(["-g", i["geno"], "-o", o, "-f", "phased", "-T", "1", "--popsFile", i["pops"], "--windType", "sites", "-w",
                         "1000", "--overlap", "250", "-m", "20", "--minData", "0.9", "--fixed", "--writeFailedWindows",
                         "--addWindowID"] + pops4)
    _same_rows(open(o).read(), CLI2["fourPopWindows_fixed_sites"], 7, 1.0001e-4, rtol=0)


def test_freq_rows_bit_exact(inputs):
    from genomics_general_b200.cli import freq
    i = inputs["four_pops"]
    o = os.path.join(i["dir"], "f.tsv")
    base = ["-g", i["geno"], "-o", o, "-f", "phased", "-t", "1", "--popsFile", i["pops"]] + i["popargs"]
    freq.main(base)
    txt = open(o).read().splitlines()
    assert txt[:400] == CLI["four_pops"]["freq_head"] and len(txt) == CLI["four_pops"]["freq_nlines"]
    assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == CLI["four_pops"]["freq_sha256"]
    for key, extra in (("freq_derived", ["--target", "derived"]), ("freq_derived_counts", ["--target", "derived", "--asCounts"]),
                       ("freq_derived_keepnan_mindata", ["--target", "derived", "--keepNanLines", "--minData", "11"]),
                       ("freq_derived_threshold", ["--target", "derived", "--threshold", "0.5"])):
        freq.main(base + extra)
        txt = open(o).read().splitlines()
        assert txt[:300] == CLI2[key + "_head"] and len(txt) == CLI2[key + "_nlines"], key
        assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == CLI2[key + "_sha256"], key
    freq.main(["-g", i["geno"], "-o", o, "-f", "phased", "-t", "1", "--indFreqs"])
    txt = open(o).read().splitlines()
    assert txt[:50] == CLI["four_pops"]["freq_indFreqs_head"]
    assert hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest() == CLI["four_pops"]["freq_indFreqs_sha256"]


def _floats(text):
    out = []
    for x in text.split():
        try:
            out.append(float(x))
        except ValueError:
            pass
    return np.array(out)


def test_distMat_outputs(inputs):
    from genomics_general_b200.cli import distMat
    i = inputs["four_pops"]
    c = i["cfg"]
    o = os.path.join(i["dir"], "d.txt")
    distMat.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "-T", "1", "--outFormat", "raw",
                  "--roundTo", "10"])
    ours, ref = open(o).read(), CLI["four_pops"]["distMat_raw"]
    assert ours.count("\n") == ref.count("\n")
    assert_close(_floats(ours), _floats(ref), "raw", rtol=1e-6, atol=2e-10)
    distMat.main(["--windType", "cat", "-g", i["geno"], "-o", o, "-f", "phased", "--outFormat", "phylip", "--roundTo", "8"])
    ours, ref = open(o).read(), CLI["four_pops"]["distMat_cat_phylip"]
    assert [l.split()[0] for l in ours.splitlines()] == [l.split()[0] for l in ref.splitlines()]
    assert_close(_floats(ours), _floats(ref), "phylip", rtol=1e-6, atol=2e-8)
    wd = os.path.join(i["dir"], "wd.txt")
    distMat.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o, "-f", "phased", "--outFormat", "nexus",
                  "--roundTo", "7", "--includeSameWithSame", "--windowDataOutFile", wd, "--samples"]
                 + CLI["four_pops"]["distMat_subset_samples"])
    assert open(wd).read() == CLI["four_pops"]["distMat_windowData"]
    ours, ref = open(o).read().splitlines(), CLI["four_pops"]["distMat_nexus_subset"].splitlines()
    assert len(ours) == len(ref)
    assert [l for l in ours if not l.startswith("[")] == [l for l in ref if not l.startswith("[")]
    assert_close(_floats("\n".join(l.split("    ", 1)[1] for l in ours if l.startswith("[") and "    " in l)),
                 _floats("\n".join(l.split("    ", 1)[1] for l in ref if l.startswith("[") and "    " in l)), "nexus",
                 rtol=1e-6, atol=2e-7)


@pytest.mark.parametrize("key", [k for k in CLI2 if k.startswith("sfs_") and not k.startswith(("sfs_base", "sfs_target"))
                                 and k + "_args" in CLI2])
def test_sfs_genotype_command_line(key, tmp_path, capsys):
    from genomics_general_b200 import synth
    from genomics_general_b200.cli import sfs as sfs_cli
    from test_oracle_golden2 import sfs_inputs
    c = CLI2["sfs_cfg"]
    spec, g, scaf = sfs_inputs()
    path = str(tmp_path / "sfs.geno")
    synth.write_geno(path, g, synth.synth_positions(c["S"], seed=c["seed"]), ["chr%d" % (k + 1) for k in scaf], spec.sample_names())
    pops = str(tmp_path / "sfs.pops")
    with open(pops, "wt") as f:
        for i, nm in enumerate(spec.sample_names()):
            f.write("%s pop%d\n" % (nm, i // c["spp"]))
    capsys.readouterr()
    sfs_cli.main(["-i", path, "--inputType", "genotypes", "--popsFile", pops, "--pipe", "-p", "pop0", "-p", "pop1", "-p", "pop2",
                  "-p", "pop3"] + CLI2[key + "_args"])
    assert capsys.readouterr().out == CLI2[key]


def test_gbin_cache_and_timing_file(inputs, tmp_path):
    """--cache writes <geno>.gbin after the first ingest and loads it on the next run (same rows, no tokenisation); a stale or
    mismatching cache is ignored; --timing writes the phase / kernel times as JSON"""
    import shutil
    from genomics_general_b200.cli import popgenWindows
    i = inputs["two_pops"] if "two_pops" in inputs else inputs[list(inputs)[0]]
    c = i["cfg"]
    g = str(tmp_path / "c.geno")
    shutil.copy(i["geno"], g)
    o2, tj = str(tmp_path / "o2.csv"), str(tmp_path / "t.json")
    from oracle_engine import OracleEngine
    # direct test of the cache functions (engine-independent)
    from genomics_general_b200 import geno_io
    gd = geno_io.parse_geno(g, geno_format="phased")
    cache = g + ".gbin"
    geno_io.save_gbin(cache, gd, g, "phased")
    eng = OracleEngine()
    back = geno_io.load_gbin(cache, eng, g, "phased")
    assert back is not None and back.geno is None
    assert np.array_equal(eng.g, gd.geno) and np.array_equal(back.pos, gd.pos) and np.array_equal(back.scaf_ids, gd.scaf_ids)
    assert back.scaf_names == gd.scaf_names and back.names == gd.names
    assert geno_io.load_gbin(cache, eng, g, "phased", samples=gd.names[:3]) is None          # other sample selection
    os.utime(g, ns=(1, 1))
    assert geno_io.load_gbin(cache, eng, g, "phased") is None                                # the source changed
    popgenWindows.main(["-w", str(c["w"]), "-m", str(c["m"]), "-g", i["geno"], "-o", o2, "-f", "phased", "-T", "1", "--popsFile",
                        i["pops"], "--timing", tj] + i["popargs"])
    t = json.load(open(tj))
    assert set(t["phases_s"]) == {"ingest", "windows", "statistics", "rows"} and t["windows"] > 0 and t["sites"] == c["S"]
