"""Parity corners of the next-rank rows against fixtures produced by the unmodified reference
(oracle/make_golden4.py): sfs.py --regions / --regionsFile, sfs.py --subsample (numpy's seeded global stream, drawn in
the reference's order), popgenWindows.py --inferPloidy on a file with haploid, diploid and triploid samples.

Every test runs twice: on the CPU with the oracle-backed engine (host logic of the command lines) and, marked gpu, on
the real engine through the C-ABI."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN, assert_close
from oracle_engine import OracleEngine

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import make_golden4 as mg4  # noqa: E402   (input generators only; the reference is not imported)

C4 = json.load(open(os.path.join(GOLDEN, "cases4.json")))


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def engine_kind(request, monkeypatch):
    from genomics_general_b200.cli import _common, popgenWindows, sfs
    if request.param == "oracle":
        for mod in (popgenWindows, sfs):
            monkeypatch.setattr(mod, "Engine", OracleEngine)
        real = _common.load_geno
        monkeypatch.setattr(_common, "load_geno", lambda args, samples, pl, header=None, engine=None: real(args, samples, pl, header, None))
    return request.param


@pytest.fixture(scope="module")
def inputs4(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("golden4"))
    assert mg4.SFS_CFG == C4["sfs_cfg"] and mg4.PLOIDY_CFG == C4["ploidy_cfg"]
    si = mg4.sfs_input(d)
    rfile = os.path.join(d, "regions.txt")
    with open(rfile, "wt") as f:
        f.write(C4["regions_file"])
    # freq.py's default rows, written from the oracle's counts (tests/test_oracle_golden2.py pins that format)
    from oracle import dense_oracle as do
    hp = np.repeat(np.arange(4, dtype=np.int32), 8)
    base = do.site_counts(si["g"], hp, 4)
    bc = os.path.join(d, "base.tsv")
    with open(bc, "wt") as f:
        f.write("scaffold\tposition\tpop0\tpop1\tpop2\tpop3\n")
        for s in range(len(base)):
            f.write("%s\t%d\t%s\n" % (si["scaf"][s], si["pos"][s], "\t".join(",".join(str(v) for v in base[s, x]) for x in range(4))))
    return dict(sfs=si, regions=rfile, base=bc, ploidy=mg4.ploidy_input(d))


SFS_KEYS = [k for k in C4 if k.startswith(("regions_", "subsample_")) and k + "_args" in C4]


@pytest.mark.parametrize("key", SFS_KEYS)
def test_sfs_regions_and_subsample(engine_kind, inputs4, key, capsys):
    """byte for byte the reference's --pipe output: one count column per interval, keys in order of first appearance in
    any interval; the down-sampled counts follow numpy's legacy stream (sfs.py:23-24)"""
    from genomics_general_b200.cli import sfs as sfs_cli
    extra = [inputs4["regions"] if x == "@REGIONS@" else x for x in C4[key + "_args"]]
    if C4[key + "_input"] == "geno":
        si = inputs4["sfs"]
        argv = ["-i", si["geno"], "--inputType", "genotypes", "--popsFile", si["pops"], "-p", "pop0", "-p", "pop1", "-p", "pop2",
                "-p", "pop3"]
    else:
        argv = ["-i", inputs4["base"], "--inputType", "baseCounts"]
    capsys.readouterr()
    sfs_cli.main(argv + ["--pipe"] + extra)
    assert capsys.readouterr().out == C4[key]


def test_sfs_region_parsing():
    """genomics.parseRegionText / Intervals (genomics.py:2323-2336, 2361-2367)"""
    import argparse
    from genomics_general_b200.cli import sfs as sfs_cli
    assert sfs_cli.parse_region_text("chr1:10-20") == ("chr1", 10, 20)
    assert sfs_cli.parse_region_text("chr1:20-10") == ("chr1", 10, 20)
    assert sfs_cli.parse_region_text("chr1:7") == ("chr1", 7, None)
    assert sfs_cli.parse_region_text("chr1") == ("chr1", None, None)
    a = argparse.Namespace(regions=["c:5-9", "c:7", "d"], regionsFile=None)
    iv = sfs_cli.read_intervals(a)
    assert iv == [("c", 5, 9), ("c", 7, 7), ("d", 0, None)]
    sc = np.array(["c", "c", "c", "d", "e"])
    m = sfs_cli.interval_masks(iv, sc, [4, 7, 9, 10 ** 9, 7], np.array([1, 1, 0, 1, 1], dtype=np.uint8))
    assert [list(x) for x in m] == [[0, 1, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 0, 1, 0]]


def test_sfs_subsample_individuals_is_refused():
    from genomics_general_b200.cli import sfs as sfs_cli
    import argparse
    with pytest.raises(NotImplementedError):
        sfs_cli.subsample_sizes(argparse.Namespace(subsampleIndividuals=True, subsample=[2]), ["a"])


def _table(text):
    lines = text.strip("\n").split("\n")
    hdr = lines[0].split(",")
    return hdr, [dict(zip(hdr, l.split(","))) for l in lines[1:]]


@pytest.mark.parametrize("key", [k for k in C4 if k.startswith("infer_ploidy_") and k + "_args" in C4])
def test_popgenWindows_infer_ploidy(engine_kind, inputs4, key, tmp_path):
    """--inferPloidy: the token widths decide (haploid, diploid and triploid samples in one phased file)"""
    from genomics_general_b200.cli import popgenWindows
    pi = inputs4["ploidy"]
    o = str(tmp_path / "o.csv")
    popgenWindows.main(["-o", o, "-T", "1", "--roundTo", "9", "-g", pi["geno"], "-f", "phased", "--popsFile", pi["pops"],
                        "-p", "pop0", "-p", "pop1", "--inferPloidy"] + C4[key + "_args"])
    h1, r1 = _table(open(o).read())
    h2, r2 = _table(C4[key])
    assert h1 == h2
    assert len(r1) == len(r2)
    for a, b in zip(r1, r2):
        for k in h2[:5]:
            assert a[k] == b[k], (k, a[k], b[k])
        assert_close([float(a[k]) for k in h2[5:]], [float(b[k]) for k in h2[5:]], "row " + a["start"], rtol=1e-6, atol=2e-9)


def test_infer_ploidy_from_token_widths(tmp_path):
    import argparse
    from genomics_general_b200.cli import _common as C
    p = str(tmp_path / "x.geno")
    with open(p, "wt") as f:
        f.write("#CHROM\tPOS\ta\tb\tc\n\n# comment\nchr1\t5\tA\tA/T\tA|C|G\n")
    a = argparse.Namespace(genoFile=p, header=None, genoFormat="phased", inferPloidy=True, ploidy=None, ploidyFile=None)
    assert C.ploidy_dict(a, ["a", "b", "c"], None) == dict(a=1, b=2, c=3)
    a.genoFormat = "diplo"
    assert C.ploidy_dict(a, ["a", "c"], None) == dict(a=2, c=2)
    a.genoFormat = "pairs"
    assert C.ploidy_dict(a, ["b"], None) == dict(b=3)
