"""Host logic of the multi-GPU command lines (genomics_general_b200/mgpu.py) on the CPU: byte ranges cut at line starts,
window ownership / halo, the order of the gathered table, and the file rendezvous between rank processes."""
import os
import subprocess
import sys

import numpy as np

from genomics_general_b200 import mgpu


def test_byte_ranges_cut_at_line_starts(tmp_path):
    rng = np.random.default_rng(1)
    lines = [b"#CHROM\tPOS\ta\tb\n"] + [b"chr1\t%d\t%s\n" % (i, b"A|C\t" * int(rng.integers(1, 9))) for i in range(1000)]
    data = b"".join(lines)
    p = str(tmp_path / "x.geno")
    open(p, "wb").write(data)
    body = len(lines[0])
    for world in (1, 2, 3, 8, 64):
        rs = mgpu.byte_ranges(p, body, world)
        assert rs[0][0] == body and rs[-1][1] == len(data)
        for (a, b), (c, d) in zip(rs[:-1], rs[1:]):
            assert b == c
        for a, b in rs:
            assert a == len(data) or a == body or data[a - 1:a] == b"\n"
        assert b"".join(data[a:b] for a, b in rs) == data[body:]


def test_window_ownership_halo_and_gather_order():
    starts = np.array([0, 100, 100, 250, 400])            # rank 1 has no sites
    lo = np.array([0, 50, 90, 100, 240, 250, 399, 400])
    hi = np.array([50, 90, 130, 240, 260, 399, 400, 400])  # window 2 reaches 30 sites into the next share; the last is empty
    owned = [mgpu.assign_windows(lo, hi, starts, r) for r in range(4)]
    assert [list(o[0]) for o in owned] == [[0, 1, 2], [], [3, 4], [5, 6, 7]]
    assert [o[3] for o in owned] == [30, 0, 10, 0]
    assert list(owned[2][1]) == [0, 140] and list(owned[2][2]) == [140, 160]
    w_max, row_of = mgpu.gathered_order([o[0] for o in owned])
    assert w_max == 3 and [row_of[w] for w in range(8)] == [0, 1, 2, 6, 7, 9, 10, 11]


def test_rendezvous_between_processes(tmp_path):
    d = str(tmp_path / "rdv")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from genomics_general_b200 import mgpu; "
            "r = mgpu.Rendezvous(int(sys.argv[1]), 3, %r, timeout=60); "
            "got = r.allgather('x', np.arange(4) * (r.rank + 1)); "
            "assert [int(g[3]) for g in got] == [3, 6, 9]; r.barrier('b'); print('ok')"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, text=True) for r in range(3)]
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out


# ------------------------------------------------------------------------------------------------------------------
# The complete `--devices N` command lines on the CPU: N rank processes (tests/_mgpu_cpu_worker.py) whose engine is the
# oracle-backed stand-in (tests/oracle_engine_mg.py: host tokenizer for the rank's byte range, the all-gather through
# files), everything else — byte ranges, the global picture, window ownership, halo sites, the gathered table's order,
# rank 0 writing — the product's own multi-GPU path.  The output must equal the single-device command line's.
# ------------------------------------------------------------------------------------------------------------------
import pytest  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mg_input(tmp_path_factory):
    from genomics_general_b200 import synth
    d = tmp_path_factory.mktemp("mg_cpu")
    spec = synth.SynthSpec(4, 3, seed=77, miss=0.03)
    S = 3600
    g = synth.synth_genotypes(spec, 0, S)
    scafs = ["chr1"] * 1500 + ["chr2"] * 900 + ["chr3"] * 1200
    pos = np.concatenate([synth.synth_positions(n, seed=77 + k) for k, n in enumerate((1500, 900, 1200))])
    geno = str(d / "mg.geno")
    synth.write_geno(geno, g, pos, scafs, spec.sample_names())
    pops = str(d / "mg.pops")
    with open(pops, "wt") as f:
        for i, n in enumerate(spec.sample_names()):
            f.write("%s pop%d\n" % (n, i // 3))
    coords = str(d / "mg.windows")
    with open(coords, "wt") as f:            # predefined windows: file order, one empty, one spanning most of a scaffold
        f.write("chr1 1 4000 a\nchr1 3000 9000 b\nchr1 9001 9002 c\nchr2 100 8000 d\nchr3 1 3000 e\nchr3 2500 12000 f\n")
    excl = str(d / "mg.exclude")
    with open(excl, "wt") as f:
        f.write("chr2\n")
    nohdr = str(d / "mg_nohdr.geno")                  # the same genotypes without the header line (--header supplies it)
    with open(geno) as f, open(nohdr, "wt") as o:
        header = f.readline().rstrip("\n")
        o.write(f.read())
    return dict(geno=geno, pops=pops, dir=str(d), coords=coords, exclude=excl, nohdr=nohdr, header=header)


def _single_device(module, argv, monkeypatch):
    import importlib
    from oracle_engine import OracleEngine
    from genomics_general_b200.cli import _common
    mod = importlib.import_module("genomics_general_b200.cli." + module)
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    real = _common.load_geno
    monkeypatch.setattr(_common, "load_geno", lambda args, samples, pl, header=None, engine=None: real(args, samples, pl, header, None))
    mod.main(argv)


def _ranks(module, argv, world, rdv_dir):
    os.makedirs(rdv_dir, exist_ok=True)
    procs = []
    for r in range(world):
        env = dict(os.environ, PG_MG_RANK=str(r), PG_MG_WORLD=str(world), PG_MG_DIR=rdv_dir, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_mgpu_cpu_worker.py"), module] + argv, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out[-3000:]


POPS4 = ["-p", "pop0", "-p", "pop1", "-p", "pop2", "-p", "pop3"]
P4 = ["-P1", "pop0", "-P2", "pop1", "-P3", "pop2", "-O", "pop3"]
MG_CASES = {
    "popgen_coordinate": ("popgenWindows", ["-w", "9000", "-m", "50", "-f", "phased", "--roundTo", "9"] + POPS4),
    "popgen_sites_overlap": ("popgenWindows", ["--windType", "sites", "-w", "400", "-O", "150", "-m", "100", "-f", "phased",
                                               "--writeFailedWindows", "--addWindowID"] + POPS4),
    "popgen_predefined": ("popgenWindows", ["--windType", "predefined", "--windCoords", "@COORDS@", "-m", "20", "-f", "phased",
                                            "--writeFailedWindows", "--addWindowID"] + POPS4),
    "popgen_exclude_file": ("popgenWindows", ["-w", "5000", "-m", "50", "-f", "phased", "--exclude", "@EXCLUDE@", "--addWindowID"]
                            + POPS4),
    "abba_headerless_file": ("ABBABABAwindows", ["-w", "8000", "-m", "30", "-f", "phased", "--minData", "0.5", "--header", "@HEADER@",
                                                 "@NOHDR@"] + P4),
    "popgen_popfreq": ("popgenWindows", ["-w", "9000", "-m", "50", "-f", "phased", "--analysis", "popFreq", "popDist", "popPairDist",
                                         "--writeFailedWindows"] + POPS4),
    "popgen_all_analyses": ("popgenWindows", ["-w", "9000", "-m", "50", "-f", "phased", "--analysis", "popDist", "popPairDist",
                                              "indPairDist", "indHet", "hapStats", "--hapDist", "0.02"] + POPS4),
    "popgen_pairwise_only": ("popgenWindows", ["--windType", "sites", "-w", "700", "-m", "100", "-f", "phased", "--analysis",
                                               "indPairDist", "hapStats", "--writeFailedWindows"] + POPS4),
    "abba_coordinate": ("ABBABABAwindows", ["-w", "12000", "-m", "30", "-f", "phased", "--minData", "0.5"] + P4),
    "fourpop_sites": ("fourPopWindows", ["--windType", "sites", "-w", "500", "--overlap", "100", "-m", "30", "-f", "phased",
                                         "--minData", "0.5", "--polarize"] + P4),
    "distmat_raw_windowdata": ("distMat", ["-w", "9000", "-m", "50", "-f", "phased", "--outFormat", "raw", "--roundTo", "8",
                                           "--windowDataOutFile", "@WDATA@", "--addWindowID"]),
    "distmat_phylip_sites": ("distMat", ["--windType", "sites", "-w", "600", "-O", "200", "-m", "100", "-f", "phased", "--outFormat",
                                         "phylip", "--includeSameWithSame", "--minPerInd", "560", "--writeFailedWindows"]),
    "distmat_cat_nexus": ("distMat", ["--windType", "cat", "-f", "phased", "--outFormat", "nexus", "--roundTo", "9", "--minPerInd",
                                      "3000"]),
    "sfs_polarized_pairs_regions": ("sfs", ["--inputType", "genotypes", "--polarized", "--doPairs", "--exclude", "chr2", "--regions",
                                            "chr1:1-9000", "chr3:2000-11000", "chr1:5000-15000"] + POPS4),
    "freq_counts": ("freq", ["-f", "phased"] + POPS4),
    "freq_target": ("freq", ["-f", "phased", "--target", "derived", "--minData", "0.5"] + POPS4),
}


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("case", list(MG_CASES))
def test_command_lines_on_n_ranks_equal_one_device_cpu(mg_input, case, world, monkeypatch, tmp_path):
    module, argv = MG_CASES[case]
    sub = {"@COORDS@": mg_input["coords"], "@EXCLUDE@": mg_input["exclude"], "@HEADER@": mg_input["header"]}
    base = [sub.get(x, x) for x in argv if x != "@NOHDR@"]
    base += ["-g", mg_input["nohdr"] if "@NOHDR@" in argv else mg_input["geno"], "--popsFile", mg_input["pops"]]
    if module == "distMat":
        base = [x for x in base if x != "--popsFile" and x != mg_input["pops"]]
    if module == "sfs":                            # -i / -p / --pref instead of -g / -p / -o
        base = ["-i" if x == "-g" else x for x in base]
    one = str(tmp_path / "one.txt")
    many = str(tmp_path / "many.txt")
    if module == "sfs":                            # one file per spectrum: concatenate them for the comparison
        import glob
        for pref, run in ((one, lambda a: _single_device(module, a, monkeypatch)),
                          (many, lambda a: _ranks(module, a, world, str(tmp_path / "rdv")))):
            run(base + ["--pref", pref + ".", "--suff", ".sfs"])
            files = sorted(glob.glob(pref + ".*.sfs"))
            assert len(files) == 3 + 3          # polarized: three in-group populations, their three pairs
            with open(pref, "wt") as f:
                for x in files:
                    f.write(os.path.basename(x)[len(os.path.basename(pref)):] + "\n" + open(x).read())
        base = None
    else:
        _single_device(module, [x if x != "@WDATA@" else one + ".w" for x in base] + ["-o", one], monkeypatch)
        _ranks(module, [x if x != "@WDATA@" else many + ".w" for x in base] + ["-o", many], world, str(tmp_path / "rdv"))
    a, b = open(one).read(), open(many).read()
    assert len(a.splitlines()) > 3
    assert a == b
    if base and "@WDATA@" in base:
        assert open(one + ".w").read() == open(many + ".w").read() and open(one + ".w").read().count("\n") > 2


def test_devices_launch_path_relaunches_the_command_line(tmp_path):
    """mgpu.init(module, argv, N): rank 0 starts N-1 copies of the command line with PG_MG_RANK / PG_MG_WORLD / PG_MG_DIR,
    they meet in the exchange directory, finish() waits for them and removes it"""
    import json
    out = str(tmp_path / "seen.json")
    env = dict(os.environ, PYTHONPATH=HERE + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for k in ("PG_MG_RANK", "PG_MG_WORLD", "PG_MG_DIR", "RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "_mgpu_dummy_cli", "3", out, "--flag", "x y"], env=env, cwd=HERE,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    seen = json.load(open(out))
    assert seen["ranks"] == [0, 1, 2]
    assert seen["argv"] == [["3", out, "--flag", "x y"]] * 3
    assert not os.path.exists(seen["dir"])
