#!/usr/bin/env python
"""TEST INFRASTRUCTURE (build container only: needs /root/reference) — differential fuzzing of the command lines' HOST logic:
random small inputs and flag combinations, the UNMODIFIED reference script as a subprocess against this package's command line
on the oracle-backed engine (tests/oracle_engine.py).  What it exercises is everything around the kernels — windows of every
type, gates, prefixes, rounding, number formatting, the sfs interval / subsample plumbing — on cases nobody wrote by hand.

    python oracle/fuzz_vs_reference.py popgen 0 24      # popgenWindows: window types x analyses x flags (rows to 1e-6)
    python oracle/fuzz_vs_reference.py abba 0 12        # ABBABABAwindows: population orders, minData, window types
    python oracle/fuzz_vs_reference.py sfs 0 25         # sfs.py: --regions / --subsample / --exclude (byte for byte)
    python oracle/fuzz_vs_reference.py freq_distmat 0 10   # freq.py modes (byte for byte) + distMat formats / cat / windows

Round 2: 24 + 12 + 37 + 20 cases, no mismatch (distMat windows use -m 1: the reference hangs on a failed window under numpy 2).  A reference worker that dies leaves its parent waiting: every run has a timeout."""
import contextlib
import io
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from genomics_general_b200 import synth  # noqa: E402
from genomics_general_b200.cli import ABBABABAwindows as ab, _common, distMat as dm, freq as fq, popgenWindows as pg, sfs as sfs_cli  # noqa: E402
from oracle_engine import OracleEngine  # noqa: E402

for _m in (ab, dm, fq, pg, sfs_cli):
    _m.Engine = OracleEngine
_real = _common.load_geno
_common.load_geno = lambda args, samples, pl, header=None, engine=None: _real(args, samples, pl, header, None)
REF = "/root/reference"
TMP = tempfile.mkdtemp(prefix="pg_fuzz_")


def fuzz_popgen(lo, hi):
    bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(1000 + seed)
        npops = int(rng.integers(2, 4)); spp = int(rng.integers(2, 5)); S = int(rng.integers(800, 2500))
        miss = float(rng.choice([0.0, 0.02, 0.15]))
        spec = synth.SynthSpec(npops, spp, seed=seed, miss=miss)
        g = synth.synth_genotypes(spec, 0, S)
        nsc = int(rng.integers(1, 4))
        sizes = np.diff(np.concatenate([[0], np.sort(rng.choice(np.arange(1, S), nsc - 1, replace=False)), [S]])) if nsc > 1 else np.array([S])
        scaf = []; pos = []
        for k, n in enumerate(sizes):
            scaf += ["sc%d" % (k + 1)] * int(n); pos.append(synth.synth_positions(int(n), seed=seed * 7 + k))
        pos = np.concatenate(pos)
        d = TMP
        gp = os.path.join(d, "f.geno"); synth.write_geno(gp, g, pos, scaf, spec.sample_names())
        pp = os.path.join(d, "f.pops")
        with open(pp, "wt") as f:
            for i, n in enumerate(spec.sample_names()): f.write("%s pop%d\n" % (n, i // spp))
        argv = ["-g", gp, "-f", "phased", "--popsFile", pp, "--roundTo", "8", "-T", "1"]
        for k in range(npops): argv += ["-p", "pop%d" % k]
        wt = rng.choice(["coordinate", "sites", "sites_overlap", "coord_step"])
        if wt == "coordinate": argv += ["-w", str(int(rng.integers(2000, 9000))), "-m", str(int(rng.integers(1, 200)))]
        elif wt == "coord_step":
            w = int(rng.integers(3000, 9000)); argv += ["-w", str(w), "-s", str(int(w * rng.choice([0.25, 0.5, 1.5]))), "-m", str(int(rng.integers(1, 100)))]
        elif wt == "sites": argv += ["--windType", "sites", "-w", str(int(rng.integers(100, 500))), "-m", str(int(rng.integers(1, 120)))]
        else:
            w = int(rng.integers(150, 500)); argv += ["--windType", "sites", "-w", str(w), "-O", str(int(w * rng.choice([0.2, 0.5, 0.8]))), "-m", str(int(rng.integers(1, 100)))]
            if rng.random() < 0.5: argv += ["-D", str(int(rng.integers(1000, 6000)))]
        if rng.random() < 0.4: argv += ["--writeFailedWindows"]
        if rng.random() < 0.4: argv += ["--addWindowID"]
        if rng.random() < 0.3: argv += ["--minData", str(float(rng.choice([0.3, 0.8, 0.99])))]
        an = rng.choice(["default", "popFreq", "indHet", "ipd"])
        if an == "popFreq": argv += ["--analysis", "popFreq", "popDist", "popPairDist"]
        elif an == "indHet": argv += ["--analysis", "popDist", "indHet"]
        elif an == "ipd": argv += ["--analysis", "indPairDist", "popPairDist"]
        oref, oours = os.path.join(d, "ref.csv"), os.path.join(d, "ours.csv")
        try:
            r = subprocess.run([sys.executable, os.path.join(REF, "popgenWindows.py")] + argv + ["-o", oref], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            print(seed, "REF TIMEOUT (worker died)", argv[8:]); continue
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                pg.main(argv + ["-o", oours])
            err = None
        except BaseException as e:
            err = repr(e)
        if r.returncode != 0:
            print(seed, "REF FAILED", r.stderr[-200:].replace("\n", " | "), "ours:", err); continue
        if err: bad += 1; print(seed, "OURS FAILED", err, argv[8:]); continue
        A, B = open(oours).read().strip().split("\n"), open(oref).read().strip().split("\n")
        ha, hb = A[0].split(","), B[0].split(",")
        ok = sorted(ha) == sorted(hb) and len(A) == len(B)
        if ok:
            for x, y in zip(A[1:], B[1:]):
                dx, dy = dict(zip(ha, x.split(","))), dict(zip(hb, y.split(",")))
                for k in hb:
                    if dx[k] == dy[k]: continue
                    try:
                        fx, fy = float(dx[k]), float(dy[k])
                        if (np.isnan(fx) and np.isnan(fy)) or abs(fx - fy) <= 2e-7 + 1e-6 * abs(fy): continue
                    except ValueError: pass
                    ok = False; print("   diff", k, dx[k], dy[k], "row", x[:40]); break
                if not ok: break
        if not ok:
            bad += 1; print(seed, "MISMATCH", len(A), len(B), argv[8:])
        else:
            print(seed, "ok", len(A), wt, an)
    print("bad", bad)

    return bad


def fuzz_abba(lo, hi):
    bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(5000 + seed)
        spp = int(rng.integers(2, 5)); S = int(rng.integers(800, 2500)); miss = float(rng.choice([0.0, 0.03, 0.2]))
        spec = synth.SynthSpec(4, spp, seed=seed, miss=miss)
        g = synth.synth_genotypes(spec, 0, S); pos = synth.synth_positions(S, seed=seed)
        d = TMP
        gp = os.path.join(d, "a.geno"); synth.write_geno(gp, g, pos, ["chr1"] * S, spec.sample_names())
        pp = os.path.join(d, "a.pops")
        with open(pp, "wt") as f:
            for i, n in enumerate(spec.sample_names()): f.write("%s pop%d\n" % (n, i // spp))
        order = list(rng.permutation(4))
        argv = ["-g", gp, "-f", "phased", "--popsFile", pp, "-T", "1", "-P1", "pop%d" % order[0], "-P2", "pop%d" % order[1], "-P3", "pop%d" % order[2],
                "-O", "pop%d" % order[3], "--minData", str(float(rng.choice([0.01, 0.5, 1.0])))]
        if rng.random() < 0.5: argv += ["-w", str(int(rng.integers(3000, 9000))), "-m", str(int(rng.integers(1, 60)))]
        else: argv += ["--windType", "sites", "-w", str(int(rng.integers(150, 600))), "-m", str(int(rng.integers(1, 60)))]
        if rng.random() < 0.5: argv += ["--writeFailedWindows"]
        oref, oours = os.path.join(d, "aref.csv"), os.path.join(d, "aours.csv")
        try:
            r = subprocess.run([sys.executable, os.path.join(REF, "ABBABABAwindows.py")] + argv + ["-o", oref], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            print(seed, "REF TIMEOUT"); continue
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            ab.main(argv + ["-o", oours])
        A, B = open(oours).read().strip().split("\n"), open(oref).read().strip().split("\n")
        ok = A[0] == B[0] and len(A) == len(B)
        if ok:
            for x, y in zip(A[1:], B[1:]):
                x, y = x.split(","), y.split(",")
                if x[:6] != y[:6]: ok = False; print("  prefix", x[:6], y[:6]); break
                fx, fy = np.array([float(v) for v in x[6:]]), np.array([float(v) for v in y[6:]])
                if not np.allclose(fx, fy, rtol=0, atol=1.01e-4, equal_nan=True): ok = False; print("  vals", x, y); break
        bad += not ok
        print(seed, "ok" if ok else "MISMATCH", len(A), argv[8:])
    print("bad", bad)

    return bad


def fuzz_sfs(lo, hi):
    bad = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(seed)
        npops = int(rng.integers(2, 5)); spp = int(rng.integers(2, 5)); S = int(rng.integers(300, 900))
        miss = float(rng.choice([0.0, 0.03, 0.1]))
        spec = synth.SynthSpec(npops, spp, seed=seed, miss=miss)
        g = synth.synth_genotypes(spec, 0, S)
        pos = synth.synth_positions(S, seed=seed)
        nsc = int(rng.integers(1, 4))
        cuts = np.sort(rng.choice(np.arange(1, S), nsc - 1, replace=False)) if nsc > 1 else np.array([], dtype=int)
        scaf_id = np.searchsorted(cuts, np.arange(S), side="right")
        scaf = ["chr%d" % (k + 1) for k in scaf_id]
        d = TMP
        gp = os.path.join(d, "f.geno"); synth.write_geno(gp, g, pos, scaf, spec.sample_names())
        pp = os.path.join(d, "f.pops")
        with open(pp, "wt") as f:
            for i, n in enumerate(spec.sample_names()): f.write("%s pop%d\n" % (n, i // spp))
        argv = ["-i", gp, "--inputType", "genotypes", "--popsFile", pp, "--pipe"]
        for k in range(npops): argv += ["-p", "pop%d" % k]
        extra = ["--polarized"]
        if rng.random() < 0.5: extra += ["--doPairs"]
        if rng.random() < 0.6:
            regs = []
            for _ in range(int(rng.integers(1, 5))):
                c = int(rng.integers(1, nsc + 1)); a, b = sorted(int(x) for x in rng.choice(pos, 2))
                kind = rng.integers(0, 4)
                regs.append("chr%d:%d-%d" % (c, a, b) if kind < 2 else ("chr%d:%d-%d" % (c, b, a) if kind == 2 else "chr%d:%d" % (c, a)))
            extra += ["--regions"] + regs
        if rng.random() < 0.6:
            n_in = npops - 1
            sub = [int(rng.integers(1, 2 * spp + 1))] if rng.random() < 0.5 else [int(rng.integers(1, 2 * spp + 1)) for _ in range(n_in)]
            extra += ["--subsample"] + [str(x) for x in sub] + ["--seed", str(int(rng.integers(0, 100)))]
        if nsc > 1 and rng.random() < 0.3: extra += ["--exclude", "chr1"]
        r = subprocess.run([sys.executable, os.path.join(REF, "sfs.py")] + argv + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                sfs_cli.main(argv + extra)
            ours = buf.getvalue(); err = None
        except BaseException as e:
            ours = None; err = repr(e)
        if r.returncode != 0:
            print(seed, "REF FAILED", r.stderr[-300:].replace("\n", " | "), "ours:", err, extra); continue
        if ours != r.stdout:
            bad += 1
            print(seed, "MISMATCH", extra, "err:", err)
            if ours is not None:
                a, b = ours.split("\n"), r.stdout.split("\n")
                print("  ours %d lines, ref %d lines" % (len(a), len(b)))
                for x, y in zip(a, b):
                    if x != y: print("   ", repr(x), "!=", repr(y)); break
        else:
            print(seed, "ok", len(ours), extra)
    print("bad", bad)

    return bad


def fuzz_freq_distmat(lo, hi):
    bad = 0
    d = TMP
    for seed in range(lo, hi):
        rng = np.random.default_rng(9000 + seed)
        npops = int(rng.integers(2, 5)); spp = int(rng.integers(2, 4)); S = int(rng.integers(400, 1500)); miss = float(rng.choice([0.0, 0.05, 0.3]))
        spec = synth.SynthSpec(npops, spp, seed=seed, miss=miss)
        g = synth.synth_genotypes(spec, 0, S); pos = synth.synth_positions(S, seed=seed)
        nsc = int(rng.integers(1, 3)); cut = int(rng.integers(1, S)) if nsc == 2 else S
        scaf = ["c1"] * cut + ["c2"] * (S - cut)
        gp = os.path.join(d, "fd.geno"); synth.write_geno(gp, g, pos, scaf, spec.sample_names())
        pp = os.path.join(d, "fd.pops")
        with open(pp, "wt") as f:
            for i, n in enumerate(spec.sample_names()): f.write("%s pop%d\n" % (n, i // spp))
        # ---- freq ----
        argv = ["-g", gp, "-f", "phased", "--popsFile", pp, "-t", "1"]
        for k in rng.permutation(npops): argv += ["-p", "pop%d" % k]
        mode = rng.choice(["counts", "derived", "derived_counts", "derived_thr", "keepnan"])
        if mode == "derived": argv += ["--target", "derived"]
        elif mode == "derived_counts": argv += ["--target", "derived", "--asCounts"]
        elif mode == "derived_thr": argv += ["--target", "derived", "--threshold", "0.5", "--minData", "0.5"]
        elif mode == "keepnan": argv += ["--target", "derived", "--keepNanLines", "--minData", "0.8"]
        oref, oours = os.path.join(d, "fref.tsv"), os.path.join(d, "fours.tsv")
        r = subprocess.run([sys.executable, os.path.join(REF, "freq.py")] + argv + ["-o", oref], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            fq.main(argv + ["-o", oours])
        same = r.returncode == 0 and open(oref).read() == open(oours).read()
        bad += not same
        print(seed, "freq", mode, "ok" if same else "MISMATCH rc=%d" % r.returncode, os.path.getsize(oours))
        # ---- distMat (windows with -m 1 so that no window fails: the reference hangs on failed windows under numpy 2) ----
        fmt = str(rng.choice(["raw", "phylip", "nexus"]))
        argv = ["-g", gp, "-f", "phased", "-T", "1", "--outFormat", fmt, "--roundTo", str(int(rng.choice([4, 8])))]
        if rng.random() < 0.5: argv += ["-w", str(int(rng.integers(3000, 8000))), "-m", "1"]
        else: argv += ["--windType", "cat"]
        if rng.random() < 0.5: argv += ["--includeSameWithSame"]
        oref, oours = os.path.join(d, "dref.txt"), os.path.join(d, "dours.txt")
        try:
            r = subprocess.run([sys.executable, os.path.join(REF, "distMat.py")] + argv + ["-o", oref], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=90)
        except subprocess.TimeoutExpired:
            print(seed, "distMat REF TIMEOUT", argv[5:]); continue
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            dm.main(argv + ["-o", oours])
        A, B = open(oours).read(), open(oref).read()
        same = A == B
        if not same:      # allow last-digit differences of the rounding
            ta, tb = A.split(), B.split()
            same = len(ta) == len(tb)
            for x, y in zip(ta, tb):
                if x == y: continue
                try:
                    if abs(float(x) - float(y)) > 2e-4: same = False; break
                except ValueError:
                    same = False; break
        bad += not same
        print(seed, "distMat", fmt, "ok" if same else "MISMATCH", len(A), argv[5:])
    print("bad", bad)

    return bad


if __name__ == "__main__":
    which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    sys.exit(1 if {"popgen": fuzz_popgen, "abba": fuzz_abba, "sfs": fuzz_sfs, "freq_distmat": fuzz_freq_distmat}[which](lo, hi) else 0)
