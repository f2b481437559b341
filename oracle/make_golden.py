#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports /root/reference/genomics.py in-process (un-rounded doubles from
groupDistStats / ABBABABA / siteFreqs / indPairDists and the window generators) and
shells out to the reference CLIs (popgenWindows.py, ABBABABAwindows.py, freq.py,
distMat.py) on small deterministic inputs.  Inputs and outputs are committed as
fixtures; nothing under tests/ reads /root/reference at run time.
"""
from __future__ import annotations

import io
import json
import os
import subprocess
import sys
import tempfile
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

import genomics as ref  # noqa: E402  (the reference)
from genomics_general_b200 import synth  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
BASES = np.array(list("ACGTN"))


def geno_tokens(g, ploidies, sep="/"):
    """int8 [L,H] -> list (per site) of list (per sample) of phased tokens."""
    L, H = g.shape
    ch = BASES[np.where(g < 0, 4, g)]
    toks = []
    for s in range(L):
        row, h = [], 0
        for p in ploidies:
            row.append(sep.join(ch[s, h:h + p]))
            h += p
        toks.append(row)
    return toks


def ref_alignment(g, sample_names, ploidies, pop_names, pop_inds, geno_format="phased"):
    """Build the reference Alignment exactly the way popgenWindows' worker does
    (GenoWindow.seqDict -> genoToAlignment), popgenWindows.py:44."""
    toks = geno_tokens(g, ploidies)
    ploidy_dict = dict(zip(sample_names, ploidies))
    sd = ref.SampleData(indNames=list(sample_names), popNames=list(pop_names),
                        popInds=[list(p) for p in pop_inds], ploidyDict=ploidy_dict)
    win = ref.GenoWindow(scaffold="chr1", limits=[1, 10 ** 9], sites=toks, names=list(sample_names),
                         positions=list(range(1, g.shape[0] + 1)))
    if geno_format == "haplo":
        aln = ref.genoToAlignment(win.seqDict(), sd, genoFormat="haplo")
    else:
        aln = ref.genoToAlignment(win.seqDict(), sd, genoFormat="phased")
    return aln, sd


def random_case(rng, L, pop_sizes, ploidies_per_pop, miss_geno=0.0, miss_allele=0.0, p_var=0.5,
                p_third=0.05, allmiss_sites=0.0, out_fixed=0.0):
    """Random window: returns g [L,H] int8 in FILE order, sample names, ploidies, pops."""
    names, ploidies, pop_inds = [], [], []
    for x, (n, pl) in enumerate(zip(pop_sizes, ploidies_per_pop)):
        inds = ["s%d_%02d" % (x, i) for i in range(n)]
        names += inds
        ploidies += [pl] * n
        pop_inds.append(inds)
    # shuffle file order so pops are interleaved
    order = rng.permutation(len(names))
    names = [names[i] for i in order]
    ploidies = [ploidies[i] for i in order]
    H = sum(ploidies)
    samp_pop = {n: x for x, inds in enumerate(pop_inds) for n in inds}
    hap_pop_file = np.concatenate([[samp_pop[n]] * p for n, p in zip(names, ploidies)])
    P = len(pop_sizes)
    ref_a = rng.integers(0, 4, L)
    alt_a = (ref_a + rng.integers(1, 4, L)) % 4
    var = rng.random(L) < p_var
    freq = rng.random((L, P)) * var[:, None]
    fix = rng.random(L) < out_fixed
    freq[fix, P - 1] = 0
    g = np.where(rng.random((L, H)) < freq[:, hap_pop_file], alt_a[:, None], ref_a[:, None])
    third = (rng.random((L, H)) < 0.3) & (rng.random(L) < p_third)[:, None] & var[:, None]
    g = np.where(third, (alt_a[:, None] + 1 + (ref_a[:, None] == (alt_a[:, None] + 1) % 4)) % 4, g)
    g = g.astype(np.int8)
    if miss_geno > 0:
        h = 0
        for p in ploidies:
            m = rng.random(L) < miss_geno
            g[m, h:h + p] = -1
            h += p
    if miss_allele > 0:
        g[rng.random((L, H)) < miss_allele] = -1
    if allmiss_sites > 0:
        g[rng.random(L) < allmiss_sites, :] = -1
    return g, names, ploidies, ["pop%d" % x for x in range(P)], pop_inds


def aln_to_arrays(aln, pop_names):
    """Reference Alignment -> (int8 [L,H] in ALIGNMENT order, hap_pop, names, sampleNames)."""
    g = aln.numArray.T.copy()
    g[g < 0] = -1
    hap_pop = np.array([pop_names.index(x) if x in pop_names else -1 for x in aln.groups], dtype=np.int32)
    return g.astype(np.int8), hap_pop, [str(n) for n in aln.names], [str(n) for n in aln.sampleNames]


def make_window_cases():
    rng = np.random.default_rng(20260923)
    specs = [
        dict(name="two_pops_nomiss", L=300, pop_sizes=[5, 5], pl=[2, 2], minSites=50, minData=0.01),
        dict(name="two_pops_genomiss", L=300, pop_sizes=[5, 5], pl=[2, 2], miss_geno=0.05, minSites=50, minData=0.01),
        dict(name="three_pops_allelemiss", L=257, pop_sizes=[4, 3, 6], pl=[2, 2, 2], miss_allele=0.1, minSites=10,
             minData=0.01),
        dict(name="single_hap_pop", L=120, pop_sizes=[1, 4, 3], pl=[1, 2, 2], miss_geno=0.03, minSites=1, minData=0.01),
        dict(name="allmiss_sites_only", L=200, pop_sizes=[4, 4], pl=[2, 2], allmiss_sites=0.2, minSites=100,
             minData=0.01),
        dict(name="minsites_masks_pairs", L=64, pop_sizes=[3, 3], pl=[2, 2], miss_geno=0.3, minSites=30, minData=0.01),
        dict(name="high_mindata", L=150, pop_sizes=[10, 10], pl=[2, 2], miss_geno=0.02, minSites=10, minData=0.99),
        dict(name="mid_mindata", L=40, pop_sizes=[4, 4], pl=[2, 2], miss_geno=0.4, minSites=12, minData=0.8),
        dict(name="four_pops_abba", L=400, pop_sizes=[5, 5, 5, 3], pl=[2, 2, 2, 2], miss_geno=0.05, minSites=1,
             minData=0.01, out_fixed=0.8),
        dict(name="four_pops_abba_allelemiss", L=333, pop_sizes=[3, 4, 5, 2], pl=[2, 2, 2, 2], miss_allele=0.15,
             minSites=1, minData=0.01, out_fixed=0.7),
        dict(name="haploid_mixed", L=180, pop_sizes=[4, 4, 4, 2], pl=[1, 2, 1, 2], miss_geno=0.05, minSites=5,
             minData=0.01, out_fixed=0.8),
        dict(name="tiny_window", L=3, pop_sizes=[2, 2], pl=[2, 2], minSites=1, minData=0.01),
        dict(name="all_missing_window", L=20, pop_sizes=[2, 2], pl=[2, 2], allmiss_sites=1.0, minSites=1,
             minData=0.01),
    ]
    arrays, meta = {}, []
    for sp in specs:
        g_file, names, ploidies, pop_names, pop_inds = random_case(
            rng, sp["L"], sp["pop_sizes"], sp["pl"], miss_geno=sp.get("miss_geno", 0.0),
            miss_allele=sp.get("miss_allele", 0.0), allmiss_sites=sp.get("allmiss_sites", 0.0),
            out_fixed=sp.get("out_fixed", 0.0))
        aln, sd = ref_alignment(g_file, names, ploidies, pop_names, pop_inds)
        g_aln, hap_pop, hap_names, hap_samples = aln_to_arrays(aln, pop_names)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gds = aln.groupDistStats(doPairs=True, minSites=sp["minSites"], minData=sp["minData"])
        alnf, _ = ref_alignment(g_file, names, ploidies, pop_names, pop_inds)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                gfs = alnf.groupFreqStats()
            except ZeroDivisionError:       # the reference raises for a single-haplotype population (TajimaD, 621-626)
                gfs = None
        entry = dict(name=sp["name"], sample_names=names, ploidies=ploidies, pop_names=pop_names,
                     pop_inds=pop_inds, minSites=sp["minSites"], minData=sp["minData"],
                     hap_names=hap_names, hap_samples=hap_samples,
                     groupDistStats={k: float(v) for k, v in gds.items()},
                     groupFreqStats=None if gfs is None else {k: float(v) for k, v in gfs.items()})
        # per-pop site counts (freq.py default path, freq.py:52-58)
        sc = np.stack([aln.subset(groups=[p]).siteFreqs(asCounts=True) for p in pop_names], axis=1)
        arrays[sp["name"] + "__site_counts"] = sc.astype(np.int32)
        # pair matrices
        aln2, _ = ref_alignment(g_file, names, ploidies, pop_names, pop_inds)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dm = aln2.distMatrix()
            pn = aln2.pairNonNan()
        arrays[sp["name"] + "__distMatrix"] = dm
        arrays[sp["name"] + "__pairNonNan"] = pn.astype(np.int32)
        # indPairDists (distMat.py:42-45) in file sample order, both diagonal modes
        for inc in (False, True):
            aln3, _ = ref_alignment(g_file, names, ploidies, pop_names, pop_inds)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pdd = aln3.indPairDists(includeSameWithSame=inc)
            m = np.array([[pdd[a][b] for b in names] for a in names], dtype=np.float64)
            arrays[sp["name"] + "__indPairDists_%d" % int(inc)] = m
        # ABBA-BABA where 4 pops exist
        if len(pop_names) >= 4:
            ab = {}
            for md in (0.0, 0.01, 0.5, 1.0):
                aln4, _ = ref_alignment(g_file, names, ploidies, pop_names, pop_inds)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    r = ref.ABBABABA(aln4, "pop0", "pop1", "pop2", "pop3", md)
                ab[str(md)] = {k: float(v) for k, v in r.items()}
            entry["ABBABABA"] = ab
        arrays[sp["name"] + "__g_file"] = g_file
        arrays[sp["name"] + "__g_aln"] = g_aln
        arrays[sp["name"] + "__hap_pop"] = hap_pop
        meta.append(entry)
    np.savez_compressed(os.path.join(GOLD, "window_cases.npz"), **arrays)
    with open(os.path.join(GOLD, "window_cases.json"), "wt") as f:
        json.dump(meta, f, indent=1)
    print("window cases:", len(meta))


def geno_text(scafs, positions, names, toks):
    out = io.StringIO()
    out.write("#CHROM\tPOS\t" + "\t".join(names) + "\n")
    for sc, p, t in zip(scafs, positions, toks):
        out.write("%s\t%d\t%s\n" % (sc, p, "\t".join(t)))
    return out.getvalue()


def make_generator_cases():
    """Window generators over a 4-scaffold layout with gaps, dense runs and a 1-site scaffold."""
    rng = np.random.default_rng(7)
    scafs, positions = [], []
    for sc, n, span in (("chrA", 180, 2000), ("chrB", 1, 50), ("chrC", 60, 5000), ("chrD", 97, 400)):
        pos = np.sort(rng.choice(np.arange(1, span + 1), size=n, replace=False))
        if sc == "chrC":
            pos = pos[(pos < 1200) | (pos > 3100)]          # a gap => empty coordinate windows
        scafs += [sc] * len(pos)
        positions += [int(p) for p in pos]
    names = ["a", "b", "c"]
    g = rng.integers(0, 4, (len(positions), 6)).astype(np.int8)
    toks = geno_tokens(g, [2, 2, 2])
    text = geno_text(scafs, positions, names, toks)
    cases = []

    def record(kind, params, gen):
        wins = []
        for w in gen:
            lim = [None if (isinstance(l, float) and np.isinf(l)) else int(l) for l in w.limits]
            wins.append(dict(scaffold=w.scaffold, limits=lim, positions=[int(p) for p in w.positions],
                             ID=w.ID if isinstance(w.ID, (str, type(None))) else int(w.ID)))
        cases.append(dict(kind=kind, params=params, windows=wins))

    for ws, st in ((500, None), (500, 250), (300, 700), (1000, 100), (10000, None), (37, 37)):
        record("coordinate", dict(windSize=ws, stepSize=st),
               ref.slidingCoordWindows(io.StringIO(text), ws, st if st else ws, names=names))
    record("coordinate", dict(windSize=500, stepSize=None, exclude=["chrC"]),
           ref.slidingCoordWindows(io.StringIO(text), 500, 500, names=names, exclude=["chrC"]))
    for ws, ov, md, ms in ((50, 0, None, None), (50, 10, None, None), (50, 0, None, 20), (50, 25, None, 10),
                           (20, 0, 150, 5), (20, 5, 100, 20), (7, 3, None, 1), (200, 0, None, 1)):
        record("sites", dict(windSites=ws, overlap=ov, maxDist=md, minSites=ms),
               ref.slidingSitesWindows(io.StringIO(text), ws, ov, md if md else np.inf, ms, names=names))
    for ex in (["chrB"], ["chrC"], ["chrB", "chrC"]):            # the duplicate window after a skipped scaffold keeps its ID
        record("sites", dict(windSites=50, overlap=0, maxDist=None, minSites=20, exclude=ex),
               ref.slidingSitesWindows(io.StringIO(text), 50, 0, np.inf, 20, names=names, exclude=ex))
    record("coordinate", dict(windSize=300, stepSize=None, exclude=["chrB"]),
           ref.slidingCoordWindows(io.StringIO(text), 300, 300, names=names, exclude=["chrB"]))
    coords = [("chrA", 1, 400, "w1"), ("chrA", 300, 900, "w2"), ("chrA", 1500, 1600, "w3"), ("chrC", 1, 5000, "w4"),
              ("chrD", 100, 200, "w5"), ("chrD", 150, 160, "w6"), ("chrD", 390, 400, "w7")]
    record("predefined", dict(windCoords=[list(c) for c in coords]),
           ref.predefinedCoordWindows(io.StringIO(text), coords, names=names))
    coords2 = [("chrC", 1000, 4000), ("chrA", 1, 100), ("chrD", 1, 50)]     # chrA after chrC: never reached
    record("predefined", dict(windCoords=[list(c) for c in coords2]),
           ref.predefinedCoordWindows(io.StringIO(text), coords2, names=names))
    with open(os.path.join(GOLD, "generator_cases.json"), "wt") as f:
        json.dump(dict(scaffolds=scafs, positions=positions, names=names, cases=cases), f)
    print("generator cases:", len(cases))


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
    if r.returncode != 0:
        raise RuntimeError("reference CLI failed: %s\n%s" % (" ".join(cmd), r.stderr[-2000:]))
    return r


def make_cli_cases():
    """Reference CLIs on synthetic .geno files regenerated at test time from (spec, seed)."""
    out = {}
    tmp = tempfile.mkdtemp(prefix="golden_")
    cfgs = [
        # BASELINE.json config 1 (C1): 2 pops x 10 diploid, 100k sites, -w 50000 -m 100
        dict(name="c1_miss0", n_pops=2, spp=10, S=100000, miss=0.0, seed=20260924, w=50000, m=100),
        dict(name="c1_miss2", n_pops=2, spp=10, S=100000, miss=0.02, seed=20260924, w=50000, m=100),
        # 4 pops for ABBA-BABA / freq / distMat, 3 scaffolds, smaller
        dict(name="four_pops", n_pops=4, spp=6, S=12000, miss=0.03, seed=11, w=20000, m=50, scaffolds=3),
    ]
    for c in cfgs:
        spec = synth.SynthSpec(c["n_pops"], c["spp"], seed=c["seed"], miss=c["miss"])
        g = synth.synth_genotypes(spec, 0, c["S"])
        nsc = c.get("scaffolds", 1)
        per = c["S"] // nsc
        scafs, pos = [], []
        for k in range(nsc):
            n = per if k < nsc - 1 else c["S"] - per * (nsc - 1)
            scafs += ["chr%d" % (k + 1)] * n
            pos.append(synth.synth_positions(n, seed=c["seed"] + k))
        pos = np.concatenate(pos)
        path = os.path.join(tmp, c["name"] + ".geno")
        synth.write_geno(path, g, pos, scafs, spec.sample_names())
        pops_file = os.path.join(tmp, c["name"] + ".pops")
        with open(pops_file, "wt") as f:
            for i, n in enumerate(spec.sample_names()):
                f.write("%s pop%d\n" % (n, i // c["spp"]))
        res = dict(cfg=c)
        popargs = []
        for p in spec.pop_names():
            popargs += ["-p", p]
        o = os.path.join(tmp, "o.csv")
        run([sys.executable, os.path.join(REF, "popgenWindows.py"), "-w", str(c["w"]), "-m", str(c["m"]), "-g", path,
             "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file, "--roundTo", "12"] + popargs)
        res["popgenWindows_roundTo12"] = open(o).read()
        run([sys.executable, os.path.join(REF, "popgenWindows.py"), "-w", str(c["w"]), "-m", str(c["m"]), "-g", path,
             "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file, "--writeFailedWindows"] + popargs)
        res["popgenWindows_default"] = open(o).read()
        if c["name"] == "four_pops":
            run([sys.executable, os.path.join(REF, "popgenWindows.py"), "--windType", "sites", "-w", "500", "-O", "100",
                 "-m", "200", "-g", path, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file,
                 "--roundTo", "10"] + popargs)
            res["popgenWindows_sites"] = open(o).read()
            run([sys.executable, os.path.join(REF, "popgenWindows.py"), "-w", str(c["w"]), "-m", str(c["m"]), "-g", path,
                 "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file, "--roundTo", "8",
                 "--analysis", "popFreq", "popDist", "popPairDist", "indPairDist"] + popargs)
            res["popgenWindows_popFreq_indPairDist"] = open(o).read()
            run([sys.executable, os.path.join(REF, "ABBABABAwindows.py"), "-w", str(c["w"]), "-m", str(c["m"]),
                 "-g", path, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file, "--minData", "0.5",
                 "-P1", "pop0", "-P2", "pop1", "-P3", "pop2", "-O", "pop3"])
            res["ABBABABAwindows"] = open(o).read()
            run([sys.executable, os.path.join(REF, "freq.py"), "-g", path, "-o", o, "-f", "phased", "-t", "1",
                 "--popsFile", pops_file] + popargs)
            txt = open(o).read().splitlines()
            res["freq_head"] = txt[:400]
            import hashlib
            res["freq_sha256"] = hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest()
            res["freq_nlines"] = len(txt)
            run([sys.executable, os.path.join(REF, "distMat.py"), "-w", str(c["w"]), "-m", str(c["m"]), "-g", path,
                 "-o", o, "-f", "phased", "-T", "1", "--outFormat", "raw", "--roundTo", "10"])
            res["distMat_raw"] = open(o).read()
            run([sys.executable, os.path.join(REF, "distMat.py"), "--windType", "cat", "-g", path,
                 "-o", o, "-f", "phased", "-T", "1", "--outFormat", "phylip", "--roundTo", "8"])
            res["distMat_cat_phylip"] = open(o).read()
            # ---- more flag coverage (same input) ----
            import gzip as _gz
            gzpath = path + ".gz"
            with open(path, "rb") as fi, _gz.open(gzpath, "wb") as fo:
                fo.write(fi.read())
            coords = os.path.join(tmp, "coords.txt")
            with open(coords, "wt") as f:
                f.write("chr1 1 15000 wA\nchr1 10000 30000 wB\nchr2 500 2500 wC\nchr3 1 100000 wD\n")
            run([sys.executable, os.path.join(REF, "popgenWindows.py"), "--windType", "predefined", "--windCoords", coords,
                 "-m", "10", "-g", gzpath, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file, "--roundTo", "9",
                 "--addWindowID", "--writeFailedWindows"] + popargs)
            res["popgenWindows_predefined_gz_id"] = open(o).read()
            res["coords_file"] = open(coords).read()
            # haploid samples: two samples of pop0 declared haploid need one-letter tokens -> use a diplo-coded file
            dpath = os.path.join(tmp, "four_pops_diplo.geno")
            synth.write_geno(dpath, g, pos, scafs, spec.sample_names(), fmt="diplo")
            run([sys.executable, os.path.join(REF, "popgenWindows.py"), "-w", str(c["w"]), "-s", "10000", "-m", str(c["m"]),
                 "-g", dpath, "-o", o, "-f", "diplo", "-T", "1", "--popsFile", pops_file, "--roundTo", "9"] + popargs)
            res["popgenWindows_diplo_step"] = open(o).read()
            run([sys.executable, os.path.join(REF, "ABBABABAwindows.py"), "--windType", "sites", "-w", "1000", "--overlap",
                 "250", "-m", "100", "-g", path, "-o", o, "-f", "phased", "-T", "1", "--popsFile", pops_file,
                 "--minData", "0.9", "-P1", "pop1", "-P2", "pop0", "-P3", "pop2", "-O", "pop3", "--writeFailedWindows",
                 "--addWindowID"])
            res["ABBABABAwindows_sites_overlap"] = open(o).read()
            sub = spec.sample_names()[1::3]
            wdo = os.path.join(tmp, "wd.txt")
            run([sys.executable, os.path.join(REF, "distMat.py"), "-w", str(c["w"]), "-m", str(c["m"]), "-g", path,
                 "-o", o, "-f", "phased", "-T", "1", "--outFormat", "nexus", "--roundTo", "7", "--includeSameWithSame",
                 "--windowDataOutFile", wdo, "--samples"] + sub)
            res["distMat_nexus_subset"] = open(o).read()
            res["distMat_windowData"] = open(wdo).read()
            res["distMat_subset_samples"] = sub
            run([sys.executable, os.path.join(REF, "freq.py"), "-g", path, "-o", o, "-f", "phased", "-t", "1", "--indFreqs"])
            txt = open(o).read().splitlines()
            res["freq_indFreqs_head"] = txt[:50]
            res["freq_indFreqs_sha256"] = hashlib.sha256(("\n".join(txt) + "\n").encode()).hexdigest()
        out[c["name"]] = res
        print("cli case", c["name"], "done")
    with open(os.path.join(GOLD, "cli_cases.json"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["windows", "generators", "cli"]
    if "windows" in which:
        make_window_cases()
    if "generators" in which:
        make_generator_cases()
    if "cli" in which:
        make_cli_cases()
