"""The oracle's restatements of sampleHet / H12stats / fourPop / freq.py --target against fixtures produced by the
reference itself (oracle/make_golden2.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import dense_oracle as do

GOLD = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(GOLD, "window_cases2.json")))
ARR = np.load(os.path.join(GOLD, "window_cases2.npz"))
IDS = [m["name"] for m in META]


def hap_ind_of(m):
    """individual index (file sample order) of each alignment haplotype"""
    idx = {n: k for k, n in enumerate(m["sample_names"])}
    return np.array([idx[s] for s in m["hap_samples"]], dtype=np.int32)


@pytest.mark.parametrize("m", META, ids=IDS)
def test_sample_het(m):
    if not m["sampleHet"]:
        pytest.skip("reference raises for haploid samples")
    g = ARR[m["name"] + "__g_aln"]
    hi = hap_ind_of(m)
    n = len(m["sample_names"])
    for key, masked in (("alone", None), ("after_popDist", m["minSites"]), ("after_indPairDist", None)):
        got = do.sample_het(g, hi, n, masked_min_sites=masked)
        want = np.array([m["sampleHet"][key][s] for s in m["sample_names"]])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, equal_nan=True, err_msg=key)


@pytest.mark.parametrize("m", META, ids=IDS)
def test_h12(m):
    g = ARR[m["name"] + "__g_aln"]
    hp = ARR[m["name"] + "__hap_pop"]
    P = len(m["pop_names"])
    for key, want in m["H12stats"].items():
        state, md = key.rsplit("_", 1)
        masked = m["minSites"] if state == "after_popDist" else None
        diag = state != "alone"
        got = do.h12_stats(g, hp, P, max_dist=float(md), masked_min_sites=masked, diag_nan=diag)
        for x, pn in enumerate(m["pop_names"]):
            for k, stat in enumerate(("H1", "H12", "H2")):
                np.testing.assert_allclose(got[x, k], want[stat + "_" + pn], rtol=1e-12, atol=1e-15,
                                           err_msg="%s %s %s" % (key, stat, pn))


@pytest.mark.parametrize("m", [m for m in META if "fourPop" in m], ids=[m["name"] for m in META if "fourPop" in m])
def test_four_pop(m):
    hp = ARR[m["name"] + "__hap_pop"]
    for key, want in m["fourPop"].items():
        # default mode: fixtures were made on the input with exactly-tied sites blanked (see make_golden2.py)
        g = ARR[m["name"] + ("__g_aln_notie" if key.startswith(("default", "perm")) else "__g_aln")]
        if key.startswith("perm"):
            sel, md, kw = (2, 0, 1, 3), 0.4, {}
        else:
            mode, md = key.rsplit("_", 1)
            sel, md, kw = (0, 1, 2, 3), float(md), ({} if mode == "default" else {mode: True})
        got = do.four_pop(g, hp, *sel, md, **kw)
        for k in do.FOURPOP_KEYS:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-10, atol=1e-14, equal_nan=True, err_msg="%s %s" % (key, k))


@pytest.mark.parametrize("m", META, ids=IDS)
def test_target_freqs(m):
    g = ARR[m["name"] + "__g_aln"]
    hp = ARR[m["name"] + "__hap_pop"]
    P = len(m["pop_names"])
    for target in ("derived", "minor"):
        for md in (0.0, 3.0):
            for asCounts in (False, True):
                want = ARR["%s__tf_%s_%g_%d" % (m["name"], target, md, int(asCounts))]
                got, tie = do.target_freqs(g, hp, P, target, min_data=md, as_counts=asCounts)
                if target == "minor":
                    assert np.array_equal(tie, ARR[m["name"] + "__minor_tie"])
                np.testing.assert_allclose(got, want, rtol=1e-15, atol=0, equal_nan=True,
                                           err_msg="%s %g %d" % (target, md, asCounts))


def test_sample_het_quirk_is_exercised():
    """Some diploid individual with data must come out nan because bit 1 of n_ij is clear (genomics.py:924)."""
    seen = False
    for m in META:
        if m["sampleHet"]:
            seen |= bool(np.isnan([m["sampleHet"]["alone"][s] for s in m["sample_names"]]).any())
    assert seen


# ------------------------------------------------------------------------------------------------
# sfs.py (genotype input): the oracle against the reference script's own --pipe output
# ------------------------------------------------------------------------------------------------
CLI2 = json.load(open(os.path.join(GOLD, "cli_cases2.json")))["four_pops"]


def sfs_inputs():
    from genomics_general_b200 import synth
    c = CLI2["sfs_cfg"]
    spec = synth.SynthSpec(c["n_pops"], c["spp"], seed=c["seed"], miss=0.0)
    g = synth.synth_genotypes(spec, 0, c["S"])
    for a, b in CLI2["sfs_missing"]:
        g[a, b] = -1
    g[np.array(CLI2["sfs_blank"], dtype=np.int64)] = -1
    scaf = np.repeat(np.arange(len(c["scaffolds"])), c["scaffolds"])
    return spec, g, scaf


def sfs_plan(extra):
    """population order / groups / mask the reference derives from the flags of one golden run (sfs.py:369-407)"""
    import itertools
    pops = ["pop0", "pop1", "pop2", "pop3"]
    outgroup = None
    if "--polarized" in extra:
        outgroup = pops[-1]
    if "--outgroup" in extra:
        outgroup = extra[extra.index("--outgroup") + 1]
    inpops = [p for p in pops if p != outgroup]
    if "--FSpops" in extra:
        groups, cur = [], None
        for tok in extra:
            if tok == "--FSpops":
                cur = []
                groups.append(cur)
            elif tok.startswith("--"):
                cur = None
            elif cur is not None:
                cur.append(tok)
    else:
        groups = [[p] for p in inpops]
        for flag, k in (("--doPairs", 2), ("--doTrios", 3), ("--doQuartets", 4)):
            if flag in extra:
                groups += [list(c) for c in itertools.combinations(inpops, k)]
    keep = None
    for flag, inc in (("--include", True), ("--exclude", False)):
        if flag in extra:
            names = []
            for tok in extra[extra.index(flag) + 1:]:
                if tok.startswith("--"):
                    break
                names.append(tok)
            keep = (inc, [int(n[3:]) - 1 for n in names])
    return inpops, outgroup, groups, keep


GENO_SFS_KEYS = [k for k in CLI2 if k.startswith("sfs_") and not k.startswith(("sfs_base", "sfs_target")) and k + "_args" in CLI2]


@pytest.mark.parametrize("key", GENO_SFS_KEYS)
def test_sfs_oracle_matches_reference_output(key):
    spec, g, scaf = sfs_inputs()
    extra = CLI2[key + "_args"]
    inpops, outgroup, groups, keep = sfs_plan(extra)
    order = inpops + ([outgroup] if outgroup else [])
    remap = {int(p[3:]): k for k, p in enumerate(order)}
    hp = np.array([remap[x] for x in spec.hap_pop()], dtype=np.int32)
    mask = None
    if keep is not None:
        mask = np.isin(scaf, keep[1]) if keep[0] else ~np.isin(scaf, keep[1])
    chains, _ = do.sfs(g, hp, len(inpops), [tuple(inpops.index(p) for p in grp) for grp in groups],
                       outgroup=len(inpops) if outgroup else -1, site_mask=mask)
    text = "".join("\n".join("\t".join(str(x) for x in list(k) + [v]) for k, v in ch) + "\n" for ch in chains)
    assert text == CLI2[key]


@pytest.mark.parametrize("key", GENO_SFS_KEYS)
def test_sfs_row_order_from_dense_histograms(key):
    """Host logic of the sfs command line (no GPU): dense spectrum + first-site array -> the reference's sparse rows in
    nested-dict insertion order (cli/sfs.py::ordered_chains), against the reference script's own output."""
    from genomics_general_b200.cli.sfs import ordered_chains
    spec, g, scaf = sfs_inputs()
    extra = CLI2[key + "_args"]
    inpops, outgroup, groups, keep = sfs_plan(extra)
    order = inpops + ([outgroup] if outgroup else [])
    remap = {int(p[3:]): k for k, p in enumerate(order)}
    hp = np.array([remap[x] for x in spec.hap_pop()], dtype=np.int32)
    tc, used = do.sfs_target_counts(g, hp, len(inpops), len(inpops) if outgroup else -1)
    if keep is not None:
        used &= np.isin(scaf, keep[1]) if keep[0] else ~np.isin(scaf, keep[1])
    sizes = [int((hp == x).sum()) for x in range(len(inpops))]
    text = ""
    for grp in groups:
        gi = [inpops.index(p) for p in grp]
        shape = tuple(sizes[x] + 1 for x in gi)
        hist = np.zeros(shape, dtype=np.int64)
        first = np.full(shape, -1, dtype=np.int64)
        for s in np.where(used)[0]:                     # what pg_sfs's atomicAdd / atomicMin leave behind
            cell = tuple(int(tc[s, x]) for x in gi)
            hist[cell] += 1
            if first[cell] < 0:
                first[cell] = s
        text += "\n".join("\t".join(str(x) for x in row) for row in ordered_chains(hist, first)) + "\n"
    assert text == CLI2[key]


# ------------------------------------------------------------------------------------------------
# sfs.py on count tables (freq.py -> sfs.py): oracle vs the reference pipeline's output
# ------------------------------------------------------------------------------------------------
def sfs_table_plan(key):
    """(table kind, population columns in engine order, n_in, outgroup index, groups, scaffold filter) of one golden run"""
    import itertools
    extra = CLI2[key + "_args"]
    kind = CLI2[key + "_input"]
    names = ["pop0", "pop1", "pop2", "pop3"]
    pops = [extra[i + 1] for i, t in enumerate(extra) if t == "-p"]
    fsp = []
    cur = None
    for tok in extra:
        if tok == "--FSpops":
            cur = []
            fsp.append(cur)
        elif tok.startswith("-"):
            cur = None
        elif cur is not None:
            cur.append(tok)
    for p in [x for g in fsp for x in g]:
        if p not in pops:
            pops.append(p)
    if not pops:
        pops = list(names)
    outgroup = pops[-1] if (kind == "base" and "--polarized" in extra) else None
    inpops = [p for p in pops if p != outgroup]
    if fsp:
        groups = fsp
    else:
        groups = [[p] for p in inpops]
        for flag, k in (("--doPairs", 2), ("--doTrios", 3)):
            if flag in extra:
                groups += [list(c) for c in itertools.combinations(inpops, k)]
    excl = [int(extra[i + 1][3:]) - 1 for i, t in enumerate(extra) if t == "--exclude"]
    return kind, inpops + ([outgroup] if outgroup else []), len(inpops), (len(inpops) if outgroup else -1), groups, inpops, excl


def sfs_tables():
    """the two tables the reference's freq.py writes for the sfs input file: base counts [S,4,4] and derived-allele counts"""
    spec, g, scaf = sfs_inputs()
    hp = spec.hap_pop()
    base = do.site_counts(g, hp, 4)
    target, _ = do.target_freqs(g, hp, 4, "derived", as_counts=True)
    return base, target.astype(np.int64), scaf


@pytest.mark.parametrize("key", [k for k in CLI2 if k.startswith(("sfs_base", "sfs_target")) and k + "_args" in CLI2])
def test_sfs_tables_oracle_matches_reference_pipeline(key):
    base, target, scaf = sfs_tables()
    kind, order, n_in, og, groups, inpops, excl = sfs_table_plan(key)
    cols = [int(p[3:]) for p in order]
    used_mask = ~np.isin(scaf, excl) if excl else np.ones(len(scaf), dtype=bool)
    if kind == "base":
        tc, used = do.sfs_target_counts_from_counts(base[:, cols, :], n_in, og)
    else:
        tc, used = target[:, cols], np.ones(len(scaf), dtype=bool)
    chains = do.sfs_chains(tc, used & used_mask, [tuple(inpops.index(p) for p in grp) for grp in groups])
    text = "".join("\n".join("\t".join(str(x) for x in list(k) + [v]) for k, v in ch) + "\n" for ch in chains)
    assert text == CLI2[key]


def test_sfs_write_spectra_files_and_pipe(tmp_path, capsys):
    """host logic of the sfs command line: file naming (<pref><pops>_<...><suff>) and --pipe"""
    import argparse
    from genomics_general_b200.cli import sfs as sfs_cli
    hist = np.zeros((3, 4), dtype=np.int64)
    first = np.full((3, 4), -1, dtype=np.int64)
    for k, (cell, n) in enumerate([((2, 1), 5), ((0, 3), 1), ((2, 0), 7)]):
        hist[cell] = n
        first[cell] = 10 - k                      # the later cells were seen first
    args = argparse.Namespace(pipe=False, pref=str(tmp_path / "x_"), suff=".sfs")
    sfs_cli.write_spectra(args, [["a", "b"]], [hist], [first])
    assert open(str(tmp_path / "x_a_b.sfs")).read() == "2\t0\t7\n2\t1\t5\n0\t3\t1\n"
    args.pipe = True
    sfs_cli.write_spectra(args, [["a", "b"]], [hist], [first])
    assert capsys.readouterr().out == "2\t0\t7\n2\t1\t5\n0\t3\t1\n"


class _FakeEngine:
    """Stands in for the GPU engine in the CPU test below: same sfs_tables contract, histogram built with the oracle."""

    def __init__(self, device=0):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def sfs_tables(self, kind, table, n_in, groups, outgroup=-1, site_mask=None):
        table = np.asarray(table)
        n = table.shape[0]
        if kind == "base":
            tc, used = do.sfs_target_counts_from_counts(table, n_in, outgroup)
            dims = table.sum(axis=2).max(axis=0) + 1
        else:
            tc, used = table, np.ones(n, dtype=bool)
            dims = table.max(axis=0) + 1
        if site_mask is not None:
            used = used & np.asarray(site_mask, dtype=bool)
        hists, firsts = [], []
        for grp in groups:
            shape = tuple(int(dims[x]) for x in grp)
            h = np.zeros(shape, dtype=np.int64)
            f = np.full(shape, -1, dtype=np.int64)
            for s in np.where(used)[0]:
                cell = tuple(int(tc[s, x]) for x in grp)
                h[cell] += 1
                if f[cell] < 0:
                    f[cell] = s
            hists.append(h)
            firsts.append(f)
        return hists, firsts, int(used.sum())


@pytest.mark.parametrize("key", [k for k in CLI2 if k.startswith(("sfs_base", "sfs_target")) and k + "_args" in CLI2])
def test_sfs_table_command_line_host_logic(key, tmp_path, capsys, monkeypatch):
    """cli/sfs.py on count tables with the engine replaced by the oracle: table parsing, population / outgroup / FSpops
    plumbing, --exclude and the output order against the reference pipeline's text (the GPU tests run the real engine)."""
    from genomics_general_b200.cli import sfs as sfs_cli
    monkeypatch.setattr(sfs_cli, "Engine", _FakeEngine)
    base, target, scaf = sfs_tables()
    pos = np.arange(1, len(scaf) + 1)
    tab = str(tmp_path / "t.tsv")
    with open(tab, "wt") as f:
        f.write("scaffold\tposition\tpop0\tpop1\tpop2\tpop3\n")
        for s in range(len(scaf)):
            if CLI2[key + "_input"] == "base":
                cols = [",".join(str(v) for v in base[s, x]) for x in range(4)]
            else:
                cols = [str(v) for v in target[s]]
            f.write("chr%d\t%d\t%s\n" % (scaf[s] + 1, pos[s], "\t".join(cols)))
    capsys.readouterr()
    sfs_cli.main(["-i", tab, "--pipe"] + CLI2[key + "_args"])
    assert capsys.readouterr().out == CLI2[key]
