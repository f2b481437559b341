#!/usr/bin/env python
"""Drop-in for the reference's sfs.py (flags sfs.py:158-236): site frequency spectra of each population and, on request,
the joint spectra of pairs / trios / quartets, computed on the GPU — from genotypes (`--inputType genotypes`: per-site
counts -> target allele -> dense histograms, `pg_sfs`) or from tables of counts (`baseCounts`: the rows freq.py writes;
`targetCounts`, the script's default: e.g. freq.py --target derived --asCounts; `pg_sfs_tables`) — and written in the
reference's sparse format and order.

Not covered: `--subsample` (the reference draws with numpy's global RNG per site), `--regions`.  Where the reference's choice of the minor allele depends on numpy's unstable
sort (two alleles with exactly equal counts, sfs.py:90) the lower allele is used.
"""
from __future__ import annotations

import argparse
import itertools
import sys

import numpy as np

from .. import genomics
from ..engine import Engine
from . import _common as C


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("-i", "--inputFile")
    p.add_argument("--inputType", choices=("genotypes", "baseCounts", "targetCounts"), default="targetCounts")
    p.add_argument("--scafCol", type=int, default=0)
    p.add_argument("--posCol", type=int, default=1)
    p.add_argument("--firstSampleCol", type=int, default=2)
    p.add_argument("--header")
    p.add_argument("--genoFormat", choices=("phased", "diplo", "alleles"), default="phased")
    p.add_argument("-p", "--pop", action="append", nargs="+", metavar=("popName", "[samples]"))
    p.add_argument("--popsFile")
    p.add_argument("--ploidy", type=int, nargs="+")
    p.add_argument("--ploidyFile")
    p.add_argument("--FSpops", action="append", type=str, nargs="+")
    p.add_argument("--doPairs", action="store_true")
    p.add_argument("--doTrios", action="store_true")
    p.add_argument("--doQuartets", action="store_true")
    p.add_argument("--subsample", nargs="+", type=int)
    p.add_argument("--subsampleIndividuals", action="store_true")
    p.add_argument("--pref", default="")
    p.add_argument("--suff", default=".sfs")
    p.add_argument("--pipe", action="store_true")
    p.add_argument("--polarized", action="store_true")
    p.add_argument("--outgroup")
    p.add_argument("--include", nargs="+")
    p.add_argument("--includeFile")
    p.add_argument("--exclude", nargs="+")
    p.add_argument("--excludeFile")
    p.add_argument("--regions", nargs="+")
    p.add_argument("--regionsFile")
    p.add_argument("-R", "--report", default=100000)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--seed", type=int, default=42)
    C.add_engine_args(p)
    return p


def ordered_chains(hist, first):
    """Dense spectrum + first-site array -> rows [k1, .., kd, count] in the order the reference writes its nested
    SparseFS dictionaries (sfs.py:117-125): at every nesting level, keys in order of first appearance."""
    nz = np.argwhere(hist > 0)
    if len(nz) == 0:
        return []
    f = first[tuple(nz.T)]
    keys = []
    for lev in range(nz.shape[1]):
        _, inv = np.unique(nz[:, :lev + 1], axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        m = np.full(inv.max() + 1, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(m, inv, f)
        keys.append(m[inv])
    order = np.lexsort(tuple(reversed(keys)))
    cnt = hist[tuple(nz.T)]
    return [list(map(int, nz[i])) + [int(cnt[i])] for i in order]


def write_spectra(args, FSpops, hists, firsts):
    """one file per spectrum (<pref><pops joined by _><suff>, sfs.py:495-497) or everything to stdout (--pipe, 491-493)"""
    for i, grp in enumerate(FSpops):
        text = "\n".join("\t".join(str(x) for x in row) for row in ordered_chains(hists[i], firsts[i])) + "\n"
        if args.pipe:
            sys.stdout.write(text)
        else:
            with open(args.pref + "_".join(grp) + args.suff, "w") as out:
                out.write(text)


def main_tables(args, include, exclude):
    """--inputType baseCounts | targetCounts (sfs.py:330-365, 456-474): one column per population."""
    import gzip
    import io
    import pandas as pd
    opener = gzip.open if args.inputFile and args.inputFile.endswith(".gz") else open
    with (opener(args.inputFile, "rt") if args.inputFile else sys.stdin) as f:
        header = args.header if args.header else f.readline()
        body = "".join(line for line in f if line[:1] != "#")
    names = header.split()[2:]
    popNames = []
    if args.pop or args.FSpops:
        for pop in args.pop or []:
            popNames.append(pop[0])
        for pop in [p for pops in (args.FSpops or []) for p in pops]:
            if pop not in popNames:
                popNames.append(pop)
    else:
        popNames = list(names)
    sys.stderr.write("\nPopulations:\n" + " ".join(popNames) + "\n")
    outgroup = None
    inPopNames = list(popNames)
    if args.inputType == "baseCounts" and (args.polarized or args.outgroup):
        outgroup = args.outgroup if args.outgroup else popNames[-1]
        inPopNames = [pn for pn in popNames if pn != outgroup]
        sys.stderr.write("\nFrequencies will be polarized assuming outgroup is {}\n".format(outgroup))
    if args.FSpops:
        FSpops = [list(g) for g in args.FSpops]
    else:
        FSpops = [[pn] for pn in inPopNames]
        if args.doPairs:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 2)]
        if args.doTrios:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 3)]
        if args.doQuartets:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 4)]
    df = pd.read_csv(io.StringIO(body), sep=r"\s+", header=None, names=["scaffold", "position"] + names, dtype=str)
    order = inPopNames + ([outgroup] if outgroup else [])
    mask = None
    if include or exclude:
        sc = df["scaffold"].to_numpy()
        mask = np.array([(not include or x in include) and (x not in exclude) for x in sc], dtype=np.uint8)
    if args.inputType == "baseCounts":
        cols = []
        for pn in order:
            parts = df[pn].str.split(",", expand=True).to_numpy(dtype=np.float64)      # "a,c,g,t" (floats allowed, 459)
            cols.append(parts.astype(np.int64))
        table = np.stack(cols, axis=1) if len(df) else np.zeros((0, len(order), 4), dtype=np.int64)
        assert table.max(initial=0) <= 65535, "counts above 65535 are not supported"
        kind = "base"
    else:
        table = df[order].to_numpy(dtype=np.int64) if len(df) else np.zeros((0, len(order)), dtype=np.int64)
        kind = "target"
    groups = [tuple(inPopNames.index(pn) for pn in grp) for grp in FSpops]
    with Engine(args.device) as eng:
        hists, firsts, _ = eng.sfs_tables(kind, table, len(inPopNames), groups, outgroup=len(inPopNames) if outgroup else -1,
                                          site_mask=mask)
    write_spectra(args, FSpops, hists, firsts)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.subsample or args.subsampleIndividuals:
        raise NotImplementedError("--subsample draws with numpy's global RNG per site in the reference; not supported")
    if args.regions or args.regionsFile:
        raise NotImplementedError("--regions is not supported")
    assert (args.scafCol, args.posCol, args.firstSampleCol) == (0, 1, 2), "non-default column layout is not supported"
    if not args.polarized and args.outgroup is None and args.inputType != "targetCounts":
        sys.stderr.write("\nNo outgroup provided. Minor allele frequency will be used.\n")
    include = set(args.include or [])
    exclude = set(args.exclude or [])
    if args.includeFile:
        include |= set(open(args.includeFile, "rt").read().split())
    if args.excludeFile:
        exclude |= set(open(args.excludeFile, "rt").read().split())
    if args.inputType != "genotypes":
        return main_tables(args, include, exclude)

    headerInds = C.header_names(args.inputFile) if args.header is None else args.header.split()[2:]
    popNames, popDict = [], {}
    if args.pop or args.FSpops:                                         # sfs.py:291-309
        for pop in args.pop or []:
            popNames.append(pop[0])
            popDict[pop[0]] = [] if len(pop) == 1 else pop[1].split(",")
        for pop in [p for pops in (args.FSpops or []) for p in pops]:
            if pop not in popNames:
                popNames.append(pop)
                popDict[pop] = []
        if args.popsFile:
            with open(args.popsFile, "rt") as pf:
                for line in pf:
                    if not line.strip():
                        continue
                    ind, pop = line.split()
                    if pop in popDict and ind not in popDict[pop]:
                        popDict[pop].append(ind)
    else:
        popNames, popDict = ["all"], {"all": list(headerInds)}
    for pn in popNames:
        assert len(popDict[pn]) >= 1, "Population {} has no samples".format(pn)
    allSamples = [s for pn in popNames for s in popDict[pn]]
    if args.ploidy is not None:
        ploidy = args.ploidy if len(args.ploidy) != 1 else args.ploidy * len(allSamples)
        assert len(ploidy) == len(allSamples), "Incorrect number of ploidy values supplied."
        ploidyDict = dict(zip(allSamples, ploidy))
    elif args.ploidyFile is not None:
        with open(args.ploidyFile, "rt") as pf:
            ploidyDict = dict([[s[0], int(s[1])] for s in [l.split() for l in pf if l.strip()]])
    else:
        ploidyDict = dict(zip(allSamples, [2] * len(allSamples)))
    sys.stderr.write("\nPopulations:\n" + " ".join(popNames) + "\n")
    outgroup = None
    inPopNames = list(popNames)
    if args.polarized or args.outgroup:                                  # sfs.py:369-373
        outgroup = args.outgroup if args.outgroup else popNames[-1]
        inPopNames = [pn for pn in popNames if pn != outgroup]
        sys.stderr.write("\nFrequencies will be polarized assuming outgroup is {}\n".format(outgroup))
    if args.FSpops:
        FSpops = [list(g) for g in args.FSpops]
    else:
        FSpops = [[pn] for pn in inPopNames]
        if args.doPairs:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 2)]
        if args.doTrios:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 3)]
        if args.doQuartets:
            FSpops += [list(c) for c in itertools.combinations(inPopNames, 4)]

    # engine populations: the in-group first, the outgroup last
    enginePops = inPopNames + ([outgroup] if outgroup else [])
    sampleData = genomics.SampleData(indNames=list(allSamples), popNames=enginePops,
                                     popInds=[popDict[pn] for pn in enginePops], ploidyDict=ploidyDict)
    args.genoFile, args.hostParse = args.inputFile, getattr(args, "hostParse", False)
    eng = Engine(args.device)
    gd = C.load_geno(args, sampleData.indNames, ploidyDict, header=args.header, engine=eng)
    mask = None
    if include or exclude:
        ok = np.array([(not include or n in include) and (n not in exclude) for n in gd.scaf_names], dtype=np.uint8)
        mask = ok[gd.scaf_ids]
    with eng:
        C.ensure_resident(eng, gd)
        hp = C.hap_pop_vector(gd, enginePops, [popDict[pn] for pn in enginePops])
        eng.set_pops(hp, len(enginePops))
        sizes = [int((hp == x).sum()) for x in range(len(enginePops))]
        groups = [tuple(inPopNames.index(pn) for pn in grp) for grp in FSpops]
        hists, firsts, _ = eng.sfs(len(inPopNames), groups, sizes, outgroup=len(inPopNames) if outgroup else -1, site_mask=mask)
    write_spectra(args, FSpops, hists, firsts)


if __name__ == "__main__":
    main()
