#!/usr/bin/env python
"""Drop-in for the reference's distMat.py (flags 116-159, worker 28-60, writers genomics.py:2288-2306), on the GPU."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .. import genomics, windows as W
from ..engine import Engine
from . import _common as C


def build_parser():
    p = argparse.ArgumentParser()
    C.add_window_args(p, overlap_short=True, cat=True)
    p.add_argument("-Mi", "--minPerInd", type=int, metavar="sites")
    p.add_argument("--samples", nargs="+", metavar="sample names")
    p.add_argument("--includeSameWithSame", action="store_true")
    p.add_argument("--ploidy", type=int, nargs="+")
    p.add_argument("--ploidyFile")
    p.add_argument("--haploid", nargs="+", metavar="sample names")
    p.add_argument("--inferPloidy", action="store_true")
    p.add_argument("-g", "--genoFile")
    p.add_argument("-o", "--outFile")
    p.add_argument("--windowDataOutFile")
    p.add_argument("-f", "--genoFormat", choices=("phased", "pairs", "haplo", "diplo"), required=True)
    p.add_argument("--outFormat", choices=("raw", "phylip", "nexus"), default="phylip")
    p.add_argument("--headers", nargs="+")
    p.add_argument("--roundTo", type=int, default=4)
    p.add_argument("--exclude")
    p.add_argument("--include")
    p.add_argument("-T", "--threads", type=int, default=1, metavar="threads")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--addWindowID", action="store_true")
    p.add_argument("--writeFailedWindows", action="store_true")
    C.add_engine_args(p)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    minSites = args.minSites
    coords = None
    if args.windType == "cat":
        minSites = 1
    else:
        minSites, coords = C.check_window_args(args)
    if args.samples:
        samples = args.samples
    elif args.headers:
        samples = args.headers[2:]
    else:
        assert args.genoFile, "If piping from stdin, you need to specify either --samples or --headers"
        samples = C.header_names(args.genoFile)
    ploidyDict = C.ploidy_dict(args, samples, args.haploid)
    sampleData = genomics.SampleData(indNames=list(samples), ploidyDict=ploidyDict)
    header = "\t".join(args.headers) if args.headers else None
    eng = Engine(args.device)
    gd = C.load_geno(args, sampleData.indNames, ploidyDict, header=header, engine=eng)
    if args.windType == "cat":
        ws = W.WindowSet()
        ws.add(None, -np.inf, np.inf, 0, gd.n_sites, None)               # parseGenoFile: one window, positions ignored
    else:
        ws = C.make_windows(args, gd, minSites, coords, C.read_scaffold_list(args.include), C.read_scaffold_list(args.exclude))
    out = C.open_out(args.outFile)
    wout = None
    if args.windowDataOutFile:
        wout = C.open_out(args.windowDataOutFile)
        wout.write("scaffold,start,end,mid,sites," if not args.addWindowID else "windowID,scaffold,start,end,mid,sites,")
    lo, hi = ws.ranges()
    nInd = len(sampleData.indNames)
    hap_ind = gd.hap_sample()
    with eng:
        C.ensure_resident(eng, gd)
        eng.set_windows(lo, hi)
        if args.windType == "cat":
            dcat, ntot = eng.pairdist_cat(hap_ind, nInd, args.includeSameWithSame)      # chunked over the site axis
            r = dict(dist=dcat[None], sites=np.array([ntot], dtype=np.int64), pos_sum=np.zeros(1, dtype=np.int64))
        else:
            r = eng.pairdist(hap_ind, nInd, args.includeSameWithSame)
        per_ind_ok = np.ones(len(ws), dtype=bool)
        if args.minPerInd:
            per_ind_ok = eng.seq_nonnan().min(axis=1) >= args.minPerInd             # min(aln.seqNonNan()) (distMat.py:40)
    for k in range(len(ws)):
        sites = int(r["sites"][k])
        good = sites >= minSites and bool(per_ind_ok[k])
        m = r["dist"][k] if good else np.full((nInd, nInd), np.nan)
        if args.outFormat == "nexus":
            s = genomics.makeDistMatNexusString(m, names=sampleData.indNames, roundTo=args.roundTo)
        elif args.outFormat == "phylip":
            s = genomics.makeDistMatPhylipString(m, names=sampleData.indNames, roundTo=args.roundTo)
        else:
            s = genomics.makeDistMatString(m, roundTo=args.roundTo) + "\n"
        if good or args.writeFailedWindows:
            out.write(s)
            if wout is not None:
                if args.windType == "cat":
                    pre = [None, -np.inf, np.inf, np.nan, sites]
                else:
                    pre = C.window_prefix(args, ws, k, gd, sites, r["pos_sum"][k])
                wout.write("\t".join(str(x) for x in (([] if not args.addWindowID else [ws.ID[k]]) + pre)) + "\n")
    if out is not sys.stdout:
        out.close()
    if wout is not None and wout is not sys.stdout:
        wout.close()


if __name__ == "__main__":
    main()
