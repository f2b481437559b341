#!/usr/bin/env python
"""Drop-in for the reference's distMat.py (flags 116-159, worker 28-60, writers genomics.py:2288-2306), on the GPU."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .. import genomics, mgpu, windows as W
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
    # --devices N: every rank tokenises its share of the file.  Window modes: a rank computes + formats the windows that start
    # in its share; there is no collective — the formatted windows travel through the exchange directory and rank 0 writes
    # them in window order (distMat.py:317-353's worker pool and sorter).  `cat` (one window over every site): the ranks hold
    # shards of the SITES and pg_pairdist_cat adds the integer pair matrices with ONE ncclAllReduce before the division.
    rdv = mgpu.init("genomics_general_b200.cli.distMat", argv, args.devices)
    eng = Engine(args.device if rdv is None else mgpu.device_for(rdv, args.device))
    if rdv is None:
        gd = C.load_geno(args, sampleData.indNames, ploidyDict, header=header, engine=eng)
    elif args.windType == "cat":
        gd = mgpu.local_ingest(eng, rdv, args.genoFile, args.genoFormat, sampleData.indNames, ploidyDict, header)
    else:
        gd, starts, off_all = mgpu.sharded_ingest(eng, rdv, args.genoFile, args.genoFormat, sampleData.indNames, ploidyDict, header)
    if args.windType == "cat":
        ws = W.WindowSet()
        ws.add(None, -np.inf, np.inf, 0, gd.n_sites, None)               # parseGenoFile: one window, positions ignored
    else:
        ws = C.make_windows(args, gd, minSites, coords, C.read_scaffold_list(args.include), C.read_scaffold_list(args.exclude))
    writer = rdv is None or rdv.rank == 0
    out = C.open_out(args.outFile) if writer else None
    wout = None
    if args.windowDataOutFile and writer:
        wout = C.open_out(args.windowDataOutFile)
        wout.write("scaffold,start,end,mid,sites," if not args.addWindowID else "windowID,scaffold,start,end,mid,sites,")
    lo, hi = ws.ranges()
    nInd = len(sampleData.indNames)
    hap_ind = np.repeat(np.arange(len(gd.names), dtype=np.int32), np.asarray(gd.ploidy, dtype=np.int64))
    mine = np.arange(len(ws))                     # windows this process computes (all of them on one device)
    with eng:
        if rdv is None:
            C.ensure_resident(eng, gd)
            eng.set_windows(lo, hi)
        elif args.windType == "cat":
            eng.set_windows(lo, hi)               # the rank's shard of the one window
            mgpu.nccl_connect(eng, rdv)
        else:
            mine, llo, lhi, halo = mgpu.assign_windows(lo, hi, starts, rdv.rank)
            mgpu.fetch_halo(eng, args.genoFile, gd, starts, off_all, rdv.rank, halo, args.genoFormat, ploidyDict)
            eng.set_windows(llo, lhi)
        if len(mine) == 0:
            r = dict(dist=np.zeros((0, nInd, nInd)), sites=np.zeros(0, np.int64), pos_sum=np.zeros(0, np.int64))
            per_ind_ok = np.ones(0, dtype=bool)
        elif args.windType == "cat":
            dcat, ntot = eng.pairdist_cat(hap_ind, nInd, args.includeSameWithSame)      # chunked over the site axis
            r = dict(dist=dcat[None], sites=np.array([ntot], dtype=np.int64), pos_sum=np.zeros(1, dtype=np.int64))
        else:
            r = eng.pairdist(hap_ind, nInd, args.includeSameWithSame)
        if len(mine):
            per_ind_ok = np.ones(len(mine), dtype=bool)
            if args.minPerInd:
                nn = eng.seq_nonnan()
                if rdv is not None and args.windType == "cat":                      # the counts of the shards add up
                    nn = np.sum(rdv.allgather("cat_nonnan", nn), axis=0)
                per_ind_ok = nn.min(axis=1) >= args.minPerInd                        # min(aln.seqNonNan()) (distMat.py:40)
        if rdv is not None and args.windType == "cat":
            eng.nccl_finalize()
            if rdv.rank != 0:                     # every rank holds the same matrix; rank 0 writes it
                rdv.finish()
                return
            rdv.finish()
            rdv = None
    texts, wrows = {}, {}                         # window index -> matrix text / window-data row of the windows that are written
    for j, k in enumerate(mine):
        k = int(k)
        sites = int(r["sites"][j])
        good = sites >= minSites and bool(per_ind_ok[j])
        m = r["dist"][j] if good else np.full((nInd, nInd), np.nan)
        if args.outFormat == "nexus":
            s = genomics.makeDistMatNexusString(m, names=sampleData.indNames, roundTo=args.roundTo)
        elif args.outFormat == "phylip":
            s = genomics.makeDistMatPhylipString(m, names=sampleData.indNames, roundTo=args.roundTo)
        else:
            s = genomics.makeDistMatString(m, roundTo=args.roundTo) + "\n"
        if good or args.writeFailedWindows:
            wrow = None
            if args.windowDataOutFile:
                if args.windType == "cat":
                    pre = [None, -np.inf, np.inf, np.nan, sites]
                else:
                    pre = C.window_prefix(args, ws, k, gd, sites, r["pos_sum"][j])
                wrow = "\t".join(str(x) for x in (([] if not args.addWindowID else [ws.ID[k]]) + pre)) + "\n"
            if rdv is None:                       # one device: written as they come
                out.write(s)
                if wout is not None:
                    wout.write(wrow)
            else:
                texts[k] = s
                if wrow is not None:
                    wrows[k] = wrow
    if rdv is not None:
        # a rank's windows as one blob: window indices, text lengths, the texts back to back (same for the window-data rows)
        keys = sorted(texts)
        rdv.put("dm_idx", np.array(keys, dtype=np.int64))
        rdv.put("dm_len", np.array([len(texts[k].encode()) for k in keys], dtype=np.int64))
        rdv.put_bytes("dm_txt", "".join(texts[k] for k in keys).encode())
        rdv.put("dm_wlen", np.array([len(wrows[k].encode()) if k in wrows else 0 for k in keys], dtype=np.int64))
        rdv.put_bytes("dm_wtxt", "".join(wrows.get(k, "") for k in keys).encode())
        if rdv.rank != 0:
            rdv.finish()
            return
        texts, wrows = {}, {}
        for q in range(rdv.world):
            keys, blob, wblob = rdv.get("dm_idx", q), rdv.get_bytes("dm_txt", q), rdv.get_bytes("dm_wtxt", q)
            o = wo = 0
            for k, n, wn in zip(keys, rdv.get("dm_len", q), rdv.get("dm_wlen", q)):
                texts[int(k)] = blob[o:o + int(n)].decode()
                o += int(n)
                if wn:
                    wrows[int(k)] = wblob[wo:wo + int(wn)].decode()
                    wo += int(wn)
    for k in sorted(texts):
        out.write(texts[k])
        if wout is not None and k in wrows:
            wout.write(wrows[k])
    if out is not sys.stdout:
        out.close()
    if wout is not None and wout is not sys.stdout:
        wout.close()
    if rdv is not None:
        rdv.finish()


if __name__ == "__main__":
    main()
