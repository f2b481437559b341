#!/usr/bin/env python
"""Drop-in for the reference's fourPopWindows.py (flags 107-150, header 245-250, worker 27-52), on the GPU.
(The reference script itself stops on numpy >= 2 at its `np.NaN`, fourPopWindows.py:36; genomics.fourPop runs.)"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .. import genomics, mgpu, multigpu
from ..engine import Engine
from . import _common as C

STATS = ["ABBA", "BABA", "ABAA", "BAAA", "D", "fd", "fd'", "fdm", "fdm'", "fdh", "fdh2", "fh"]     # fourPopWindows.py:248


def build_parser():
    p = argparse.ArgumentParser()
    C.add_window_args(p, overlap_short=False)        # -O is the outgroup here
    p.add_argument("--minData", help="Min proportion of samples genotped per site", type=float, default=0.01,
                   metavar="proportion")
    for flag, long in (("-P1", "--pop1"), ("-P2", "--pop2"), ("-P3", "--pop3"), ("-O", "--outgroup")):
        p.add_argument(flag, long, help="Pop name and optionally sample names (separated by commas)", required=True,
                       nargs="+", metavar=("popName", "[samples]"))
    p.add_argument("--popsFile", help="Optional file of sample names and populations")
    p.add_argument("--ploidy", type=int, nargs="+")
    p.add_argument("--ploidyFile")
    p.add_argument("--haploid", metavar="sample names")
    p.add_argument("--inferPloidy", action="store_true")
    p.add_argument("--polarize", help="Ensure outgroup is fixed for ancestral allele", action="store_true")
    p.add_argument("--fixed", help="Only count fixed SNPs", action="store_true")
    p.add_argument("-g", "--genoFile")
    p.add_argument("-o", "--outFile")
    p.add_argument("--exclude")
    p.add_argument("--include")
    p.add_argument("-f", "--genoFormat", choices=("phased", "pairs", "haplo", "diplo"), required=True)
    p.add_argument("--header")
    p.add_argument("-T", "--Threads", type=int, default=1, metavar="threads")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--addWindowID", action="store_true")
    p.add_argument("--writeFailedWindows", action="store_true")
    C.add_engine_args(p)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    minSites, coords = C.check_window_args(args, with_id=True)
    assert 0 <= args.minData <= 1, "minimum data per site must be between 0 and 1."
    popNames, popInds = C.parse_pop_args([args.pop1, args.pop2, args.pop3, args.outgroup], args.popsFile)
    allInds = sorted(set(i for p in popInds for i in p))
    ploidyDict = C.ploidy_dict(args, allInds, args.haploid.split(",") if args.haploid else None)
    sampleData = genomics.SampleData(indNames=allInds, popNames=popNames, popInds=popInds, ploidyDict=ploidyDict)
    # --devices N: this process becomes rank 0 of N (fourPopWindows.py:253-330's worker pool, one process per GPU here)
    rdv = mgpu.init("genomics_general_b200.cli.fourPopWindows", argv, args.devices)
    out = C.open_out(args.outFile) if (rdv is None or rdv.rank == 0) else open(os.devnull, "wt")
    out.write(",".join((["windowID"] if args.addWindowID else []) + ["scaffold", "start", "end", "mid", "sites", "sitesUsed"]
                       + STATS) + "\n")
    eng = Engine(args.device if rdv is None else mgpu.device_for(rdv, args.device))
    if rdv is None:
        gd = C.load_geno(args, sampleData.indNames, ploidyDict, header=args.header, engine=eng)
    else:
        gd, starts, off_all = mgpu.sharded_ingest(eng, rdv, args.genoFile, args.genoFormat, sampleData.indNames, ploidyDict,
                                                  args.header)
    ws = C.make_windows(args, gd, minSites, coords, C.read_scaffold_list(args.include), C.read_scaffold_list(args.exclude))
    lo, hi = ws.ranges()
    written = 0
    with eng:
        if rdv is None:
            C.ensure_resident(eng, gd)
            eng.set_windows(lo, hi)
            eng.set_pops(C.hap_pop_vector(gd, popNames, popInds), 4)
            r = eng.fourpop(0, 1, 2, 3, args.minData, polarize=args.polarize, fixed=args.fixed)
        else:
            idx, llo, lhi, halo = mgpu.assign_windows(lo, hi, starts, rdv.rank)
            mgpu.fetch_halo(eng, args.genoFile, gd, starts, off_all, rdv.rank, halo, args.genoFormat, ploidyDict)
            eng.set_windows(llo, lhi)
            eng.set_pops(C.hap_pop_vector(gd, popNames, popInds), 4)
            w_max, row_of = mgpu.gathered_order([mgpu.assign_windows(lo, hi, starts, q)[0] for q in range(rdv.world)])
            mgpu.nccl_connect(eng, rdv)
            table = np.zeros((rdv.world * w_max, 17), dtype=np.float64)
            eng.fourpop_allgather(0, 1, 2, 3, args.minData, w_max, table, polarize=args.polarize,
                                  fixed=args.fixed)                             # ONE ncclAllGather of the records
            eng.nccl_finalize()
            rows = np.array([row_of[w] for w in range(len(ws))], dtype=np.int64)
            r = multigpu.unpack_fourpop_records(table[rows] if len(rows) else table[:0])
            if rdv.rank != 0:
                rdv.finish()
                return
    for k in range(len(ws)):
        pre = C.window_prefix(args, ws, k, gd, r["sites"][k], r["pos_sum"][k])
        sitesUsed = np.nan
        good = False
        vals = [np.nan] * len(STATS)
        if pre[4] >= minSites:
            sitesUsed = int(r["sitesUsed"][k])
            if sitesUsed >= minSites:
                good = True
                vals = [round(np.float64(r[s][k]), 4) for s in STATS]
        if good or args.writeFailedWindows:
            row = ([] if not args.addWindowID else [ws.ID[k]]) + pre + [sitesUsed] + vals
            out.write(",".join(str(x) for x in row) + "\n")
            written += 1
    if out is not sys.stdout:
        out.close()
    if rdv is not None:
        rdv.finish()
    sys.stderr.write("%d windows were tested.\n%d results were written.\n" % (len(ws), written))


if __name__ == "__main__":
    main()
