#!/usr/bin/env python
"""Drop-in for the reference's freq.py (flags 187-222), on the GPU: the default mode (per-site per-population A,C,G,T
counts, freq.py:52-58,100-111) and --target derived|minor (frequency or count of one allele per population,
freq.py:62-98; the reference's random pick between exactly tied minor alleles becomes "the lower allele")."""
from __future__ import annotations

import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import genomics, geno_io, mgpu
from ..engine import Engine
from . import _common as C


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("-g", "--genoFile")
    p.add_argument("-o", "--outFile")
    p.add_argument("-f", "--genoFormat", choices=("phased", "diplo", "alleles"), default="phased")
    p.add_argument("-p", "--population", action="append", nargs="+", metavar=("popName", "[samples]"))
    p.add_argument("--popsFile")
    p.add_argument("--indFreqs", action="store_true")
    p.add_argument("--target", choices=("minor", "derived"), default=None)
    p.add_argument("--asCounts", action="store_true")
    p.add_argument("--ploidy", type=int, nargs="+")
    p.add_argument("--ploidyFile")
    p.add_argument("--haploid", nargs="+")
    p.add_argument("--minData", type=float, default=0, metavar="proportion")
    p.add_argument("--threshold", type=float, metavar="proportion")
    p.add_argument("--keepNanLines", action="store_true")
    p.add_argument("-t", "--threads", type=int, default=1)
    p.add_argument("--sliceSize", type=int, default=1000000)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--test", action="store_true")
    C.add_engine_args(p)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    headerInds = C.header_names(args.genoFile)
    if not args.indFreqs and not args.population:
        if args.target == "derived":       # freq.py:238-242
            sys.stderr.write("\nNo populations specified. Assuming the final individual is the outgroup for polarising.\n")
            popNames, popInds = ["ingroup", "outgroup"], [headerInds[:-1], [headerInds[-1]]]
        else:
            popNames, popInds = ["all"], [headerInds]
    elif args.indFreqs:
        popNames, popInds = list(headerInds), [[i] for i in headerInds]
    else:
        popNames, popInds = C.parse_pop_args(args.population, args.popsFile)
    allInds = sorted(set(i for p in popInds for i in p))
    args.inferPloidy = False
    ploidyDict = C.ploidy_dict(args, allInds, args.haploid)
    for s in args.haploid or []:           # freq.py:283-284: --haploid also overrides --ploidy / --ploidyFile
        ploidyDict[s] = 1
    sampleData = genomics.SampleData(indNames=allInds, popNames=popNames, popInds=popInds, ploidyDict=ploidyDict)
    # --devices N: every rank tokenises its share of the file and formats its own rows (freq.py:328-360's slices, one
    # process per GPU here); rank 0 then appends the parts in rank order
    rdv = mgpu.init("genomics_general_b200.cli.freq", argv, args.devices)
    if rdv is None or rdv.rank == 0:
        out = C.open_out(args.outFile)
        out.write("scaffold\tposition\t" + "\t".join(popNames) + "\n")
    else:
        out = open(os.path.join(rdv.dir, "rows.r%d.part" % rdv.rank), "wb")
    tm = C.Timing(args.timing if (rdv is None or rdv.rank == 0) else None)
    eng = Engine(args.device if rdv is None else mgpu.device_for(rdv, args.device))
    if rdv is None:
        gd = C.load_geno(args, sampleData.indNames, ploidyDict, engine=eng)
    else:
        gd = mgpu.local_ingest(eng, rdv, args.genoFile, args.genoFormat, sampleData.indNames, ploidyDict)
    P = len(popNames)
    tm.mark("ingest", eng)
    busy = dict(fetch=0.0, format=0.0, write=0.0)
    with eng:
        C.ensure_resident(eng, gd)
        eng.set_pops(C.hap_pop_vector(gd, popNames, popInds), P)
        # rows are formatted by native host threads (pg_format_freq_rows) and written as bytes
        raw = out if (rdv is not None and rdv.rank != 0) else (out.buffer if hasattr(out, "buffer") else None)
        if raw is not None:
            out.flush()

        def fetch(s, n):
            """one slab of per-site values off the device: (formatter mode, values, keep mask)"""
            t0 = time.perf_counter()
            try:
                if not args.target:
                    return 0, eng.site_counts(s, n), None
                # freq.py:302-304: with a target the user's --asCounts / --keepNanLines / --minData apply
                v, _ = eng.site_target_freqs(args.target, s, n, min_data=args.minData, as_counts=args.asCounts)
                if args.asCounts:
                    return 2, v, (None if args.keepNanLines else ~np.all(v == 0, axis=1))
                v = np.around(v, 4)                                       # freq.py:91
                if args.threshold:                                        # freq.py:96-98
                    hi, lo = v >= args.threshold, v < args.threshold
                    v[hi] = 1
                    v[lo] = 0
                return 1, v, (None if args.keepNanLines else ~np.all(np.isnan(v), axis=1))
            finally:
                busy["fetch"] += time.perf_counter() - t0

        def emit(segments):
            t0 = time.perf_counter()
            for seg in segments:
                if raw is not None:
                    raw.write(seg)
                else:
                    out.write(bytes(seg).decode())
            busy["write"] += time.perf_counter() - t0

        # three stages in flight: the device pass + D2H of slab k+1 and the write of slab k-1 run on two helper threads
        # under the formatting of slab k (every stage leaves the interpreter lock: ctypes calls and file writes)
        slab = 1 << 18
        starts = list(range(0, gd.n_sites, slab))
        with ThreadPoolExecutor(1) as fetcher, ThreadPoolExecutor(1) as writer:
            nxt = fetcher.submit(fetch, starts[0], min(slab, gd.n_sites - starts[0])) if starts else None
            wrote = None
            for k, s in enumerate(starts):
                n = min(slab, gd.n_sites - s)
                mode, v, keep = nxt.result()
                if k + 1 < len(starts):
                    nxt = fetcher.submit(fetch, starts[k + 1], min(slab, gd.n_sites - starts[k + 1]))
                t0 = time.perf_counter()
                segs = geno_io.format_freq_rows(mode, v, gd.pos[s:s + n], gd.scaf_ids[s:s + n], gd.scaf_names, keep)
                busy["format"] += time.perf_counter() - t0
                if wrote is not None:
                    wrote.result()
                wrote = writer.submit(emit, segs)
            if wrote is not None:
                wrote.result()
    tm.mark("rows")
    tm.write(sites=int(gd.n_sites), haplotypes=int(gd.n_haps), populations=P, stage_busy_s=busy,
             devices=(1 if rdv is None else rdv.world))
    if rdv is not None:
        if rdv.rank != 0:
            out.close()
            rdv.put_bytes("rows_done", b"1")
            rdv.finish()
            return
        out.flush()
        sink = out.buffer if hasattr(out, "buffer") else None
        for q in range(1, rdv.world):
            rdv.get_bytes("rows_done", q)
            with open(os.path.join(rdv.dir, "rows.r%d.part" % q), "rb") as part:
                while True:
                    blk = part.read(1 << 24)
                    if not blk:
                        break
                    if sink is not None:
                        sink.write(blk)
                    else:
                        out.write(blk.decode())
    if out is not sys.stdout:
        out.close()
    if rdv is not None:
        rdv.finish()
    sys.stderr.write("\nDone\n")


if __name__ == "__main__":
    main()
