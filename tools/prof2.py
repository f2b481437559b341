"""One pass of each hot kernel for ncu (round 1, second half): K1 popgen / ABBA / fourPop on the C2 shape, K1 popgen and
counts on the C5 row shape, the device-side text tokenizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import geno_io, synth
from genomics_general_b200.engine import Engine

with Engine(0) as eng:
    S = 10_000_000
    spec = synth.SynthSpec(4, 50, miss=0.0, seed=20260925)
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, np.minimum(lo + 5000, S))
    eng.popgen(100, 0.01)
    eng.abbababa(0, 1, 2, 3, 0.5)
    eng.fourpop(0, 1, 2, 3, 0.5)
    print("C2", eng.last_timings())
    S5 = 5_000_000
    spec5 = synth.SynthSpec(8, 100, miss=0.0, seed=5)
    eng.synth_fill(spec5, S5)
    eng.set_pops(spec5.hap_pop(), 8)
    lo = np.arange(0, S5, 5000, dtype=np.int64)
    eng.set_windows(lo, np.minimum(lo + 5000, S5))
    eng.popgen(100, 0.01)
    print("C5", eng.last_timings())
    eng.site_counts(0, 2_000_000)
    # text tokenizer: 200 k lines of the C2 width
    St = 200_000
    spec_t = synth.SynthSpec(4, 50, miss=0.0, seed=9)
    gt = synth.synth_genotypes(spec_t, 0, St)
    pos_t = synth.synth_positions(St, seed=9)
    nS = 200
    width = 5 + 9 + nS * 4 + 1
    txt = np.empty((St, width), dtype=np.uint8)
    txt[:, :5] = np.frombuffer(b"chr1\t", dtype=np.uint8)
    txt[:, 5:14] = (pos_t[:, None].astype(np.int64) // 10 ** np.arange(8, -1, -1)[None, :]) % 10 + 48
    ch = np.frombuffer(b"ACGTN", dtype=np.uint8)[np.where(gt < 0, 4, gt)]
    v = txt[:, 14:14 + nS * 4].reshape(St, nS, 4)
    v[:, :, 0] = 9
    v[:, :, 1] = ch[:, 0::2]
    v[:, :, 2] = ord("|")
    v[:, :, 3] = ch[:, 1::2]
    txt[:, -1] = 10
    text = ("#CHROM\tPOS\t" + "\t".join(spec_t.sample_names()) + "\n").encode() + txt.tobytes()
    geno_io.ingest_geno(eng, text, geno_format="phased")
    print("ingest", eng.last_timings())
