"""Text-ingest throughput against the number of host threads that fill the pinned staging buffers (PG_INGEST_THREADS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import geno_io, synth
from genomics_general_b200.engine import Engine

St, nS = 1_000_000, 200
spec_t = synth.SynthSpec(4, 50, miss=0.0, seed=9)
gt = synth.synth_genotypes(spec_t, 0, St)
pos_t = synth.synth_positions(St, seed=9)
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
path = "/tmp/pg_ingest_threads.geno"
with open(path, "wb") as f:
    f.write(text)
with Engine(0) as eng:
    for nt in (4, 8, 16, 24, 32, 48, 64):
        os.environ["PG_INGEST_THREADS"] = str(nt)
        for src, name in ((text, "memory"), (path, "file")):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                geno_io.ingest_geno(eng, src, geno_format="phased")
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            print("threads %2d %-6s %.1f ms  %.1f GB/s  %s" % (nt, name, best * 1e3, len(text) / best / 1e9,
                  {k: round(x["ms"], 1) for k, x in eng.last_timings().items()}), flush=True)
os.remove(path)
