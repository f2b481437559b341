"""K1 sweep over instantiation knobs (PG_K1_NW, PG_K1_NO_BYTES, PG_K1_G, ...) for several shapes, one process."""
import os, sys, subprocess
code = r'''
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine
P, spp, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfgs = sys.argv[4].split(";")
with Engine(0) as eng:
    spec = synth.SynthSpec(P, spp, miss=0.0, seed=5)
    eng.synth_fill(spec, S)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    H = spec.n_haps
    for cfg in cfgs:
        for k in ("PG_K1_NW", "PG_K1_NO_BYTES", "PG_K1_G", "PG_K1_WPT", "PG_K1_I", "PG_K1_TILE_KB", "PG_K1_STAGES", "PG_K1_LANEPOP"):
            os.environ.pop(k, None)
        for kv in cfg.split():
            k, v = kv.split("="); os.environ[k] = v
        eng.set_pops(spec.hap_pop(), P)          # new epoch -> the launch plan is rebuilt with the knobs
        eng.set_windows(lo, np.minimum(lo + 5000, S))
        out = []
        for fq in (False, True):
            eng.set_freqstats(fq)
            for _ in range(4): eng.popgen(100, 0.01)
            a = eng.last_timings()["k1_popgen"]["ms"]
            out.append("%s %.3f ms %4.0f GB/s" % ("popfreq" if fq else "popgen", a, S*(H+4)/a/1e6))
        eng.set_freqstats(False)
        if P >= 4:
            for _ in range(3): eng.abbababa(0, 1, 2, P - 1, 0.5)
            b = eng.last_timings()["k1_abba"]["ms"]
            out.append("abba %.3f ms %4.0f GB/s" % (b, S*(H+4)/b/1e6))
            for _ in range(3): eng.fourpop(0, 1, 2, P - 1, 0.5)
            b = eng.last_timings()["k1_fourpop"]["ms"]
            out.append("fourpop %.3f ms %4.0f GB/s" % (b, S*(H+4)/b/1e6))
        print("P=%d spp=%d S=%d %-40s %s" % (P, spp, S, cfg or "(default)", " | ".join(out)), flush=True)
'''
shapes = [("8", "100", "8000000"), ("4", "50", "10000000"), ("2", "10", "20000000")]
cfgs = ["", "PG_K1_NW=8", "PG_K1_NW=12", "PG_K1_NO_BYTES=1", "PG_K1_G=1", "PG_K1_G=4", "PG_K1_TILE_KB=32", "PG_K1_TILE_KB=128"]
if len(sys.argv) > 1:
    cfgs = sys.argv[1].split(";")
if len(sys.argv) > 2:
    shapes = [tuple(x.split(",")) for x in sys.argv[2].split(";")]
for sh in shapes:
    r = subprocess.run([sys.executable, "-c", code] + list(sh) + [";".join(cfgs)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout.strip()[-6000:], flush=True)
