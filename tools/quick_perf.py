"""Quick device-side timing probe (not the bench contract): config-2 shape, K1 and K2 paths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
P, spp = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 50)
with Engine(0) as eng:
    for miss in (0.0, 0.02):
        spec = synth.SynthSpec(P, spp, miss=miss, seed=20260925)
        t = time.time(); eng.synth_fill(spec, S); print("synth_fill %.2fs" % (time.time() - t), flush=True)
        eng.set_pops(spec.hap_pop(), P)
        lo = np.arange(0, S, 5000, dtype=np.int64); hi = np.minimum(lo + 5000, S)
        eng.set_windows(lo, hi)
        for rep in range(3):
            t = time.time(); r = eng.popgen(100, 0.01); dt = time.time() - t
            tm = eng.last_timings()
            print("miss=%g popgen wall %.1f ms  paths=%s  %s" % (miss, dt * 1e3, np.bincount(r["path"], minlength=3).tolist(),
                  {k: round(v["ms"], 3) for k, v in tm.items()}), flush=True)
        if "k1_popgen" in tm:
            ms = tm["k1_popgen"]["ms"]
            print("  K1: %.1f Gsites/s, %.0f GB/s algorithmic" % (S / ms / 1e6, S * (spec.n_haps + 4) / ms / 1e6))
        for rep in range(2):
            t = time.time(); a = eng.abbababa(0, 1, 2, P - 1, 0.5); dt = time.time() - t
            tm = eng.last_timings()
            print("miss=%g abba wall %.1f ms %s" % (miss, dt * 1e3, {k: round(v["ms"], 3) for k, v in tm.items()}), flush=True)
        print("  pi[0]=%s D[0]=%s" % (r["pi"][0], a["D"][0]))
