"""One pass of each path on the config-2 shape, for ncu (short)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine
S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
miss = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
what = sys.argv[3] if len(sys.argv) > 3 else "popgen"
with Engine(0) as eng:
    spec = synth.SynthSpec(4, 50, miss=miss, seed=20260925)
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, np.minimum(lo + 5000, S))
    for _ in range(2):
        if what == "popgen": eng.popgen(100, 0.01)
        else: eng.abbababa(0, 1, 2, 3, 0.5)
    print(eng.last_timings())
