"""K2 pair kernel with and without the carry-save step (PG_K2_CSA) on the C2 shape with 2 % missing genotypes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine
S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
with Engine(0) as eng:
    spec = synth.SynthSpec(4, 50, miss=0.02, seed=20260925)
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64)
    eng.set_windows(lo, np.minimum(lo + 5000, S))
    out = {}
    for csa in ("0", "1", "0", "1"):
        os.environ["PG_K2_CSA"] = csa
        r = eng.popgen(100, 0.01)
        t = eng.last_timings()
        print("PG_K2_CSA=%s" % csa, {k: round(v["ms"], 3) for k, v in t.items()}, flush=True)
        out[csa] = r
    for k in ("pi", "dxy", "fst"):
        assert np.array_equal(out["0"][k], out["1"][k], equal_nan=True)
    print("identical results")
