"""GPU probe: popgen through the tensor pairwise path vs the POPC path at growing sizes; lists the windows that differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine

def run(eng, S, env=None):
    spec = synth.SynthSpec(4, 50, miss=0.02, seed=20260925)
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, 5000, dtype=np.int64); hi = np.minimum(lo + 5000, S)
    eng.set_windows(lo, hi)
    os.environ["PG_K2_POPC"] = "1"
    ref = eng.popgen(100, 0.01)
    os.environ.pop("PG_K2_POPC", None)
    for k, v in (env or {}).items(): os.environ[k] = v
    out = []
    for rep in range(3):
        r = eng.popgen(100, 0.01)
        bad = [w for w in range(len(lo)) if not (np.array_equal(r["pi"][w], ref["pi"][w], equal_nan=True) and np.array_equal(r["dxy"][w], ref["dxy"][w], equal_nan=True))]
        out.append(bad)
    for k in (env or {}): os.environ.pop(k, None)
    print("S=%d env=%s windows=%d bad per rep: %s" % (S, env, len(lo), [(len(b), b[:12]) for b in out]), flush=True)
    if out[0]:
        w = out[0][0]
        print("   window %d pi tensor %s\n             pi popc   %s" % (w, r["pi"][w], ref["pi"][w]))

with Engine(0) as eng:
    for S in (6000, 70000, 300000, 1000000, 3000000):
        run(eng, S)
    run(eng, 1000000, {"PG_K2T_NRAW": "1"})
    run(eng, 1000000, {"PG_K2T_NRAW": "2", "PG_K2T_NSTAGES": "1"})
    run(eng, 3000000, {"PG_K2T_NRAW": "1", "PG_K2T_NSTAGES": "1"})
