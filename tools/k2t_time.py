"""Kernel times of one 2 %-missing popgen pass on the C2 shape (10 M sites, 2000 windows) — for sweeps of the PG_K2T_* knobs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from genomics_general_b200 import synth  # noqa: E402
from genomics_general_b200.engine import Engine  # noqa: E402

S = int(float(os.environ.get("K2T_SITES", "10000000")))
W = int(os.environ.get("K2T_WIN", "5000"))
with Engine(0) as eng:
    spec = synth.SynthSpec(4, 50, miss=0.02, seed=20260925)
    eng.synth_fill(spec, S)
    eng.set_pops(spec.hap_pop(), 4)
    lo = np.arange(0, S, W, dtype=np.int64)
    eng.set_windows(lo, np.minimum(lo + W, S))
    for _ in range(3):
        r = eng.popgen(100, 0.01)
    k = {a: round(b["ms"], 3) for a, b in eng.last_timings().items()}
    print(json.dumps(dict(env={a[7:]: b for a, b in os.environ.items() if a.startswith("PG_K2T")}, gram_n=k.get("k2t_gram_n"),
                          gram_diff=k.get("k2t_gram_diff"), total=round(sum(k.values()), 3),
                          check=float(np.nansum(r["pi"]) + np.nansum(r["dxy"])))))
