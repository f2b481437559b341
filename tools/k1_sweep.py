"""K1 geometry sweep for one shape (env knobs PG_K1_G / PG_K1_WPT / PG_K1_TILE_KB / PG_K1_STAGES)."""
import os, sys, subprocess, json
shape = sys.argv[1:4] if len(sys.argv) > 3 else ["8", "100", "5000000"]
code = r'''
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine
P, spp, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
with Engine(0) as eng:
    spec = synth.SynthSpec(P, spp, miss=0.0, seed=5)
    eng.synth_fill(spec, S); eng.set_pops(spec.hap_pop(), P)
    lo = np.arange(0, S, 5000, dtype=np.int64); eng.set_windows(lo, np.minimum(lo + 5000, S))
    for _ in range(3): eng.popgen(100, 0.01)
    a = eng.last_timings()["k1_popgen"]["ms"]
    for _ in range(3): eng.abbababa(0, 1, 2, P - 1, 0.5)
    b = eng.last_timings()["k1_abba"]["ms"]
    H = spec.n_haps
    print("popgen %.3f ms %.0f GB/s | abba %.3f ms %.0f GB/s" % (a, S*(H+4)/a/1e6, b, S*(H+4)/b/1e6))
'''
for cfg in ["", "PG_K1_G=2", "PG_K1_G=2 PG_K1_WPT=2", "PG_K1_G=2 PG_K1_WPT=1", "PG_K1_TILE_KB=32", "PG_K1_G=4 PG_K1_WPT=4", "PG_K1_WPT=2 PG_K1_TILE_KB=128"]:
    env = dict(os.environ)
    for kv in cfg.split():
        k, v = kv.split("="); env[k] = v
    r = subprocess.run([sys.executable, "-c", code] + shape, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print("%-32s %s" % (cfg or "(default)", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output"), flush=True)
