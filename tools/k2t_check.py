"""GPU probe: the tensor-core pairwise path (k2t.cu) against numpy and against the POPC kernels, with mismatch
details, then timings of the 2 %-missing C2 pass on both paths.  Development tool, not the bench contract."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from genomics_general_b200 import synth
from genomics_general_b200.engine import Engine


def ref_counts(g):
    """g int8 [L, H] -> diff, n int64 [H, H]"""
    v = (g >= 0)
    n = v.T.astype(np.int64) @ v.astype(np.int64)
    same = np.zeros_like(n)
    for a in range(4):
        x = (g == a).astype(np.int64)
        same += x.T @ x
    return n - same, n


def tm(eng):
    return {k: round(v["ms"], 3) for k, v in eng.last_timings().items()}


def check_shape(eng, P, spp, S, miss, wins, label, p_third=0.01):
    spec = synth.SynthSpec(P, spp, miss=miss, seed=1234 + P * 7 + spp, p_third=p_third)
    eng.synth_fill(spec, S)
    g, _ = eng.download(0, S)
    lo = np.array([w[0] for w in wins], dtype=np.int64)
    hi = np.array([w[1] for w in wins], dtype=np.int64)
    eng.set_windows(lo, hi)
    ok = True
    for w in range(len(wins)):
        rd, rn = ref_counts(g[lo[w]:hi[w]])
        os.environ.pop("PG_K2_POPC", None)
        d, n = eng.pair_counts(w)
        bad_n = np.argwhere(n != rn)
        bad_d = np.argwhere(d != rd)
        if len(bad_n) or len(bad_d):
            ok = False
            print("MISMATCH %s window %d [%d,%d): n bad %d, diff bad %d of %d" % (label, w, lo[w], hi[w], len(bad_n), len(bad_d), n.size))
            for name, bad, got, ref in (("n", bad_n, n, rn), ("diff", bad_d, d, rd)):
                for (i, j) in bad[:6]:
                    print("   %s[%d,%d] = %d, expected %d" % (name, i, j, got[i, j], ref[i, j]))
                if len(bad):
                    print("   %s bad rows: %s ... cols: %s ..." % (name, sorted(set(bad[:, 0].tolist()))[:20], sorted(set(bad[:, 1].tolist()))[:20]))
        os.environ["PG_K2_POPC"] = "1"
        d2, n2 = eng.pair_counts(w)
        os.environ.pop("PG_K2_POPC", None)
        if not (np.array_equal(d2, rd) and np.array_equal(n2, rn)):
            print("   (POPC path also differs from numpy!)")
    print("%s: %s" % (label, "ok" if ok else "FAILED"), flush=True)
    return ok


def main():
    allok = True
    with Engine(0) as eng:
        allok &= check_shape(eng, 2, 10, 3000, 0.05, [(0, 3000), (100, 164), (5, 70), (64, 128), (1000, 1001)], "H=40")
        allok &= check_shape(eng, 2, 10, 3000, 0.0, [(0, 3000), (17, 2100)], "H=40 no missing")
        allok &= check_shape(eng, 3, 22, 2500, 0.1, [(0, 2500), (63, 1999)], "H=132", p_third=0.2)
        allok &= check_shape(eng, 4, 50, 6000, 0.02, [(0, 5000), (5000, 6000), (123, 4567)], "H=400")
        allok &= check_shape(eng, 1, 300, 1500, 0.03, [(0, 1500), (200, 900)], "H=600")
        allok &= check_shape(eng, 1, 500, 1200, 0.02, [(0, 1200)], "H=1000")
        print("ALL OK" if allok else "SOME FAILED", flush=True)
        # timings: C2 shape with 2 % missing genotypes, both paths
        S = int(float(os.environ.get("K2T_SITES", "10000000")))
        spec = synth.SynthSpec(4, 50, miss=0.02, seed=20260925)
        eng.synth_fill(spec, S)
        eng.set_pops(spec.hap_pop(), 4)
        lo = np.arange(0, S, 5000, dtype=np.int64)
        hi = np.minimum(lo + 5000, S)
        eng.set_windows(lo, hi)
        res = {}
        for path in ("tensor", "popc"):
            if path == "popc":
                os.environ["PG_K2_POPC"] = "1"
            else:
                os.environ.pop("PG_K2_POPC", None)
            for rep in range(3):
                t = time.perf_counter()
                r = eng.popgen(100, 0.01)
                wall = (time.perf_counter() - t) * 1e3
            res[path] = r
            k = tm(eng)
            print(json.dumps(dict(path=path, S=S, wall_ms=round(wall, 2), kernel_ms=k, total_kernel_ms=round(sum(k.values()), 3),
                                  paths=np.bincount(r["path"], minlength=3).tolist())), flush=True)
        os.environ.pop("PG_K2_POPC", None)
        for key in ("pi", "dxy", "fst"):
            a, b = res["tensor"][key], res["popc"][key]
            print("tensor vs popc %s: identical=%s maxabs=%g maxrel=%g" % (key, np.array_equal(a, b, equal_nan=True), np.nanmax(np.abs(a - b)), np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
