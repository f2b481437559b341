"""GPU probe: the complete popgenWindows / freq command lines on the FULL C2 file (10 M sites, 8.1 GB of .geno text,
2 % missing genotypes), single device and --devices N.  Prints one JSON line."""
import json, os, sys, tempfile, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

S = int(float(os.environ.get("FT_SITES", "10000000")))
tdir = tempfile.mkdtemp(prefix="pg_full_", dir=os.environ.get("FT_DIR", "/dev/shm" if os.path.isdir("/dev/shm") else None))
t0 = time.perf_counter()
gpath, ppath = bench.write_workload_files(tdir, S, 0.02, bench.SEED + 9)
res = {"sites": S, "text_bytes": os.path.getsize(gpath), "write_s": round(time.perf_counter() - t0, 1), "dir": tdir}
popargs = []
for k in range(bench.N_POPS):
    popargs += ["-p", "pop%d" % k]
env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for name, mod, argv in (("popgenWindows", "popgenWindows", ["-w", "50000", "-m", "100", "-f", "phased", "--popsFile", ppath] + popargs),
                        ("freq", "freq", ["-f", "phased", "--popsFile", ppath] + popargs)):
    for dev in [[]] + ([["--devices", os.environ["FT_DEVICES"]]] if os.environ.get("FT_DEVICES") else []):
        o = os.path.join(tdir, "out_%s_%d.txt" % (name, len(dev)))
        best = None
        for rep in range(2):
            t1 = time.perf_counter()
            r = subprocess.run([sys.executable, "-m", "genomics_general_b200.cli." + mod, "-g", gpath, "-o", o, "--timing",
                                o + ".json"] + argv + dev, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            dt = time.perf_counter() - t1
            assert r.returncode == 0, r.stderr[-2000:]
            best = dt if best is None else min(best, dt)
        key = name + (" --devices " + dev[1] if dev else "")
        res[key] = {"wall_s": round(best, 3), "sites_per_s": round(S / best), "output_bytes": os.path.getsize(o)}
        if os.path.exists(o + ".json"):
            res[key]["timing"] = json.load(open(o + ".json"))
        os.remove(o)
for p in (gpath, ppath):
    os.remove(p)
print(json.dumps(res))
