"""CPU test of the N>1 path's host logic: window sharding + the single all-gather of rows, with the gloo
backend at world_size 2 (the GPU path uses the same code over NCCL)."""
import os
import subprocess
import sys

import numpy as np

from genomics_general_b200 import multigpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_windows_balances_sites():
    lo = np.arange(0, 100000, 1000)
    hi = lo + 1000
    for world in (1, 2, 3, 8):
        sh = multigpu.shard_windows(lo, hi, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(lo)
        sizes = [e - b for b, e in sh]
        assert max(sizes) - min(sizes) <= 1
    # more ranks than windows: trailing ranks get empty shards
    sh = multigpu.shard_windows(lo[:3], hi[:3], 8)
    assert sum(e - b for b, e in sh) == 3


def test_world2_gloo_gather():
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-3000:]
