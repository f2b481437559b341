"""2-GPU test (skipped on a single-GPU box): window shards on two ranks + the native NCCL all-gather equal the
single-GPU result."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_native_nccl_gather_two_gpus():
    import ctypes as C
    from genomics_general_b200 import _lib
    n = C.c_int(0)
    _lib.lib().pg_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "_nccl_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_GATHER_OK" in r.stdout, r.stdout[-4000:]
