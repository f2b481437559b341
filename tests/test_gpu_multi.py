"""2-GPU test (skipped on a single-GPU box): window shards on two ranks + the native NCCL all-gather equal the
single-GPU result."""
import os
import subprocess
import sys

import numpy as np
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


def _two_gpus():
    import ctypes as C
    from genomics_general_b200 import _lib
    n = C.c_int(0)
    _lib.lib().pg_device_count(C.byref(n))
    return n.value >= 2


def _write_inputs(tmp_path, miss, scaffolds=3):
    import numpy as np
    from genomics_general_b200 import synth
    spec = synth.SynthSpec(4, 6, miss=miss, seed=11)
    S = 24000
    g = synth.synth_genotypes(spec, 0, S)
    per = S // scaffolds
    scafs, pos = [], []
    for k in range(scaffolds):
        n = per if k < scaffolds - 1 else S - per * (scaffolds - 1)
        scafs += ["chr%d" % (k + 1)] * n
        pos.append(synth.synth_positions(n, seed=5 + k))
    path = str(tmp_path / "m.geno")
    synth.write_geno(path, g, np.concatenate(pos), scafs, spec.sample_names())
    pops = str(tmp_path / "m.pops")
    with open(pops, "wt") as f:
        for i, n in enumerate(spec.sample_names()):
            f.write("%s pop%d\n" % (n, i // 6))
    return path, pops


@pytest.mark.parametrize("miss", [0.0, 0.03])
def test_command_lines_on_two_devices_equal_one_device(tmp_path, miss):
    """popgenWindows / ABBABABAwindows / freq with --devices 2 (each rank tokenises its byte range of the file, windows that
    straddle the cut fetch their halo sites, one NCCL all-gather, rank 0 writes) == the single-device output, byte for byte"""
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    path, pops = _write_inputs(tmp_path, miss)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    pp = ["-p", "pop0", "-p", "pop1", "-p", "pop2", "-p", "pop3", "--popsFile", pops]
    runs = [("popgenWindows", ["-w", "7000", "-m", "50", "-f", "phased", "--roundTo", "10", "--writeFailedWindows"] + pp),
            ("popgenWindows", ["--windType", "sites", "-w", "900", "-O", "300", "-m", "100", "-f", "phased", "--roundTo", "10"] + pp),
            ("ABBABABAwindows", ["-w", "7000", "-m", "50", "-f", "phased", "--minData", "0.5", "--popsFile", pops, "-P1", "pop0",
                                 "-P2", "pop1", "-P3", "pop2", "-O", "pop3", "--writeFailedWindows"]),
            ("freq", ["-f", "phased"] + pp)]
    for mod, argv in runs:
        outs = []
        for dev in ([], ["--devices", "2"]):
            o = str(tmp_path / ("o_%s_%d.txt" % (mod, len(dev))))
            r = subprocess.run([sys.executable, "-m", "genomics_general_b200.cli." + mod, "-g", path, "-o", o] + argv + dev,
                               env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-3000:]
            outs.append(open(o).read())
        assert outs[0].count("\n") > 3
        if mod == "ABBABABAwindows":        # fp64 sums of a window differ in the last bits when a rank tiles its sites differently
            a, b = outs[0].strip().split("\n"), outs[1].strip().split("\n")
            assert len(a) == len(b) and a[0] == b[0]
            for x, y in zip(a[1:], b[1:]):
                x, y = x.split(","), y.split(",")
                assert x[:6] == y[:6]
                assert np.allclose([float(v) for v in x[6:]], [float(v) for v in y[6:]], rtol=0, atol=1.0001e-4, equal_nan=True)
        else:
            assert outs[0] == outs[1], (mod, argv)
