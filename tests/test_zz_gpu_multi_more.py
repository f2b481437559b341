"""More `--devices 2` command lines on two GPUs (runs last: these cases were added after the round's last 2-GPU session
and have only run on the CPU stand-in of tests/test_mgpu_cpu.py so far): popgenWindows with popFreq and with the pairwise
analyses, fourPopWindows, distMat (windows and cat), sfs — each must equal its single-device output."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from test_gpu_multi import HERE, _two_gpus, _write_inputs

pytestmark = pytest.mark.gpu


def _run(mod, argv, env):
    r = subprocess.run([sys.executable, "-m", "genomics_general_b200.cli." + mod] + argv, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.parametrize("miss", [0.0, 0.03])
def test_more_command_lines_on_two_devices_equal_one_device(tmp_path, miss):
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    path, pops = _write_inputs(tmp_path, miss)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    pp = ["-p", "pop0", "-p", "pop1", "-p", "pop2", "-p", "pop3", "--popsFile", pops]
    p4 = ["--popsFile", pops, "-P1", "pop0", "-P2", "pop1", "-P3", "pop2", "-O", "pop3"]
    runs = [("popgenWindows", ["-w", "7000", "-m", "50", "-f", "phased", "--roundTo", "10", "--analysis", "popFreq", "popDist",
                               "popPairDist"] + pp),
            ("popgenWindows", ["-w", "7000", "-m", "50", "-f", "phased", "--roundTo", "10", "--analysis", "popDist", "popPairDist",
                               "indPairDist", "indHet", "hapStats", "--hapDist", "0.02"] + pp),
            ("fourPopWindows", ["-w", "7000", "-m", "50", "-f", "phased", "--minData", "0.5", "--polarize", "--writeFailedWindows"] + p4),
            ("distMat", ["-w", "7000", "-m", "50", "-f", "phased", "--outFormat", "raw", "--roundTo", "10"]),
            ("distMat", ["--windType", "cat", "-f", "phased", "--outFormat", "phylip", "--roundTo", "10"])]
    for n, (mod, argv) in enumerate(runs):
        outs = []
        for dev in ([], ["--devices", "2"]):
            o = str(tmp_path / ("o_%d_%s_%d.txt" % (n, mod, len(dev))))
            _run(mod, ["-g", path, "-o", o] + argv + dev, env)
            outs.append(open(o).read())
        assert outs[0].count("\n") > 3
        if mod == "fourPopWindows":          # fp64 sums of a window differ in the last bits when a rank tiles its sites differently
            a, b = outs[0].strip().split("\n"), outs[1].strip().split("\n")
            assert len(a) == len(b) and a[0] == b[0]
            for x, y in zip(a[1:], b[1:]):
                x, y = x.split(","), y.split(",")
                assert x[:6] == y[:6]
                assert np.allclose([float(v) for v in x[6:]], [float(v) for v in y[6:]], rtol=0, atol=1.0001e-4, equal_nan=True)
        else:
            assert outs[0] == outs[1], (mod, argv)
    # sfs: one file per spectrum
    for dev in ([], ["--devices", "2"]):
        _run("sfs", ["-i", path, "--inputType", "genotypes", "--polarized", "--doPairs", "--pref", str(tmp_path / ("s%d." % len(dev)))]
             + pp + dev, env)
    one, two = sorted(glob.glob(str(tmp_path / "s0.*"))), sorted(glob.glob(str(tmp_path / "s2.*")))
    assert len(one) == 6 and len(two) == 6
    for a, b in zip(one, two):
        assert open(a).read() == open(b).read()
