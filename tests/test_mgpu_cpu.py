"""Host logic of the multi-GPU command lines (genomics_general_b200/mgpu.py) on the CPU: byte ranges cut at line starts,
window ownership / halo, the order of the gathered table, and the file rendezvous between rank processes."""
import os
import subprocess
import sys

import numpy as np

from genomics_general_b200 import mgpu


def test_byte_ranges_cut_at_line_starts(tmp_path):
    rng = np.random.default_rng(1)
    lines = [b"#CHROM\tPOS\ta\tb\n"] + [b"chr1\t%d\t%s\n" % (i, b"A|C\t" * int(rng.integers(1, 9))) for i in range(1000)]
    data = b"".join(lines)
    p = str(tmp_path / "x.geno")
    open(p, "wb").write(data)
    body = len(lines[0])
    for world in (1, 2, 3, 8, 64):
        rs = mgpu.byte_ranges(p, body, world)
        assert rs[0][0] == body and rs[-1][1] == len(data)
        for (a, b), (c, d) in zip(rs[:-1], rs[1:]):
            assert b == c
        for a, b in rs:
            assert a == len(data) or a == body or data[a - 1:a] == b"\n"
        assert b"".join(data[a:b] for a, b in rs) == data[body:]


def test_window_ownership_halo_and_gather_order():
    starts = np.array([0, 100, 100, 250, 400])            # rank 1 has no sites
    lo = np.array([0, 50, 90, 100, 240, 250, 399, 400])
    hi = np.array([50, 90, 130, 240, 260, 399, 400, 400])  # window 2 reaches 30 sites into the next share; the last is empty
    owned = [mgpu.assign_windows(lo, hi, starts, r) for r in range(4)]
    assert [list(o[0]) for o in owned] == [[0, 1, 2], [], [3, 4], [5, 6, 7]]
    assert [o[3] for o in owned] == [30, 0, 10, 0]
    assert list(owned[2][1]) == [0, 140] and list(owned[2][2]) == [140, 160]
    w_max, row_of = mgpu.gathered_order([o[0] for o in owned])
    assert w_max == 3 and [row_of[w] for w in range(8)] == [0, 1, 2, 6, 7, 9, 10, 11]


def test_rendezvous_between_processes(tmp_path):
    d = str(tmp_path / "rdv")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from genomics_general_b200 import mgpu; "
            "r = mgpu.Rendezvous(int(sys.argv[1]), 3, %r, timeout=60); "
            "got = r.allgather('x', np.arange(4) * (r.rank + 1)); "
            "assert [int(g[3]) for g in got] == [3, 6, 9]; r.barrier('b'); print('ok')"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, text=True) for r in range(3)]
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out
