"""Native row formatter (pg_format_freq_rows, host only) against the strings the reference's numpy/Python code produces
(freq.py:100-111: np.column_stack of astype(str) columns joined with tabs)."""
import numpy as np

from genomics_general_b200 import geno_io


def _rows(segs):
    return b"".join(bytes(s) for s in segs).decode().splitlines()


def test_counts_rows():
    rng = np.random.default_rng(1)
    n, P = 1003, 5
    c = rng.integers(0, 400, (n, P, 4)).astype(np.uint16)
    c[5] = 0
    c[7, 2] = 65535
    pos = rng.integers(1, 2 ** 31 - 1, n).astype(np.int32)
    ids = np.sort(rng.integers(0, 3, n)).astype(np.int32)
    names = ["chr1", "scaffold_0000123", "x"]
    for threads in (1, 4, 7):
        got = _rows(geno_io.format_freq_rows(0, c, pos, ids, names, threads=threads))
        want = [names[ids[i]] + "\t" + str(pos[i]) + "\t" + "\t".join(",".join(str(v) for v in c[i, x]) for x in range(P))
                for i in range(n)]
        assert got == want


def test_float_rows_print_like_numpy():
    rng = np.random.default_rng(2)
    n, P = 2000, 4
    v = np.around(rng.random((n, P)), 4)
    v[rng.random((n, P)) < 0.2] = np.nan
    v[0] = [0.0, 1.0, 0.0001, 0.5]
    v[1] = [1e-05, 123456.789, 1e16, 0.1 + 0.2]
    v[2] = [-0.0, 3.0, np.inf, -np.inf]
    v[3] = [2.5e-7, 1 / 3, 12345678901234567.0, 0.9999]
    pos = np.arange(1, n + 1, dtype=np.int32)
    ids = np.zeros(n, dtype=np.int32)
    keep = ~np.all(np.isnan(v), axis=1)
    got = _rows(geno_io.format_freq_rows(1, v, pos, ids, ["c"], keep, threads=3))
    vs = v.astype(str)
    want = ["c\t%d\t%s" % (pos[i], "\t".join(vs[i])) for i in range(n) if keep[i]]
    assert got == want


def test_integer_rows_and_keep_mask():
    v = np.array([[0, 0], [3, 0], [0, 12], [0, 0]], dtype=np.float64)
    keep = ~np.all(v == 0, axis=1)
    got = _rows(geno_io.format_freq_rows(2, v, np.array([5, 6, 7, 8], np.int32), np.array([0, 0, 1, 1], np.int32), ["a", "b"], keep))
    assert got == ["a\t6\t3\t0", "b\t7\t0\t12"]
    assert geno_io.format_freq_rows(0, np.zeros((0, 2, 4), np.uint16), np.zeros(0, np.int32), np.zeros(0, np.int32), ["a"]) == []


def test_distmat_strings_equal_the_numpy_formatting():
    """makeDistMat*String (genomics.py:2288-2306) through the native printer == the reference's own expressions."""
    from genomics_general_b200 import genomics as G
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 60):
        m = rng.random((n, n)) * rng.choice([1.0, 1e-3, 1e-6], (n, n))
        m[rng.random((n, n)) < 0.1] = np.nan
        np.fill_diagonal(m, 0.0)
        names = ["ind_%d" % i for i in range(n)]
        for rt in (3, 7, 10):
            rows = [" ".join(r) for r in m.round(rt).astype(str)]
            assert G.makeDistMatString(m, roundTo=rt) == "\n".join(rows)
            assert G.makeDistMatPhylipString(m, names, roundTo=rt) == "%d\n" % n + "".join(
                "%s  %s\n" % (nm, r) for nm, r in zip(names, rows))
            body = "".join("[%d] '%s'    %s\n" % (i + 1, nm, r) for i, (nm, r) in enumerate(zip(names, rows)))
            assert body in G.makeDistMatNexusString(m, names, roundTo=rt)
