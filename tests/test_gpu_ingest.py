"""Device-side .geno tokenizer (pg_ingest_text) against the host tokenizer (pg_geno_parse) — which the CPU tests pin
against the reference's own parser outputs — on every format, on sample subsets / re-ordering, mixed ploidy, comment
and blank lines, CRLF, a missing final newline, several scaffolds, and on malformed input."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from genomics_general_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _both(eng, text, **kw):
    from genomics_general_b200 import geno_io
    host = geno_io.parse_geno(io.BytesIO(text), **kw)
    dev = geno_io.ingest_geno(eng, io.BytesIO(text), **kw)
    assert dev.geno is None
    assert dev.n_sites == host.n_sites and dev.n_haps == host.n_haps
    g, p = eng.download(0, host.n_sites) if host.n_sites else (np.zeros((0, host.n_haps), np.int8), np.zeros(0, np.int32))
    assert np.array_equal(g, host.geno)
    assert np.array_equal(p, host.pos) and np.array_equal(dev.pos, host.pos)
    assert np.array_equal(dev.scaf_ids, host.scaf_ids) and dev.scaf_names == host.scaf_names
    assert dev.names == host.names and np.array_equal(dev.ploidy, host.ploidy)
    return host


def _text(fmt, S, names, ploidy, seed, scaffolds=("chr1",), sep="|", miss=0.1, eol="\n", final_nl=True):
    from genomics_general_b200 import synth
    rng = np.random.default_rng(seed)
    lines = ["#CHROM\tPOS\t" + "\t".join(names)]
    per = max(1, S // len(scaffolds))
    dip = {"AA": "A", "CC": "C", "GG": "G", "TT": "T", "GT": "K", "TG": "K", "AC": "M", "CA": "M", "CG": "S", "GC": "S",
           "AG": "R", "GA": "R", "AT": "W", "TA": "W", "CT": "Y", "TC": "Y"}
    pos = 0
    for s in range(S):
        if s % per == 0:
            pos = 0
        pos += int(rng.integers(1, 2000))
        toks = []
        for k in range(len(names)):
            al = ["ACGTN"[x] if rng.random() > miss else "N" for x in rng.integers(0, 4, ploidy[k])]
            if fmt == "phased":
                toks.append(sep.join(al))
            elif fmt == "pairs":
                toks.append("".join(al))
            elif fmt == "haplo":
                toks.append(al[0])
            else:
                a2 = (al + al)[:2]
                toks.append(dip.get(a2[0] + a2[1], "N"))
        lines.append("%s\t%d\t%s" % (scaffolds[min(s // per, len(scaffolds) - 1)], pos, "\t".join(toks)))
    t = eol.join(lines) + (eol if final_nl else "")
    return t.encode()


def test_phased_diploid_basic_and_subsets(eng):
    names = ["s%02d" % i for i in range(13)]
    text = _text("phased", 3000, names, [2] * 13, 1, scaffolds=("chr1", "chr2", "scaf_with_long_name_0003"))
    _both(eng, text, geno_format="phased")
    _both(eng, text, geno_format="phased", samples=names[::-1])
    _both(eng, text, geno_format="phased", samples=[names[7], names[2], names[11]])
    _both(eng, text, geno_format="phased", samples=[names[12]])


def test_all_formats_and_mixed_ploidy(eng):
    names = ["a", "b", "c", "d", "e"]
    _both(eng, _text("phased", 700, names, [2, 1, 2, 3, 1], 2, sep="/"), geno_format="phased",
          ploidy=dict(zip(names, [2, 1, 2, 3, 1])))
    _both(eng, _text("pairs", 500, names, [2] * 5, 3), geno_format="pairs")
    _both(eng, _text("haplo", 500, names, [1] * 5, 4), geno_format="haplo")
    _both(eng, _text("diplo", 900, names, [2] * 5, 5), geno_format="diplo")
    # --haploid samples in a diplo file: forceHomo keeps homozygous calls only (genomics.py:407)
    _both(eng, _text("diplo", 900, names, [2] * 5, 6), geno_format="diplo", ploidy=dict(zip(names, [2, 1, 2, 1, 2])))


def test_line_layout_edge_cases(eng):
    names = ["x", "y", "z"]
    base = _text("phased", 400, names, [2, 2, 2], 7).decode().split("\n")
    hdr, rows = base[0], [r for r in base[1:] if r]
    # comment lines, blank lines, lines of blanks, CRLF, extra blanks between fields, no final newline
    mixed = [hdr]
    for i, r in enumerate(rows):
        if i % 37 == 0:
            mixed.append("# a comment " + r)
        if i % 41 == 0:
            mixed.append("")
        if i % 43 == 0:
            mixed.append("   \t ")
        if i % 5 == 0:
            r = r.replace("\t", "  \t ", 2)
        if i % 7 == 0:
            r = "  " + r
        mixed.append(r + ("\r" if i % 3 == 0 else ""))
    text = "\n".join(mixed).encode()
    h = _both(eng, text, geno_format="phased")
    assert h.n_sites == len(rows)
    _both(eng, text + b"\n\n\n", geno_format="phased")
    # a file with a header only, and an empty body
    _both(eng, (hdr + "\n").encode(), geno_format="phased")
    # long lines crossing many 128-byte steps, tiny lines
    many = ["n%03d" % i for i in range(300)]
    _both(eng, _text("phased", 64, many, [2] * 300, 8), geno_format="phased")
    _both(eng, _text("haplo", 64, ["q"], [1], 9), geno_format="haplo")
    # the header passed separately (--header)
    body = "\n".join(rows).encode()
    _both(eng, body, geno_format="phased", header=hdr)


def test_statistics_from_text_equal_statistics_from_matrix(eng):
    from genomics_general_b200 import geno_io, synth
    spec = synth.SynthSpec(3, 6, miss=0.03, seed=17)
    S = 8000
    g = synth.synth_genotypes(spec, 0, S)
    pos = synth.synth_positions(S, seed=17)
    lut = np.array(list("ACGTN"))
    ch = lut[np.where(g < 0, 4, g)]
    lines = ["#CHROM\tPOS\t" + "\t".join(spec.sample_names())]
    for s in range(S):
        lines.append("chr1\t%d\t%s" % (pos[s], "\t".join(ch[s, 2 * k] + "|" + ch[s, 2 * k + 1] for k in range(18))))
    text = ("\n".join(lines) + "\n").encode()
    lo = np.arange(0, S, 1000, dtype=np.int64)
    hi = lo + 1000
    gd = geno_io.ingest_geno(eng, text, geno_format="phased")
    eng.set_pops(spec.hap_pop(), 3)
    eng.set_windows(lo, hi)
    a = eng.popgen(100, 0.01)
    eng.upload(g, pos)
    eng.set_pops(spec.hap_pop(), 3)
    eng.set_windows(lo, hi)
    b = eng.popgen(100, 0.01)
    for k in ("pi", "dxy", "fst", "sites", "pos_sum", "path"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_malformed_input_is_reported(eng):
    from genomics_general_b200 import geno_io
    from genomics_general_b200._lib import PgError
    names = ["a", "b", "c"]
    good = _text("phased", 50, names, [2, 2, 2], 11).decode().split("\n")
    bad = list(good)
    bad[20] = bad[20].rsplit("\t", 1)[0] + "\tA"                 # haploid token for a diploid sample
    with pytest.raises(PgError, match="data line 20.*ploidy"):
        geno_io.ingest_geno(eng, "\n".join(bad).encode(), geno_format="phased")
    bad = list(good)
    bad[31] = "\t".join(bad[31].split("\t")[:4])                 # a short line
    with pytest.raises(PgError, match="data line 31"):
        geno_io.ingest_geno(eng, "\n".join(bad).encode(), geno_format="phased")
    bad = list(good)
    f = bad[5].split("\t")
    f[1] = "12x"
    bad[5] = "\t".join(f)
    geno_io.ingest_geno(eng, "\n".join(bad).encode(), geno_format="phased")      # int("12x") fails in the reference; digits are read
    f[1] = "pos"
    bad[5] = "\t".join(f)
    with pytest.raises(PgError, match="data line 5.*position"):
        geno_io.ingest_geno(eng, "\n".join(bad).encode(), geno_format="phased")


def test_file_path_ingest_equals_bytes_ingest(eng, tmp_path):
    """pg_ingest_file: the library reads the file itself (pread into pinned staging), header handled by the caller."""
    from genomics_general_b200 import geno_io
    names = ["s%02d" % i for i in range(9)]
    text = _text("phased", 5000, names, [2] * 9, 21, scaffolds=("chrA", "chrB"))
    path = str(tmp_path / "x.geno")
    with open(path, "wb") as f:
        f.write(text)
    host = geno_io.parse_geno(path, geno_format="phased", samples=names[2:7])
    dev = geno_io.ingest_geno(eng, path, geno_format="phased", samples=names[2:7])
    g, p = eng.download(0, host.n_sites)
    assert np.array_equal(g, host.geno) and np.array_equal(p, host.pos)
    assert np.array_equal(dev.scaf_ids, host.scaf_ids) and dev.scaf_names == host.scaf_names
    # gzip goes through host memory
    import gzip
    with gzip.open(path + ".gz", "wb") as f:
        f.write(text)
    dev2 = geno_io.ingest_geno(eng, path + ".gz", geno_format="phased", samples=names[2:7])
    g2, _ = eng.download(0, host.n_sites)
    assert np.array_equal(g2, host.geno) and dev2.scaf_names == host.scaf_names
