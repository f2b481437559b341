"""Whole-file .geno ingest: text -> dense int8 [sites x haplotypes] + positions + scaffold runs.

Replaces, for a whole file at once, the reference's streaming reader and per-window conversion
(genomics.py:1884-1945 parseGenoLine/GenoFileReader, 390-396 splitSeq, 1101-1127 genoToAlignment,
74-77 seqArrayToNumArray).  Tokenising runs in native code (csrc/geno_parse.cpp, multi-threaded);
this module only handles files, the header line and sample selection.
"""
from __future__ import annotations

import ctypes as C
import gzip
import io
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PgError, check

FORMATS = {"phased": 0, "diplo": 1, "pairs": 2, "haplo": 3, "alleles": 0}


@dataclass
class GenoData:
    geno: np.ndarray          # int8 [S, H]  A0 C1 G2 T3, -1 missing; haplotypes of sample k at hap_off[k]..+ploidy[k]
                              # (None after ingest_geno: the matrix was built on the device and lives there)
    pos: np.ndarray           # int32 [S]
    scaf_ids: np.ndarray      # int32 [S] run index (a scaffold that re-appears later starts a new run)
    scaf_names: list          # name of each run
    names: list               # sample names (columns kept, in the order requested)
    ploidy: np.ndarray        # int8 per sample
    hap_off: np.ndarray       # first haplotype column of each sample
    header: str

    @property
    def n_sites(self):
        return int(self.pos.shape[0])

    @property
    def n_haps(self):
        return int(self.ploidy.astype(np.int64).sum())

    def hap_sample(self):
        """sample index of every haplotype column"""
        return np.repeat(np.arange(len(self.names), dtype=np.int32), self.ploidy.astype(np.int64))


def _lib_parse():
    return _lib.lib()


def read_bytes(source):
    """source: path (optionally .gz), bytes, or a text/binary file object."""
    if isinstance(source, (bytes, bytearray)):
        return bytes(source)
    if isinstance(source, str):
        if source.endswith(".gz"):
            with gzip.open(source, "rb") as f:
                return f.read()
        with open(source, "rb") as f:
            return f.read()
    data = source.read()
    return data.encode() if isinstance(data, str) else data


def _select(header, geno_format, samples, ploidy):
    """header line + requested samples -> (file column names, samples, format code, ploidies, file column of each sample)"""
    file_names = header.split()[2:]
    if samples is None:
        samples = list(file_names)
    col = {}
    for i, n in enumerate(file_names):
        col.setdefault(n, i)
    missing = [s for s in samples if s not in col]
    if missing:
        raise KeyError("samples not in the genotype file header: %s" % ", ".join(missing[:5]))
    fmt = FORMATS[geno_format]
    default_pl = 1 if geno_format == "haplo" else 2
    pl = np.array([int((ploidy or {}).get(s, default_pl) or default_pl) for s in samples], dtype=np.int8)
    col_take = np.array([col[s] for s in samples], dtype=np.int32)
    return file_names, list(samples), fmt, pl, col_take


def _prepare(source, geno_format, samples, ploidy, header):
    """shared front end of parse_geno / ingest_geno: bytes, header line, sample -> column selection.
    Returns the whole data and the offset of the first data line (no copy of the body is made)."""
    data = read_bytes(source)
    off = 0
    if header is None:
        nl = data.find(b"\n")
        if nl < 0:
            nl = len(data)
        header = data[:nl].decode()
        off = min(nl + 1, len(data))
    return (data, off, header) + _select(header, geno_format, samples, ploidy)


def _scaffold_runs(newsc, off, name_at):
    S = len(newsc)
    scaf_ids = (np.cumsum(newsc.astype(np.int64)) - 1).astype(np.int32) if S else np.zeros(0, dtype=np.int32)
    scaf_names = [name_at(int(off[s])).split(None, 1)[0].decode() for s in np.flatnonzero(newsc)]
    return scaf_ids, scaf_names


def _column_maps(file_names, samples, pl, col_take):
    """file column -> first output haplotype (-1: not wanted) and ploidy, for the device tokenizer"""
    n_cols = len(file_names)
    col_hap = np.full(max(n_cols, 1), -1, dtype=np.int32)
    col_pl = np.ones(max(n_cols, 1), dtype=np.int8)
    hap_off = np.concatenate([[0], np.cumsum(pl.astype(np.int64))[:-1]]).astype(np.int32) if len(pl) else np.zeros(0, np.int32)
    for k, c in enumerate(col_take):
        if col_hap[c] >= 0:
            raise ValueError("sample %s requested twice" % samples[k])
        col_hap[c] = hap_off[k]
        col_pl[c] = pl[k]
    return col_hap, col_pl, hap_off, int(pl.astype(np.int64).sum())


def ingest_geno(eng, source, geno_format="phased", samples=None, ploidy=None, header=None) -> GenoData:
    """Like parse_geno, but the text is tokenised ON THE DEVICE (pg_ingest_text / pg_ingest_file): the file's bytes are
    copied to the GPU as they are and the resident matrix of `eng` is built there.  The returned GenoData has
    geno = None.  A plain file path is read straight into pinned staging buffers by the library."""
    from_file = isinstance(source, str) and not source.endswith(".gz")
    if from_file:
        boff = 0
        if header is None:
            with open(source, "rb") as f:
                first = f.readline()
            header = first.decode()
            boff = len(first)
        file_names, samples, fmt, pl, col_take = _select(header, geno_format, samples, ploidy)
        data = None
    else:
        data, boff, header, file_names, samples, fmt, pl, col_take = _prepare(source, geno_format, samples, ploidy, header)
    col_hap, col_pl, hap_off, H = _column_maps(file_names, samples, pl, col_take)
    if from_file:
        S = eng.ingest_file(source, boff, fmt, col_hap, col_pl, H)
    else:
        S = eng.ingest_text(data, fmt, col_hap, col_pl, H, offset=boff)
    pos, newsc, off = eng.ingest_meta(S)
    if from_file:
        with open(source, "rb") as f:
            def name_at(o):
                f.seek(boff + o)
                return f.read(256)
            scaf_ids, scaf_names = _scaffold_runs(newsc, off, name_at)
    else:
        scaf_ids, scaf_names = _scaffold_runs(newsc, off, lambda o: data[boff + o:boff + o + 256])
    return GenoData(geno=None, pos=pos, scaf_ids=scaf_ids, scaf_names=scaf_names, names=samples, ploidy=pl,
                    hap_off=hap_off, header=header)


def parse_geno(source, geno_format="phased", samples=None, ploidy=None, header=None, threads=None) -> GenoData:
    """Parse a whole .geno file on the host (native multi-threaded tokenizer).

    samples: sample names to keep (default: every column of the header, genomics.py:1918);
    ploidy: dict sample -> ploidy (default 2; 1 for -f haplo, popgenWindows.py:302);
    header: header text when the file has none (--header)."""
    data, boff, header, file_names, samples, fmt, pl, col_take = _prepare(source, geno_format, samples, ploidy, header)
    body = data[boff:] if boff else data
    H = int(pl.astype(np.int64).sum())
    L = _lib_parse()
    n = C.c_int64(0)
    check(L.pg_geno_count_lines(body, len(body), C.byref(n)), "pg_geno_count_lines")
    S = int(n.value)
    geno = np.empty((S, H), dtype=np.int8)
    pos = np.empty(S, dtype=np.int32)
    newsc = np.empty(S, dtype=np.int8)
    off = np.empty(S, dtype=np.int64)
    if threads is None:
        threads = min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    check(L.pg_geno_parse(body, len(body), fmt, len(samples), col_take.ctypes.data_as(C.c_void_p),
                          pl.ctypes.data_as(C.c_void_p), H, S, geno.ctypes.data_as(C.c_void_p),
                          pos.ctypes.data_as(C.c_void_p), newsc.ctypes.data_as(C.c_void_p),
                          off.ctypes.data_as(C.c_void_p), int(threads)), "pg_geno_parse")
    scaf_ids, scaf_names = _scaffold_runs(newsc, off, lambda o: body[o:o + 256])
    hap_off = np.concatenate([[0], np.cumsum(pl.astype(np.int64))[:-1]]).astype(np.int32) if len(pl) else np.zeros(0, np.int32)
    return GenoData(geno=geno, pos=pos, scaf_ids=scaf_ids, scaf_names=scaf_names, names=list(samples), ploidy=pl,
                    hap_off=hap_off, header=header)


def format_freq_rows(mode, data, pos, scaf_ids, scaf_names, keep=None, threads=None):
    """freq.py's output rows for a block of sites, formatted by native host threads (pg_format_freq_rows).
    mode 0: data uint16 [n,P,4] counts; 1: float64 [n,P] (numpy's float -> str); 2: float64 [n,P] as integers.
    Returns a list of memoryviews to write in order."""
    n, P = int(data.shape[0]), int(data.shape[1])
    if n == 0:
        return []
    if threads is None:
        threads = min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    threads = max(1, min(int(threads), n))
    data = np.ascontiguousarray(data, dtype=np.uint16 if mode == 0 else np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    scaf_ids = np.ascontiguousarray(scaf_ids, dtype=np.int32)
    names = (C.c_char_p * len(scaf_names))(*[s.encode() for s in scaf_names])
    longest = max((len(s) for s in scaf_names), default=0)
    per_row = longest + 16 + P * 48
    seg_cap = ((n + threads - 1) // threads + 1) * per_row
    buf = np.empty(threads * seg_cap, dtype=np.uint8)
    seg_len = np.zeros(threads, dtype=np.uint64)
    kp = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    check(_lib.lib().pg_format_freq_rows(int(mode), data.ctypes.data_as(C.c_void_p), n, P, pos.ctypes.data_as(C.c_void_p),
                                         scaf_ids.ctypes.data_as(C.c_void_p), C.cast(names, C.c_void_p),
                                         None if kp is None else kp.ctypes.data_as(C.c_void_p),
                                         buf.ctypes.data_as(C.c_void_p), seg_cap, threads, seg_len.ctypes.data_as(C.c_void_p)),
          "pg_format_freq_rows")
    mv = memoryview(buf)
    return [mv[t * seg_cap: t * seg_cap + int(seg_len[t])] for t in range(threads)]


def format_matrix_rows(m, sep=" ", prefixes=None, threads=None) -> str:
    """Rows of a float64 matrix as newline-terminated text, numbers printed like numpy's float64 -> str
    (pg_format_matrix_rows); `prefixes` = one string per row or None."""
    m = np.ascontiguousarray(m, dtype=np.float64)
    rows, cols = m.shape
    if rows == 0:
        return ""
    if threads is None:
        threads = min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    threads = max(1, min(int(threads), rows, max(1, rows * cols // 4096)))
    pre = None
    longest = 0
    if prefixes is not None:
        enc = [p.encode() for p in prefixes]
        longest = max(len(e) for e in enc)
        pre = (C.c_char_p * rows)(*enc)
    seg_cap = ((rows + threads - 1) // threads + 1) * (longest + cols * 42 + 2)
    buf = np.empty(threads * seg_cap, dtype=np.uint8)
    seg_len = np.zeros(threads, dtype=np.uint64)
    check(_lib.lib().pg_format_matrix_rows(m.ctypes.data_as(C.c_void_p), rows, cols, ord(sep),
                                           None if pre is None else C.cast(pre, C.c_void_p), buf.ctypes.data_as(C.c_void_p),
                                           seg_cap, threads, seg_len.ctypes.data_as(C.c_void_p)), "pg_format_matrix_rows")
    return b"".join(buf[t * seg_cap: t * seg_cap + int(seg_len[t])].tobytes() for t in range(threads)).decode()


# ------------------------------------------------------------------------------------------------
# .gbin — binary cache of an ingested file: the reference re-parses the text on every run
# (genomics.py:1884-1904); a second run on the same file loads the dense matrix instead.
# ------------------------------------------------------------------------------------------------
GBIN_MAGIC = b"PGWINGB1"


def _gbin_key(source, geno_format, names, ploidy):
    st = os.stat(source)
    return dict(source=os.path.abspath(source), size=int(st.st_size), mtime_ns=int(st.st_mtime_ns), geno_format=geno_format,
                names=list(names), ploidy=[int(p) for p in ploidy])


def save_gbin(path, gd: GenoData, source, geno_format, eng=None):
    """Write `<path>`: magic | uint64 header length | JSON header | pos int32 [S] | scaf_ids int32 [S] | geno int8 [S x H]
    (reference codes A0 C1 G2 T3, -1 missing).  With gd.geno None the matrix is read back from the engine slab by slab."""
    import json
    S, H = gd.n_sites, gd.n_haps
    head = dict(_gbin_key(source, geno_format, gd.names, gd.ploidy), n_sites=S, n_haps=H, scaf_names=list(gd.scaf_names),
                header=gd.header)
    blob = json.dumps(head).encode()
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(GBIN_MAGIC)
        f.write(np.uint64(len(blob)).tobytes())
        f.write(blob)
        f.write(np.ascontiguousarray(gd.pos, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(gd.scaf_ids, dtype=np.int32).tobytes())
        if gd.geno is not None:
            f.write(np.ascontiguousarray(gd.geno, dtype=np.int8).tobytes())
        else:
            slab = 1 << 20
            for s in range(0, S, slab):
                g, _ = eng.download(s, min(slab, S - s), want_pos=False)
                f.write(g.tobytes())
    os.replace(tmp, path)


def load_gbin(path, eng, source, geno_format, samples=None, ploidy=None, header=None):
    """-> GenoData with the matrix uploaded to `eng` (geno = None), or None when the cache does not match the request
    (other file / size / mtime / format / samples / ploidy)."""
    import json
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        if f.read(8) != GBIN_MAGIC:
            return None
        n = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
        head = json.loads(f.read(n).decode())
        data_off = 16 + n
    if header is None:
        header = head["header"]
    file_names, want, fmt, pl, col_take = _select(header, geno_format, samples, ploidy)
    try:
        key = _gbin_key(source, geno_format, want, pl)
    except OSError:
        return None
    if any(head.get(k) != v for k, v in key.items()):
        return None
    S, H = int(head["n_sites"]), int(head["n_haps"])
    pos = np.fromfile(path, dtype=np.int32, count=S, offset=data_off)
    scaf_ids = np.fromfile(path, dtype=np.int32, count=S, offset=data_off + 4 * S)
    geno = np.memmap(path, dtype=np.int8, mode="r", offset=data_off + 8 * S, shape=(S, H)) if S else np.zeros((0, H), np.int8)
    eng.upload(geno, pos)                      # slabs through the pinned staging buffers of pg_upload
    del geno
    pl = np.asarray(pl, dtype=np.int8)
    hap_off = np.concatenate([[0], np.cumsum(pl.astype(np.int64))[:-1]]).astype(np.int32) if len(pl) else np.zeros(0, np.int32)
    return GenoData(geno=None, pos=pos, scaf_ids=scaf_ids, scaf_names=list(head["scaf_names"]), names=list(want), ploidy=pl,
                    hap_off=hap_off, header=header)
