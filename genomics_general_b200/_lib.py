"""ctypes binding of libpgwin.so (C-ABI: include/pgwin.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is created, this raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgwin.so")
CSRC = os.path.join(_HERE, "csrc")

_lib = None


class PgError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libpgwin.so (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h", ".cpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "pgwin.h"))
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = ["make", "-C", CSRC, "-j4"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise PgError("building libpgwin.so failed (nvcc for sm_100a)")
    return LIB_PATH


_SIGS = {
    "pg_version": (C.c_int, []),
    "pg_last_error": (C.c_char_p, []),
    "pg_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pg_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pg_ctx_destroy": (C.c_int, [C.c_void_p]),
    "pg_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "pg_host_free": (C.c_int, [C.c_void_p]),
    "pg_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "pg_alloc_sites": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32]),
    "pg_upload_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "pg_append_sites": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pg_synth_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64,
                                C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32]),
    "pg_download": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "pg_set_pops": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pg_set_windows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pg_popgen": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "pg_popgen_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]),
    "pg_popgen_freqstats": (C.c_int, [C.c_void_p] * 6),
    "pg_set_freqstats": (C.c_int, [C.c_void_p, C.c_int32]),
    "pg_abbababa": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "pg_fourpop": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "pg_site_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "pg_site_target_freqs": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    "pg_sfs": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.POINTER(C.c_int64)]),
    "pg_sfs_tables": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "pg_pairdist": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "pg_pairdist_cat": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]),
    "pg_seq_nonnan": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pg_ind_het": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "pg_hapstats": (C.c_int, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]),
    "pg_pair_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pg_last_timings": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "pg_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "pg_debug_k1_plan": (C.c_int, [C.c_int64, C.c_int32] + [C.POINTER(C.c_int32)] * 5),
    "pg_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "pg_nccl_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "pg_nccl_finalize": (C.c_int, [C.c_void_p]),
    "pg_popgen_gather_begin": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int64, C.c_int32]),
    "pg_popgen_gather_end": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pg_popgen_allgather": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_int64, C.c_void_p,
                                      C.POINTER(C.c_int64)]),
    "pg_ingest_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_int64)]),
    "pg_ingest_file_range": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.POINTER(C.c_int64)]),
    "pg_ingest_text": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_int64)]),
    "pg_ingest_meta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pg_ingest_release": (C.c_int, [C.c_void_p]),
    "pg_format_freq_rows": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "pg_format_matrix_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_int32, C.c_void_p]),
    "pg_abbababa_allgather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int64,
                                        C.c_void_p]),
    "pg_fourpop_allgather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                       C.c_int64, C.c_void_p]),
    "pg_geno_count_lines": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int64)]),
    "pg_geno_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load libpgwin.so (once).  Raises PgError when it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PgError("libpgwin.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                      "`make -C genomics_general_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)          # AttributeError here means the header and the library diverged
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().pg_last_error()
        raise PgError("%s failed: %s" % (what or "libpgwin call", msg.decode() if msg else "unknown error"))
