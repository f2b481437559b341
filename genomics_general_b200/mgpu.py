"""Multi-GPU command lines (`--devices N`): one process per GPU, no torch.

The reference parallelises in the script — one producer parses the file and feeds `-T` worker processes through queues,
a sorter restores the window order (popgenWindows.py:398-446, ABBABABAwindows.py:276-353, freq.py:328-360).  Here every rank

  1. tokenises ITS byte range of the .geno file on its GPU (cut at line starts; pg_ingest_file_range),
  2. publishes positions / scaffold runs / line offsets; every rank assembles the same global picture and runs the same
     window generator (windows.py) on it,
  3. owns the windows whose first site lies in its range; the few sites such a window needs from the next rank's range are
     re-read from the file (host tokenizer) and appended (pg_append_sites),
  4. computes its windows and takes part in ONE ncclAllGather of the fixed-width records (pg_*_allgather); rank 0 writes rows.

Ranks find each other through a directory in /dev/shm (arrays are published as .npy files, the 128-byte NCCL id as raw
bytes): `--devices N` makes the command line re-launch itself N-1 times with PG_MG_RANK / PG_MG_WORLD / PG_MG_DIR set; under
torchrun the RANK / WORLD_SIZE / MASTER_PORT variables are used instead.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
import time

import numpy as np

from . import geno_io


class Rendezvous:
    """File-based exchange between the ranks of one box."""

    def __init__(self, rank: int, world: int, directory: str, timeout: float = 900.0):
        self.rank, self.world, self.dir, self.timeout = int(rank), int(world), directory, timeout
        os.makedirs(directory, exist_ok=True)
        self.children = []

    def _path(self, name, r, ext):
        return os.path.join(self.dir, "%s.r%d.%s" % (name, r, ext))

    def _wait(self, path):
        t0 = time.time()
        while not os.path.exists(path):
            for c in self.children:                     # a dead child must not leave the others waiting forever
                if c.poll() not in (None, 0):
                    raise RuntimeError("rank process %d exited with status %s" % (c.pid, c.returncode))
            if time.time() - t0 > self.timeout:
                raise TimeoutError("timed out waiting for %s" % path)
            time.sleep(0.002)

    def put(self, name, arr):
        p = self._path(name, self.rank, "npy")
        with open(p + ".tmp", "wb") as f:
            np.save(f, np.asarray(arr))
        os.rename(p + ".tmp", p)

    def get(self, name, r):
        p = self._path(name, r, "npy")
        self._wait(p)
        return np.load(p, allow_pickle=False)

    def put_bytes(self, name, data: bytes):
        p = self._path(name, self.rank, "bin")
        with open(p + ".tmp", "wb") as f:
            f.write(data)
        os.rename(p + ".tmp", p)

    def get_bytes(self, name, r):
        p = self._path(name, r, "bin")
        self._wait(p)
        with open(p, "rb") as f:
            return f.read()

    def allgather(self, name, arr):
        self.put(name, arr)
        return [self.get(name, r) for r in range(self.world)]

    def barrier(self, name):
        self.put_bytes("bar_" + name, b"1")
        for r in range(self.world):
            self.get_bytes("bar_" + name, r)

    def finish(self):
        """rank 0: wait for the other ranks, remove the directory.  A rank writes a last file once it has passed the
        barrier, i.e. needs nothing more from the directory; rank 0 removes the files only after it has seen all of them —
        under torchrun the other ranks are not its children, and deleting right after its own barrier would leave a rank
        that is still polling for the barrier files waiting forever."""
        self.barrier("done")
        self.put_bytes("bye", b"1")
        rc = 0
        if self.rank == 0:
            for r in range(self.world):
                self.get_bytes("bye", r)
        for c in self.children:
            rc = rc or c.wait()
        if self.rank == 0:
            for f in os.listdir(self.dir):
                try:
                    os.remove(os.path.join(self.dir, f))
                except OSError:
                    pass
            try:
                os.rmdir(self.dir)
            except OSError:
                pass
        if rc:
            raise RuntimeError("a rank process failed (status %d)" % rc)


def init(module: str, argv, devices: int):
    """-> Rendezvous or None (single device).  `module` is the command line's module name (re-launched for ranks 1..N-1)."""
    env = os.environ
    if "PG_MG_RANK" in env:                               # a rank started by the parent command line
        return Rendezvous(int(env["PG_MG_RANK"]), int(env["PG_MG_WORLD"]), env["PG_MG_DIR"])
    if devices is None and "RANK" in env and int(env.get("WORLD_SIZE", "1")) > 1:       # torchrun
        # the ranks of one launch share their parent (the torchrun agent): its pid keeps the files of an earlier launch that
        # died before cleaning up (same port, same run id) out of this one's exchange
        d = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                         "pgwin_%s_%s_%d" % (env.get("MASTER_PORT", "0"), env.get("TORCHELASTIC_RUN_ID", "run"), os.getppid()))
        return Rendezvous(int(env["RANK"]), int(env["WORLD_SIZE"]), d)
    if not devices or devices <= 1:
        return None
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    d = tempfile.mkdtemp(prefix="pgwin_mg_", dir=base)
    rdv = Rendezvous(0, devices, d)
    argv = list(sys.argv[1:] if argv is None else argv)
    for r in range(1, devices):
        cenv = dict(env, PG_MG_RANK=str(r), PG_MG_WORLD=str(devices), PG_MG_DIR=d)
        rdv.children.append(subprocess.Popen([sys.executable, "-m", module] + argv, env=cenv, stdout=subprocess.DEVNULL))
    return rdv


def device_for(rdv, base_device: int = 0) -> int:
    return base_device + rdv.rank


def nccl_connect(eng, rdv):
    """rank 0 creates the NCCL id and publishes it; every rank joins the communicator of its engine"""
    if rdv.rank == 0:
        rdv.put_bytes("ncclid", eng.nccl_unique_id())
    eng.nccl_init(rdv.world, rdv.rank, rdv.get_bytes("ncclid", 0))


# ------------------------------------------------------------------------------------------------
# byte ranges
# ------------------------------------------------------------------------------------------------
def byte_ranges(path: str, body_off: int, world: int):
    """`world` contiguous byte ranges of the data lines, cut at line starts."""
    size = os.path.getsize(path)
    cuts = [body_off]
    with open(path, "rb") as f:
        for r in range(1, world):
            c = body_off + (size - body_off) * r // world
            c = max(c, cuts[-1])
            if c > body_off:                                  # first line start at or after c
                f.seek(c - 1)
                while True:
                    buf = f.read(1 << 16)
                    if not buf:
                        c = size
                        break
                    k = buf.find(b"\n")
                    if k >= 0:
                        c = c - 1 + k + 1
                        break
                    c += len(buf)
            cuts.append(min(c, size))
    cuts.append(size)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class GlobalGeno:
    """What the window generators and the row writers need from the WHOLE file (the matrix itself stays sharded)."""

    def __init__(self, pos, scaf_ids, scaf_names, names, ploidy, header):
        self.pos, self.scaf_ids, self.scaf_names = pos, scaf_ids, scaf_names
        self.names, self.ploidy, self.header = names, ploidy, header
        self.geno = None

    @property
    def n_sites(self):
        return int(self.pos.shape[0])

    @property
    def n_haps(self):
        return int(np.asarray(self.ploidy, dtype=np.int64).sum())


def sharded_ingest(eng, rdv, path, geno_format, samples, ploidy, header=None):
    """Step 1 + 2 of the module docstring.  Returns (GlobalGeno, starts [world + 1] global site index of each rank's first
    site, line_off_all int64 [S] global byte offset of every data line)."""
    assert isinstance(path, str) and not path.endswith(".gz"), "--devices needs an uncompressed .geno file on disk"
    body_off = 0
    if header is None:
        with open(path, "rb") as f:
            first = f.readline()
        header = first.decode()
        body_off = len(first)
    file_names, samples, fmt, pl, col_take = geno_io._select(header, geno_format, samples, ploidy)
    col_hap, col_pl, hap_off, H = geno_io._column_maps(file_names, samples, pl, col_take)
    lo, hi = byte_ranges(path, body_off, rdv.world)[rdv.rank]
    S = eng.ingest_file_range(path, lo, hi, fmt, col_hap, col_pl, H)
    pos, newsc, off = eng.ingest_meta(S)
    # names of this rank's scaffold runs (a run = maximal block of lines with the same first field)
    with open(path, "rb") as f:
        def name_at(o):
            f.seek(lo + o)
            return f.read(256)
        run_names = [name_at(int(off[s])).split(None, 1)[0].decode() for s in np.flatnonzero(newsc)]
    rdv.put("pos", pos)
    rdv.put("newsc", newsc)
    rdv.put("off", off + lo)
    rdv.put("runs", np.array(run_names, dtype=np.str_) if run_names else np.zeros(0, dtype="<U1"))
    parts_pos, parts_new, parts_off, names_all = [], [], [], []
    starts = [0]
    for r in range(rdv.world):
        p, n, o = rdv.get("pos", r), rdv.get("newsc", r).copy(), rdv.get("off", r)
        rn = [str(x) for x in rdv.get("runs", r)]
        if len(p) and names_all and rn and rn[0] == names_all[-1]:
            n[0] = 0                                   # the rank starts inside the previous rank's last scaffold run
            rn = rn[1:]
        names_all += rn
        parts_pos.append(p)
        parts_new.append(n)
        parts_off.append(o)
        starts.append(starts[-1] + len(p))
    pos_all = np.concatenate(parts_pos) if parts_pos else np.zeros(0, np.int32)
    new_all = np.concatenate(parts_new) if parts_new else np.zeros(0, np.int8)
    off_all = np.concatenate(parts_off) if parts_off else np.zeros(0, np.int64)
    scaf_ids = (np.cumsum(new_all.astype(np.int64)) - 1).astype(np.int32) if len(new_all) else np.zeros(0, np.int32)
    gd = GlobalGeno(pos_all, scaf_ids, names_all, samples, pl, header)
    gd.hap_off = hap_off
    return gd, np.array(starts, dtype=np.int64), off_all


def local_ingest(eng, rdv, path, geno_format, samples, ploidy, header=None):
    """This rank's byte range only (per-site outputs such as freq.py need no global picture): a GenoData of the local sites."""
    assert isinstance(path, str) and not path.endswith(".gz"), "--devices needs an uncompressed .geno file on disk"
    body_off = 0
    if header is None:
        with open(path, "rb") as f:
            first = f.readline()
        header = first.decode()
        body_off = len(first)
    file_names, samples, fmt, pl, col_take = geno_io._select(header, geno_format, samples, ploidy)
    col_hap, col_pl, hap_off, H = geno_io._column_maps(file_names, samples, pl, col_take)
    lo, hi = byte_ranges(path, body_off, rdv.world)[rdv.rank]
    S = eng.ingest_file_range(path, lo, hi, fmt, col_hap, col_pl, H)
    pos, newsc, off = eng.ingest_meta(S)
    with open(path, "rb") as f:
        def name_at(o):
            f.seek(lo + o)
            return f.read(256)
        scaf_ids, scaf_names = geno_io._scaffold_runs(newsc, off, name_at)
    return geno_io.GenoData(geno=None, pos=pos, scaf_ids=scaf_ids, scaf_names=scaf_names, names=samples, ploidy=pl,
                            hap_off=hap_off, header=header)


def assign_windows(lo, hi, starts, rank):
    """Windows owned by `rank`: those whose first site lies in its range (an empty window at the very end goes to the last
    rank).  Returns (window indices, local lo, local hi, halo = sites needed past the rank's own last site)."""
    lo = np.asarray(lo, dtype=np.int64)
    hi = np.asarray(hi, dtype=np.int64)
    world = len(starts) - 1
    owner = np.clip(np.searchsorted(starts, lo, side="right") - 1, 0, world - 1)
    # ranks without sites own nothing
    sizes = np.diff(starts)
    for r in range(world):
        if sizes[r] == 0:
            nxt = [q for q in range(r + 1, world) if sizes[q] > 0]
            owner[owner == r] = nxt[0] if nxt else max([q for q in range(world) if sizes[q] > 0] or [0])
    idx = np.flatnonzero(owner == rank)
    a, b = int(starts[rank]), int(starts[rank + 1])
    need = int(hi[idx].max()) if len(idx) else b
    halo = max(0, need - b)
    return idx, lo[idx] - a, hi[idx] - a, halo


def fetch_halo(eng, path, gd, starts, off_all, rank, halo, geno_format, ploidy_dict):
    """Append the `halo` sites that follow this rank's range (read from the file, host tokenizer)."""
    if halo <= 0:
        return
    b = int(starts[rank + 1])
    S = gd.n_sites
    byte_lo = int(off_all[b])
    byte_hi = int(off_all[b + halo]) if b + halo < S else os.path.getsize(path)
    with open(path, "rb") as f:
        f.seek(byte_lo)
        data = f.read(byte_hi - byte_lo)
    part = geno_io.parse_geno(data, geno_format=geno_format, samples=gd.names, ploidy=ploidy_dict, header=gd.header)
    assert part.n_sites == halo, (part.n_sites, halo)
    eng.append_sites(part.geno, part.pos)


def gathered_order(all_idx):
    """all_idx[r] = window indices of rank r (rank order in the gathered table) -> (table row of each window)"""
    w_max = max(max((len(i) for i in all_idx), default=0), 1)
    rows = {}
    for r, idx in enumerate(all_idx):
        for k, w in enumerate(idx):
            rows[int(w)] = r * w_max + k
    return w_max, rows
