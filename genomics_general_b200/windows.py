"""Window boundaries over dense site arrays, with the reference generators' exact semantics.

The reference builds windows by streaming text lines through Python state machines
(genomics.py:1971-2027 slidingCoordWindows, 2032-2108 slidingSitesWindows, 2112-2171
predefinedCoordWindows).  Here the whole file is a pair of arrays (scaffold id per site,
position per site), every window is a half-open site-index range [lo, hi), and the ranges are
computed with searchsorted per scaffold run — O(W log S) instead of O(S) interpreter steps.
The ranges are what ``pg_set_windows`` (include/pgwin.h) consumes.

Checked against the reference generators in tests/test_windows.py (golden fixtures) and against
the oracle's literal restatement on random inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class WindowSet:
    scaffold: list = field(default_factory=list)      # scaffold name per window
    start: list = field(default_factory=list)         # limits[0]  (None for sites windows)
    end: list = field(default_factory=list)           # limits[1]
    lo: list = field(default_factory=list)            # first site index
    hi: list = field(default_factory=list)            # one past the last site index
    ID: list = field(default_factory=list)

    def add(self, scaffold, start, end, lo, hi, ID=None):
        self.scaffold.append(scaffold)
        self.start.append(start)
        self.end.append(end)
        self.lo.append(int(lo))
        self.hi.append(int(hi))
        self.ID.append(ID)

    def __len__(self):
        return len(self.lo)

    def ranges(self):
        return np.asarray(self.lo, dtype=np.int64), np.asarray(self.hi, dtype=np.int64)


def scaffold_runs(scaf_ids):
    """Maximal runs of equal scaffold id -> (ids, run_lo, run_hi)."""
    scaf_ids = np.asarray(scaf_ids)
    S = len(scaf_ids)
    if S == 0:
        return np.zeros(0, dtype=scaf_ids.dtype), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    change = np.flatnonzero(scaf_ids[1:] != scaf_ids[:-1]) + 1
    lo = np.concatenate([[0], change]).astype(np.int64)
    hi = np.concatenate([change, [S]]).astype(np.int64)
    return scaf_ids[lo], lo, hi


def _wanted(name, include, exclude):
    return (not include and not exclude) or (bool(include) and name in include) or (bool(exclude) and name not in exclude)


def _dup_after_skip(ws, k):
    """Reference quirk (genomics.py:2001-2027 / 2062-2104): after skipping excluded scaffolds the generator
    re-enters its loop with the previous window object untouched and yields it a second time (unless the
    file ended).  Drop-in means the duplicate row is reproduced."""
    ws.add(ws.scaffold[k], ws.start[k], ws.end[k], ws.lo[k], ws.hi[k], ws.ID[k])


def sliding_coord_windows(scaf_ids, scaf_names, pos, wind_size, step_size=None, include=None, exclude=None):
    """genomics.py:1971-2027.  Per scaffold run: windows [1+k*step, k*step+windSize], k = 0..k_last where
    k_last is the first window whose upper limit reaches the run's last position; empty windows are
    emitted; a new run restarts at [1, windSize]."""
    if not step_size:
        step_size = wind_size
    pos = np.asarray(pos, dtype=np.int64)
    ws = WindowSet()
    ids, rlo, rhi = scaffold_runs(scaf_ids)
    wid = 0
    pending_dup = None          # reference quirk: see _dup_after_skip
    for sid, a, b in zip(ids, rlo, rhi):
        name = scaf_names[int(sid)]
        if not _wanted(name, include, exclude):
            if len(ws) and pending_dup is None:
                pending_dup = len(ws) - 1
            continue
        if pending_dup is not None:
            _dup_after_skip(ws, pending_dup)
            wid += 1
            pending_dup = None
        p = pos[a:b]
        last = int(p[-1])
        k_last = max(0, -(-(last - wind_size) // step_size))
        k = np.arange(k_last + 1, dtype=np.int64)
        starts = 1 + k * step_size
        ends = k * step_size + wind_size
        los = a + np.searchsorted(p, starts, side="left")
        his = a + np.searchsorted(p, ends, side="right")
        his = np.maximum(his, los)
        for i in range(len(k)):
            wid += 1
            ws.add(name, int(starts[i]), int(ends[i]), los[i], his[i], wid)
    return ws


def sliding_sites_windows(scaf_ids, scaf_names, pos, wind_sites, overlap=0, max_dist=None, min_sites=None,
                          include=None, exclude=None):
    """genomics.py:2032-2108.  Windows of wind_sites consecutive sites inside a scaffold run; the next
    window keeps the last `overlap` sites; a run's remainder is emitted iff it has >= min_sites sites;
    max_dist bounds position span (sites are then dropped from the left one at a time)."""
    if not min_sites:
        min_sites = wind_sites
    if not overlap:
        overlap = 0
    if max_dist is None:
        max_dist = math.inf
    assert overlap < wind_sites, "overlap must be smaller than the window"
    pos = np.asarray(pos, dtype=np.int64)
    ws = WindowSet()
    ids, rlo, rhi = scaffold_runs(scaf_ids)
    wid = 0
    pending_dup = None
    run_end_emitted = None      # index of the window emitted when the previous wanted run ended (or None)
    for sid, ra, rb in zip(ids, rlo, rhi):
        name = scaf_names[int(sid)]
        if not _wanted(name, include, exclude):
            if run_end_emitted is not None and pending_dup is None:
                pending_dup = run_end_emitted
            continue
        if pending_dup is not None:
            _dup_after_skip(ws, pending_dup)
            wid += 1                                # windowsDone += 1 for the duplicate too (genomics.py:2062-2066)
            pending_dup = None
        run_end_emitted = None
        n = int(rb - ra)
        p = pos[ra:rb]
        finite = not math.isinf(max_dist)
        a = b = 0                                   # the window object holds sites [a, b) of this run
        while True:
            # fill: add sites while the window is not full and the span limit allows (2052)
            if b < n and b - a < wind_sites:
                if b == a:
                    b += 1                          # the first site is always accepted
                cap = min(a + wind_sites, n)
                if finite:
                    cap = min(cap, int(np.searchsorted(p, p[a] + max_dist, side="right")))
                b = max(b, cap)
            if b - a >= min_sites:
                wid += 1
                ws.add(name, None, None, ra + a, ra + b, wid)
                if b >= n:
                    run_end_emitted = len(ws) - 1
                    break                           # scaffold finished (2077-2088)
                if b - a - overlap == 0:
                    raise RuntimeError("sites windows do not advance: overlap >= sites in the window "
                                       "(the reference loops forever here, genomics.py:1779-1788)")
                a = b - overlap                     # trim(leave=overlap) (2072)
            else:
                if b >= n:
                    break
                a += 1                              # trim(remove=1) (2090-2091)
                if a > b:
                    b = a
    return ws


def predefined_coord_windows(scaf_ids, scaf_names, pos, wind_coords):
    """genomics.py:2112-2171.  wind_coords: list of (scaffold, start, end[, ID]) in file order."""
    pos = np.asarray(pos, dtype=np.int64)
    ids, rlo, rhi = scaffold_runs(scaf_ids)
    run_names = [scaf_names[int(s)] for s in ids]
    S = len(pos)
    all_scafs = [w[0] for w in wind_coords]
    scafs = sorted(set(all_scafs), key=lambda x: all_scafs.index(x))
    sidx = {s: k for k, s in enumerate(scafs)}
    ws = WindowSet()
    i = 0                       # index of the site in hand
    run = 0                     # run containing i
    w_scaf = None
    held_lo = held_hi = 0       # sites currently held by the window object: [held_lo, held_hi)

    def cur_run():
        nonlocal run
        while run < len(ids) and i >= rhi[run]:
            run += 1
        return run

    for w in wind_coords:
        name, start, end = w[0], int(w[1]), int(w[2])
        ID = w[3] if len(w) > 3 else "NA"
        if w_scaf is not None and w_scaf == name:
            # slide(newLimits): drop held sites left of the new start (sites right of `end` stay — as the reference)
            k = held_lo
            while k < held_hi and pos[k] < start:
                k += 1
            held_lo = k
        else:
            w_scaf = name
            held_lo = held_hi = i
        wsi = sidx[name]
        # skip whole runs whose scaffold is not wanted or precedes this window's scaffold
        while i < S:
            r = cur_run()
            rn = run_names[r]
            if rn not in sidx or sidx[rn] < wsi:
                i = int(rhi[r])
            else:
                break
        if i < S:
            r = cur_run()
            if run_names[r] == name:
                # skip sites left of the window, then take sites inside it
                a, b = i, int(rhi[r])
                i = a + int(np.searchsorted(pos[a:b], start, side="left"))
                j = a + int(np.searchsorted(pos[a:b], end, side="right"))
                j = max(j, i)
                if held_hi == held_lo:
                    held_lo = i
                    held_hi = i
                if j > i:
                    if held_hi != i:        # nothing is lost between the held sites and the new ones
                        held_lo = i if held_hi == held_lo else held_lo
                    held_hi = j
                    i = j
        ws.add(name, start, end, held_lo, held_hi, ID)
        if i >= S:
            break
    return ws


def mid_pos(pos_sum: int, n: int):
    """GenoWindow.midPos (genomics.py:1795-1797): int(round(sum/len)) with Python-3 rounding; nan if empty."""
    if n == 0:
        return float("nan")
    return int(round(int(pos_sum) / int(n)))
