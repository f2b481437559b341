"""Per-instruction execution counts / stall samples of one kernel from `ncu -i REP --page source --csv --print-source sass`:
prints the instruction ranges (between branch targets / waits) that execute most."""
import csv
import sys


def kernels(path):
    cur, hdr, out = None, None, {}
    for row in csv.reader(open(path)):
        if row and row[0] == "Kernel Name":
            cur = row[1]
            out[cur] = []
            hdr = None
        elif cur is not None and hdr is None and row and row[0] == "Address":
            hdr = row
        elif cur is not None and hdr is not None and row:
            out[cur].append(dict(zip(hdr, row)))
    return out


def main(path, pick, top=40):
    for name, rows in kernels(path).items():
        if pick not in name:
            continue
        tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
        samp = sum(int(r["# Samples"] or 0) for r in rows)
        print("== %s: %d SASS lines, %d warp instructions, %d samples" % (name, len(rows), tot, samp))
        # regions: split at waits / barriers / branches
        reg, cur = [], []
        for i, r in enumerate(rows):
            cur.append((i, r))
            s = r["Source"]
            if "BRA" in s or "SYNCS" in s or "BAR." in s or "EXIT" in s:
                reg.append(cur)
                cur = []
        if cur:
            reg.append(cur)
        agg = []
        for g in reg:
            ex = sum(int(r["Instructions Executed"] or 0) for _, r in g)
            sm = sum(int(r["# Samples"] or 0) for _, r in g)
            agg.append((ex, sm, g))
        agg.sort(key=lambda t: -t[0])
        for ex, sm, g in agg[:top]:
            i0, r0 = g[0]
            i1, r1 = g[-1]
            print("  lines %5d-%5d  n=%3d  exec=%10d (%4.1f%%)  per-line=%9d  samples=%6d   last: %s" % (
                i0, i1, len(g), ex, 100.0 * ex / max(tot, 1), int(r1["Instructions Executed"] or 0), sm, r1["Source"][:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
