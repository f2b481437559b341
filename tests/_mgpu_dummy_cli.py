"""TEST INFRASTRUCTURE — a stand-in command line for the `--devices N` launch path of genomics_general_b200/mgpu.py: rank 0
(started by the test) re-launches this module N-1 times through mgpu.init; every rank publishes its rank and argv, rank 0
writes what it gathered to the file named by argv[2]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from genomics_general_b200 import mgpu  # noqa: E402


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    rdv = mgpu.init("_mgpu_dummy_cli", argv, int(argv[0]))
    got = rdv.allgather("who", np.array([rdv.rank, len(argv)]))
    rdv.put_bytes("argv", json.dumps(argv).encode())
    if rdv.rank == 0:
        seen = [json.loads(rdv.get_bytes("argv", r).decode()) for r in range(rdv.world)]
        with open(argv[1], "wt") as f:
            json.dump(dict(ranks=[int(g[0]) for g in got], argv=seen, dir=rdv.dir), f)
    rdv.finish()


if __name__ == "__main__":
    main()
