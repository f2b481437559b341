#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE — stages the UNMODIFIED reference scripts for the CPU arm of bench.py.

    python oracle/build_ref.py            (called by __graft_entry__.build() when /root/reference is present)

The reference is pure Python; "building" it means placing its own files, byte for byte, under the git-ignored directory
oracle/_ref/ so that they travel to the GPU box with the snapshot (the box has no /root/reference).  Nothing under
oracle/_ref/ is ever imported by the product; bench.py --impl reference runs `python oracle/_ref/popgenWindows.py ...`
as a subprocess, exactly as a user of the reference would.  A manifest with the sha256 of every staged file is written
next to them so that "unmodified" can be checked.
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["genomics.py", "popgenWindows.py", "ABBABABAwindows.py", "freq.py", "distMat.py"]


def stage(verbose=True):
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        src = os.path.join(REF, f)
        if not os.path.exists(src):
            continue
        shutil.copyfile(src, os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "wt") as m:
        json.dump(dict(source=REF, sha256=manifest), m, indent=1)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
