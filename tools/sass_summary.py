"""SASS mnemonic counts per kernel of libpgwin.so (cuobjdump -sass): tcgen05 / TMEM / TMA / mbarrier evidence for profiles/.
Usage: python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTCIMMA[.\w]*|UTCHMMA[.\w]*|LDTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UBLKCP[.\w]*|UTMALDG[.\w]*|SYNCS[.\w]*|"
                 r"FENCE\.VIEW\.ASYNC[.\w]*|IDP\.4A[.\w]*|POPC|PRMT|REDUX[.\w]*|ELECT[.\w]*)\b")
KEEP = ("k2t_gram", "k1_site_pass<0, 4, 12, true", "k1_site_pass<1, 4, 12", "k1_site_pass<4, 4, 12", "k1_site_pass_lp<0, 8",
        "k2t_valid_class<true", "k2t_build_pq", "k_parse_lines")


def main(out=None):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(REPO, "genomics_general_b200", "libpgwin.so")],
                          stdout=subprocess.PIPE, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
        elif cur:
            for x in PAT.findall(line):
                cnt[cur][x] += 1
    lines = ["# SASS mnemonic counts per kernel of libpgwin.so (cuobjdump -sass, sm_100a; tools/sass_summary.py): tcgen05 (UTCIMMA =",
             "# tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc), TMA bulk copies (UBLKCP), mbarriers",
             "# (SYNCS), the async-proxy fence, IDP.4A / POPC / PRMT / REDUX of the site-pass, plane and tokenizer kernels.", ""]
    for k, c in cnt.items():
        name = subprocess.run(["c++filt", k], stdout=subprocess.PIPE, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        if c and any(t in name for t in KEEP):
            lines += ["## " + name, "   " + "  ".join("%s x%d" % ab for ab in sorted(c.items()))]
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
