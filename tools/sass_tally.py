"""Blackwell-instruction tally per kernel of libsnerf_b200.so (runs anywhere: `cuobjdump -sass`, no GPU):

    python tools/sass_tally.py [path/to/libsnerf_b200.so]  > profiles/r2_sass_tally.md

Counts, for every kernel in the sm_100a cubin, the SASS mnemonics that show which Blackwell units it drives:
UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st: TMEM <-> registers), UTMALDG (TMA tensor
load), UBLKCP (bulk async copy), SYNCS (mbarrier), plus FFMA / HFMA2 / RED for orientation.  The built library is
git-ignored; this summary is what makes the Blackwell-native claim checkable from the tree."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA", "RED", "ATOM", "MUFU", "BAR"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "snerf_b200", "libsnerf_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur[w] += 1
    names = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic tally per kernel (cuobjdump -sass snerf_b200/libsnerf_b200.so, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, "
          "UBLKCP = bulk async copy, SYNCS = mbarrier ops.\n")
    print("| kernel | instr | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    rows = sorted(zip(names, per.values()), key=lambda r: (-r[1]["UTCHMMA"], -r[1]["total"]))
    tot = collections.Counter()
    for name, c in rows:
        short = re.sub(r"\(.*", "", name).replace("snerf::", "").replace("(anonymous namespace)::", "")
        print(f"| `{short[:90]}` | {c['total']} | " + " | ".join(str(c[w]) for w in WATCH) + " |")
        tot.update(c)
    print(f"| **all {len(rows)} kernels** | {tot['total']} | " + " | ".join(str(tot[w]) for w in WATCH) + " |")


if __name__ == "__main__":
    main()
