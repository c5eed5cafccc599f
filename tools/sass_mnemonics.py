#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor load), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), LDGSTS (cp.async), HMMA (legacy mma.sync — expected 0), FFMA.

    python tools/sass_mnemonics.py [path/to/lib.so] > profiles/sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "doc2tex_b200", "libd2t_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, k = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[k]
            k += 1
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("d2t::", "").replace("(anonymous namespace)::", "")
            if cur not in counts:
                counts[cur] = collections.Counter()
                order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in KEYS:
                counts[cur][op] += 1
            counts[cur]["_total"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — SASS mnemonic counts per kernel (static instruction counts)")
    print(f"{'kernel':<64}" + "".join(f"{k_:>9}" for k_ in KEYS) + f"{'instrs':>9}")
    tot = collections.Counter()
    for name in order:
        c = counts[name]
        tot.update(c)
        print(f"{name[:63]:<64}" + "".join(f"{c[k_]:>9}" for k_ in KEYS) + f"{c['_total']:>9}")
    print(f"{'TOTAL':<64}" + "".join(f"{tot[k_]:>9}" for k_ in KEYS) + f"{tot['_total']:>9}")


if __name__ == "__main__":
    main()
