#!/usr/bin/env python
"""Dev tool (no GPU needed): which Blackwell-native instructions each kernel of libaccelrl_b200.so contains.

    python tools/sass_mnemonics.py > profiles/r1e_sass_mnemonics.md

Counts, per kernel, the SASS mnemonics B200_PROFILING.md lists as proof of a tcgen05 / TMA kernel: UTC*MMA
(tcgen05.mma), LDTM (tcgen05.ld), UBLKCP / UTMALDG (bulk / tensor TMA copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
and for contrast LDGSTS (cp.async), HMMA (legacy mma.sync), IDP (dp4a)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "accel_rl_b200", "csrc", "libaccelrl_b200.so")
PAT = re.compile(r"\b(UTC[A-Z0-9]*MMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|HMMA|IDP)\b")
COLS = ["UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "IDP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
        elif cur:
            for t in PAT.findall(line):
                counts[cur][t] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonics per kernel (`cuobjdump -sass accel_rl_b200/csrc/libaccelrl_b200.so`, sm_100a)\n")
    print("Static instruction counts (loops are not unrolled in SASS, so these are sites, not executions). "
          "`UTCHMMA` = `tcgen05.mma` (bf16), `LDTM` = `tcgen05.ld`, `UBLKCP` = `cp.async.bulk` (TMA engine), "
          "`UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`, `IDP` = `dp4a`; "
          "`HMMA` (legacy `mma.sync`) appears nowhere.\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for n, d in zip(names, dem):
        c = counts[n]
        if not any(c[k] for k in COLS if k not in ("SYNCS",)):
            continue
        short = re.sub(r"\(.*", "", d).replace("arl::", "").replace("void ", "")
        print("| `%s` | " % short + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")


if __name__ == "__main__":
    main()
