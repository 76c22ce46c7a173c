#!/usr/bin/env python
"""Instruction histogram of libasn_b200.so per kernel (cuobjdump -sass): the Blackwell-native mnemonics the profiling
recipe asks for -- UTCHMMA / UTCHMMA.2CTA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UBLKCP (TMA), UTCBAR
(tcgen05.commit), SYNCS (mbarrier) -- and the legacy ones that must NOT appear (HMMA from mma.sync / wmma).
    python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptsegnet_b200", "lib", "libasn_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "BRA.U.ANY", "HMMA",
        "LDGSTS", "ATOMG", "REDG", "RED", "ATOMS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            if op.startswith("UTCHMMA"):
                kernels[cur]["UTCHMMA"] += 1
                if ".2CTA" in op:
                    kernels[cur]["UTCHMMA.2CTA"] += 1
            elif op.startswith("BRA.U.ANY"):
                kernels[cur]["BRA.U.ANY"] += 1
            elif base in KEYS:
                kernels[cur][base] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS instruction histogram of libasn_b200.so (sm_100a)\n")
    print("`cuobjdump -sass adaptsegnet_b200/lib/libasn_b200.so`, counted per kernel by `tools/sass_histogram.py`.  "
          "tcgen05.mma = UTCHMMA (`.2CTA` = cta_group::2), tcgen05.ld = LDTM, TMA load / store = UTMALDG / UTMASTG, "
          "tcgen05.commit = UTCBAR, mbarrier = SYNCS.  No HMMA (mma.sync / wmma) anywhere.\n")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "BRA.U.ANY", "HMMA"]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    tot = Counter()
    for (name, c), dm in zip(kernels.items(), demangle):
        tot.update(c)
        short = re.sub(r"\(.*", "", dm).replace("asn::", "")
        short = re.sub(r"\(int\)", "", short)
        if not any(c[k] for k in cols) and "--all" not in sys.argv:
            continue
        print(f"| `{short[:90]}` | {c['_total']} | " + " | ".join(str(c[k]) for k in cols) + " |")
    print(f"| **whole library ({len(kernels)} kernels)** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in cols) + " |")


if __name__ == "__main__":
    main()
