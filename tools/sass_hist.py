#!/usr/bin/env python
"""Instruction histogram of the built library's SASS (cuobjdump -sass libdavi.so): per kernel, the counts of
the mnemonics that prove which hardware path it uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, cp.async -> LDGSTS, legacy mma.sync -> HMMA) plus
registers / spills from the same build (nvcc -Xptxas -v is printed by `python -m deep_attentive_vi_b200.build -v`).

    python tools/sass_hist.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deep_attentive_vi_b200", "libdavi.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "LDGSTS", "HMMA",
       "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "RED", "ATOMG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels; counts of selected mnemonics")
    print(f"# {'kernel':<78} total  " + " ".join(f"{k:>7}" for k in KEY))
    tot = collections.Counter()
    for k, c in kernels.items():
        name = re.sub(r"\((int|bool)\)", "", names[k])
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("davi::", "")[:78]
        print(f"{name:<80} {sum(c.values()):>5}  " + " ".join(f"{c.get(x, 0):>7}" for x in KEY))
        tot.update(c)
    print(f"{'ALL KERNELS':<80} {sum(tot.values()):>5}  " + " ".join(f"{tot.get(x, 0):>7}" for x in KEY))


if __name__ == "__main__":
    sys.exit(main())
