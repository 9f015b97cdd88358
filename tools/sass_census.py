#!/usr/bin/env python
"""SASS census of libstereo_b200.so: per kernel, how often the Blackwell-specific mnemonics occur.

    python tools/sass_census.py > profiles/r02_sass_census.txt

UBLKCP = cp.async.bulk (1-D TMA bulk copy), UTMALDG = tensor-map TMA load, SYNCS = mbarrier ops, FADD2/FMUL2/FFMA2 =
packed fp32x2 arithmetic, USETMAXREG = setmaxnreg, UTC*MMA = tcgen05 (none expected: no step of the path is a
contraction), LDS/STS/LDG/STG totals for orientation.  Works without a GPU (cuobjdump on the built library)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stereo_depth_b200", "csrc", "libstereo_b200.so")
MNEMONICS = ["UBLKCP", "UTMALDG", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FADD", "FFMA", "USETMAXREG", "UTCMMA", "UTCHMMA",
             "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOM", "RED", "MEMBAR", "CCTL", "ACQBULK", "UCGABAR", "total"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            depth, cut = 0, len(name)
            for i, ch in enumerate(name):   # cut the parameter list: the first '(' outside template brackets
                if ch == "<":
                    depth += 1
                elif ch == ">":
                    depth -= 1
                elif ch == "(" and depth == 0:
                    cut = i
                    break
            cur = per.setdefault(name[:cut], collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            cur[op] += 1
    cols = [c for c in MNEMONICS if any(v[c] for v in per.values()) or c in ("UTMALDG", "UTCMMA")]
    print("SASS census of stereo_depth_b200/csrc/libstereo_b200.so (cuobjdump -sass, sm_100a; base mnemonic before the first '.')")
    print(f"{'kernel':78s} " + " ".join(f"{c:>9s}" for c in cols))
    tot = collections.Counter()
    for name, cnt in per.items():
        print(f"{name[:78]:78s} " + " ".join(f"{cnt[c]:9d}" for c in cols))
        tot.update(cnt)
    print(f"{'ALL KERNELS':78s} " + " ".join(f"{tot[c]:9d}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
