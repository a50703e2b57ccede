"""Per-kernel SASS opcode histogram of libcellseg_b200.so (cuobjdump -sass), the evidence that
the hot kernels are tcgen05 / TMEM / TMA code (B200_PROFILING.md, "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA would be the legacy mma.sync path.

  python profiles/sass_hist.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cellsegmentation_b200", "csrc", "libcellseg_b200.so")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "HMMA",
       "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "ATOMS", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
    names = demangle(list(kernels))
    print("# SASS opcode histogram per kernel (`cuobjdump -sass libcellseg_b200.so`, sm_100a)\n")
    print("| kernel | instr | " + " | ".join(KEY) + " |")
    print("|---|---|" + "---|" * len(KEY))
    for k, c in kernels.items():
        n = re.sub(r"\(anonymous namespace\)::|cs::|void ", "", names[k])
        n = re.sub(r"\(.*", "", n)
        print("| `%s` | %d | %s |" % (n[:70], sum(v for o, v in c.items() if o != "UTCHMMA.2CTA"),
                                      " | ".join(str(c.get(o, 0)) for o in KEY)))


if __name__ == "__main__":
    main()
