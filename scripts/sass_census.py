#!/usr/bin/env python
"""SASS opcode census of libdafk.so per kernel (missing 8): counts of the instructions that prove the
Blackwell data path -- UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG
(TMA tensor load / store), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync), LDSM
(ldmatrix), REDG / ATOMG (global reductions / atomics).

    python scripts/sass_census.py [path/to/libdafk.so] > profiles/r2_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal_segmentation_b200", "libdafk.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDSM",
       "REDG", "ATOMG", "USETMAXREG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in OPS:
            counts[kern][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[kern]["UTCHMMA.2CTA"] += 1
        counts[kern]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass %s | scripts/sass_census.py" % os.path.relpath(LIB, ROOT))
    print("# %d kernels; columns: instruction counts per kernel (static SASS, sm_100a)" % len(counts))
    print("%-90s %7s " % ("kernel", "instrs") + " ".join("%7s" % o[:7] for o in OPS))
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("dafk::", "")
        if not any(c[o] for o in OPS):
            continue
        print("%-90s %7d " % (name[:90], c["_total"]) + " ".join("%7d" % c[o] for o in OPS))
        tot.update(c)
    print("%-90s %7d " % ("TOTAL (kernels listed above)", tot["_total"]) + " ".join("%7d" % tot[o] for o in OPS))


if __name__ == "__main__":
    main()
