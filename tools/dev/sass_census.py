"""SASS opcode census of libflowtrack_b200.so (runs without a GPU): per kernel the counts of the Blackwell-native instructions
that prove the path (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA tensor loads / stores = UTMALDG / UTMASTG, bulk copies =
UBLKCP, tcgen05.cp = UTCCP, mbarrier = SYNCS, cluster barriers = UCGABAR, programmatic dependent launch = ACQBULK / PREEXIT-class).
Usage: python tools/dev/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "flowtrack", "pytorch_b200", "libflowtrack_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCCP", "UTCBAR", "SYNCS", "UCGABAR", "ELECT", "ACQBULK", "PREEXIT", "LDGDEPBAR", "HMMA", "FFMA", "F2FP", "STG", "LDG", "STS", "LDS", "ATOMG", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + ".") or op.startswith(o + "_") or (o == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    kernels[cur][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print("# SASS census of %s (cuobjdump -sass, sm_100a); columns: instruction count per kernel" % os.path.relpath(LIB, ROOT))
    print("%-100s %7s " % ("kernel", "instr") + " ".join("%7s" % o[:7] for o in OPS[:12]))
    for (k, c), name in zip(kernels.items(), demangle):
        name = re.sub(r"\(ft::\w+\)|\(.*\)$", "", name.replace("void ", "").replace("ft::", "").replace("(int)", "").replace("(bool)", ""))
        print("%-100s %7d " % (name[:100], c["_total"]) + " ".join("%7d" % c[o] for o in OPS[:12]))
        total.update(c)
    print("%-100s %7d " % ("TOTAL (%d kernels)" % len(kernels), total["_total"]) + " ".join("%7d" % total[o] for o in OPS[:12]))
    print("other opcodes, whole library: " + ", ".join("%s %d" % (o, total[o]) for o in OPS[12:]))


if __name__ == "__main__":
    sys.exit(main())
