#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware paths a kernel uses (B200_PROFILING guide: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP / UTMALDG = bulk-async / TMA copies, HMMA = legacy mma.sync, REDG...F32x2 / F32x4 = 8- and 16-byte vector reductions,
REDUX / SHFL = warp aggregation).  usage: sass_evidence.py [library] > profiles/rNN_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "rnb-neus2_b200", "librnb_b200.so")
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("UBLKCP/UTMA", r"\bUBLKCP|\bUTMALDG|\bUTMASTG"),
                               ("HMMA", r"\bHMMA"), ("REDG", r"\bREDG?\."), ("REDG.x2", r"\bREDG?\.\S*F32x2"), ("REDG.x4", r"\bREDG?\.\S*F32x4"), ("REDUX", r"\bREDUX"), ("SHFL", r"\bSHFL\."), ("LDG", r"\bLDG\."), ("USETMAXREG", r"\bUSETMAXREG"), ("SYNCS", r"\bSYNCS\."), ("instr", r"^\s+/\*[0-9a-f]{4,}\*/")])


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("void ", "").replace("rnb::", "") for o in out]


def main(lib=DEFAULT_LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts = collections.OrderedDict(); cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); continue
        if cur:
            for k, p in PAT.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    names = list(counts); pretty = demangle(names)
    print("# cuobjdump -sass %s: instructions per kernel by mnemonic (static counts; sm_100a)" % os.path.relpath(lib, ROOT))
    print("%-44s %s" % ("kernel", " ".join("%11s" % k for k in PAT)))
    for n, p in sorted(zip(names, pretty), key=lambda x: x[1]):
        print("%-44s %s" % (p[:44], " ".join("%11d" % counts[n][k] for k in PAT)))
    return counts, dict(zip(names, pretty))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB)
