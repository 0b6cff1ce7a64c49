"""Opcode evidence of the in-tree CUDA library: per kernel, how often the Blackwell-specific SASS opcodes occur
(cuobjdump -sass rebel_b200/libcfrb200.so).  UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (the TMA engine without a tensor map), SYNCS = mbarrier ops, FFMA2 / FMUL2 / FADD2 = packed fp32 pairs,
HFMA2 / HMUL2 = packed fp16, MUFU.TANH, DFMA / DADD / DMUL = fp64.   Writes profiles/sass_histogram.txt.

    python scripts/sass_histogram.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rebel_b200", "libcfrb200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HMUL2", "MUFU.TANH", "MUFU",
        "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ACQBULK"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_histogram.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(.*", "", fn)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[fn][k] += 1
                    break
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} : occurrences of selected opcodes per kernel (static counts)\n")
        for fn, c in counts.items():
            if not total[fn]:
                continue
            f.write(f"{fn}  [{total[fn]} instructions]\n    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
