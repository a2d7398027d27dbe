"""Per-kernel SASS opcode summary of libdiffsplit_b200.so: which kernels carry tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG /
UBLKCP / UTMAPF) and mbarrier (SYNCS) instructions.  Usage: python tools/sass_summary.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffsplitting_b200", "libdiffsplit_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS",
        "ATOMG", "REDG", "RED", "BAR", "ACQBULK", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts.setdefault(kern, collections.Counter())
            total.setdefault(kern, 0)
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            total[kern] += 1
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS") and op.startswith(w)):
                    counts[kern][w] += 1
                    break
    print(f"# SASS opcode counts per kernel, {os.path.basename(LIB)} (cuobjdump -sass, sm_100a)")
    print("# tcgen05: UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld; TMA: UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk;")
    print("# SYNCS = mbarrier ops.  Kernels without any of these are CUDA-core / bandwidth kernels.")
    for k, c in counts.items():
        keys = [w for w in WANT if c.get(w)]
        print(f"{k:60s} {total[k]:6d} instr  " + "  ".join(f"{w}={c[w]}" for w in keys))


if __name__ == "__main__":
    sys.exit(main())
