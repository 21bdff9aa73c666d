"""Per-kernel SASS evidence of the Blackwell-native paths: counts of the tensor-core / TMEM / TMA mnemonics in every kernel
of libgtav_b200.so (cuobjdump -sass).  UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = TMA tensor load,
UBLKCP = bulk copy, HMMA = legacy mma.sync.  Writes profiles/<round>/sass_summary.txt.

    python scripts/sass_summary.py [profiles/r02/sass_summary.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ai-generated-gtav_b200", "libgtav_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass_summary.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", ""))[:110]
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("SASS mnemonic counts per kernel of libgtav_b200.so (cuobjdump -sass; sm_100a)\n")
        f.write("UTCHMMA=tcgen05.mma  LDTM=tcgen05.ld  UTMALDG=TMA load  UBLKCP=bulk copy  SYNCS=mbarrier  HMMA=mma.sync\n\n")
        f.write(f"{'kernel':<112}{'instr':>7}" + "".join(f"{m:>9}" for m in MNEMONICS) + "\n")
        for k in order:
            c = counts[k]
            f.write(f"{k:<112}{c['_total']:>7}" + "".join(f"{c[m]:>9}" for m in MNEMONICS) + "\n")
    print(open(out_path).read())


if __name__ == "__main__":
    main()
