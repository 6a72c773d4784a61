"""Opcode histogram per kernel of the built library (cuobjdump -sass): which kernels are tcgen05 / TMEM / TMA code and which still run
on mma.sync or plain FMA.  Usage: python tools/sass_histogram.py [libmsg_b200.so] > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "FFMA", "FFMA2", "DFMA", "DADD",
       "MUFU", "RED", "ATOM", "ATOMG", "BAR", "LDS", "STS", "LDG", "STG", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multi_style_transfer_gan_b200", "libmsg_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
            cur["_total"] += 1
    names = demangle(list(per))
    print("# SASS opcode counts per kernel (`cuobjdump -sass libmsg_b200.so`, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / st (tensor memory), `UTMALDG` / `UTMASTG` / "
          "`UTMAREDG` = TMA tensor load / store / reduce, `UTMAPF` = TMA prefetch, `SYNCS` = mbarrier, `HMMA` = mma.sync, `LDGSTS` = cp.async.\n")
    print("| kernel | instr | " + " | ".join(KEY) + " |")
    print("|---|---:|" + "---:|" * len(KEY))
    rows = []
    for mangled, c in per.items():
        n = names[mangled]
        n = n.replace("void ", "").replace("msg::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("msg::", "")
        n = re.sub(r"\((int|bool|unsigned int)\)", "", n)
        n = re.sub(r"\(.*$", "", n)
        rows.append((n, c))
    for n, c in sorted(rows, key=lambda t: (-t[1]["UTCHMMA"], -t[1]["HMMA"], t[0])):
        print(f"| `{n[:70]}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEY) + " |")


if __name__ == "__main__":
    main()
