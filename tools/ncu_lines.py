"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm's line info of the same kernel
(cuobjdump -xelf all obj.o; nvdisasm -g -c obj.cubin) and aggregates executed instructions / stall samples per CUDA source
line and per user-given line range.

    python tools/ncu_lines.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-substring> [lo-hi:name ...]
"""
import collections
import csv
import re
import sys


def sass_lines(path, kern):
    """offset -> (file line of the OUTERMOST frame in the .cu being analysed, opcode)"""
    out, cur, on = {}, None, False
    for l in open(path):
        if l.startswith("//-") and ".text." in l:
            on = kern in l
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            # with inlining nvdisasm prints the innermost location first, then "inlined at" frames: keep the last .cu frame
            frames = re.findall(r'File "([^"]+)", line (\d+)', l)
            cu = [(f, int(n)) for f, n in frames if f.endswith(".cu")]
            cur = cu[-1][1] if cu else None
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    src, sass, kern = sys.argv[1:4]
    ranges = []
    for a in sys.argv[4:]:
        r, name = a.split(":")
        lo, hi = r.split("-")
        ranges.append((int(lo), int(hi), name))
    lines = sass_lines(sass, kern)
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(rows[hdr])}
    base = None
    per_line = collections.defaultdict(lambda: [0, 0])
    per_op = collections.defaultdict(int)
    stall_cols = [h for h in rows[hdr] if h.startswith("stall_") and "Not Issued" not in h]
    per_line_stall = collections.defaultdict(lambda: collections.defaultdict(int))
    for r in rows[hdr + 1:]:
        if not r or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        ln, op = lines.get(a - base, (None, r[1].strip()))
        inst = int(r[ix["Instructions Executed"]] or 0)
        smp = int(r[ix["# Samples"]] or 0)
        per_line[ln][0] += inst
        per_line[ln][1] += smp
        per_op[op.split()[0] if not op.startswith("@") else op.split()[1]] += inst
        for s in stall_cols:
            v = int(r[ix[s]] or 0)
            if v:
                per_line_stall[ln][s] += v
    tot = sum(v[0] for v in per_line.values())
    tots = sum(v[1] for v in per_line.values())
    print(f"instructions executed {tot}, stall samples {tots}")
    if ranges:
        agg = collections.defaultdict(lambda: [0, 0])
        stall = collections.defaultdict(lambda: collections.defaultdict(int))
        for ln, v in per_line.items():
            name = next((n for lo, hi, n in ranges if ln is not None and lo <= ln <= hi), "other")
            agg[name][0] += v[0]
            agg[name][1] += v[1]
            for s, c in per_line_stall[ln].items():
                stall[name][s] += c
        for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            top = sorted(stall[n].items(), key=lambda kv: -kv[1])[:4]
            print(f"  {n:12s} inst {100 * v[0] / tot:5.1f} %   samples {100 * v[1] / max(1, tots):5.1f} %   " +
                  ", ".join(f"{s[6:]} {100 * c / max(1, v[1]):.0f}%" for s, c in top))
    print("top lines by instructions:")
    for ln, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:25]:
        top = sorted(per_line_stall[ln].items(), key=lambda kv: -kv[1])[:3]
        print(f"  line {ln}: inst {100 * v[0] / tot:5.1f} %  samples {100 * v[1] / max(1, tots):5.1f} %  " + ", ".join(f"{s[6:]} {c}" for s, c in top))
    print("top opcodes:")
    for op, c in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {op:28s} {100 * c / tot:5.1f} %")


if __name__ == "__main__":
    main()
