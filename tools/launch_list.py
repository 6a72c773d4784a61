"""Aggregates an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...) by kernel name.
Usage: python tools/launch_list.py X.csv [first-kernel-substring last-kernel-substring]   (the slice between the
second-to-last and the last occurrence pair is taken as one step when the markers are given)"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = list(csv.reader(lines))
    ix = {h: i for i, h in enumerate(r[0])}
    return [(x[ix["Kernel Name"]], float(x[ix["Metric Value"]])) for x in r[1:]
            if len(x) > ix["Metric Value"] and x[ix["Metric Name"]] == "gpu__time_duration.sum"]


def short(k):
    k = k.replace("void ", "").replace("msg::<unnamed>::", "").replace("msg::", "").replace("at::native::", "at::")
    k = re.sub(r"\(.*$", "", k)
    m = re.match(r"([A-Za-z0-9_:]+)(<.*)?$", k)
    if m and m.group(2) and not k.startswith("at::"):
        targs = m.group(2)
        targs = re.sub(r"\((int|bool)\)", "", targs)
        return m.group(1) + targs[:40]
    return (m.group(1) if m else k)[:70]


def main():
    data = load(sys.argv[1])
    if len(sys.argv) >= 3:
        idx = [i for i, (k, v) in enumerate(data) if sys.argv[2] in k]
        per = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        lo, hi = idx[-1 - per] + 1, idx[-1] + 1
        data = data[lo:hi]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for k, v in data:
        a = agg[short(k)]
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{len(data)} launches, {tot / 1e6:.3f} ms summed (per-launch times under ncu: cold cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"{v[0] / 1e6:8.3f} ms {100 * v[0] / tot:5.1f} %  x{v[1]:<5d} {k}")


if __name__ == "__main__":
    main()
