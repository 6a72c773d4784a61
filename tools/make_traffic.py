"""profiles/traffic.json from an ncu launch list of the stylise step that also carries DRAM bytes:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/stylise_launches.csv python bench.py --batch 32 --micro-batch 32 --steps 1 --warmup 1 --no-extras --no-cpu-baseline
    python tools/make_traffic.py gpurun_out/stylise_launches.csv profiles/r2_stylise_launch_list.md 32

The LAST forward (from the last nchw_to_nhwc launch to the last blend) is taken as the sample; the tensor-core family is every
conv_tma / conv_slab / conv_shift / conv_tc / msb_ring / convt_ring / la_stage / local_attn_fwd_tc launch in it (bench.py's roofline family)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FAMILY = re.compile(r"conv_tma_kernel|conv_slab_kernel|conv_shift_kernel|conv_tc_kernel|msb64_ring_kernel|msb_ring_kernel|convt_ring_kernel|out7_ring_kernel|down_ring_kernel|la_stage_kernel|local_attn_fwd_tc_kernel")


def main():
    src, out_md = sys.argv[1], sys.argv[2]
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else 32       # images per forward of the captured run (--batch / --micro-batch)
    lines = [l for l in open(src) if not l.startswith("==")]
    r = list(csv.reader(lines))
    ix = {h: i for i, h in enumerate(r[0])}
    launches = collections.OrderedDict()
    for x in r[1:]:
        if len(x) <= ix["Metric Value"]:
            continue
        rec = launches.setdefault(x[ix["ID"]], {"name": x[ix["Kernel Name"]]})
        v, unit = float(x[ix["Metric Value"]].replace(",", "")), x[ix["Metric Unit"]]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        rec[x[ix["Metric Name"]]] = v * scale
    L = list(launches.values())
    starts = [i for i, l in enumerate(L) if "nchw_to_nhwc" in l["name"]]
    ends = [i for i, l in enumerate(L) if "blend" in l["name"]]
    lo = starts[-1]
    hi = max(e for e in ends if e > lo) if any(e > lo for e in ends) else len(L) - 1
    fwd = L[lo:hi + 1]
    fam = [l for l in fwd if FAMILY.search(l["name"])]
    tot_ns = sum(l["gpu__time_duration.sum"] for l in fwd)
    fam_ns = sum(l["gpu__time_duration.sum"] for l in fam)
    fam_bytes = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in fam)
    import bench
    tj = {"csrc_sha": bench.csrc_digest(), "ncu_file": os.path.relpath(out_md, ROOT),
          "micro_batch": mb, "conv_dram_bytes_per_launch": fam_bytes / len(fam), "family_launches_per_forward": len(fam),
          "family_share_of_forward_time": fam_ns / tot_ns}
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    agg = collections.OrderedDict()
    for l in fwd:
        k = re.sub(r"\(.*$", "", l["name"].replace("void ", "").replace("msg::<unnamed>::", "").replace("msg::", ""))
        k = re.sub(r"\((int|bool)\)", "", k)[:60]
        a = agg.setdefault(k, [0.0, 0, 0.0])
        a[0] += l["gpu__time_duration.sum"]
        a[1] += 1
        a[2] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list of one generator forward ({mb} images of the 512x512 step, bf16) + blend\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` on\n"
                f"`python bench.py --batch {mb} --micro-batch {mb} --steps 1 --warmup 1 --no-extras --no-cpu-baseline`; last forward of the run.\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"{len(fwd)} launches, {tot_ns / 1e6:.3f} ms summed; tensor-core family: {len(fam)} launches, {fam_ns / 1e6:.3f} ms "
                f"({100 * fam_ns / tot_ns:.1f} %), DRAM {fam_bytes / 1e9:.3f} GB = {fam_bytes / len(fam) / 1e6:.1f} MB per launch "
                f"(csrc {tj['csrc_sha']})\n\n| kernel | launches | ms | share | DRAM GB |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| `{k}` | {a[1]} | {a[0] / 1e6:.3f} | {100 * a[0] / tot_ns:.1f} % | {a[2] / 1e9:.3f} |\n")
    print(json.dumps(tj))


if __name__ == "__main__":
    main()
