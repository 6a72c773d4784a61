"""bench.py's host-side bookkeeping (no GPU): the byte count of the stand-alone InstanceNorm applies that are left, and the rule that a
measured-traffic figure is only quoted for the sources and the micro-batch it was captured on."""
import json
import os

from tests.conftest import ROOT


def test_standalone_apply_bytes():
    import bench
    c, H, W = 64, 512, 512
    msb3 = 3 * (2 * 128 * 256 * 256 + 256 * 128 * 128)                 # down1, up1 (C = 128 at H/2) and down2 (C = 256 at H/4): 2 reads + 1 write
    assert bench.standalone_apply_elems(3, c, H, W) == msb3
    assert bench.standalone_apply_elems(5, c, H, W) == msb3 + 2 * c * H * W + 3 * c * H * W      # + initial IN (1R + 1W) + up2's block


def test_measured_traffic_is_tied_to_sources_and_micro_batch():
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    v, note = bench.measured_traffic("conv_dram_bytes_per_launch", d.get("micro_batch", 16))
    if d["csrc_sha"] == bench.csrc_digest():
        assert v == d["conv_dram_bytes_per_launch"] and d["ncu_file"] in note
        other, why = bench.measured_traffic("conv_dram_bytes_per_launch", d.get("micro_batch", 16) * 2)
        assert other is None and "micro-batch" in why
    else:
        assert v is None and "stale" in note
