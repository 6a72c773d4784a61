"""Per-entry-point CUDA-event timing of the C-ABI calls (bench.py's roofline / breakdown).

start() arms a hook in _lib.call that brackets every launch with two CUDA events recorded on the
launching (current) stream; stop() synchronises and returns {entry point: {"ms", "launches"}}.
Event pairs cost ~2 us of host time each and nothing on the device, so the timed region's
throughput is not disturbed at the batch sizes benchmarked.
"""
import torch

from . import _lib

_records = []


def _hook(name):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _records.append((name, e0, e1))
    return e1


def start():
    _records.clear()
    _lib.prof_hook = _hook


def stop():
    _lib.prof_hook = None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _records:
        d = out.setdefault(name, {"ms": 0.0, "launches": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["launches"] += 1
    _records.clear()
    return out
