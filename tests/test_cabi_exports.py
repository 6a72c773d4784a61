"""CPU: the C-ABI library builds, loads and exports every symbol include/msg_b200.h declares
(no compute calls without a GPU), and the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from tests.conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "msg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multi_style_transfer_gan_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/msg_b200.h but not exported"
    # every declared function is bound in _lib.SIGNATURES (except msg_last_error)
    for s in syms:
        assert s == "msg_last_error" or s in _lib.SIGNATURES, s
    assert ctypes.CDLL(path).msg_version() >= 100


def test_conv_desc_layout_matches_header():
    from multi_style_transfer_gan_b200._lib import ConvDesc
    assert ctypes.sizeof(ConvDesc) == 26 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from multi_style_transfer_gan_b200 import _lib
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedDiscriminator, EnhancedGenerator
    with pytest.raises(_lib.MsgError):
        EnhancedGenerator(8, 1)(torch.zeros(1, 3, 32, 32))
    with pytest.raises(_lib.MsgError):
        EnhancedDiscriminator(8)(torch.zeros(1, 3, 32, 32))
    # the device check itself refuses to run without an sm_100 GPU
    assert _lib.load().msg_check_device() != 0
    assert b"CUDA" in _lib.load().msg_last_error() or b"sm_" in _lib.load().msg_last_error()


def test_state_dict_contract():
    """convert_model.py:12-29 / pth_info.py:7-14 / direct_transform.py:25-30 contracts."""
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedDiscriminator, EnhancedGenerator
    from oracle import restate as R
    G = EnhancedGenerator(16, 1)
    sd = G.state_dict()
    assert list(sd.keys()) == R.generator_state_dict_keys()
    assert next(iter(sd.keys())) == "initial.0.weight" and sd["initial.0.weight"].shape[0] == 16
    assert sd["up1.0.weight"].shape == (64, 32, 4, 4) and sd["up2.0.weight"].shape == (32, 16, 4, 4)
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert sum(p.numel() for p in G.parameters()) == 168611
    assert sum(p.numel() for p in EnhancedGenerator(64, 3).parameters()) == 2628227
    D = EnhancedDiscriminator(16)
    assert len(D.state_dict()) == 28
    assert sum(p.numel() for p in D.parameters()) == 324722
    # checkpoint wrappers the callers accept
    for wrap in ({"epoch": 1, "G_AB_state_dict": sd}, {"epoch": 1, "model_state_dict": sd}, sd):
        inner = wrap.get("G_AB_state_dict", wrap.get("model_state_dict", wrap))
        EnhancedGenerator(16, 1).load_state_dict(inner, strict=True)
    assert hasattr(G, "gradient_checkpointing_enable")


STRUCTS = {"msg_conv_desc": "ConvDesc", "msg_slab_desc": "SlabDesc", "msg_shift_desc": "ShiftDesc", "msg_msb_ring_desc": "MsbRingDesc",
           "msg_convt_ring_desc": "ConvtRingDesc", "msg_out7_ring_desc": "Out7RingDesc", "msg_down_ring_desc": "DownRingDesc",
           "msg_sn_batch": "SnBatch"}


def test_every_descriptor_struct_matches_its_ctypes_mirror(tmp_path):
    """include/msg_b200.h is plain C: a C compiler's sizeof / offsetof of every descriptor struct and field equals the ctypes
    Structure the Python host binds with (a reordered or missing field would silently shift every argument behind it)."""
    import subprocess
    from multi_style_transfer_gan_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "msg_b200.h")).read()
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "msg_b200.h")}"', "int main(void) {"]
    fields = {}
    for cname, pyname in STRUCTS.items():
        end = hdr.index("} " + cname + ";")
        start = hdr.rfind("typedef struct {", 0, end)
        assert start >= 0, cname
        body = re.sub(r"/\*.*?\*/", "", hdr[start + len("typedef struct {"):end], flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(unsigned\s+int|unsigned\s+char|unsigned|long\s+long|int|float|double|void|u?int\d+_t|size_t|char)\b\s*\**\s*",
                          "", decl, count=1)      # drop the type
            names += [re.sub(r"\[.*?\]", "", n).replace("*", "").strip() for n in decl.split(",")]
        fields[cname] = names
        src.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for n in names:
            src.append(f'  printf("{cname}.{n} %zu\\n", offsetof({cname}, {n}));')
    src += ["  return 0;", "}"]
    c = tmp_path / "abi.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(c)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, pyname in STRUCTS.items():
        cls = getattr(_lib, pyname)
        assert ctypes.sizeof(cls) == int(out[cname]), (cname, ctypes.sizeof(cls), out[cname])
        py_fields = [f[0] for f in cls._fields_]
        assert py_fields == fields[cname], (cname, py_fields, fields[cname])
        for n in py_fields:
            assert getattr(cls, n).offset == int(out[f"{cname}.{n}"]), (cname, n)
