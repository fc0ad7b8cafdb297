"""CPU: the C-ABI library loads and exports every symbol include/j2kgpu.h declares; struct layouts of the
Python mirror match the header; without a device the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "j2kgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(j2kgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(j2k):
    L = j2k.lib()
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), n
    assert sorted(j2k.EXPORTS) == names
    assert L.j2kgpu_abi_version() == 4


def test_struct_layouts(j2k):
    assert C.sizeof(j2k.Image) == 28 and j2k.Image.coef_bits.offset == 24
    assert C.sizeof(j2k.TileComp) == 32 and j2k.TileComp.coeff_off.offset == 24
    assert C.sizeof(j2k.CBlk) == 40 and j2k.CBlk.step.offset == 28 and j2k.CBlk.band.offset == 24 and j2k.CBlk.len_cleanup.offset == 32
    assert C.sizeof(j2k.BlkJob) == 32 and j2k.BlkJob.len_cleanup.offset == 24
    assert C.sizeof(j2k.BatchItem) == 104 and j2k.BatchItem.out_stride.offset == 88 and j2k.BatchItem.flags.offset == 96


def test_strerror(j2k):
    L = j2k.lib()
    assert L.j2kgpu_strerror(0) == b"ok"
    assert b"unsupported" in L.j2kgpu_strerror(j2k.E_UNSUPPORTED)


def test_no_device_fails_loudly(j2k):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(j2k.J2KError) as e:
        j2k.Context(0)
    assert e.value.code == j2k.E_NODEVICE


def test_product_never_references_oracle():
    """the product tree must not import, link or mention the oracle or the input generator"""
    pkg = os.path.join(ROOT, "go-jpeg2000_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".cu", ".h", ".cuh", ".py", ".cpp", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(dp, f)
                assert "datagen" not in txt, os.path.join(dp, f)
