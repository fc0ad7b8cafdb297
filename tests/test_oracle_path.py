"""CPU: the oracle's whole REF path (orc_decode_image) inverts the reference encoder pipeline restated in
datagen/ -- the transitive pin of SURVEY.md 8c (encoder -> decoder == input) at whole-image level."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from datagen import jobs


def oracle_decode(job, threads=2):
    img = O.Image()
    img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
    for c in range(job["ncomp"]):
        img.prec[c] = job["prec"]
        img.sgnd[c] = job["sgnd"]
    img.mct, img.reversible, img.nlevels, img.ht = job["mct"], job["reversible"], job["nlevels"], job["ht"]
    bpp = (1 if job["prec"] <= 8 else 2) if job["ncomp"] == 1 else (4 if job["prec"] <= 8 else 8)
    stride = job["width"] * bpp
    tcs = jobs.as_ctypes(job["tilecomps"], O.TileComp)
    cbs = jobs.as_ctypes(job["cblks"], O.CBlk)
    return O.decode_image(img, tcs, cbs, job["blob"], stride, stride * job["height"], threads), bpp


@pytest.mark.parametrize("w,h,ncomp,prec,tw,th,levels", [
    (96, 80, 3, 8, None, None, 3), (130, 70, 3, 8, 64, 64, 2), (64, 64, 1, 8, None, None, 5),
    (100, 60, 1, 12, 48, 32, 2), (70, 50, 3, 16, None, None, 3),
])
def test_lossless_path_is_identity(w, h, ncomp, prec, tw, th, levels):
    s = jobs.synth_image(w, h, ncomp, prec, seed=1)
    job = jobs.build_ref_job(s, prec, tw, th, nlevels=levels, reversible=True, threads=2)
    pix, bpp = oracle_decode(job)
    if prec <= 8:
        got = pix.reshape(h, w, bpp)
        for c in range(ncomp):
            assert np.array_equal(got[:, :, c], s[c].astype(np.uint8)), c
        if ncomp == 3:
            assert (got[:, :, 3] == 255).all()
    else:
        got = pix.reshape(h, w, bpp // 2, 2).astype(np.int64)
        val = (got[..., 0] << 8) | got[..., 1]
        maxv = (1 << prec) - 1
        for c in range(ncomp):
            want = s[c].astype(np.int64) * 65535
            want = ((want + 2**31) % 2**32 - 2**31)                         # int32 wrap (decoder.go:464)
            want = (np.abs(want) // maxv * np.sign(want)) & 0xFFFF
            assert np.array_equal(val[:, :, c], want), c


def test_lossy_path_is_close():
    w, h, prec = 96, 64, 8
    s = jobs.synth_image(w, h, 3, prec, seed=2)
    job = jobs.build_ref_job(s, prec, nlevels=3, reversible=False, quality=1.0, threads=2)
    pix, bpp = oracle_decode(job)
    got = pix.reshape(h, w, 4)[:, :, :3].astype(int)
    want = np.moveaxis(s, 0, 2).astype(int)
    # sanity only: the reference's lossy chain is loose by construction (unit-step quantiser on K-scaled
    # sub-bands, truncating int32(v+0.5) twice, 5-digit ICT constants whose forward/inverse pair is only
    # 1e-2 accurate, mct_test.go:42-69)
    assert np.abs(got - want).max() <= 12
    assert np.abs(got - want).mean() < 2.5


def test_job_tables_are_abi_sized():
    assert jobs.CBLK_DT.itemsize == C.sizeof(O.CBlk) == 40
    assert jobs.TILECOMP_DT.itemsize == C.sizeof(O.TileComp) == 32
    assert C.sizeof(O.Image) == 28                      # == j2k_image_t (include/j2kgpu.h)
