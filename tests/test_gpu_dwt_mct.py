"""GPU parity, DWT / MCT / pack stages through the C ABI against the oracle: bit-exact for int32 5-3,
bit-exact for float64 9-7 as well (same operation order, no FMA), exact for RCT / ICT / DC / createImage."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

SHAPES = [(8, 8, 1), (8, 8, 2), (16, 16, 3), (32, 32, 4), (64, 64, 5),          # dwt_test.go:152-187
          (2, 2, 1), (1, 1, 1), (1, 9, 2), (9, 1, 2), (5, 7, 2), (37, 23, 3), (3, 130, 4),
          (64, 33, 1), (65, 64, 2), (130, 70, 3), (257, 255, 5), (512, 512, 5), (640, 360, 5)]


@pytest.mark.parametrize("w,h,levels", SHAPES)
def test_reconstruct53_bit_exact(gpu_ctx, w, h, levels):
    rng = np.random.default_rng(w * 1000 + h)
    c = rng.integers(-2000, 2000, w * h).astype(np.int32)
    assert np.array_equal(gpu_ctx.reconstruct_multilevel53(c, w, h, levels), O.reconstruct53(c, w, h, levels))


def test_reconstruct53_roundtrip_and_wraparound(gpu_ctx):
    d = (np.arange(64 * 64) % 256).astype(np.int32)                              # dwt_test.go:159-164 pattern
    assert np.array_equal(gpu_ctx.reconstruct_multilevel53(O.decompose53(d, 64, 64, 5), 64, 64, 5), d)
    rng = np.random.default_rng(5)
    c = rng.integers(-2**31, 2**31 - 1, 40 * 24).astype(np.int64).astype(np.int32)   # Go int32 wrap-around
    assert np.array_equal(gpu_ctx.reconstruct_multilevel53(c, 40, 24, 3), O.reconstruct53(c, 40, 24, 3))


@pytest.mark.parametrize("w,h,levels", SHAPES)
def test_reconstruct97_bit_exact_f64(gpu_ctx, w, h, levels):
    rng = np.random.default_rng(w * 1000 + h + 7)
    c = rng.normal(0, 500, w * h)
    got, want = gpu_ctx.reconstruct_multilevel97(c, w, h, levels), O.reconstruct97(c, w, h, levels)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("w,h,levels,rev", [(64, 64, 5, 1), (64, 64, 5, 0), (100, 37, 3, 0), (512, 512, 5, 0),
                                             (33, 65, 2, 1), (16, 16, 0, 0), (16, 16, 0, 1)])
def test_apply_inverse_dwt(gpu_ctx, w, h, levels, rev):
    """tcd.go:416-437 incl. the truncating int32(v + 0.5) (and its effect with zero levels)"""
    rng = np.random.default_rng(w + h + levels)
    c = rng.integers(-3000, 3000, w * h).astype(np.int32)
    assert np.array_equal(gpu_ctx.apply_inverse_dwt(c, w, h, levels, rev), O.apply_inverse_dwt(c, w, h, levels, rev))


def test_dwt_levels_zero_and_empty(gpu_ctx):
    c = np.arange(12, dtype=np.int32)
    assert np.array_equal(gpu_ctx.reconstruct_multilevel53(c, 4, 3, 0), c)
    assert gpu_ctx.reconstruct_multilevel53(np.zeros(0, np.int32), 0, 0, 3).size == 0


def test_inverse_rct(gpu_ctx):
    """mct_test.go:8-40, 533-598"""
    rng = np.random.default_rng(31)
    for n in (0, 1, 3, 4, 1000, 70001):
        y, u, v = (rng.integers(-70000, 70000, n).astype(np.int32) for _ in range(3))
        for a, b in zip(gpu_ctx.inverse_rct(y, u, v), O.inv_rct(y, u, v)):
            assert np.array_equal(a, b)
    r, g, b = [100, 150, 200, 50], [110, 140, 190, 60], [120, 130, 180, 70]
    got = gpu_ctx.inverse_rct(*O.fwd_rct(r, g, b))
    assert [x.tolist() for x in got] == [r, g, b]


def test_inverse_ict_bit_exact(gpu_ctx):
    rng = np.random.default_rng(32)
    y, cb, cr = (rng.normal(0, 200, 5000) for _ in range(3))
    for a, b in zip(gpu_ctx.inverse_ict(y, cb, cr), O.inv_ict(y, cb, cr)):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("prec", [1, 4, 8, 10, 12, 16])
def test_dc_level_shift_inverse(gpu_ctx, prec):
    """mct_test.go:71-99, 681-717"""
    d = np.array([-(1 << (prec - 1)), -1, 0, 1, (1 << (prec - 1)) - 1], np.int32)
    assert np.array_equal(gpu_ctx.dc_level_shift_inverse(d, prec), O.dc_shift_inverse(d, prec))


@pytest.mark.parametrize("ncomp,prec", [(1, 8), (3, 8), (4, 8), (1, 5), (3, 7), (1, 12), (3, 12), (4, 10), (1, 16), (3, 16)])
def test_create_image_all_layouts(gpu_ctx, ncomp, prec):
    """decoder.go:417-588: Gray / RGBA / Gray16 / RGBA64 (big-endian), scaling and the int32 overflow quirk"""
    rng = np.random.default_rng(ncomp * 100 + prec)
    w, h = 37, 11
    lo, hi = -(1 << (prec - 1)), (1 << prec) + (1 << (prec - 1))
    comps = [rng.integers(lo, hi, w * h).astype(np.int32) for _ in range(ncomp)]
    for c in comps:
        c[:4] = [0, (1 << prec) - 1, (1 << prec) // 2, (1 << prec) // 2 + 1]
    want, bpp = O.create_image(comps, w, h, prec)
    assert np.array_equal(gpu_ctx.create_image(comps, w, h, prec), want)


def test_create_image_bad_component_count(gpu_ctx, j2k):
    """decoder.go:585-586 "unsupported number of components" """
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.create_image([np.zeros(4, np.int32)] * 2, 2, 2, 8)
    assert e.value.code == j2k.E_UNSUPPORTED and "unsupported number of components" in str(e.value)


@pytest.mark.parametrize("ncomp,prec,rev,sgnd", [(3, 8, 1, 0), (3, 8, 0, 0), (3, 12, 0, 0), (1, 8, 1, 0), (3, 8, 1, 1), (4, 8, 1, 0)])
def test_decoder_tail_then_pack(gpu_ctx, j2k, ncomp, prec, rev, sgnd):
    """decoder.go:321-348 followed by createImage, fused in one kernel"""
    rng = np.random.default_rng(ncomp + prec + rev)
    w, h = 50, 9
    comps = [rng.integers(-(1 << (prec - 1)) - 50, (1 << (prec - 1)) + 50, w * h).astype(np.int32) for _ in range(ncomp)]
    img = j2k.make_image(w, h, ncomp, prec, sgnd=sgnd, mct=1, reversible=rev)
    got = gpu_ctx.mct_dc_pack(img, comps, apply_tail=True)
    after = O.decoder_tail(comps, 1, rev, [prec] * ncomp, [sgnd] * ncomp)
    want, _ = O.create_image(after, w, h, prec)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ncomp,prec,cs,mct,rev", [(3, 8, 1, 0, 1), (3, 8, 2, 0, 1), (3, 12, 1, 0, 0), (4, 8, 2, 0, 1), (3, 16, 2, 1, 0),
                                                   (1, 8, 1, 0, 1), (3, 8, 3, 0, 1), (3, 12, 3, 0, 0), (3, 8, 4, 0, 1), (4, 8, 5, 0, 1),
                                                   (4, 12, 6, 0, 0), (3, 8, 5, 0, 1), (3, 8, 6, 0, 1)])
def test_colour_conversion_then_pack(gpu_ctx, j2k, ncomp, prec, cs, mct, rev):
    """decoder.go:321-356 followed by createImage: the conversions of colorspace.go that need no pow() run in the same
    epilogue (values outside the nominal range included: the conversions clamp; CMYK / YCCK with 3 components: no-op)"""
    rng = np.random.default_rng(10 * ncomp + prec + cs)
    w, h = 61, 7
    comps = [rng.integers(-(1 << (prec - 1)) - 90, (1 << (prec - 1)) + 90, w * h).astype(np.int32) for _ in range(ncomp)]
    img = j2k.make_image(w, h, ncomp, prec, sgnd=0, mct=mct, reversible=rev, colorspace=cs)
    got = gpu_ctx.mct_dc_pack(img, comps, apply_tail=True)
    after = O.decoder_tail(comps, mct, rev, [prec] * ncomp, [0] * ncomp)
    after = O.colour_convert(after, prec, cs)
    want, _ = O.create_image(after, w, h, prec)
    assert np.array_equal(got, want)


def _unpack(pix, prec):
    a = np.asarray(pix, np.uint8).astype(np.int32)
    return a if prec <= 8 else (a[0::2] << 8) | a[1::2]


@pytest.mark.parametrize("ncomp,prec,cs,mct,rev", [(3, 8, 7, 0, 1), (3, 12, 7, 0, 0), (3, 16, 8, 0, 1), (4, 8, 8, 0, 1), (3, 8, 9, 0, 1),
                                                   (3, 12, 9, 1, 0), (3, 8, 10, 0, 1), (3, 16, 10, 0, 1), (4, 8, 10, 1, 1), (1, 8, 7, 0, 1)])
def test_pow_colour_conversion_then_pack(gpu_ctx, j2k, ncomp, prec, cs, mct, rev):
    """CIELab / CIEJab / e-sRGB / ROMM-RGB (colorspace.go:250-427) in the pixel epilogue: CUDA's pow() where Go has math.Pow,
    so the tolerance is 1 LSB (north_star's floating-point tolerance); every other operation is in the reference's order"""
    rng = np.random.default_rng(10 * ncomp + prec + cs)
    w, h = 97, 61
    comps = [rng.integers(-(1 << (prec - 1)) - 30, (1 << (prec - 1)) + 30, w * h).astype(np.int32) for _ in range(ncomp)]
    img = j2k.make_image(w, h, ncomp, prec, sgnd=0, mct=mct, reversible=rev, colorspace=cs)
    got = gpu_ctx.mct_dc_pack(img, comps, apply_tail=True)
    after = O.decoder_tail(comps, mct, rev, [prec] * ncomp, [0] * ncomp)
    after = O.colour_convert(after, prec, cs)
    want, _ = O.create_image(after, w, h, prec)
    g, w_ = _unpack(got, prec), _unpack(want, prec)
    d = np.abs(g - w_)
    # a NaN of the ROMM conversion (negative input) packs to the same bytes on both sides; everything else within 1 LSB
    assert d.max() <= 1, (int(d.max()), int((d > 0).sum()))
    assert (d > 0).sum() <= max(1, d.size // 10000)


def test_colour_conversion_not_built(gpu_ctx, j2k):
    img = j2k.make_image(8, 8, 3, 8, colorspace=11)
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.mct_dc_pack(img, [np.zeros(64, np.int32)] * 3)
    assert e.value.code == j2k.E_UNSUPPORTED
