"""CPU checks of the oracle's image tail (no GPU)."""
import numpy as np
import pytest

import oracle_lib as O


@pytest.mark.parametrize("prec,cs", [(8, 1), (8, 2), (12, 1), (16, 2)])
def test_colour_conversion_matches_numpy_restatement(prec, cs):
    """colorspace.go:90-140, 429-452, 483-491 in float64 numpy (IEEE, no FMA) against the C restatement"""
    rng = np.random.default_rng(prec + cs)
    n = 5000
    comps = [rng.integers(-300, (1 << prec) + 300, n).astype(np.int32) for _ in range(3)]
    got = O.colour_convert(comps, prec, cs)
    maxv, half = float((1 << prec) - 1), float(1 << (prec - 1))
    y, cb, cr = comps[0].astype(np.float64), comps[1].astype(np.float64) - half, comps[2].astype(np.float64) - half
    k = (1.5748, 0.1873, 0.4681, 1.8556) if cs == 1 else (1.402, 0.344136, 0.714136, 1.772)
    chans = (y + k[0] * cr, y - k[1] * cb - k[2] * cr, y + k[3] * cb)
    for g, v in zip(got, chans):
        want = np.where(v < 0, 0, np.where(v > maxv, int(maxv), np.trunc(v + 0.5))).astype(np.int32)
        assert np.array_equal(g, want)
    # fewer than three components: no-op (colorspace.go:93-95)
    one = O.colour_convert([comps[0]], prec, cs)
    assert np.array_equal(one[0], comps[0])


@pytest.mark.parametrize("prec,cs,ncomp", [(8, 3, 3), (12, 3, 3), (8, 4, 3), (8, 5, 4), (12, 5, 4), (8, 6, 4), (16, 6, 4), (8, 5, 3)])
def test_other_colour_conversions_match_numpy_restatement(prec, cs, ncomp):
    """PhotoYCC, CMY, CMYK, YCCK (colorspace.go:142-250) in float64 numpy against the C restatement"""
    rng = np.random.default_rng(100 * prec + cs)
    n = 4000
    comps = [rng.integers(-50, (1 << prec) + 50, n).astype(np.int32) for _ in range(ncomp)]
    got = O.colour_convert(comps, prec, cs)
    maxv = float((1 << prec) - 1)
    f = [c.astype(np.float64) for c in comps]
    if cs in (5, 6) and ncomp < 4:                                   # colorspace.go:194, 222: fewer than 4 components -> no-op
        for g, c in zip(got, comps):
            assert np.array_equal(g, c)
        return
    if cs == 4:
        want = [((1 << prec) - 1 - c).astype(np.int32) for c in comps[:3]]
    else:
        if cs == 5:
            k = f[3] / maxv
            chans = [(1 - f[i] / maxv) * (1 - k) * maxv for i in range(3)]
        else:
            scale = maxv / 255.0
            y, c1, c2 = f[0] / scale, f[1] / scale - 156.0, f[2] / scale - 156.0
            chans = [y + 1.3584 * c2, y - 0.4302 * c1 - 0.7915 * c2, y + 2.2179 * c1]
            if cs == 6:
                k = f[3] / maxv
                chans = [v * scale * (1 - k) for v in chans]
            else:
                chans = [v * scale for v in chans]
        want = [np.where(v < 0, 0, np.where(v > maxv, int(maxv), np.trunc(v + 0.5))).astype(np.int32) for v in chans]
    for g, w_ in zip(got[:3], want):
        assert np.array_equal(g, w_)
    if ncomp == 4:
        assert np.array_equal(got[3], comps[3])                      # the 4th component stays (and becomes alpha)


def _go_consts():
    from fractions import Fraction as F
    return float(F(6, 29)), float(F(4, 29)), float(F(108, 841)), float(F(10, 24))


def test_go_constant_expressions_round_like_c():
    """Go evaluates 6.0/29.0, 4.0/29.0, 3*delta*delta and 1.0/2.4 exactly and rounds once; the C / CUDA double expressions
    round to the same float64 values"""
    d, c, k, g = _go_consts()
    assert (6.0 / 29.0, 4.0 / 29.0, 3 * (6.0 / 29.0) * (6.0 / 29.0), 1.0 / 2.4) == (d, c, k, g)


def _srgb_gamma(lin):
    with np.errstate(invalid="ignore"):
        return np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * np.power(np.where(lin <= 0.0031308, 1.0, lin), 1.0 / 2.4) - 0.055)


def pow_conversion_numpy(comps, prec, cs):
    """colorspace.go:250-427 in float64 numpy, statement by statement (np.power = the libm's pow)"""
    d, c4, k, _ = _go_consts()
    maxv = float((1 << prec) - 1)
    f = [c.astype(np.float64) for c in comps[:3]]
    inv = lambda t: np.where(t > d, t * t * t, k * (t - c4))
    def xyz(x, y, z, clamp):
        lin = [3.2404542 * x - 1.5371385 * y - 0.4985314 * z, -0.9692660 * x + 1.8760108 * y + 0.0415560 * z,
               0.0556434 * x - 0.2040259 * y + 1.0572252 * z]
        if clamp:
            lin = [np.clip(v, 0, 1) for v in lin]
        return [_srgb_gamma(v) * maxv for v in lin]
    if cs in (7, 8):
        L, a, b = f[0] / maxv * 100.0, f[1] / maxv * 255.0 - 128.0, f[2] / maxv * 255.0 - 128.0
        fy = (L + 16.0) / 116.0
        fx, fz = a / 500.0 + fy, fy - b / 200.0
        chans = xyz(0.96422 * inv(fx), 1.0 * inv(fy), 0.82521 * inv(fz), False)
    elif cs == 9:
        chans = [_srgb_gamma(np.clip(v / maxv * 1.25 - 0.25, 0, 1)) * maxv for v in f]
    else:
        r, g, b = (np.power(v / maxv, 1.8) for v in f)
        chans = xyz(0.7977 * r + 0.1352 * g + 0.0313 * b, 0.2880 * r + 0.7119 * g + 0.0001 * b, 0.0 * r + 0.0 * g + 0.8249 * b, True)
    return [np.where(v < 0, 0, np.where(v > maxv, int(maxv), np.trunc(v + 0.5))).astype(np.int32) for v in chans]


@pytest.mark.parametrize("prec,cs", [(8, 7), (12, 7), (16, 8), (8, 8), (8, 9), (12, 9), (8, 10), (16, 10)])
def test_pow_colour_conversions_match_numpy_restatement(prec, cs):
    """CIELab, CIEJab, e-sRGB, ROMM-RGB (colorspace.go:250-427) in float64 numpy against the C restatement; ROMM with
    non-negative inputs (a negative one is NaN in Go and in C alike: checked separately)"""
    rng = np.random.default_rng(100 * prec + cs)
    n = 6000
    lo = 0 if cs == 10 else -40
    comps = [rng.integers(lo, (1 << prec) + 40, n).astype(np.int32) for _ in range(3)]
    got = O.colour_convert(comps, prec, cs)
    for g, w_ in zip(got, pow_conversion_numpy(comps, prec, cs)):
        assert np.array_equal(g, w_)
    assert len(set(np.concatenate(got).tolist())) > 100            # not a constant image
    four = O.colour_convert(comps + [comps[0]], prec, cs)            # a 4th component stays
    assert np.array_equal(four[3], comps[0]) and np.array_equal(four[0], got[0])


def test_cielab_known_points():
    """L* = 100, a* = b* = 0 is the D50 white taken through the reference's (unadapted) XYZ -> sRGB matrix, L* = 0 is black;
    CIEJab is the same arithmetic (colorspace.go:319-359)"""
    prec = 8
    white = [np.array([255], np.int32), np.array([128], np.int32), np.array([128], np.int32)]
    black = [np.array([0], np.int32), np.array([128], np.int32), np.array([128], np.int32)]
    w7, w8 = O.colour_convert(white, prec, 7), O.colour_convert(white, prec, 8)
    assert [int(c[0]) for c in w7] == [int(c[0]) for c in w8]
    assert w7[0][0] == 255 and w7[1][0] >= 240 and 200 <= w7[2][0] <= 255     # D50 white: reddish in an unadapted D65 space
    assert [int(c[0]) for c in O.colour_convert(black, prec, 7)] == [0, 0, 0]


def test_romm_negative_input_is_go_nan_conversion():
    """math.Pow(negative, 1.8) is NaN; clampToInt32 lets NaN through both comparisons and int32(NaN) is 0x80000000 on amd64"""
    got = O.colour_convert([np.array([-5], np.int32), np.array([10], np.int32), np.array([10], np.int32)], 8, 10)
    assert got[0][0] == -2147483648
