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
