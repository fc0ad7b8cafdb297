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
