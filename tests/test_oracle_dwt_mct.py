"""Pins the oracle's DWT / MCT / DC-shift / createImage restatement with the reference's tests
(internal/dwt/dwt_test.go, internal/mct/mct_test.go) plus independent numpy re-derivations."""
import numpy as np
import pytest

import oracle_lib as O


# ---------------------------------------------------------------- 1-D 5-3 (dwt_test.go:8-46)
@pytest.mark.parametrize("data", [
    [42], [10, 20], [1, 2, 3, 4], [1, 2, 3, 4, 5, 6, 7, 8], [1, 2, 3, 4, 5, 6, 7],
    [0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100], [50] * 8, [-10, 10] * 4,
])
def test_53_1d_roundtrip_reference_cases(data):
    assert O.inv53(O.fwd53(data)).tolist() == data


def _np_inv53(c):
    """independent numpy statement of dwt.go:122-147 (L..H.. layout in, samples out)"""
    n = len(c)
    if n < 2:
        return np.array(c, np.int64)
    half = (n + 1) // 2
    x = np.zeros(n, np.int64)
    x[0::2] = c[:half]
    x[1::2] = c[half:]
    for i in range(0, n, 2):
        l = x[i - 1] if i - 1 >= 0 else x[i + 1]
        r = x[i + 1] if i + 1 < n else x[i - 1]
        x[i] -= (l + r + 2) >> 2
    for i in range(1, n, 2):
        if i + 1 < n:
            x[i] += (x[i - 1] + x[i + 1]) >> 1
        else:
            x[i] += x[i - 1]
    return x


def test_53_1d_matches_numpy_statement():
    rng = np.random.default_rng(1)
    for n in list(range(1, 20)) + [63, 64, 65, 255, 256, 8192]:       # dwt_test.go:455-495 uses 8192
        c = rng.integers(-5000, 5000, n)
        assert np.array_equal(O.inv53(c), _np_inv53(c)), n
        assert np.array_equal(O.inv53(O.fwd53(c)), c), n


def test_53_int32_wraps_like_go():
    c = np.array([2**31 - 1, -2**31, 2**31 - 5, 2**30, -2**30, 7], np.int32)
    assert np.array_equal(O.inv53(O.fwd53(c)), c)


# ---------------------------------------------------------------- 1-D 9-7 (dwt_test.go:48-79)
@pytest.mark.parametrize("data", [
    [42.0], [10.0, 20.0], [1.0, 2.0, 3.0, 4.0], [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0],
    [0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100],
])
def test_97_1d_roundtrip_reference_cases(data):
    out = O.inv97(O.fwd97(data))
    assert np.max(np.abs(out - np.array(data, float))) < 1e-10


def _np_inv97(c):
    """independent statement of dwt.go:213-262 with the same operation order (numpy float64 never
    fuses multiply-add)"""
    a, b, g, d = -1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971
    K, Ki = 1.230174104914001, 0.812893066115961
    n = len(c)
    if n < 2:
        return np.array(c, np.float64)
    half = (n + 1) // 2
    x = np.zeros(n, np.float64)
    x[0::2] = c[:half]
    x[1::2] = c[half:]
    x[0::2] *= K
    x[1::2] *= Ki

    def step(start, coef):
        for i in range(start, n, 2):
            if 0 < i < n - 1:
                x[i] -= coef * (x[i - 1] + x[i + 1])
            elif i == 0:
                x[0] -= (2 * coef) * x[1]
            else:
                x[n - 1] -= (2 * coef) * x[n - 2]
    step(0, d); step(1, g); step(0, b); step(1, a)
    return x


def test_97_1d_bit_exact_vs_numpy_statement():
    rng = np.random.default_rng(2)
    for n in list(range(1, 14)) + [64, 65, 127, 1000]:
        c = rng.normal(0, 300, n)
        got, want = O.inv97(c), _np_inv97(c)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), n   # same bits, no FMA anywhere


# ---------------------------------------------------------------- 2-D / multi-level
@pytest.mark.parametrize("w,h", [(2, 2), (4, 4), (8, 8), (16, 16), (8, 4), (4, 8), (5, 7), (1, 9), (9, 1), (1, 1)])
def test_53_2d_roundtrip(w, h):
    """dwt_test.go:81-116 (+ odd and degenerate shapes)"""
    d = np.arange(w * h, dtype=np.int32) * 10
    assert np.array_equal(O.inv2d53(O.fwd2d53(d, w, h), w, h), d)


def test_53_2d_is_columns_then_rows():
    """dwt.go:410-429: inverse = all columns, then all rows, on the L|H / L-over-H layout"""
    rng = np.random.default_rng(4)
    w, h = 11, 6
    c = rng.integers(-999, 999, (h, w))
    t = np.stack([_np_inv53(c[:, x]) for x in range(w)], axis=1)
    want = np.stack([_np_inv53(t[y]) for y in range(h)], axis=0)
    assert np.array_equal(O.inv2d53(c, w, h).reshape(h, w), want)


@pytest.mark.parametrize("w,h,levels", [(8, 8, 1), (8, 8, 2), (16, 16, 3), (32, 32, 4), (64, 64, 5),
                                         (37, 23, 3), (512, 512, 5)])
def test_53_multilevel_roundtrip(w, h, levels):
    """dwt_test.go:152-187 TestMultiLevel53_Roundtrip (+ ragged and cfg1-sized cases)"""
    d = (np.arange(w * h) % 256).astype(np.int32)
    assert np.array_equal(O.reconstruct53(O.decompose53(d, w, h, levels), w, h, levels), d)


def test_53_multilevel_is_dense_prefix_not_mallat():
    """dwt.go:524-548: level l works on data[0:w_l*h_l] with stride w_l"""
    rng = np.random.default_rng(6)
    w, h, L = 16, 12, 3
    c = rng.integers(-500, 500, w * h).astype(np.int32)
    want = c.copy()
    dims = [(w, h)]
    for _ in range(L - 1):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
    for (lw, lh) in reversed(dims):
        want[: lw * lh] = O.inv2d53(want[: lw * lh], lw, lh)
    assert np.array_equal(O.reconstruct53(c, w, h, L), want)


@pytest.mark.parametrize("w,h,levels", [(8, 8, 1), (8, 8, 2), (16, 16, 3), (32, 32, 4), (45, 31, 3)])
def test_97_multilevel_roundtrip(w, h, levels):
    """dwt_test.go:275-309 TestMultiLevel97_Roundtrip (1e-9)"""
    d = (np.arange(w * h) % 256).astype(np.float64)
    out = O.reconstruct97(O.decompose97(d, w, h, levels), w, h, levels)
    assert np.max(np.abs(out - d)) < 1e-9


def test_quantize_known_answers():
    """dwt_test.go:204-223 TestQuantize_Dequantize"""
    data = [0.0, 1.5, -2.3, 100.7, -50.2]
    q = O.quantize(data, 0.5)
    assert q.tolist() == [0, 3, -5, 201, -100]


def test_apply_inverse_dwt_97_truncating_round():
    """tcd.go:428-435: int32(v + 0.5) truncates toward zero (negatives round up)"""
    w = h = 8
    rng = np.random.default_rng(8)
    c = rng.integers(-300, 300, w * h).astype(np.int32)
    f = O.reconstruct97(c.astype(np.float64), w, h, 2)
    want = np.trunc(f + 0.5).astype(np.int32)
    assert np.array_equal(O.apply_inverse_dwt(c, w, h, 2, False), want)
    assert (want != np.floor(f + 0.5)).any()          # the quirk is actually exercised
    assert np.array_equal(O.apply_inverse_dwt(c, w, h, 2, True), O.reconstruct53(c, w, h, 2))


# ---------------------------------------------------------------- MCT (mct_test.go)
def test_rct_roundtrip_reference_case():
    """mct_test.go:8-40"""
    r, g, b = [100, 150, 200, 50], [110, 140, 190, 60], [120, 130, 180, 70]
    y, u, v = O.fwd_rct(r, g, b)
    rr, gg, bb = O.inv_rct(y, u, v)
    assert (rr.tolist(), gg.tolist(), bb.tolist()) == (r, g, b)


@pytest.mark.parametrize("r,g,b", [
    ([0, 0, 0], [0, 0, 0], [0, 0, 0]), ([255] * 3, [255] * 3, [255] * 3),
    ([-128, -64, 0], [-128, -64, 0], [-128, -64, 0]), ([-100, 0, 100], [50, -50, 150], [-50, 100, -100]),
    ([128], [128], [128]), ([], [], []),
])
def test_rct_edge_cases(r, g, b):
    """mct_test.go:533-598 TestForwardRCT_EdgeCases (incl. empty slices)"""
    y, u, v = O.fwd_rct(r, g, b)
    rr, gg, bb = O.inv_rct(y, u, v)
    assert (rr.tolist(), gg.tolist(), bb.tolist()) == (r, g, b)


def test_inverse_rct_formula():
    """mct.go:56-66: g = y - ((u+v)>>2); r = v+g; b = u+g"""
    rng = np.random.default_rng(10)
    y, u, v = (rng.integers(-4096, 4096, 1000) for _ in range(3))
    r, g, b = O.inv_rct(y, u, v)
    gg = y - ((u + v) >> 2)
    assert np.array_equal(g, gg) and np.array_equal(r, v + gg) and np.array_equal(b, u + gg)


def test_ict_roundtrip_and_formula():
    """mct_test.go:42-69 (1e-2) and mct.go:43-53 evaluated left to right without FMA"""
    r, g, b = [100.0, 150.0, 200.0, 50.0], [110.0, 140.0, 190.0, 60.0], [120.0, 130.0, 180.0, 70.0]
    y, cb, cr = O.fwd_ict(r, g, b)
    rr, gg, bb = O.inv_ict(y, cb, cr)
    assert max(np.abs(rr - r).max(), np.abs(gg - g).max(), np.abs(bb - b).max()) < 1e-2
    rng = np.random.default_rng(12)
    y, cb, cr = (rng.normal(0, 100, 500) for _ in range(3))
    rr, gg, bb = O.inv_ict(y, cb, cr)
    assert np.array_equal(rr, y + 1.402 * cr)
    assert np.array_equal(gg, (y - 0.34413 * cb) - 0.71414 * cr)
    assert np.array_equal(bb, y + 1.772 * cb)


@pytest.mark.parametrize("prec", [1, 4, 8, 10, 12, 16])
def test_dc_shift_precisions(prec):
    """mct_test.go:71-99, 681-717"""
    d = np.array([0, 1, (1 << prec) - 1], np.int32)
    s = O.dc_shift_forward(d, prec)
    assert np.array_equal(s, d - (1 << (prec - 1)))
    assert np.array_equal(O.dc_shift_inverse(s, prec), d)


# ---------------------------------------------------------------- decoder tail / createImage
def test_decoder_tail_ict_double_rounding():
    """decoder.go:326-340: ICT on float64 copies of the already-rounded planes, int32(v+0.5) truncating"""
    rng = np.random.default_rng(14)
    y, cb, cr = (rng.integers(-2000, 2000, 4096).astype(np.int32) for _ in range(3))
    out = O.decoder_tail([y, cb, cr], mct=1, reversible=0, prec=[12, 12, 12], sgnd=[0, 0, 0])
    yf, cbf, crf = y.astype(float), cb.astype(float), cr.astype(float)
    want = [np.trunc(yf + 1.402 * crf + 0.5), np.trunc((yf - 0.34413 * cbf) - 0.71414 * crf + 0.5),
            np.trunc(yf + 1.772 * cbf + 0.5)]
    for o, w_ in zip(out, want):
        assert np.array_equal(o, w_.astype(np.int32) + 2048)


def test_decoder_tail_signed_components_not_shifted():
    """decoder.go:344-348: DC shift only for unsigned components; no MCT below 3 components"""
    d = np.array([-5, 0, 5], np.int32)
    out = O.decoder_tail([d], mct=1, reversible=1, prec=[8], sgnd=[1])
    assert np.array_equal(out[0], d)
    out = O.decoder_tail([d], mct=1, reversible=1, prec=[8], sgnd=[0])
    assert np.array_equal(out[0], d + 128)


def _go_pack(comps, prec):
    """independent python statement of decoder.go:417-588 with int32 wrap-around"""
    maxv = (1 << prec) - 1
    n = len(comps[0])
    nch = 1 if len(comps) == 1 else 4
    out = []
    for i in range(n):
        for c in range(nch):
            if c < len(comps):
                v = min(max(int(comps[c][i]), 0), maxv)
                if prec <= 8:
                    if prec != 8:
                        v = int(np.int32(np.int64(v * 255) & 0xFFFFFFFF if v * 255 < 2**31 else v * 255 - 2**32))
                        v = int(v / maxv) if v >= 0 else -int(-v / maxv)
                else:
                    p = (v * 65535) & 0xFFFFFFFF
                    if p >= 2**31:
                        p -= 2**32
                    v = abs(p) // maxv * (1 if p >= 0 else -1)          # Go division truncates toward zero
            else:
                v = 255 if prec <= 8 else 65535
            if prec <= 8:
                out.append(v & 0xFF)
            else:
                out += [(v >> 8) & 0xFF, v & 0xFF]
    return np.array(out, np.uint8)


@pytest.mark.parametrize("ncomp,prec", [(1, 8), (3, 8), (4, 8), (1, 5), (3, 7), (1, 12), (3, 12), (4, 10), (1, 16), (3, 16)])
def test_create_image_layouts(ncomp, prec):
    """decoder.go:417-588: Gray / RGBA (A=255) / Gray16 / RGBA64 big-endian, precision scaling,
    int32 overflow of v*65535 at 16 bit"""
    rng = np.random.default_rng(100 + ncomp * 17 + prec)
    w, h = 13, 7
    lo, hi = -(1 << (prec - 1)), (1 << prec) + (1 << (prec - 1))
    comps = [rng.integers(lo, hi, w * h).astype(np.int32) for _ in range(ncomp)]
    for c in comps:                      # make sure the interesting values are present
        c[:4] = [0, (1 << prec) - 1, (1 << prec) // 2, (1 << prec) // 2 + 1]
    pix, bpp = O.create_image(comps, w, h, prec)
    assert bpp == {(1, True): 1, (1, False): 2}.get((ncomp, prec <= 8), 4 if prec <= 8 else 8)
    assert np.array_equal(pix, _go_pack(comps, prec))


def test_create_image_16bit_overflow_quirk_values():
    """decoder.go:464: v*65535 overflows int32 for v >= 32769 at prec 16; the low 16 bits of the
    truncated quotient are what lands in Pix"""
    comps = [np.array([0, 1, 32767, 32768, 32769, 40000, 65535], np.int32)]
    pix, bpp = O.create_image(comps, 7, 1, 16)
    vals = (pix[0::2].astype(int) << 8) | pix[1::2]
    want = []
    for v in comps[0]:
        p = (int(v) * 65535) & 0xFFFFFFFF
        if p >= 2**31:
            p -= 2**32
        q = abs(p) // 65535 * (1 if p >= 0 else -1)
        want.append(q & 0xFFFF)
    assert vals.tolist() == want
    assert vals[0] == 0 and vals[1] == 1 and vals[2] == 32767


def test_create_image_rejects_bad_component_count():
    """decoder.go:585-586"""
    with pytest.raises(ValueError):
        O.create_image([np.zeros(4, np.int32)] * 2, 2, 2, 8)
