"""CPU: code-block styles of ISO/IEC 15444-1 Table A.19 in the ISO-mode checker (oracle/iso_t1.c) and in the product's
tier-2 (host code, no GPU).  The streams are written by OpenJPEG itself through its C API (datagen/opj_direct.py: Pillow
does not pass the style byte through) and OpenJPEG's own decode of them is the expected image."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import jobs

PIL_Image = pytest.importorskip("PIL.Image")
opj = pytest.importorskip("datagen.opj_direct")


def _have_openjpeg():
    try:
        opj.lib()
        return True
    except OSError:
        return False


HAVE_OPENJPEG = _have_openjpeg()


@pytest.fixture(autouse=True)
def _needs_openjpeg_encoder(request):
    """the streams are written at test time by libopenjp2 (bundled with Pillow); the golden-vector cases do not need it"""
    if not HAVE_OPENJPEG and "golden" not in request.node.name and "segment" not in request.node.name:
        pytest.skip("libopenjp2 not found next to Pillow")

RESET, VCAUSAL, PREDTERM, SEGSYM = 0x02, 0x08, 0x10, 0x20
STYLES = [RESET, VCAUSAL, SEGSYM, PREDTERM, RESET | VCAUSAL, VCAUSAL | SEGSYM, RESET | VCAUSAL | PREDTERM | SEGSYM]


def opj_decode(data):
    im = PIL_Image.open(io.BytesIO(data))
    im.load()
    a = np.array(im)
    return a[None] if a.ndim == 2 else np.moveaxis(a, 2, 0)


def test_direct_encoder_writes_what_was_asked():
    s = jobs.synth_image(96, 80, 3, 8, seed=1)
    for mode, mct in ((0, 0), (RESET | SEGSYM, 1)):
        d = opj.encode(s, mode=mode, num_resolutions=3, mct=mct)
        k = d.index(b"\xff\x52")
        assert d[k + 8] == mct and d[k + 9] == 2 and d[k + 12] == mode
        assert np.array_equal(opj_decode(d), s)


@pytest.mark.parametrize("style", STYLES)
@pytest.mark.parametrize("w,h,nc,kw", [
    (200, 150, 3, dict(num_resolutions=4)),
    (131, 77, 1, dict(num_resolutions=3, cblk=(32, 32))),
    (256, 192, 3, dict(num_resolutions=5, tile=(128, 128), rates=[30, 8, 1])),
])
def test_styles_reversible_checker_equals_openjpeg(style, w, h, nc, kw):
    s = jobs.synth_image(w, h, nc, 8, seed=style + w)
    data = opj.encode(s, mode=style, **kw)
    job = jobs.build_iso_job_from_codestream(data)
    assert job["cblk_style"] == style
    got = O.iso_decode_job(job).reshape(h, w, -1)[:, :, :nc]
    assert np.array_equal(np.moveaxis(got, 2, 0), opj_decode(data))
    if kw.get("rates", [1])[-1] == 1:
        assert np.array_equal(np.moveaxis(got, 2, 0), s)


@pytest.mark.parametrize("style", [RESET, VCAUSAL | SEGSYM, RESET | VCAUSAL | PREDTERM | SEGSYM])
def test_styles_irreversible_checker_equals_openjpeg(style):
    w, h = 240, 160
    s = jobs.synth_image(w, h, 3, 8, seed=style)
    data = opj.encode(s, mode=style, irreversible=True, num_resolutions=4, rates=[25, 6])
    job = jobs.build_iso_job_from_codestream(data)
    got = O.iso_decode_job(job).reshape(h, w, -1)[:, :, :3]
    assert np.abs(np.moveaxis(got, 2, 0).astype(int) - opj_decode(data).astype(int)).max() == 0


def test_style_matters():
    """decoding a RESET / VCAUSAL / SEGSYM stream as the default style gives another image: the tests above are not vacuous"""
    s = jobs.synth_image(128, 128, 1, 8, seed=5)
    for style in (RESET, VCAUSAL, SEGSYM):
        job = jobs.build_iso_job_from_codestream(opj.encode(s, mode=style, num_resolutions=3))
        job["cblk_style"] = 0
        assert not np.array_equal(O.iso_decode_job(job).reshape(128, 128), s[0])


@pytest.mark.parametrize("style", STYLES)
def test_product_tier2_carries_the_style(j2k, style):
    s = jobs.synth_image(150, 100, 3, 8, seed=2)
    data = opj.encode(s, mode=style, num_resolutions=3)
    p = j2k.Parsed(data)
    try:
        assert p.image.cblk_style == style and p.image.ht == 0
    finally:
        p.close()


BYPASS, TERMALL = 0x01, 0x04
SEGMENTED = [TERMALL, BYPASS, BYPASS | TERMALL, TERMALL | RESET, BYPASS | VCAUSAL | SEGSYM, 0x3F, BYPASS | TERMALL | PREDTERM]


def checker_from_product_tier2(j2k, data):
    """the product's tier-2 (host code) fills the tables, the CPU checker decodes them"""
    p = j2k.Parsed(data)
    try:
        im = p.image
        tcs, cbs, blob = p.tables()
        job = dict(width=im.width, height=im.height, ncomp=im.ncomp, prec=im.prec[0], sgnd=im.sgnd[0], mct=im.mct, reversible=im.reversible,
                   nlevels=im.nlevels, ht=im.ht, tilecomps=tcs, cblks=cbs, blob=np.concatenate([blob, np.zeros(8, np.uint8)]),
                   cblk_style=im.cblk_style, colorspace=im.colorspace)
        return O.iso_decode_job(job).reshape(im.height, im.width, -1)[:, :, :im.ncomp], im.cblk_style
    finally:
        p.close()


@pytest.mark.parametrize("style", SEGMENTED)
@pytest.mark.parametrize("w,h,nc,kw", [
    (200, 150, 3, dict(num_resolutions=4)),
    (131, 77, 1, dict(num_resolutions=3, cblk=(32, 32))),
    (256, 192, 3, dict(num_resolutions=5, tile=(128, 128), rates=[30, 8, 1])),
    (240, 160, 3, dict(num_resolutions=4, irreversible=True, rates=[25, 6])),
])
def test_segmented_styles_tier2_and_checker_equal_openjpeg(j2k, style, w, h, nc, kw):
    """selective arithmetic-coding bypass and termination on each coding pass: several codeword segments per block, their
    lengths signalled per segment in the packet headers (B.10.7.2) -- product tier-2 + checker == OpenJPEG's own decode"""
    s = jobs.synth_image(w, h, nc, 8, seed=style + w)
    data = opj.encode(s, mode=style, **kw)
    got, st = checker_from_product_tier2(j2k, data)
    assert st == style
    assert np.array_equal(np.moveaxis(got, 2, 0), opj_decode(data))


def test_segment_count():
    L = O.lib()
    assert [L.iso_t1_num_segments(0, n) for n in (1, 10, 40)] == [1, 1, 1]
    assert [L.iso_t1_num_segments(TERMALL, n) for n in (1, 2, 10, 11)] == [1, 2, 10, 11]
    # bypass: passes 0..9 | (10, 11) raw | 12 cleanup | (13, 14) raw | 15 cleanup ...
    assert [L.iso_t1_num_segments(BYPASS, n) for n in (1, 10, 11, 12, 13, 14, 15, 16)] == [1, 1, 2, 2, 3, 4, 4, 5]


GOLD = ["reset", "vcausal", "segsym", "all_four_layers_tiles", "all_four_lossy_97", "termall", "bypass",
        "all_six_layers_tiles", "bypass_termall_lossy_97"]


@pytest.mark.parametrize("name", GOLD)
def test_golden_styled_streams_checker(j2k, name):
    """tests/golden/iso_styles.npz (tests/golden/make_golden_styles.py): bytes OpenJPEG wrote, pixels OpenJPEG decoded;
    the product's tier-2 fills the tables, the CPU checker decodes them"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "iso_styles.npz"))
    data, ref = g[name + "_j2k"].tobytes(), g[name + "_pix"]
    got, _ = checker_from_product_tier2(j2k, data)
    assert np.array_equal(got, ref.reshape(got.shape))
