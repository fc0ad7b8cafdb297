"""GPU: code-block styles RESET / VCAUSAL / PREDTERM / SEGSYM (ISO/IEC 15444-1 Table A.19) through the C ABI, on streams
OpenJPEG wrote with those styles (datagen/opj_direct.py): block level against the checker, whole path and the codestream
front door against OpenJPEG's own decode."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import codestream as cs, jobs

pytestmark = pytest.mark.gpu
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
ISO = 1
RESET, VCAUSAL, PREDTERM, SEGSYM = 0x02, 0x08, 0x10, 0x20
STYLES = [RESET, VCAUSAL, SEGSYM, PREDTERM, RESET | VCAUSAL, VCAUSAL | SEGSYM, RESET | VCAUSAL | PREDTERM | SEGSYM]


def opj_decode(data):
    im = PIL_Image.open(io.BytesIO(data))
    im.load()
    a = np.array(im)
    return a[:, :, None] if a.ndim == 2 else a


@pytest.mark.parametrize("group", [0, 4, 32])
@pytest.mark.parametrize("style", STYLES)
def test_blocks_vs_checker(gpu_ctx, style, group):
    """every code block of a styled stream, all passes and a random truncation, k_t1_iso == oracle/iso_t1.c"""
    s = jobs.synth_image(200, 150, 3, 8, seed=style)
    h = cs.parse_codestream(opj.encode(s, mode=style, num_resolutions=4, cblk=(64, 64) if style & 1 == 0 else (32, 32)))
    rng = np.random.default_rng(style)
    blocks, want = [], []
    for b in h["blocks"]:
        if not b["passes"]:
            continue
        for npass in (b["passes"], int(rng.integers(1, b["passes"] + 1))):
            blocks.append((b["data"], b["w"], b["h"], b["num_bps"], b["band"], npass))
            v = O.iso_t1_decode(b["data"], b["w"], b["h"], b["num_bps"], npass, b["band"], style)
            want.append(np.sign(v) * (np.abs(v) >> 1))
    assert len(blocks) > 100
    with gpu_ctx.options(t1_group=group):
        got = gpu_ctx.t1_decode_blocks(blocks, mode=ISO, style=style)
    for i, (g, w_) in enumerate(zip(got, want)):
        assert np.array_equal(g, w_), i


@pytest.mark.parametrize("style", STYLES)
@pytest.mark.parametrize("w,h,nc,kw", [
    (200, 150, 3, dict(num_resolutions=4)),
    (131, 77, 1, dict(num_resolutions=3, cblk=(32, 32))),
    (256, 192, 3, dict(num_resolutions=5, tile=(128, 128), rates=[30, 8, 1])),
    (240, 160, 3, dict(num_resolutions=4, irreversible=True, rates=[25, 6])),
])
def test_codestream_in_pixels_out(j2k, gpu_ctx, style, w, h, nc, kw):
    """raw bytes -> product tier-2 -> kernels == OpenJPEG's decode of its own stream (lossy 9-7 too: measured exact), and the
    same pixels from the table entry point with the harness's tier-2 and from the CPU checker"""
    s = jobs.synth_image(w, h, nc, 8, seed=style + w)
    data = opj.encode(s, mode=style, **kw)
    ref = opj_decode(data)
    got = gpu_ctx.decode_codestream(data).reshape(h, w, -1)[:, :, :nc]
    assert np.array_equal(got, ref)
    job = jobs.build_iso_job_from_codestream(data)
    for cbits in (0, job["coef_bits"] if job["reversible"] else 0):
        img = j2k.make_image(w, h, nc, 8, mct=job["mct"], reversible=job["reversible"], nlevels=job["nlevels"], ht=0, mode=ISO,
                             coef_bits=cbits, cblk_style=style)
        tab = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk), job["blob"])
        assert np.array_equal(tab.reshape(h, w, -1)[:, :, :nc], ref)
    assert np.array_equal(O.iso_decode_job(job).reshape(h, w, -1)[:, :, :nc], ref)


def test_batch_of_frames_with_different_styles(j2k, gpu_ctx):
    """the style travels per block: one batch call, every frame written with another style"""
    w, h = 192, 128
    srcs = [jobs.synth_image(w, h, 3, 8, seed=40 + i) for i in range(len(STYLES) + 1)]
    streams = [opj.encode(s, mode=m, num_resolutions=4) for s, m in zip(srcs, [0] + STYLES)]
    outs = gpu_ctx.decode_codestreams(streams)
    for s, got in zip(srcs, outs):
        assert np.array_equal(got.reshape(h, w, -1)[:, :, :3], np.moveaxis(s, 0, 2))


BYPASS, TERMALL = 0x01, 0x04
SEGMENTED = [TERMALL, BYPASS, BYPASS | TERMALL, TERMALL | RESET, BYPASS | VCAUSAL | SEGSYM, 0x3F, BYPASS | TERMALL | PREDTERM]


@pytest.mark.parametrize("style", SEGMENTED)
@pytest.mark.parametrize("w,h,nc,kw", [
    (200, 150, 3, dict(num_resolutions=4)),
    (131, 77, 1, dict(num_resolutions=3, cblk=(32, 32))),
    (256, 192, 3, dict(num_resolutions=5, tile=(128, 128), rates=[30, 8, 1])),
    (240, 160, 3, dict(num_resolutions=4, irreversible=True, rates=[25, 6])),
])
def test_segmented_styles_codestream_in_pixels_out(j2k, gpu_ctx, style, w, h, nc, kw):
    """selective bypass / termination on each pass (several codeword segments per block, raw passes): front door == OpenJPEG,
    for every lanes-per-block variant of the kernel and with int32 planes as well"""
    s = jobs.synth_image(w, h, nc, 8, seed=style + w)
    data = opj.encode(s, mode=style, **kw)
    ref = opj_decode(data)
    assert np.array_equal(gpu_ctx.decode_codestream(data).reshape(h, w, -1)[:, :, :nc], ref)
    for opts in (dict(t1_group=4), dict(t1_group=32), dict(coef32=1)):
        with gpu_ctx.options(**opts):
            assert np.array_equal(gpu_ctx.decode_codestream(data).reshape(h, w, -1)[:, :, :nc], ref), opts


def test_segment_table_outside_blob_is_refused(j2k, gpu_ctx):
    img = j2k.make_image(64, 64, 1, 8, mct=0, nlevels=0, ht=0, mode=ISO, cblk_style=TERMALL)
    tcs = (j2k.TileComp * 1)(j2k.TileComp(0, 0, 0, 64, 64, 0))
    cbs = (j2k.CBlk * 1)()
    cbs[0].w = cbs[0].h = 64
    cbs[0].num_bps, cbs[0].num_passes, cbs[0].data_len, cbs[0].step = 3, 7, 8, 1.0
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_tiles(img, tcs, cbs, np.zeros(16, np.uint8))          # 8 code bytes + 7 x 4 table bytes do not fit in 16
    assert e.value.code == j2k.E_RANGE
    img = j2k.make_image(64, 64, 1, 8, mct=0, nlevels=0, ht=0, mode=ISO, cblk_style=0x40)
    with pytest.raises(j2k.J2KError):
        gpu_ctx.decode_tiles(img, tcs, cbs, np.zeros(64, np.uint8))


def test_batch_with_segmented_and_plain_frames(j2k, gpu_ctx):
    w, h = 192, 128
    modes = [0, TERMALL, BYPASS, 0x3F, 0x02, BYPASS | TERMALL]
    srcs = [jobs.synth_image(w, h, 3, 8, seed=70 + i) for i in range(len(modes))]
    streams = [opj.encode(s, mode=m, num_resolutions=4) for s, m in zip(srcs, modes)]
    for s, got in zip(srcs, gpu_ctx.decode_codestreams(streams)):
        assert np.array_equal(got.reshape(h, w, -1)[:, :, :3], np.moveaxis(s, 0, 2))


def test_4k_rgb_all_styles_lossless(j2k, gpu_ctx):
    """BASELINE cfg2's geometry (3840x2160 RGB, 5 levels, 64x64 blocks) written by OpenJPEG with all six style bits (BYPASS |
    RESET | TERMALL | VCAUSAL | PREDTERM | SEGSYM): the front door returns the source"""
    s = jobs.synth_image(3840, 2160, 3, 8, seed=9)
    data = opj.encode(s, mode=0x3F, num_resolutions=6, tile=(1024, 1024))
    got = gpu_ctx.decode_codestream(data).reshape(2160, 3840, -1)[:, :, :3]
    assert np.array_equal(got, np.moveaxis(s, 0, 2))


@pytest.mark.parametrize("name", ["reset", "vcausal", "segsym", "all_four_layers_tiles", "all_four_lossy_97", "termall", "bypass",
        "all_six_layers_tiles", "bypass_termall_lossy_97"])
def test_golden_styled_streams(gpu_ctx, name):
    """committed bytes OpenJPEG wrote with each style -> front door -> the committed pixels OpenJPEG decoded from them"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "iso_styles.npz"))
    data, ref = g[name + "_j2k"].tobytes(), g[name + "_pix"]
    h, w = ref.shape[:2]
    nc = 1 if ref.ndim == 2 else ref.shape[2]
    got = gpu_ctx.decode_codestream(data).reshape(h, w, -1)[:, :, :nc]
    assert np.array_equal(got, ref.reshape(h, w, nc))
