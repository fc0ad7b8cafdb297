"""GPU: the codestream front door (j2kgpu_decode_codestream / j2kgpu_decode_codestreams) -- raw codestream bytes in, pixels
out: tier-2 on host threads inside the library (host/tier2.cpp), overlapped with the copy / kernel / copy pipeline.
Checked against OpenJPEG's decode of the same bytes (independent decoder) and against the source images (lossless)."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import jobs

pytestmark = pytest.mark.gpu
PIL_Image = pytest.importorskip("PIL.Image")


def opj_decode(data, reduce=0):
    im = PIL_Image.open(io.BytesIO(data))
    if reduce:
        im.reduce = reduce
    im.load()
    return np.array(im)


def opj_encode(s, **kw):
    a = np.moveaxis(s, 0, 2).astype(np.uint8) if s.shape[0] == 3 else s[0].astype(np.uint8)
    buf = io.BytesIO()
    PIL_Image.fromarray(a).save(buf, format="JPEG2000", no_jp2=True, **kw)
    return buf.getvalue()


def pixels(got, h, w, nc):
    return got.reshape(h, w, -1)[:, :, :nc]


@pytest.mark.parametrize("w,h,nc,kw", [
    (96, 64, 1, dict(num_resolutions=3)),
    (200, 150, 3, dict(num_resolutions=4, mct=1)),
    (333, 211, 3, dict(num_resolutions=5, mct=1, tile_size=(128, 128))),
    (256, 256, 3, dict(num_resolutions=6, mct=1, quality_layers=[40, 20, 10, 5, 1])),
    (300, 200, 3, dict(num_resolutions=4, mct=1, irreversible=True, quality_layers=[30, 10])),
    (256, 192, 3, dict(num_resolutions=4, mct=1, progression="RPCL", quality_layers=[20, 5, 1])),
    (256, 192, 3, dict(num_resolutions=4, mct=1, progression="CPRL", quality_layers=[20, 1], tile_size=(128, 64))),
    (200, 150, 3, dict(num_resolutions=4, mct=1, plt=True, tile_size=(64, 64))),
    (333, 211, 3, dict(num_resolutions=5, mct=1, precinct_size=(128, 128), quality_layers=[20, 5, 1])),                       # user-defined precincts
    (333, 211, 3, dict(num_resolutions=4, mct=1, precinct_size=(64, 32), progression="PCRL", tile_size=(128, 128), quality_layers=[20, 1])),
    (300, 200, 3, dict(num_resolutions=4, mct=1, precinct_size=(64, 64), irreversible=True, progression="RPCL", quality_layers=[30, 10])),
])
def test_openjpeg_codestream_in_pixels_out(j2k, gpu_ctx, w, h, nc, kw):
    s = jobs.synth_image(w, h, nc, 8, seed=w + 1)
    data = opj_encode(s, **kw)
    got = gpu_ctx.decode_codestream(data)
    assert np.array_equal(pixels(got, h, w, nc), opj_decode(data).reshape(h, w, nc))


@pytest.mark.parametrize("passes,P", [(1, 0), (3, 1), (2, 2)])
def test_htj2k_batch_of_codestreams(j2k, gpu_ctx, passes, P):
    """a batch of distinct HTJ2K frames through one call; single-pass frames are lossless (= source)"""
    w, h, n = 384, 256, 6
    srcs = [jobs.synth_image(w, h, 3, 8, seed=100 + i) for i in range(n)]
    streams = [jobs.build_iso_job(s, 8, 128, 128, 4, ht_passes=passes, ht_plane=P)["codestream"] for s in srcs]
    with gpu_ctx.options(chunks="1,2,3"):                          # three chunks: the pipeline hands frames over in pieces
        outs = gpu_ctx.decode_codestreams(streams)
    for s, data, got in zip(srcs, streams, outs):
        assert np.array_equal(pixels(got, h, w, 3), opj_decode(data).reshape(h, w, 3))
        if passes == 1:
            assert np.array_equal(pixels(got, h, w, 3), np.moveaxis(s, 0, 2))
    again = gpu_ctx.decode_codestreams(streams)                    # default chunk plan, recycled pools: same pixels
    assert all(np.array_equal(a, b) for a, b in zip(outs, again))


@pytest.mark.parametrize("w,h,reduce", [(330, 210, 1), (330, 210, 2), (320, 200, 3)])
def test_reduce_resolution(j2k, gpu_ctx, w, h, reduce):
    s = jobs.synth_image(w, h, 3, 8, seed=5)
    data = opj_encode(s, num_resolutions=5, mct=1, tile_size=(128, 128), quality_layers=[10, 1])
    ref = opj_decode(data, reduce)
    got = gpu_ctx.decode_codestream(data, reduce)
    assert np.array_equal(pixels(got, ref.shape[0], ref.shape[1], 3), ref)


def test_page_locked_output_and_gray16(j2k, gpu_ctx):
    s = jobs.synth_image(320, 180, 1, 16, seed=3)
    from datagen import codestream as cs
    data, _ = cs.write_htj2k(s, 16, 128, 128, 4)
    out = gpu_ctx.host_alloc(320 * 180 * 2)
    try:
        got = gpu_ctx.decode_codestreams([data], outs=[out])[0].reshape(180, 320, 2)
        val = (got[:, :, 0].astype(np.uint16) << 8) | got[:, :, 1]  # Gray16 is big-endian (decoder.go:448-452)
        assert np.array_equal(val, s[0].astype(np.uint16))
    finally:
        gpu_ctx.host_free(out)


def test_jp2_file_and_alpha(j2k, gpu_ctx):
    """a JP2 file (container + RGBA codestream) straight into the front door"""
    s = jobs.synth_image(200, 150, 4, 8, seed=5)
    buf = io.BytesIO()
    PIL_Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8), "RGBA").save(buf, format="JPEG2000", num_resolutions=4)
    data = buf.getvalue()
    assert data[4:8] == b"jP  "
    assert np.array_equal(gpu_ctx.decode_codestream(data).reshape(150, 200, 4), np.moveaxis(s, 0, 2))


@pytest.mark.parametrize("rate", [0, 150])
def test_16_bit_rgb_written_by_opencv(j2k, gpu_ctx, rate):
    """16-bit RGB EBCOT (int32 planes, RGBA64 output) from a JP2 file written by OpenCV: OpenCV's own decode"""
    cv2 = pytest.importorskip("cv2")
    s = jobs.synth_image(300, 200, 3, 16, seed=6)
    bgr = np.ascontiguousarray(np.moveaxis(s, 0, 2).astype(np.uint16)[:, :, ::-1])
    ok, enc = cv2.imencode(".jp2", bgr, [cv2.IMWRITE_JPEG2000_COMPRESSION_X1000, rate] if rate else [])
    assert ok
    data = enc.tobytes()
    got = gpu_ctx.decode_codestream(data).reshape(200, 300, 4, 2)
    val = (got[..., 0].astype(np.uint16) << 8) | got[..., 1]
    assert np.array_equal(val[:, :, :3], cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)[:, :, ::-1])   # (OpenCV returns B G R)


def test_errors_surface_and_the_context_survives(j2k, gpu_ctx):
    s = jobs.synth_image(128, 128, 3, 8, seed=8)
    good = opj_encode(s, num_resolutions=3, mct=1)
    bad = bytearray(good)
    bad[good.index(b"\xff\x52") + 4 + 5] = 40                      # COD: 40 decomposition levels
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_codestreams([good, bytes(bad)])
    assert "codestream 1" in str(e.value)
    i = good.index(b"\xff\x5c")                                   # a COC that gives component 0 one level less than the others: not handled
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_codestream(good[:i] + b"\xff\x53" + (9).to_bytes(2, "big") + bytes([0, 0, 1, 4, 4, 0, 1]) + good[i:])
    assert e.value.code == j2k.E_UNSUPPORTED
    same = good[:i] + b"\xff\x53" + (9).to_bytes(2, "big") + bytes([0, 0, 2, 4, 4, 0, 1]) + good[i:]   # a COC that repeats COD: decoded
    assert np.array_equal(pixels(gpu_ctx.decode_codestream(same), 128, 128, 3), np.moveaxis(s, 0, 2))
    trunc = good[: len(good) * 2 // 3]                             # truncated: remaining packets absent, still decodes
    gpu_ctx.decode_codestream(trunc)
    assert np.array_equal(pixels(gpu_ctx.decode_codestream(good), 128, 128, 3), np.moveaxis(s, 0, 2))


def test_full_size_4k_htj2k_codestream(j2k, gpu_ctx):
    """BASELINE configs[1] as bytes: 3840x2160 RGB lossless HTJ2K, 6 resolutions, 512x512 tiles -> source pixels"""
    w, h = 3840, 2160
    s = jobs.synth_image(w, h, 3, 8, seed=77)
    data = jobs.build_iso_job(s, 8, 512, 512, 5)["codestream"]
    p = j2k.Parsed(data)
    assert p.info["tiles"] == 40 and p.info["zero_copy"] == 1
    got = gpu_ctx.decode_codestreams([data, data])
    for g in got:
        assert np.array_equal(pixels(g, h, w, 3), np.moveaxis(s, 0, 2))


@pytest.mark.parametrize("enumcs,cconv,kw", [
    (18, 1, dict(num_resolutions=4, mct=0)),                                   # sYCC, lossless 5-3
    (3, 2, dict(num_resolutions=3, mct=0, tile_size=(128, 128))),              # YCbCr(2)
    (14, 7, dict(num_resolutions=4, mct=0, irreversible=True, quality_layers=[10])),   # CIELab, 9-7
    (21, 10, dict(num_resolutions=4, mct=0)),                                  # ROMM-RGB
    (9, 3, dict(num_resolutions=4, mct=0)),                                    # PhotoYCC
    (16, 0, dict(num_resolutions=4, mct=1)),                                   # sRGB: no conversion
])
def test_jp2_file_with_colour_specification(j2k, gpu_ctx, enumcs, cconv, kw):
    """a JP2 file through the front door: the colr box's enumerated colour space selects the conversion to sRGB that
    decoder.go:350-356 applies; the pixels equal the CPU checker's with that conversion (1 LSB for the pow-based ones) and the
    raw codestream decodes to OpenJPEG's unconverted samples"""
    from datagen import codestream as cs
    w, h = 300, 200
    s = jobs.synth_image(w, h, 3, 8, seed=enumcs)
    data = opj_encode(s, **kw)
    got = pixels(gpu_ctx.decode_codestream(cs.wrap_jp2(data, enumcs, w, h)), h, w, 3)
    plain = pixels(gpu_ctx.decode_codestream(data), h, w, 3)
    assert np.array_equal(plain, opj_decode(data).reshape(h, w, 3))
    job = jobs.build_iso_job_from_codestream(data)
    job["colorspace"] = cconv
    want = pixels(O.iso_decode_job(job), h, w, 3)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= (1 if cconv >= 7 else 0)
    assert np.array_equal(got, plain) == (cconv == 0)


@pytest.mark.parametrize("kw", [dict(num_resolutions=4, mct=1), dict(num_resolutions=4, mct=1, irreversible=True, quality_layers=[20, 5])])
def test_qcc_marker_segments(j2k, gpu_ctx, kw):
    """per-component quantisation (QCC) with a QCD that says something else: the front door decodes to OpenJPEG's pixels"""
    from datagen import codestream as cs
    w, h = 200, 150
    s = jobs.synth_image(w, h, 3, 8, seed=3)
    data = cs.with_qcc(opj_encode(s, **kw))
    assert np.array_equal(pixels(gpu_ctx.decode_codestream(data), h, w, 3), opj_decode(data).reshape(h, w, 3))
    data = cs.with_coc(data)                                              # and per-component coding parameters (COC) on top
    assert np.array_equal(pixels(gpu_ctx.decode_codestream(data), h, w, 3), opj_decode(data).reshape(h, w, 3))
