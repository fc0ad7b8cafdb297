"""GPU parity, ISO mode (J2KGPU_MODE_ISO): conformant HTJ2K decode through the C ABI.

Checkers: oracle/iso_ht.c (pinned by OpenJPEG, tests/test_iso_codestream.py) for the block decoder; the ISO
inverse 5-3 = reference Inverse2D53 applied to the transpose (ISO does rows first, the reference columns first);
and, end to end, the source image of a lossless codestream AND OpenJPEG's own decode of the same bytes."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import iso_ht_encode, jobs

pytestmark = pytest.mark.gpu
ISO = 1


def test_iso_ht_blocks_vs_oracle(gpu_ctx):
    rng = np.random.default_rng(31)
    blocks, want = [], []
    for t in range(400):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(1, 16))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.97)] = 0
        enc = iso_ht_encode(d, w, h)
        nbps = int(rng.integers(1, 4))
        blocks.append((enc, w, h, nbps, 0))
        # cleanup at bit-plane P = nbps - 1 with the mid-point bit below it (OpenJPEG's reconstruction)
        want.append((d << (nbps - 1)) + np.sign(d) * ((1 << (nbps - 1)) >> 1))
    for i, (got, w_) in enumerate(zip(gpu_ctx.ht_decode_blocks(blocks, mode=ISO), want)):
        assert np.array_equal(got, w_), i
        assert np.array_equal(got, O.iso_ht_decode(blocks[i][0], blocks[i][1], blocks[i][2], blocks[i][3])[0]), i


def _rev(q):
    """quarter units -> reversible value: sign * (|Q| >> 2)"""
    q = q.astype(np.int64)
    return (np.sign(q) * (np.abs(q) >> 2)).astype(np.int32)


def test_iso_ht_sigprop_magref_blocks_vs_oracle(gpu_ctx):
    """one HT set of 1, 2 or 3 passes (cleanup at bit-plane P, SigProp / MagRef at P - 1), streams of the extended
    writer: the GPU equals the CPU checker (itself identical to OpenJPEG on whole codestreams) and the writer's own
    statement of the reconstruction"""
    from datagen import iso_ht_encode_passes
    rng = np.random.default_rng(41)
    blocks, want = [], []
    for t in range(500):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(2, 12))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.97)] = 0
        small = rng.random(w * h) < rng.uniform(0, 0.5)
        d[small] = rng.integers(-3, 4, int(small.sum()))
        P, npass = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        enc, lcup, recon = iso_ht_encode_passes(d, w, h, P, npass)
        if not enc:
            continue
        q, rc = O.iso_ht_decode_passes(enc, lcup, w, h, P + 1, npass)
        assert rc == 0 and np.array_equal(q, recon)
        blocks.append((enc, w, h, P + 1, 0, npass, lcup))
        want.append(_rev(q))
    assert len(blocks) > 400
    for i, (got, w_) in enumerate(zip(gpu_ctx.ht_decode_blocks(blocks, mode=ISO), want)):
        assert np.array_equal(got, w_), (i, blocks[i][1:])


def test_iso_ht_damaged_refinement_segments_vs_oracle(gpu_ctx):
    """valid cleanup segments followed by refinement segments no encoder produces (random bytes, 0xFF runs, too short,
    too long, empty): same values as the checker, nothing faults"""
    from datagen import iso_ht_encode_passes
    rng = np.random.default_rng(42)
    blocks, want = [], []
    for t in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        d = rng.integers(-200, 201, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0.2, 0.95)] = 0
        P, npass = int(rng.integers(1, 3)), int(rng.integers(2, 4))
        enc, lcup, _ = iso_ht_encode_passes(d, w, h, P, npass)
        if not enc:
            continue
        ref = np.frombuffer(enc[lcup:], np.uint8).copy()
        mode = int(rng.integers(0, 5))
        if mode == 0:
            ref = rng.integers(0, 256, ref.size).astype(np.uint8)
        elif mode == 1:
            ref[rng.random(ref.size) < 0.4] = 0xFF
        elif mode == 2:
            ref = ref[: int(rng.integers(0, ref.size + 1))]
        elif mode == 3:
            ref = np.concatenate([ref, rng.integers(0, 256, int(rng.integers(1, 2000))).astype(np.uint8)])
        else:
            ref = np.full(int(rng.integers(1, 1500)), 0xFF, np.uint8)
        data = enc[:lcup] + ref.tobytes()
        blocks.append((data, w, h, P + 1, 0, npass, lcup))
        want.append(_rev(O.iso_ht_decode_passes(data, lcup, w, h, P + 1, npass)[0]))
    for i, (got, w_) in enumerate(zip(gpu_ctx.ht_decode_blocks(blocks, mode=ISO), want)):
        assert np.array_equal(got, w_), (i, blocks[i][1:])


def test_iso_ht_garbage_vs_oracle(gpu_ctx):
    """malformed segments decode to zero exactly as the checker does; nothing faults"""
    rng = np.random.default_rng(32)
    blocks = [(b"", 8, 8, 1, 0), (b"\x00\x00", 8, 8, 1, 0), (b"\xff\xff\xff", 8, 8, 1, 0)]
    for _ in range(300):
        n = int(rng.integers(2, 600))
        s = rng.integers(0, 256, n).astype(np.uint8)
        if rng.random() < 0.4:
            s[rng.random(n) < 0.3] = 0xFF
        scup = int(rng.integers(2, min(n, 4079) + 1))
        s[-1], s[-2] = scup >> 4, (s[-2] & 0xF0) | (scup & 0xF)
        blocks.append((s.tobytes(), int(rng.integers(1, 65)), int(rng.integers(1, 65)), 1, 0))
    for i, ((s, w, h, nb, _), got) in enumerate(zip(blocks, gpu_ctx.ht_decode_blocks(blocks, mode=ISO))):
        assert np.array_equal(got, O.iso_ht_decode(s, w, h, nb)[0]), i


def test_iso_ht_corrupted_magsgn_vs_oracle(gpu_ctx):
    """valid MEL / VLC segments over damaged MagSgn segments (0xFF runs, random bytes, truncation): the exponent
    predictor, the U_q > 31 and Mb bounds and the stuffing all see values no encoder produces"""
    rng = np.random.default_rng(33)
    blocks = []
    for _ in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(1, 16))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.9)] = 0
        s = np.frombuffer(iso_ht_encode(d, w, h), np.uint8).copy()
        if s.size < 2:
            continue
        scup = (int(s[-1]) << 4) + (int(s[-2]) & 0x0F)
        L = s.size - scup
        if L > 0:
            ms, tail = s[:L].copy(), s[L:]
            mode = int(rng.integers(0, 4))
            if mode == 0:
                ms[rng.random(L) < 0.3] = 0xFF
            elif mode == 1:
                ms = rng.integers(0, 256, L).astype(np.uint8)
            elif mode == 2:
                ms = ms[: int(rng.integers(0, L))]
            else:
                ms = np.concatenate([ms[: L // 2], np.full(L - L // 2, 0xFF, np.uint8)])
            s = np.concatenate([ms, tail])
        blocks.append((s.tobytes(), w, h, int(rng.integers(1, 4)), 0))
    assert len(blocks) > 200
    for i, ((s, w, h, nb, _), got) in enumerate(zip(blocks, gpu_ctx.ht_decode_blocks(blocks, mode=ISO))):
        assert np.array_equal(got, O.iso_ht_decode(s, w, h, nb)[0]), i


def iso_inverse53(plane, w, h, levels):
    """ISO 15444-1 inverse 5-3 on a Mallat plane: per level, rows first then columns = Inverse2D53 of the transpose"""
    a = np.array(plane, np.int32).reshape(h, w).copy()
    dims = [(w, h)]
    for _ in range(levels - 1):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
    for (lw, lh) in reversed(dims):
        sub = np.ascontiguousarray(a[:lh, :lw].T)
        a[:lh, :lw] = O.inv2d53(sub, lh, lw).reshape(lw, lh).T
    return a.reshape(-1)


@pytest.mark.parametrize("w,h,levels", [(8, 8, 1), (64, 64, 5), (16, 12, 2), (37, 23, 3), (1, 9, 2), (130, 70, 3),
                                         (256, 256, 5), (512, 512, 5), (640, 360, 5), (264, 136, 2)])
def test_iso_idwt53_vs_transposed_reference(gpu_ctx, w, h, levels):
    rng = np.random.default_rng(w * 7 + h)
    c = rng.integers(-3000, 3000, w * h).astype(np.int32)
    assert np.array_equal(gpu_ctx.reconstruct_multilevel53(c, w, h, levels, mode=ISO), iso_inverse53(c, w, h, levels))


def iso_pixels(j2k, ctx, job, coef_bits=0):
    img = j2k.make_image(job["width"], job["height"], job["ncomp"], job["prec"], mct=job["mct"], reversible=1,
                         nlevels=job["nlevels"], ht=1, mode=ISO, coef_bits=coef_bits)
    return ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                            job["blob"])


@pytest.mark.parametrize("w,h,ncomp,prec,tw,th,nl", [
    (64, 64, 1, 8, None, None, 0), (128, 128, 1, 8, None, None, 2), (256, 256, 3, 8, None, None, 5),
    (200, 150, 3, 8, None, None, 3), (512, 512, 3, 8, 256, 256, 5), (333, 211, 3, 8, 128, 128, 4),
    (640, 360, 1, 12, None, None, 5), (300, 200, 3, 16, None, None, 4), (1024, 512, 3, 8, 512, 512, 5),
    (328, 140, 3, 8, None, None, 5), (488, 508, 1, 8, None, None, 4),        # odd level sizes above the fused pair
])
def test_iso_whole_path_lossless_htj2k(j2k, gpu_ctx, w, h, ncomp, prec, tw, th, nl):
    s = jobs.synth_image(w, h, ncomp, prec, seed=w + 3 * h)
    job = jobs.build_iso_job(s, prec, tw, th, nl)
    got = iso_pixels(j2k, gpu_ctx, job)
    if prec <= 8:
        pix = got.reshape(h, w, -1)
        for c in range(ncomp):
            assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8)), c     # decode(encode(x)) == x
        if ncomp == 3:
            assert (pix[:, :, 3] == 255).all()
        # and the same codestream decoded by OpenJPEG (independent ISO decoder) gives the same pixels
        Image = pytest.importorskip("PIL.Image")
        im = Image.open(io.BytesIO(job["codestream"]))
        im.load()
        a = np.array(im)
        a = a[:, :, None] if a.ndim == 2 else a
        assert np.array_equal(pix[:, :, :ncomp], a)
    else:
        val = got.reshape(h, w, -1, 2).astype(np.int64)
        val = (val[..., 0] << 8) | val[..., 1]
        maxv = (1 << prec) - 1
        for c in range(ncomp):
            assert np.array_equal(val[:, :, c], s[c].astype(np.int64) * 65535 // maxv)   # ISO packing: no int32 wrap


@pytest.mark.parametrize("w,h,ncomp,tw,nl,passes,P", [
    (200, 150, 1, None, 3, 3, 1), (256, 256, 3, None, 5, 3, 1), (333, 211, 3, 128, 4, 2, 1), (512, 512, 3, 256, 5, 3, 2),
    (640, 360, 3, None, 5, 2, 3), (328, 140, 3, None, 5, 1, 2), (1024, 512, 3, 512, 5, 3, 1), (96, 80, 1, 32, 2, 3, 3),
    (300, 200, 3, None, 4, 3, 1), (1024, 256, 3, 512, 5, 2, 2),
])
def test_iso_whole_path_sigprop_magref_equals_openjpeg(j2k, gpu_ctx, w, h, ncomp, tw, nl, passes, P):
    """multi-pass HTJ2K codestreams (cleanup + SigProp [+ MagRef]) through the whole path: the pixels OpenJPEG decodes
    from the same bytes, and the C checker's; with and without int16 coefficient planes"""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image(w, h, ncomp, 8, seed=w + 7 * h)
    job = jobs.build_iso_job(s, 8, tw, tw, nl, ht_passes=passes, ht_plane=P)
    assert (job["cblks"]["num_passes"] == passes).any()
    im = Image.open(io.BytesIO(job["codestream"]))
    im.load()
    a = np.array(im).reshape(h, w, ncomp)
    want = O.iso_decode_job(job)
    assert np.array_equal(want.reshape(h, w, -1)[:, :, :ncomp], a)
    for cbits in (0, job["coef_bits"]):
        got = iso_pixels(j2k, gpu_ctx, job, coef_bits=cbits)
        assert np.array_equal(got, want), cbits
    if passes == 3 and P == 1:                                     # only isolated +-1 coefficients are lost
        assert np.abs(a.astype(int) - np.moveaxis(s, 0, 2)).max() <= 3


@pytest.mark.parametrize("w,h,ncomp,tw,th,nl", [(256, 256, 3, None, None, 5), (1024, 512, 3, 512, 512, 5),
                                                 (640, 360, 1, None, None, 5), (333, 211, 3, 128, 128, 4)])
def test_iso_int16_planes_with_coef_bits(j2k, gpu_ctx, w, h, ncomp, tw, th, nl):
    """j2k_image_t.coef_bits (max Mb of the codestream) <= 14 switches the coefficient planes to int16: same pixels"""
    s = jobs.synth_image(w, h, ncomp, 8, seed=5 * w + h)
    job = jobs.build_iso_job(s, 8, tw, th, nl)
    assert job["coef_bits"] <= 14
    a = iso_pixels(j2k, gpu_ctx, job)
    b = iso_pixels(j2k, gpu_ctx, job, coef_bits=job["coef_bits"])
    assert np.array_equal(a, b)
    pix = b.reshape(h, w, -1)
    for c in range(ncomp):
        assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8))


def test_iso_coef_bits_violation_zeroes_blocks(j2k, gpu_ctx):
    """a declared magnitude bound that the stream exceeds marks the block malformed (zero), it never overflows int16"""
    s = jobs.synth_image(128, 128, 1, 8, seed=3)
    job = jobs.build_iso_job(s, 8, None, None, 2)
    got = iso_pixels(j2k, gpu_ctx, job, coef_bits=2).reshape(128, 128)
    assert got.shape == (128, 128) and not np.array_equal(got, s[0].astype(np.uint8))


# ---- classic EBCOT in ISO mode: real codestreams written by OpenJPEG ------------------------------------------------
def opj_encode(s, **kw):
    Image = pytest.importorskip("PIL.Image")
    a = np.moveaxis(s, 0, 2).astype(np.uint8) if s.shape[0] == 3 else s[0].astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="JPEG2000", no_jp2=True, **kw)
    return buf.getvalue()


def test_iso_ebcot_blocks_vs_oracle(gpu_ctx):
    """every code block of an OpenJPEG codestream, all passes and a random truncation, against oracle/iso_t1.c"""
    from datagen import codestream as cs
    s = jobs.synth_image(200, 150, 3, 8, seed=3)
    h = cs.parse_codestream(opj_encode(s, irreversible=False, num_resolutions=4, mct=1))
    rng = np.random.default_rng(0)
    blocks, want = [], []
    for b in h["blocks"]:
        if not b["passes"]:
            continue
        for npass in (b["passes"], int(rng.integers(1, b["passes"] + 1))):
            blocks.append((b["data"], b["w"], b["h"], b["num_bps"], b["band"], npass))
            v = O.iso_t1_decode(b["data"], b["w"], b["h"], b["num_bps"], npass, b["band"])
            want.append(np.sign(v) * (np.abs(v) >> 1))
    assert len(blocks) > 100
    for i, (g, w_) in enumerate(zip(gpu_ctx.t1_decode_blocks(blocks, mode=ISO), want)):
        assert np.array_equal(g, w_), i


@pytest.mark.parametrize("group", [4, 8, 16, 32])
def test_iso_ebcot_every_lanes_per_block_variant(j2k, gpu_ctx, group):
    """k_t1_iso<OT, G> for every G: block level (all passes and truncations, int32) and whole path (int16 planes, 3 layers
    and a lossy 9-7 stream with float32 planes) against the oracle / OpenJPEG"""
    from datagen import codestream as cs
    s = jobs.synth_image(160, 120, 3, 8, seed=13)
    h = cs.parse_codestream(opj_encode(s, irreversible=False, num_resolutions=3, mct=1))
    rng = np.random.default_rng(group)
    blocks, want = [], []
    for b in h["blocks"]:
        if not b["passes"]:
            continue
        npass = int(rng.integers(1, b["passes"] + 1))
        blocks.append((b["data"], b["w"], b["h"], b["num_bps"], b["band"], npass))
        v = O.iso_t1_decode(b["data"], b["w"], b["h"], b["num_bps"], npass, b["band"])
        want.append(np.sign(v) * (np.abs(v) >> 1))
    with gpu_ctx.options(t1_group=group):
        got = gpu_ctx.t1_decode_blocks(blocks, mode=ISO)
        for i, (g, w_) in enumerate(zip(got, want)):
            assert np.array_equal(g, w_), i
        for kw in (dict(num_resolutions=4, mct=1, quality_layers=[20, 5, 1]), dict(num_resolutions=4, mct=1, irreversible=True, quality_layers=[15])):
            data = opj_encode(s, **kw)
            got_px = gpu_ctx.decode_codestream(data).reshape(120, 160, 4)[:, :, :3]
            assert np.array_equal(got_px, np.array(pytest.importorskip("PIL.Image").open(io.BytesIO(data))))


def test_iso_ebcot_garbage_vs_oracle(gpu_ctx):
    rng = np.random.default_rng(7)
    blocks, want = [(b"", 8, 8, 5, 0, 0)], [np.zeros(64, np.int32)]
    for _ in range(150):
        n = int(rng.integers(1, 400))
        d = rng.integers(0, 256, n).astype(np.uint8)
        if rng.random() < 0.3:
            d[rng.random(n) < 0.3] = 0xFF
        w, h, nb, band = int(rng.integers(1, 65)), int(rng.integers(1, 65)), int(rng.integers(1, 13)), int(rng.integers(0, 4))
        npass = int(rng.integers(0, 3 * nb))
        blocks.append((d.tobytes(), w, h, nb, band, npass))
        v = O.iso_t1_decode(d.tobytes(), w, h, nb, npass if npass else 3 * nb - 2, band)
        want.append(np.sign(v) * (np.abs(v) >> 1))
    for i, (g, w_) in enumerate(zip(gpu_ctx.t1_decode_blocks(blocks, mode=ISO), want)):
        assert np.array_equal(g, w_), i


@pytest.mark.parametrize("w,h,ncomp,kw", [
    (200, 150, 3, dict(num_resolutions=4, mct=1)),
    (512, 512, 3, dict(num_resolutions=6, mct=1)),                          # BASELINE configs[0] as a real codestream
    (333, 211, 3, dict(num_resolutions=5, mct=1, tile_size=(128, 128))),
    (300, 200, 1, dict(num_resolutions=3, codeblock_size=(32, 32))),
    (256, 128, 3, dict(num_resolutions=4, mct=1, quality_mode="rates", quality_layers=[30, 10, 1])),
    (256, 192, 3, dict(num_resolutions=5, mct=1, quality_mode="rates", quality_layers=[25])),          # truncated passes
    (1024, 512, 3, dict(num_resolutions=6, mct=1, tile_size=(512, 512))),                               # wide fused kernel
])
def test_iso_whole_path_openjpeg_ebcot(j2k, gpu_ctx, w, h, ncomp, kw):
    """a classic JPEG 2000 codestream written by OpenJPEG decodes to exactly OpenJPEG's own pixels (and to the source
    image when lossless), with int32 and with int16 coefficient planes"""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data = opj_encode(s, irreversible=False, **kw)
    job = jobs.build_iso_job_from_codestream(data)
    ref = np.array(Image.open(io.BytesIO(data)))
    ref = ref[:, :, None] if ref.ndim == 2 else ref
    for cb in (0, job["coef_bits"]):
        img = j2k.make_image(w, h, ncomp, 8, mct=job["mct"], reversible=1, nlevels=job["nlevels"], ht=0, mode=ISO, coef_bits=cb)
        got = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                                   job["blob"])
        pix = got.reshape(h, w, -1)
        assert np.array_equal(pix[:, :, :ncomp], ref), cb
        if "quality_layers" not in kw or kw["quality_layers"][-1] <= 1:
            for c in range(ncomp):
                assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8))


@pytest.mark.parametrize("w,h,ncomp,kw", [
    (256, 256, 1, dict(num_resolutions=4, quality_mode="rates", quality_layers=[10])),
    (256, 256, 3, dict(num_resolutions=6, mct=1, quality_mode="rates", quality_layers=[20])),
    (200, 150, 3, dict(num_resolutions=4, mct=1, quality_mode="rates", quality_layers=[40, 20, 8])),              # 3 layers
    (333, 211, 3, dict(num_resolutions=5, mct=1, tile_size=(128, 128), quality_mode="dB", quality_layers=[38])),  # ragged tiles
    (1024, 512, 3, dict(num_resolutions=6, mct=1, tile_size=(512, 512), quality_mode="rates", quality_layers=[12])),
])
def test_iso_whole_path_openjpeg_lossy_97(j2k, gpu_ctx, w, h, ncomp, kw):
    """lossy 9-7 + ICT codestreams written by OpenJPEG decode to OpenJPEG's own pixels: max |delta| <= 1 LSB is the
    north_star tolerance for the irreversible path; the kernels use OpenJPEG's constants and operation order in
    float32, so the expected difference is zero"""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data = opj_encode(s, irreversible=True, **kw)
    job = jobs.build_iso_job_from_codestream(data)
    img = j2k.make_image(w, h, ncomp, 8, mct=job["mct"], reversible=0, nlevels=job["nlevels"], ht=0, mode=ISO)
    got = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                               job["blob"])
    pix = got.reshape(h, w, -1)
    ref = np.array(Image.open(io.BytesIO(data)))
    ref = ref[:, :, None] if ref.ndim == 2 else ref
    d = np.abs(pix[:, :, :ncomp].astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1                                                     # tolerance stated by north_star
    assert d.max() == 0                                                     # what the design achieves


@pytest.mark.parametrize("name", ["rgb_lossless", "rgb_lossless_layers_tiles", "rgb_truncated_53", "rgb_lossy_97",
                                  "gray_lossy_97_odd", "gray16_lossless"])
def test_iso_golden_openjpeg_vectors(j2k, gpu_ctx, name):
    """committed codestreams written by OpenJPEG decode on the GPU to the committed pixels OpenJPEG decoded from them"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "iso_openjpeg.npz"))
    data, ref = g[name + "_j2k"].tobytes(), g[name + "_pix"]
    job = jobs.build_iso_job_from_codestream(data)
    w, h, nc, prec = job["width"], job["height"], job["ncomp"], job["prec"]
    img = j2k.make_image(w, h, nc, prec, mct=job["mct"], reversible=job["reversible"], nlevels=job["nlevels"], ht=0, mode=ISO,
                         coef_bits=job["coef_bits"] if job["reversible"] else 0)
    got = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                               job["blob"])
    if prec > 8:
        be = got.reshape(h, w, 2).astype(np.uint16)
        pix = (be[:, :, 0] << 8) | be[:, :, 1]                                   # image.Gray16 is big-endian
        assert np.array_equal(pix, ref)
    elif nc == 1:
        assert np.array_equal(got.reshape(h, w), ref)
    else:
        assert np.array_equal(got.reshape(h, w, 4)[:, :, :3], ref)


@pytest.mark.parametrize("w,h,kw", [
    (256, 256, dict(irreversible=False, num_resolutions=6, mct=1)),
    (320, 200, dict(irreversible=False, num_resolutions=4, mct=1)),
    (256, 192, dict(irreversible=True, num_resolutions=5, mct=1, quality_mode="rates", quality_layers=[20])),
])
@pytest.mark.parametrize("reduce", [1, 2, 3])
def test_iso_reduce_resolution_matches_openjpeg(j2k, gpu_ctx, w, h, kw, reduce):
    """Config.ReduceResolution (jpeg2000.go:205-207): the host leaves out the finest resolutions when it builds the job
    tables; no kernel knows about it.  The result equals OpenJPEG's reduced decode of the same codestream."""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image(w, h, 3, 8, seed=w + reduce)
    data = opj_encode(s, **kw)
    job = jobs.build_iso_job_from_codestream(data, reduce=reduce)
    W, H = job["width"], job["height"]
    assert (W, H) == (-(-w // (1 << reduce)), -(-h // (1 << reduce)))             # decoder.go:289-295
    img = j2k.make_image(W, H, 3, 8, mct=job["mct"], reversible=job["reversible"], nlevels=job["nlevels"], ht=0, mode=ISO)
    got = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                               job["blob"]).reshape(H, W, 4)
    im = Image.open(io.BytesIO(data))
    im.reduce = reduce
    im.load()
    assert np.array_equal(got[:, :, :3], np.array(im))


@pytest.mark.parametrize("w,h,ncomp,tw,nl,step", [(128, 128, 1, None, 3, 2.0), (256, 192, 3, None, 5, 1.0),
                                                   (333, 211, 3, 128, 4, 4.0), (1024, 512, 3, 512, 5, 2.0),
                                                   (480, 270, 3, None, 5, 1.0)])      # BASELINE cfg5 geometry / 4
def test_iso_whole_path_lossy_htj2k(j2k, gpu_ctx, w, h, ncomp, tw, nl, step):
    """lossy HTJ2K (9-7, ICT, HT cleanup blocks): the GPU path gives OpenJPEG's pixels for the same codestream
    (north_star tolerance 1 LSB; achieved 0)"""
    from datagen import codestream as cs
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data, _ = cs.write_htj2k(s, 8, tw, tw, nl, lossy_step=step)
    job = jobs.build_iso_job_from_codestream(data)
    assert job["ht"] == 1 and job["reversible"] == 0
    img = j2k.make_image(w, h, ncomp, 8, mct=job["mct"], reversible=0, nlevels=nl, ht=1, mode=ISO)
    got = gpu_ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk),
                               job["blob"]).reshape(h, w, -1)
    ref = np.array(Image.open(io.BytesIO(data)))
    ref = ref[:, :, None] if ref.ndim == 2 else ref
    d = np.abs(got[:, :, :ncomp].astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1 and d.max() == 0
