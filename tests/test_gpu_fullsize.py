"""GPU parity at the FULL sizes of BASELINE.json configs[1..4] (cfg2 .. cfg5), both conformance modes, through the C ABI:
bit-exact against the CPU checkers (orc_decode_image for REF semantics, oracle/iso_path.c for ISO), against the source image
where the chain is lossless, and against OpenJPEG's own decode where OpenJPEG can read or write the stream.  A frame each:
the checkers finish in seconds."""
import io
import multiprocessing as mp
import os

import numpy as np
import pytest

import oracle_lib as O
from datagen import jobs

pytestmark = pytest.mark.gpu
ISO = 1
THREADS = max(4, min(32, os.cpu_count() or 4))


def ref_oracle(job):
    img = O.Image()
    img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
    for c in range(job["ncomp"]):
        img.prec[c], img.sgnd[c] = job["prec"], job["sgnd"]
    img.mct, img.reversible, img.nlevels, img.ht = job["mct"], job["reversible"], job["nlevels"], job["ht"]
    bpp = (1 if job["prec"] <= 8 else 2) if job["ncomp"] == 1 else (4 if job["prec"] <= 8 else 8)
    stride = job["width"] * bpp
    return O.decode_image(img, jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk),
                          job["blob"], stride, stride * job["height"], threads=THREADS)


def gpu(j2k, ctx, job, mode=0, coef_bits=0):
    img = j2k.make_image(job["width"], job["height"], job["ncomp"], job["prec"], sgnd=job["sgnd"], mct=job["mct"],
                         reversible=job["reversible"], nlevels=job["nlevels"], ht=job["ht"], mode=mode, coef_bits=coef_bits)
    return ctx.decode_tiles(img, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk), job["blob"])


def opj(data):
    Image = pytest.importorskip("PIL.Image")
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.array(im)


# ---- cfg2: 3840x2160 RGB 8-bit lossless 5-3, 6 resolutions, 512x512 tiles, RCT ------------------------------------------------
@pytest.mark.parametrize("ht", [1, 0])
def test_cfg2_ref_full_size(j2k, gpu_ctx, ht):
    """REF semantics: the reference's HT coder (ht=1) and its EBCOT coder (ht=0)"""
    s = jobs.synth_image_fast(3840, 2160, 3, 8, seed=1002)
    job = jobs.build_ref_job(s, 8, 512, 512, nlevels=5, reversible=True, ht=bool(ht), threads=THREADS)
    got = gpu(j2k, gpu_ctx, job)
    assert np.array_equal(got, ref_oracle(job))
    if not ht:                                                     # EBCOT is lossless: decode(encode(x)) == x
        pix = got.reshape(2160, 3840, 4)
        for c in range(3):
            assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8))


def test_cfg2_iso_htj2k_full_size(j2k, gpu_ctx):
    """the bench headline's frame as a conformant HTJ2K codestream: source image == OpenJPEG == CPU checker == GPU"""
    s = jobs.synth_image_fast(3840, 2160, 3, 8, seed=2002)
    job = jobs.build_iso_job(s, 8, 512, 512, 5)
    want = O.iso_decode_job(job, threads=THREADS)
    for cbits in (job["coef_bits"], 0):
        assert np.array_equal(gpu(j2k, gpu_ctx, job, ISO, cbits), want), cbits
    pix = want.reshape(2160, 3840, 4)
    for c in range(3):
        assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8))
    assert np.array_equal(opj(job["codestream"]), pix[:, :, :3])


def test_cfg2_iso_htj2k_three_passes_full_size(j2k, gpu_ctx):
    """the same frame with every block as cleanup + SigProp + MagRef (rate-controlled form of HTJ2K)"""
    s = jobs.synth_image_fast(3840, 2160, 3, 8, seed=2003)
    job = jobs.build_iso_job(s, 8, 512, 512, 5, ht_passes=3, ht_plane=1)
    want = O.iso_decode_job(job, threads=THREADS)
    assert np.array_equal(gpu(j2k, gpu_ctx, job, ISO, job["coef_bits"]), want)
    assert np.array_equal(opj(job["codestream"]), want.reshape(2160, 3840, 4)[:, :, :3])


# ---- cfg3: 3840x2160 RGB 12-bit lossy 9-7 EBCOT, ICT, 5 quality layers, LRCP ---------------------------------------------------
def test_cfg3_ref_full_size(j2k, gpu_ctx):
    """REF semantics, 12-bit, one tile, float64 9-7 + ICT with the reference's double rounding, RGBA64 output"""
    s = jobs.synth_image_fast(3840, 2160, 3, 12, seed=1003)
    job = jobs.build_ref_job(s, 12, None, None, nlevels=5, reversible=False, ht=False, threads=THREADS)
    assert np.array_equal(gpu(j2k, gpu_ctx, job), ref_oracle(job))


def test_cfg3_iso_openjpeg_stream_full_size(j2k, gpu_ctx):
    """ISO mode on a codestream OpenJPEG wrote: 4K RGB, irreversible 9-7 + ICT, 5 quality layers, LRCP, classic EBCOT.
    (8-bit: no encoder in this image writes 12-bit RGB; the 12-bit ISO case below is grey.)  GPU == CPU checker == OpenJPEG."""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image_fast(3840, 2160, 3, 8, seed=3003)
    buf = io.BytesIO()
    Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8)).save(buf, format="JPEG2000", no_jp2=True, irreversible=True, mct=1,
                                                               num_resolutions=6, quality_mode="rates", quality_layers=[80, 40, 20, 10, 5],
                                                               progression="LRCP")
    data = buf.getvalue()
    job = jobs.build_iso_job_from_codestream(data)
    assert job["layers"] == 5 and not job["reversible"]
    want = O.iso_decode_job(job, threads=THREADS)
    assert np.array_equal(want.reshape(2160, 3840, 4)[:, :, :3], opj(data))
    assert np.array_equal(gpu(j2k, gpu_ctx, job, ISO, job["coef_bits"]), want)


# ---- cfg4: 8192x8192 grey 16-bit lossless HTJ2K, 1024x1024 tiles ----------------------------------------------------------------
def test_cfg4_ref_full_size(j2k, gpu_ctx):
    s = jobs.synth_image_fast(8192, 8192, 1, 16, seed=1004)
    job = jobs.build_ref_job(s, 16, 1024, 1024, nlevels=5, reversible=True, ht=True, threads=THREADS)
    assert np.array_equal(gpu(j2k, gpu_ctx, job), ref_oracle(job))


def _cfg4_tile(t):
    s = jobs.synth_image_fast(1024, 1024, 1, 16, seed=4004 + t)
    j = jobs.build_iso_job(s, 16, None, None, 5)
    j.pop("codestream", None)
    return j


def cfg4_iso_tiles():
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        return pool.map(_cfg4_tile, range(64), chunksize=1)


def assemble_tiles(tiles, which):
    """one job table for the tiles `which` of the 8 x 8 grid (tile t lies at column t % 8, row t // 8)"""
    tcs, cbs, blobs, boff, ntc = [], [], [], 0, 0
    for t in which:
        j = tiles[t]
        tc = j["tilecomps"].copy()
        tc["x0"] += (t % 8) * 1024; tc["x1"] += (t % 8) * 1024; tc["y0"] += (t // 8) * 1024; tc["y1"] += (t // 8) * 1024
        cb = j["cblks"].copy()
        cb["tilecomp"] += ntc
        cb["data_off"] += boff
        ntc += len(tc)
        boff += j["blob"].size
        tcs.append(tc); cbs.append(cb); blobs.append(j["blob"])
    return dict(width=8192, height=8192, ncomp=1, prec=16, sgnd=0, mct=0, reversible=1, nlevels=5, ht=1, mode=1,
                tilecomps=np.concatenate(tcs), cblks=np.concatenate(cbs), blob=np.concatenate(blobs),
                coef_bits=max(tiles[t]["coef_bits"] for t in which))


def test_cfg4_iso_full_size_and_tile_sharding(j2k, gpu_ctx):
    """ISO mode, the whole image in one call; then the SAME image as two disjoint tile subsets decoded by two contexts into
    ONE host buffer (J2KGPU_ITEM_TILES_ONLY: only the owned rectangles are written) -- large-image tile sharding"""
    tiles = cfg4_iso_tiles()
    src = np.zeros((8192, 8192), np.uint16)
    for t, j in enumerate(tiles):
        src[(t // 8) * 1024:(t // 8 + 1) * 1024, (t % 8) * 1024:(t % 8 + 1) * 1024] = j["samples"][0]
    whole = assemble_tiles(tiles, range(64))
    got = gpu(j2k, gpu_ctx, whole, ISO)
    val = got.reshape(8192, 8192, 2)
    assert np.array_equal((val[:, :, 0].astype(np.uint16) << 8) | val[:, :, 1], src)
    assert np.array_equal(got, O.iso_decode_job(whole, threads=THREADS))
    # two contexts, tiles dealt by compressed size (the package's planner), one shared output buffer pre-filled with a marker
    from go_jpeg2000_b200 import shard
    plan = shard.shard_units([j["blob"].size for j in tiles], 2)
    out = np.full(8192 * 8192 * 2, 0xAB, np.uint8)
    ctx2 = j2k.Context(0)
    try:
        for ctx, which in ((gpu_ctx, plan[0]), (ctx2, plan[1])):
            part = assemble_tiles(tiles, which)
            tcs, cbs = jobs.as_ctypes(part["tilecomps"], j2k.TileComp), jobs.as_ctypes(part["cblks"], j2k.CBlk)
            img = j2k.make_image(8192, 8192, 1, 16, nlevels=5, ht=1, mode=ISO)
            item = j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), part["blob"].ctypes.data_as(j2k.u8p), part["blob"].size,
                                 out.ctypes.data_as(j2k.u8p), 8192 * 2, j2k.ITEM_TILES_ONLY, 0)
            ctx.decode_batch([item])
            if ctx is gpu_ctx:                                   # after the first half: the other tiles still hold the marker
                t_other = plan[1][0]
                blk = out.reshape(8192, 8192, 2)[(t_other // 8) * 1024:(t_other // 8 + 1) * 1024, (t_other % 8) * 1024:(t_other % 8 + 1) * 1024]
                assert (blk == 0xAB).all()
    finally:
        ctx2.close()
    assert np.array_equal(out, got)


# ---- cfg5: 1920x1080 RGB 8-bit HTJ2K lossy 9-7 frames -----------------------------------------------------------------------------
def test_cfg5_ref_full_size_batch(j2k, gpu_ctx):
    """REF semantics: float64 9-7 + ICT + the reference's HT coder, a batch of 4 frames through j2kgpu_decode_batch"""
    jl = [jobs.build_ref_job(jobs.synth_image_fast(1920, 1080, 3, 8, seed=1005 + i), 8, None, None, nlevels=5, reversible=False,
                             ht=True, threads=THREADS) for i in range(4)]
    keep, items, outs = [], [], []
    for j in jl:
        tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
        blob = np.ascontiguousarray(j["blob"])
        out = np.zeros(1920 * 1080 * 4, np.uint8)
        keep += [tcs, cbs, blob]
        outs.append(out)
        img = j2k.make_image(1920, 1080, 3, 8, reversible=0, nlevels=5, ht=1)
        items.append(j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                                   out.ctypes.data_as(j2k.u8p), 1920 * 4, 0, 0))
    gpu_ctx.decode_batch(items)
    for o, j in zip(outs, jl):
        assert np.array_equal(o, ref_oracle(j))


def test_cfg5_iso_lossy_htj2k_full_size(j2k, gpu_ctx):
    """ISO mode: irreversible 9-7 + ICT + dead-zone quantiser + HT blocks, 1080p: GPU == CPU checker == OpenJPEG (max |d| = 0,
    where north_star allows 1 LSB)"""
    from datagen import codestream as cs
    s = jobs.synth_image_fast(1920, 1080, 3, 8, seed=5005)
    data, _ = cs.write_htj2k(s, 8, None, None, 5, lossy_step=1.0)
    job = jobs.build_iso_job_from_codestream(data)
    assert job["ht"] and not job["reversible"]
    want = O.iso_decode_job(job, threads=THREADS)
    ref = opj(data)
    assert np.array_equal(want.reshape(1080, 1920, 4)[:, :, :3], ref)
    assert np.array_equal(gpu(j2k, gpu_ctx, job, ISO, job["coef_bits"]), want)
    assert 10 * np.log10(255.0 ** 2 / np.mean((ref.astype(np.float64) - np.moveaxis(s, 0, 2)) ** 2)) > 35


def test_cfg3_iso_12bit_lossy_grey_layers(j2k, gpu_ctx):
    """ISO mode at more than 8 bits with quality layers: a 16-bit-container grey image (12-bit range) written by OpenJPEG with
    the irreversible 9-7 and 5 layers; compared with OpenJPEG's decode within the 1 LSB north_star allows (measured: exact)"""
    Image = pytest.importorskip("PIL.Image")
    s = jobs.synth_image_fast(2048, 1080, 1, 12, seed=3012)
    buf = io.BytesIO()
    Image.fromarray(s[0].astype(np.uint16)).save(buf, format="JPEG2000", no_jp2=True, irreversible=True, num_resolutions=6,
                                                 quality_mode="rates", quality_layers=[60, 30, 15, 8, 4])
    data = buf.getvalue()
    job = jobs.build_iso_job_from_codestream(data)
    assert job["layers"] == 5 and job["prec"] == 16
    got = gpu(j2k, gpu_ctx, job, ISO, job["coef_bits"]).reshape(1080, 2048, 2)
    val = (got[:, :, 0].astype(np.int64) << 8) | got[:, :, 1]
    ref = opj(data).astype(np.int64)
    assert np.abs(val - ref).max() <= 1
    assert np.array_equal(gpu(j2k, gpu_ctx, job, ISO, job["coef_bits"]), O.iso_decode_job(job, threads=THREADS))
