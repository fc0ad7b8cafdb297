"""GPU parity, whole path through j2kgpu_decode / j2kgpu_decode_batch / j2kgpu_job_* (C ABI): pixels identical
to the oracle's whole REF path (orc_decode_image) on the same job, and -- lossless -- identical to the source
image (encode -> decode round trip, the size-independent property used at full BASELINE sizes)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from datagen import jobs

pytestmark = pytest.mark.gpu


def hdr(j2k, job):
    return j2k.make_image(job["width"], job["height"], job["ncomp"], job["prec"], sgnd=job["sgnd"], mct=job["mct"],
                          reversible=job["reversible"], nlevels=job["nlevels"], ht=job["ht"])


def oracle_pixels(job):
    img = O.Image()
    img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
    for c in range(job["ncomp"]):
        img.prec[c], img.sgnd[c] = job["prec"], job["sgnd"]
    img.mct, img.reversible, img.nlevels, img.ht = job["mct"], job["reversible"], job["nlevels"], job["ht"]
    bpp = (1 if job["prec"] <= 8 else 2) if job["ncomp"] == 1 else (4 if job["prec"] <= 8 else 8)
    stride = job["width"] * bpp
    return O.decode_image(img, jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk),
                          job["blob"], stride, stride * job["height"], threads=4)


def gpu_pixels(j2k, ctx, job):
    return ctx.decode_tiles(hdr(j2k, job), jobs.as_ctypes(job["tilecomps"], j2k.TileComp),
                            jobs.as_ctypes(job["cblks"], j2k.CBlk), job["blob"])


CASES = [  # w, h, ncomp, prec, tile_w, tile_h, levels, reversible, ht
    (96, 80, 3, 8, None, None, 3, 1, 0),
    (130, 70, 3, 8, 64, 64, 2, 1, 0),          # ragged last tile column / row
    (64, 64, 1, 8, None, None, 5, 1, 0),
    (100, 60, 1, 12, 48, 32, 2, 1, 0),         # Gray16
    (70, 50, 3, 16, None, None, 3, 1, 0),      # RGBA64 + overflow quirk
    (96, 64, 3, 8, None, None, 3, 0, 0),       # 9-7 + ICT
    (200, 120, 3, 12, 128, 128, 4, 0, 0),      # 12-bit lossy, tiles
    (128, 128, 3, 8, 64, 64, 3, 1, 1),         # reference "HT" coder
    (96, 96, 1, 8, None, None, 0, 1, 0),       # zero decomposition levels
    (512, 512, 3, 8, None, None, 5, 1, 0),     # BASELINE cfg1: 512x512 RGB 8-bit lossless 5-3, 1 tile, 64x64 blocks
    (256, 128, 1, 8, 128, 128, 3, 1, 0),       # streaming kernel, 1 component (Gray8)
    (256, 192, 4, 8, None, None, 4, 1, 0),     # streaming kernel, 4 components (RGBA with coded alpha)
    (384, 256, 3, 12, 128, 128, 3, 1, 0),      # streaming kernel, RGBA64 epilogue
    (264, 136, 3, 8, 128, 128, 2, 1, 0),       # 8-wide / 8-high edge tiles (one active quad pair per warp)
    (1024, 192, 3, 8, None, None, 2, 1, 0),    # several warps across one tile row (30-quad ranges + halo lanes)
    (480, 270, 3, 8, None, None, 5, 0, 1),     # cfg5 geometry / 4: reference HT coder + 9-7; its magnitudes overflow
                                               # int32(float64): amd64 gives 0x80000000, not saturation
    (512, 256, 1, 16, 256, 256, 4, 1, 1),      # cfg4 geometry / 16: Gray16, fast grayscale epilogue of the fused kernel
    (256, 128, 1, 8, None, None, 3, 1, 1),     # Gray8 fast epilogue, int32 planes
    (512, 256, 3, 12, 256, 256, 4, 0, 0),      # streaming 9-7 (float64) on every level, ICT, RGBA64
    (1024, 96, 3, 8, None, None, 3, 0, 0),     # streaming 9-7, several warps per row, RGBA8 fast store
    (384, 192, 1, 8, 192, 192, 2, 0, 0),       # streaming 9-7, one component
    (256, 128, 4, 8, None, None, 3, 0, 1),     # streaming 9-7, four components, reference HT coder
]


@pytest.mark.parametrize("w,h,ncomp,prec,tw,th,levels,rev,ht", CASES)
def test_whole_path_matches_oracle(j2k, gpu_ctx, w, h, ncomp, prec, tw, th, levels, rev, ht):
    s = jobs.synth_image(w, h, ncomp, prec, seed=1000 + w + h)
    job = jobs.build_ref_job(s, prec, tw, th, nlevels=levels, reversible=bool(rev), ht=bool(ht), threads=4)
    got = gpu_pixels(j2k, gpu_ctx, job)
    assert np.array_equal(got, oracle_pixels(job))
    if rev and not ht and prec == 8:                     # lossless round trip: decode(encode(x)) == x
        pix = got.reshape(h, w, -1)
        for c in range(ncomp):
            assert np.array_equal(pix[:, :, c], s[c].astype(np.uint8))


def test_uncoded_blocks_and_partial_coverage(j2k, gpu_ctx):
    """tcd.go:394-396: len(Data)==0 -> block stays zero; planes not covered by any block stay zero"""
    s = jobs.synth_image(128, 64, 3, 8, seed=5)
    job = jobs.build_ref_job(s, 8, nlevels=2, reversible=True, threads=2)
    cb = job["cblks"].copy()
    cb["data_len"][::3] = 0
    job["cblks"] = cb[np.arange(len(cb)) % 5 != 4]        # drop every fifth block entirely
    assert np.array_equal(gpu_pixels(j2k, gpu_ctx, job), oracle_pixels(job))


def test_batch_and_job_api(j2k, gpu_ctx):
    """j2kgpu_decode_batch and the device-resident j2kgpu_job_run agree with per-image decode"""
    import torch
    jl = [jobs.build_ref_job(jobs.synth_image(96, 64, 3, 8, seed=40 + i), 8, 64, 64, nlevels=3, reversible=True, threads=2)
          for i in range(3)]
    want = [oracle_pixels(j) for j in jl]
    keep, items, outs = [], [], []
    for j in jl:
        tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
        blob = np.ascontiguousarray(j["blob"])
        out = np.zeros(96 * 64 * 4, np.uint8)
        keep += [tcs, cbs, blob]
        outs.append(out)
        items.append(j2k.BatchItem(hdr(j2k, j), tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                                   out.ctypes.data_as(j2k.u8p), 96 * 4))
    gpu_ctx.decode_batch(items)
    for o, w_ in zip(outs, want):
        assert np.array_equal(o, w_)
    # device-resident: blobs back to back in HBM, pixels stay in HBM, torch owns memory and stream
    job = j2k.Job(gpu_ctx, items)
    d_blob = torch.from_numpy(np.concatenate([j["blob"] for j in jl])).cuda()
    d_out = torch.zeros(job.out_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    gpu_ctx.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()                              # uploads above ran on torch's default stream
    n0 = gpu_ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    job.run(d_blob.data_ptr(), d_out.data_ptr())
    ev1.record(stream)
    stream.synchronize()
    assert ev0.elapsed_time(ev1) > 0.02                   # the kernels really ran on the caller's stream
    assert job.fused_levels == 2 and job.coef_bytes == 2  # 64- and 32-wide tiles fit the fused kernel; 8-bit EBCOT fits int16
    assert gpu_ctx.launches - n0 == 3                     # 1 entropy + level 2 + fused (levels 1, 0, MCT, DC, pack)
    host = d_out.cpu().numpy()
    for i, w_ in enumerate(want):
        assert np.array_equal(host[job.out_offset(i): job.out_offset(i) + w_.size], w_)
    for o in outs:
        o[:] = 0
    job.run_host()
    for o, w_ in zip(outs, want):
        assert np.array_equal(o, w_)
    gpu_ctx.set_stream(0)
    job.close()


def _run_items(j2k, ctx, jl, env=None, mode=0, coef_bits=0):
    """decode a list of jobs as ONE batch through j2kgpu_job_create + j2kgpu_job_run_host (the pipelined host path)"""
    import os
    keep, items, outs = [], [], []
    for j in jl:
        tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
        blob = np.ascontiguousarray(j["blob"])
        bpp = j2k.fmt_bpp(j["ncomp"], j["prec"])
        out = np.zeros(j["width"] * j["height"] * bpp, np.uint8)
        keep += [tcs, cbs, blob]
        outs.append(out)
        img = j2k.make_image(j["width"], j["height"], j["ncomp"], j["prec"], sgnd=j["sgnd"], mct=j["mct"],
                             reversible=j["reversible"], nlevels=j["nlevels"], ht=j["ht"], mode=mode, coef_bits=coef_bits)
        items.append(j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                                   out.ctypes.data_as(j2k.u8p), j["width"] * bpp))
    with ctx.options(**{k[len("J2KGPU_"):].lower(): v for k, v in (env or {}).items()}):
        job = j2k.Job(ctx, items)
        flags = (job.fused_levels, job.coef_bytes, job.plan)
        job.run_host()
        job.close()
    return outs, flags


@pytest.mark.parametrize("w,h,ncomp,prec,tw,th,levels,ht", [
    (512, 256, 3, 8, 256, 256, 5, 0),      # EBCOT: int16 planes + fused kernel
    (512, 256, 3, 8, 256, 256, 5, 1),      # reference HT coder: no magnitude bound -> int32 planes + fused kernel
    (1024, 128, 3, 8, None, None, 2, 1),   # several warps per row (28-quad ranges + 2 halo lanes each side), coarsest = level 1
    (256, 64, 3, 8, None, None, 1, 0),     # a single decomposition level (no level 1 inside the fused kernel)
    (192, 96, 1, 8, 96, 96, 3, 0),         # 1 component, Gray8
    (160, 72, 4, 8, None, None, 3, 0),     # 4 components; 72 rows: strips shorter than the ring depth at the bottom
    (384, 256, 3, 12, 128, 128, 4, 0),     # RGBA64 epilogue
    (1024, 64, 3, 8, None, None, 3, 1),    # wide kernel: 64 lanes' worth of columns -> 30 owned lanes + 1 halo lane each side
    (1536, 40, 3, 8, None, None, 2, 0),    # wide kernel, three warps per row, int16 planes, short strip (20 row pairs)
    (528, 136, 3, 8, None, None, 4, 0),    # wide kernel: 33 lanes' worth -> halo path with a nearly empty second warp
    (512, 512, 3, 8, None, None, 5, 1),    # one warp per tile row, no halo lanes (the bench geometry)
    (328, 140, 3, 8, None, None, 5, 0),    # odd level sizes above the fused pair (82x35, 41x18, 21x9)
    (488, 508, 1, 8, None, None, 4, 1),    # level-2 image 122x127
])
def test_fused_and_int16_variants_agree(j2k, gpu_ctx, w, h, ncomp, prec, tw, th, levels, ht):
    """the fused levels-1+0 kernel and the int16 coefficient planes are optimisations: every combination of
    {fused, per-level} x {int16, int32} gives the oracle's pixels"""
    s = jobs.synth_image(w, h, ncomp, prec, seed=77 + w)
    job = jobs.build_ref_job(s, prec, tw, th, nlevels=levels, reversible=True, ht=bool(ht), threads=4)
    want = oracle_pixels(job)
    seen, plans = set(), set()
    for env in ({}, {"J2KGPU_NO_FUSE": "1"}, {"J2KGPU_COEF32": "1"}, {"J2KGPU_NO_FUSE": "1", "J2KGPU_COEF32": "1"},
                {"J2KGPU_NO_WIDE": "1"}, {"J2KGPU_NO_WIDE": "1", "J2KGPU_COEF32": "1"},
                {"J2KGPU_NO_WIDE": "1", "J2KGPU_NO_FAST_EPI": "1"}):
        outs, flags = _run_items(j2k, gpu_ctx, [job], env)
        assert np.array_equal(outs[0], want), env
        seen.add(flags[:2])
        plans.add(flags[2])
    if ncomp == 3 and prec == 8 and w % 16 == 0 and (tw or w) % 16 == 0:
        # the 16-columns-per-lane kernel, the 4-columns-per-lane kernel with the fast and with the generic epilogue
        assert {p & 7 for p in plans} >= {7, 3, 1, 0}
    assert (2, 4) in seen and (1, 4) in seen
    if not ht:
        assert (2, 2) in seen and (1, 2) in seen      # EBCOT magnitudes are bounded by num_bps <= 15
    else:
        assert (2, 2) not in seen                     # ht.go:664-684 has no magnitude bound: never int16


def test_pipelined_host_run_many_items(j2k, gpu_ctx):
    """j2kgpu_job_run_host cuts a batch of >= 16 items into chunks (copy-in / kernels / copy-out on three streams)"""
    jl = []
    for i in range(19):
        w, h = (64, 32) if i % 3 else (96, 64)
        jl.append(jobs.build_ref_job(jobs.synth_image(w, h, 3, 8, seed=300 + i), 8, 32, 32, nlevels=2, reversible=True,
                                     ht=bool(i % 2 == 0) and False, threads=2))
    outs, _ = _run_items(j2k, gpu_ctx, jl)
    for o, j in zip(outs, jl):
        assert np.array_equal(o, oracle_pixels(j))


@pytest.mark.parametrize("ht", [0, 1])
def test_host_run_is_independent_of_the_chunk_plan(j2k, gpu_ctx, ht):
    """the pipelined host-buffer run gives the same pixels whatever its chunk plan (planner default, one chunk, one item
    per chunk, uneven chunks): the HT scratch tables are indexed by job-wide block numbers, the arenas by item"""
    import os
    jl = [jobs.build_ref_job(jobs.synth_image(96, 64, 3, 8, seed=900 + i), 8, 32, 32, nlevels=2, reversible=True,
                             ht=bool(ht), threads=2) for i in range(7)]
    want = [oracle_pixels(j) for j in jl]
    for plan in ("", "7", "1,1,1,1,1,1,1", "2,4,1", "3"):
        outs, _ = _run_items(j2k, gpu_ctx, jl, {"J2KGPU_CHUNKS": plan})
        for o, wnt in zip(outs, want):
            assert np.array_equal(o, wnt), plan


@pytest.mark.parametrize("w,h,prec,tw,levels,rev,ht,cs", [
    (96, 80, 8, None, 3, 1, 0, 1),        # fused kernels, generic epilogue (the fast RGBA8 one has no conversion)
    (512, 256, 8, 256, 4, 1, 1, 2),       # would take the 16-columns-per-lane kernel without the conversion
    (200, 120, 12, 128, 4, 0, 0, 1),      # 9-7, RGBA64
    (130, 70, 8, 64, 2, 1, 0, 2),         # ragged tiles, tiled per-level path
    (96, 80, 8, None, 3, 1, 0, 7),        # CIELab (pow-based conversions: 1 LSB tolerance, see include/j2kgpu.h)
    (512, 256, 8, 256, 4, 1, 1, 9),       # e-sRGB
    (200, 120, 12, 128, 4, 0, 0, 10),     # ROMM-RGB, 9-7, RGBA64
    (130, 70, 8, 64, 2, 1, 0, 8),         # CIEJab
])
def test_whole_path_with_colour_conversion(j2k, gpu_ctx, w, h, prec, tw, levels, rev, ht, cs):
    """j2k_image_t.colorspace: sYCC / YCbCr images come out as sRGB exactly as decoder.go:350-356 + colorspace.go would"""
    s = jobs.synth_image(w, h, 3, prec, seed=7 * w + cs)
    job = jobs.build_ref_job(s, prec, tw, tw, nlevels=levels, reversible=bool(rev), ht=bool(ht), threads=2)
    img = O.Image()
    img.width, img.height, img.ncomp = w, h, 3
    for c in range(3):
        img.prec[c], img.sgnd[c] = prec, 0
    img.mct, img.reversible, img.nlevels, img.ht, img.colorspace = job["mct"], rev, levels, ht, cs
    bpp = 4 if prec <= 8 else 8
    want = O.decode_image(img, jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk),
                          job["blob"], w * bpp, w * bpp * h, threads=2)
    gimg = j2k.make_image(w, h, 3, prec, mct=job["mct"], reversible=rev, nlevels=levels, ht=ht, colorspace=cs)
    got = gpu_ctx.decode_tiles(gimg, jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk), job["blob"])
    if cs >= 7:                                                      # CUDA's pow() where Go has math.Pow
        un = lambda p: p.astype(np.int32) if prec <= 8 else (p[0::2].astype(np.int32) << 8) | p[1::2]
        assert np.abs(un(got) - un(want)).max() <= 1
    else:
        assert np.array_equal(got, want)
    plain = gpu_ctx.decode_tiles(j2k.make_image(w, h, 3, prec, mct=job["mct"], reversible=rev, nlevels=levels, ht=ht),
                                 jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk), job["blob"])
    assert not np.array_equal(plain, got)


def test_page_locked_host_buffers(j2k, gpu_ctx):
    """j2kgpu_host_alloc / j2kgpu_host_register (ABI v3): decoding from a page-locked blob into page-locked pixels gives
    the pixels of the pageable call; the calls fail as values on bad arguments"""
    j = jobs.build_ref_job(jobs.synth_image(96, 64, 3, 8, seed=41), 8, 32, 32, nlevels=2, reversible=True, ht=True, threads=2)
    want = oracle_pixels(j)
    tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
    blob = np.ascontiguousarray(j["blob"])
    img = j2k.make_image(96, 64, 3, 8, mct=j["mct"], reversible=1, nlevels=2, ht=1)
    pin_blob = gpu_ctx.host_alloc(blob.size)                         # allocated by the library
    pin_blob[:] = blob
    out = np.zeros(96 * 64 * 4, np.uint8)
    gpu_ctx.host_register(out)                                       # owned by the caller, page-locked in place
    try:
        item = j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), pin_blob.ctypes.data_as(j2k.u8p), blob.size,
                             out.ctypes.data_as(j2k.u8p), 96 * 4)
        job = j2k.Job(gpu_ctx, [item])
        job.run_host()
        job.close()
        assert np.array_equal(out, want)
    finally:
        gpu_ctx.host_unregister(out)
        gpu_ctx.host_free(pin_blob)
    L = j2k.lib()
    import ctypes as C
    assert L.j2kgpu_host_alloc(None, C.c_uint64(16), C.byref(C.c_void_p())) == j2k.E_ARG
    assert L.j2kgpu_host_register(gpu_ctx._h, None, C.c_uint64(16)) == j2k.E_ARG
    assert L.j2kgpu_host_free(gpu_ctx._h, None) == 0


def test_path_argument_errors(j2k, gpu_ctx):
    s = jobs.synth_image(64, 64, 3, 8, seed=9)
    job = jobs.build_ref_job(s, 8, nlevels=2, reversible=True, threads=2)
    tcs, cbs = jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk)
    bad = hdr(j2k, job)
    bad.ncomp = 2
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_tiles(bad, tcs, cbs, job["blob"])
    assert e.value.code == j2k.E_UNSUPPORTED              # decoder.go:585-586
    cb2 = job["cblks"].copy()
    cb2["data_off"][1] = len(job["blob"]) + 10
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_tiles(hdr(j2k, job), tcs, jobs.as_ctypes(cb2, j2k.CBlk), job["blob"])
    assert e.value.code == j2k.E_RANGE
    cb3 = job["cblks"].copy()
    cb3["x0"][0] = 60
    with pytest.raises(j2k.J2KError) as e:
        gpu_ctx.decode_tiles(hdr(j2k, job), tcs, jobs.as_ctypes(cb3, j2k.CBlk), job["blob"])
    assert e.value.code == j2k.E_RANGE
    # garbage bytes never fault (fuzz contract): corrupt the blob and compare with the oracle
    blob = job["blob"].copy()
    blob[::7] ^= 0xA5
    job["blob"] = blob
    assert np.array_equal(gpu_pixels(j2k, gpu_ctx, job), oracle_pixels(job))


def test_padded_stride_and_partial_tile_set(j2k, gpu_ctx):
    """out_stride > width * bpp: the padding bytes belong to the caller and are left alone; tiles that do not cover the image:
    the uncovered pixels hold what the reference's zero-initialised planes decode to (decoder.go:305-309: mid-grey after the DC
    shift), never stale device memory -- also right after a call that left other pixels in the pooled buffers"""
    w, h = 160, 96
    s = jobs.synth_image(w, h, 3, 8, seed=61)
    job = jobs.build_ref_job(s, 8, 64, 32, nlevels=2, reversible=True, threads=2)
    full = oracle_pixels(job).reshape(h, w, 4)
    gpu_pixels(j2k, gpu_ctx, job)                                # fills the pooled device buffers with this image
    # keep only the tiles whose index is not a multiple of 3
    tcs = job["tilecomps"]
    tile_id = (tcs["y0"] // 32) * 3 + tcs["x0"] // 64
    keep_tc = tile_id % 3 != 0
    remap = np.cumsum(keep_tc) - 1
    cb = job["cblks"][keep_tc[job["cblks"]["tilecomp"]]].copy()
    cb["tilecomp"] = remap[cb["tilecomp"]]
    part = dict(job, tilecomps=tcs[keep_tc].copy(), cblks=cb)
    stride = w * 4 + 48
    img = hdr(j2k, part)
    out = gpu_ctx.decode_tiles(img, jobs.as_ctypes(part["tilecomps"], j2k.TileComp), jobs.as_ctypes(part["cblks"], j2k.CBlk),
                               part["blob"], out_stride=stride).reshape(h, stride)
    assert not out[:, w * 4:].any()                              # decode_tiles hands over a zeroed buffer: padding untouched
    pix = out[:, :w * 4].reshape(h, w, 4)
    covered = np.zeros((h, w), bool)
    for t in part["tilecomps"]:
        covered[t["y0"]:t["y1"], t["x0"]:t["x1"]] = True
    assert covered.any() and not covered.all()
    assert np.array_equal(pix[covered], full[covered])
    assert (pix[~covered] == np.array([128, 128, 128, 255], np.uint8)).all()
    # the CPU checker agrees on the whole picture (its planes are zero-initialised like the reference's)
    assert np.array_equal(pix, oracle_pixels(part).reshape(h, w, 4))


def test_overlapping_blocks_are_refused(j2k, gpu_ctx):
    """block areas that add up to the plane while leaving a hole (one block duplicated, one dropped) are not mistaken for
    full coverage: overlapping blocks are an argument error (a block's samples are its own while its bit-planes
    accumulate), and the context keeps working"""
    s = jobs.synth_image(128, 64, 1, 8, seed=62)
    job = jobs.build_ref_job(s, 8, nlevels=1, reversible=True, threads=2)
    want = gpu_pixels(j2k, gpu_ctx, job)
    cb = job["cblks"].copy()
    assert len(cb) == 2 and cb["w"][0] == cb["w"][1]
    cb[1] = cb[0]                                                # block 0 twice, block 1 missing: the areas still sum to w * h
    with pytest.raises(j2k.J2KError) as e:
        gpu_pixels(j2k, gpu_ctx, dict(job, cblks=cb))
    assert e.value.code == j2k.E_ARG and "overlap" in str(e.value)
    hole = dict(job, cblks=job["cblks"][:1].copy())              # block 1 missing, no overlap: the hole reads zero coefficients
    assert np.array_equal(gpu_pixels(j2k, gpu_ctx, hole), oracle_pixels(hole))
    assert np.array_equal(gpu_pixels(j2k, gpu_ctx, job), want)


def test_packed_rgb_transfer_equals_rgba_transfer(j2k, gpu_ctx):
    """host-buffer runs of RGBA8 images move packed R G B over the link and host threads of the library fill in the alpha
    byte (host/rgb_expand.h): same bytes as the plain RGBA transfer, for one image, a batch in several chunks, a padded
    stride, and the codestream front door; padding bytes of the caller's rows stay untouched"""
    for (w, h, tile) in ((256, 128, 128), (512, 192, None)):
        srcs = [jobs.synth_image(w, h, 3, 8, seed=70 + i) for i in range(5)]
        isos = [jobs.build_iso_job(s, 8, tile, tile, 3) for s in srcs]
        for pad in (0, 64):
            stride = w * 4 + pad
            outs = {}
            for mode in (0, 1):
                bufs = [np.full(stride * h, 0xA5, np.uint8) for _ in isos]
                items = []
                keep = []
                for ij, buf in zip(isos, bufs):
                    img = j2k.make_image(w, h, 3, 8, nlevels=3, ht=1, mode=j2k.MODE_ISO, coef_bits=ij["coef_bits"])
                    tcs, cbs = jobs.as_ctypes(ij["tilecomps"], j2k.TileComp), jobs.as_ctypes(ij["cblks"], j2k.CBlk)
                    blob = np.ascontiguousarray(ij["blob"], np.uint8)
                    keep += [tcs, cbs, blob]
                    items.append(j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                                               buf.ctypes.data_as(j2k.u8p), stride, 0, 0))
                with gpu_ctx.options(host_alpha=mode, chunks="1,2,2"):
                    gpu_ctx.decode_batch(items)
                outs[mode] = bufs
            for a, b, s in zip(outs[0], outs[1], srcs):
                assert np.array_equal(a, b)
                px = b.reshape(h, stride)[:, :w * 4].reshape(h, w, 4)
                assert np.array_equal(px[:, :, :3], np.moveaxis(s, 0, 2)) and (px[:, :, 3] == 255).all()
                if pad:
                    assert (b.reshape(h, stride)[:, w * 4:] == 0xA5).all()
    s = jobs.synth_image(256, 128, 3, 8, seed=80)
    data = jobs.build_iso_job(s, 8, 128, 128, 3)["codestream"]
    with gpu_ctx.options(host_alpha=1):
        a = gpu_ctx.decode_codestreams([data, data])
    with gpu_ctx.options(host_alpha=0):
        b = gpu_ctx.decode_codestreams([data, data])
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(a[0].reshape(128, 256, 4)[:, :, :3], np.moveaxis(s, 0, 2))


def test_one_large_image_is_pipelined_by_tile_groups(j2k, gpu_ctx):
    """a single image of many tiles goes through the batch pipeline as groups of tiles (copy-in, decode and copy-out of
    neighbouring groups overlap): same pixels as the one-chunk run, for ISO HT, REF EBCOT and a padded stride"""
    s = jobs.synth_image(1280, 1024, 3, 8, seed=91)
    for kind in ("iso", "ref"):
        job = jobs.build_iso_job(s, 8, 256, 256, 4) if kind == "iso" else jobs.build_ref_job(s, 8, tile_w=256, tile_h=256, nlevels=3, reversible=True, threads=4)
        mode = j2k.MODE_ISO if kind == "iso" else j2k.MODE_REF
        img = j2k.make_image(1280, 1024, 3, 8, nlevels=job["nlevels"], ht=job["ht"], mode=mode, coef_bits=job.get("coef_bits", 0))
        tcs, cbs = jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk)
        for stride in (None, 1280 * 4 + 128):
            with gpu_ctx.options(split_min_mpixel=1):              # (the default threshold is 24 Mpixel)
                split = gpu_ctx.decode_tiles(img, tcs, cbs, job["blob"], out_stride=stride)
            with gpu_ctx.options(chunks="1"):                      # an explicit chunk plan keeps the image in one piece
                whole = gpu_ctx.decode_tiles(img, tcs, cbs, job["blob"], out_stride=stride)
            assert np.array_equal(split, whole)
            st = stride or 1280 * 4
            assert np.array_equal(split.reshape(1024, st)[:, :1280 * 4].reshape(1024, 1280, 4)[:, :, :3], np.moveaxis(s, 0, 2))


def test_two_devices_in_one_process(j2k):
    """a second context on another device of the same process gets its own constant tables (EBCOT contexts, MQ states)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one CUDA device")
    s = jobs.synth_image(96, 80, 3, 8, seed=63)
    job = jobs.build_ref_job(s, 8, nlevels=3, reversible=True, threads=2)
    want = oracle_pixels(job)
    for dev in (1, 0):
        ctx = j2k.Context(dev)
        try:
            assert np.array_equal(gpu_pixels(j2k, ctx, job), want), dev
        finally:
            ctx.close()


def test_closing_a_context_retires_its_jobs(j2k):
    """Context.close() destroys the jobs still alive on it first; their later close / finaliser is then a no-op"""
    j = jobs.build_ref_job(jobs.synth_image(64, 64, 1, 8, seed=64), 8, nlevels=2, reversible=True, threads=1)
    tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
    blob = np.ascontiguousarray(j["blob"])
    out = np.zeros(64 * 64, np.uint8)
    item = j2k.BatchItem(hdr(j2k, j), tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                         out.ctypes.data_as(j2k.u8p), 64)
    ctx = j2k.Context(0)
    job = j2k.Job(ctx, [item])
    job.run_host()
    assert np.array_equal(out, oracle_pixels(j))
    ctx.close()
    assert not job._h
    job.close()
    del job


@pytest.mark.parametrize("mode", ["ref_ebcot", "iso_ht"])
def test_job_run_replays_from_a_cuda_graph(j2k, gpu_ctx, mode):
    """j2kgpu_job_run only enqueues work on the context's stream (no allocation, no synchronisation, no host copy once the
    job exists), so a caller can capture it into a CUDA graph and replay the whole path with one launch"""
    import torch
    s = jobs.synth_image(256, 192, 3, 8, seed=71)
    if mode == "ref_ebcot":
        j = jobs.build_ref_job(s, 8, 128, 64, nlevels=3, reversible=True, threads=2)
        img = hdr(j2k, j)
    else:
        j = jobs.build_iso_job(s, 8, 128, 64, 3, ht_passes=3, ht_plane=1)
        img = j2k.make_image(256, 192, 3, 8, nlevels=3, ht=1, mode=j2k.MODE_ISO, coef_bits=j["coef_bits"])
    tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
    blob = np.ascontiguousarray(j["blob"])
    out = np.zeros(256 * 192 * 4, np.uint8)
    item = j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size, out.ctypes.data_as(j2k.u8p), 256 * 4)
    job = j2k.Job(gpu_ctx, [item])
    try:
        job.run_host()
        want = out.copy()
        assert np.array_equal(want.reshape(192, 256, 4)[:, :, :3], np.moveaxis(s, 0, 2)) or mode == "iso_ht"
        d_blob = torch.from_numpy(blob).cuda()
        d_out = torch.zeros(job.out_bytes, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.Stream()
        gpu_ctx.set_stream(stream.cuda_stream)
        torch.cuda.synchronize()
        job.run(d_blob.data_ptr(), d_out.data_ptr())                 # warm: tables, attributes
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = gpu_ctx.launches
        with torch.cuda.graph(graph, stream=stream):
            job.run(d_blob.data_ptr(), d_out.data_ptr())
        launches = gpu_ctx.launches - n0
        assert launches >= 3
        for _ in range(3):
            d_out.zero_()
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy()[: want.size], want)
    finally:
        gpu_ctx.set_stream(0)
        job.close()
