#!/usr/bin/env python3
"""Side measurements of the other BASELINE.json configs (the bench.py line is configs[1]): each config is built in
REF semantics with datagen/, decoded on cuda:0 through the C ABI (device-resident job), CHECKED bit-exactly against the
CPU oracle on the full frame, and timed with CUDA events.  One JSON object per config on stdout.

    python tools/bench_configs.py [cfg1 cfg3 cfg4 cfg5 ...] [--frames N]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (W, H, ncomp, prec, tile, levels, reversible, ht, frames, note)
    "cfg1": (512, 512, 3, 8, None, 5, 1, 0, 8, "512x512 RGB 8-bit lossless 5-3, 1 tile, 64x64 blocks, EBCOT"),
    "cfg2": (3840, 2160, 3, 8, 512, 5, 1, 1, 2, "3840x2160 RGB 8-bit lossless, 512x512 tiles, RCT, reference HT coder"),
    "cfg3": (3840, 2160, 3, 12, None, 5, 0, 0, 1, "3840x2160 RGB 12-bit lossy 9-7 EBCOT, ICT, 1 tile (REF: one quality layer)"),
    "cfg4": (8192, 8192, 1, 16, 1024, 5, 1, 1, 1, "8192x8192 grayscale 16-bit lossless, 1024x1024 tiles, reference HT coder"),
    "cfg5": (1920, 1080, 3, 8, None, 5, 0, 1, 32, "1920x1080 RGB 8-bit lossy 9-7 frames, reference HT coder (batch)"),
}


ISO_CONFIGS = {
    # name: (W, H, ncomp, Pillow/OpenJPEG encoder options, frames, note)
    "iso_cfg1": (512, 512, 3, dict(irreversible=False, num_resolutions=6, mct=1), 8,
                 "configs[0] as a real codestream: 512x512 RGB 8-bit lossless 5-3, 1 tile, 64x64 blocks, EBCOT, written by OpenJPEG"),
    "iso_4k_ebcot": (3840, 2160, 3, dict(irreversible=False, num_resolutions=6, mct=1, tile_size=(512, 512)), 2,
                     "3840x2160 RGB 8-bit lossless 5-3, 512x512 tiles, EBCOT, written by OpenJPEG"),
    "iso_4k_lossy": (3840, 2160, 3, dict(irreversible=True, num_resolutions=6, mct=1, quality_mode="rates", quality_layers=[80, 40, 20, 10, 5]), 2,
                     "3840x2160 RGB 8-bit lossy 9-7 EBCOT, ICT, 5 quality layers, LRCP, 1 tile, written by OpenJPEG (configs[2] at 8 bits)"),
    # code-block styles: OpenJPEG's encoder through its C API (datagen/opj_direct.py), tables from the product's tier-2
    "iso_4k_bypass": (3840, 2160, 3, dict(opj_mode=0x01, num_resolutions=6, mct=1, tile=(512, 512)), 2,
                      "iso_4k_ebcot written with selective arithmetic-coding bypass (raw significance / refinement passes from the 5th bit-plane on)"),
    "iso_4k_all_styles": (3840, 2160, 3, dict(opj_mode=0x3F, num_resolutions=6, mct=1, tile=(512, 512)), 2,
                          "iso_4k_ebcot written with all six code-block styles (BYPASS RESET TERMALL VCAUSAL PREDTERM SEGSYM)"),
    # written by datagen.codestream.write_htj2k (OpenJPEG 2.5 decodes HTJ2K but does not write it)
    "iso_cfg5": (1920, 1080, 3, dict(htj2k=True, lossy_step=1.0, nlevels=5), 32,
                 "configs[4] as a real codestream: 1920x1080 RGB 8-bit lossy 9-7 HTJ2K, ICT, 1 tile, 64x64 blocks (batch of frames)"),
    "iso_cfg2": (3840, 2160, 3, dict(htj2k=True, lossy_step=None, nlevels=5, tile=512), 2,
                 "configs[1] as a real codestream: 3840x2160 RGB 8-bit lossless HTJ2K, RCT, 512x512 tiles"),
}


def run_iso(name, args, j2k, ctx, stream):
    """a codestream written by OpenJPEG (through Pillow) -> harness tier-2 -> GPU decode in J2KGPU_MODE_ISO, checked
    against OpenJPEG's own decode of the same bytes"""
    import io
    import torch
    from PIL import Image
    from datagen import jobs
    W, H, nc, kw, F, note = ISO_CONFIGS[name]
    F = args.frames or F
    s = jobs.synth_image(W, H, nc, 8, seed=77)
    buf = io.BytesIO()
    t0 = time.perf_counter()
    parsed = None
    if kw.get("htj2k"):
        from datagen import codestream as cs
        data, _ = cs.write_htj2k(s, 8, kw.get("tile"), kw.get("tile"), kw["nlevels"], lossy_step=kw["lossy_step"])
    elif "opj_mode" in kw:
        from datagen import opj_direct
        data = opj_direct.encode(s, mode=kw["opj_mode"], num_resolutions=kw["num_resolutions"], mct=kw["mct"], tile=kw["tile"])
    else:
        Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8)).save(buf, format="JPEG2000", no_jp2=True, **kw)
        data = buf.getvalue()
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    if "opj_mode" in kw:                                   # the harness's Python tier-2 does not read segmented blocks: product tier-2
        parsed = j2k.Parsed(data)
        tcs_np, cbs_np, blob_np = parsed.tables()
        pim = parsed.image
        job = dict(mct=pim.mct, reversible=pim.reversible, nlevels=pim.nlevels, ht=pim.ht, coef_bits=pim.coef_bits, layers=parsed.info["layers"],
                   tilecomps=tcs_np, cblks=cbs_np, blob=np.concatenate([blob_np, np.zeros(8, np.uint8)]), cblk_style=pim.cblk_style)
    else:
        job = jobs.build_iso_job_from_codestream(data)
    t_parse = time.perf_counter() - t0
    t0 = time.perf_counter()
    im = Image.open(io.BytesIO(data))
    im.load()
    t_opj = time.perf_counter() - t0
    ref = np.array(im)
    stride = W * 4
    tcs, cbs = jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk)
    blob = np.ascontiguousarray(job["blob"])
    img = j2k.make_image(W, H, nc, 8, mct=job["mct"], reversible=job["reversible"], nlevels=job["nlevels"], ht=job["ht"], mode=1,
                         coef_bits=job["coef_bits"], cblk_style=job.get("cblk_style", 0))
    outs = [np.zeros(stride * H, np.uint8) for _ in range(F)]
    items = [j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                           o.ctypes.data_as(j2k.u8p), stride) for o in outs]
    J = j2k.Job(ctx, items)
    d_blob = torch.cat([torch.from_numpy(blob)] * F + [torch.zeros(64, dtype=torch.uint8)]).cuda()
    d_out = torch.empty(J.out_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    J.run(d_blob.data_ptr(), d_out.data_ptr())
    stream.synchronize()
    got = d_out[J.out_offset(F - 1): J.out_offset(F - 1) + stride * H].cpu().numpy().reshape(H, W, 4)
    d = np.abs(got[:, :, :nc].astype(np.int64) - ref.astype(np.int64))

    def timeit(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(args.reps):
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    ms_all = timeit(lambda: J.run(d_blob.data_ptr(), d_out.data_ptr()))
    ms_ent = timeit(lambda: J.run_entropy(d_blob.data_ptr()))
    ms_dwt = timeit(lambda: J.run_dwt_mct(d_out.data_ptr()))
    print(json.dumps(dict(config=name, workload=note, frames=F, codestream_bytes=len(data), layers=job["layers"],
                          max_abs_diff_vs_openjpeg=int(d.max()), plan=J.plan, coef_plane_bytes=J.coef_bytes,
                          code_blocks=len(job["cblks"]) * F,
                          ms=dict(whole=round(ms_all, 3), entropy=round(ms_ent, 3), dwt_mct_pack=round(ms_dwt, 3)),
                          mpixel_per_s=round(W * H * F / 1e3 / ms_all, 1),
                          openjpeg_cpu_mpixel_per_s=round(W * H / 1e6 / t_opj, 2),
                          harness_s=dict(encode=round(t_enc, 2), python_tier2_parse=round(t_parse, 2)))), flush=True)
    J.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["cfg1", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import oracle_lib as O
    from datagen import jobs
    from __graft_entry__ import load_package
    j2k = load_package()
    ctx = j2k.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    threads = os.cpu_count() or 1
    for name in args.configs:
        if name in ISO_CONFIGS:
            run_iso(name, args, j2k, ctx, stream)
            continue
        W, H, nc, prec, tile, lv, rev, ht, F, note = CONFIGS[name]
        F = args.frames or F
        t0 = time.perf_counter()
        s = jobs.synth_image(W, H, nc, prec, seed=1000 + int(name[3:]))
        job = jobs.build_ref_job(s, prec, tile, tile, nlevels=lv, reversible=bool(rev), ht=bool(ht), threads=threads)
        t_gen = time.perf_counter() - t0
        bpp = j2k.fmt_bpp(nc, prec)
        stride = W * bpp
        tcs, cbs = jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk)
        blob = np.ascontiguousarray(job["blob"])
        img = j2k.make_image(W, H, nc, prec, mct=job["mct"], reversible=rev, nlevels=lv, ht=ht)
        outs = [np.zeros(stride * H, np.uint8) for _ in range(F)]
        items = [j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                               o.ctypes.data_as(j2k.u8p), stride) for o in outs]
        J = j2k.Job(ctx, items)
        d_blob = torch.cat([torch.from_numpy(blob)] * F + [torch.zeros(64, dtype=torch.uint8)]).cuda()
        d_out = torch.empty(J.out_bytes, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        J.run(d_blob.data_ptr(), d_out.data_ptr())
        stream.synchronize()
        got = d_out[: stride * H].cpu().numpy()
        # the checker: the oracle's whole REF path on the same job, full frame
        oimg = O.Image()
        oimg.width, oimg.height, oimg.ncomp = W, H, nc
        for c in range(nc):
            oimg.prec[c], oimg.sgnd[c] = prec, 0
        oimg.mct, oimg.reversible, oimg.nlevels, oimg.ht = job["mct"], rev, lv, ht
        t0 = time.perf_counter()
        want = O.decode_image(oimg, jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk),
                              job["blob"], stride, stride * H, threads=threads)
        t_cpu = time.perf_counter() - t0
        exact = bool(np.array_equal(got, want))
        last = d_out[J.out_offset(F - 1): J.out_offset(F - 1) + stride * H].cpu().numpy()
        exact = exact and bool(np.array_equal(last, want))

        def timeit(fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(args.reps):
                a.record(stream)
                fn()
                b.record(stream)
                b.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.median(ts))

        for _ in range(2):
            J.run(d_blob.data_ptr(), d_out.data_ptr())
        ms_all = timeit(lambda: J.run(d_blob.data_ptr(), d_out.data_ptr()))
        ms_ent = timeit(lambda: J.run_entropy(d_blob.data_ptr()))
        ms_dwt = timeit(lambda: J.run_dwt_mct(d_out.data_ptr()))
        alg = (4 * W * H * nc + W * H * bpp) * F
        print(json.dumps(dict(config=name, workload=note, frames=F, bit_exact_vs_oracle=exact, plan=J.plan,
                              coef_plane_bytes=J.coef_bytes, code_blocks=len(job["cblks"]) * F,
                              ms=dict(whole=round(ms_all, 3), entropy=round(ms_ent, 3), dwt_mct_pack=round(ms_dwt, 3)),
                              mpixel_per_s=round(W * H * F / 1e3 / ms_all, 1),
                              dwt_mct_gbs=round(alg / ms_dwt / 1e6, 1), dwt_mct_frac_of_hbm_peak=round(alg / ms_dwt / 1e6 / 6537.6, 3),
                              cpu_oracle_mpixel_per_s=round(W * H / 1e6 / t_cpu, 1), cpu_threads=threads,
                              datagen_s=round(t_gen, 1))), flush=True)
        J.close()
        del d_blob, d_out
        torch.cuda.empty_cache()
    ctx.set_stream(0)
    ctx.close()


if __name__ == "__main__":
    main()
