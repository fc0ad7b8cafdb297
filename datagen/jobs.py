"""datagen.jobs -- builds REF-mode decode jobs (flat tables of include/j2kgpu.h) from synthetic images.

Pipeline = the reference encoder's own (encoder.go:216-281), applied per tile:
DC shift -> RCT (lossless) / ICT with round-half-away (lossy) -> DecomposeMultiLevel53/97 in the
reference's dense-prefix layout -> [9-7: quantiser with step 1/quality] -> code blocks -> T1.Encode /
HTEncoder.Encode.  The reference encoder cannot emit a decodable container (one tile, no
packet headers, bogus block addressing), so the job table is emitted directly:

  * tile-components: the image is cut into tile_w x tile_h tiles (reference geometry, tcd.go:241-295);
  * code blocks: a cb x cb grid over each tile-component plane, placed where they lie in the plane;
    the band type of a block is the level-1 quadrant its origin falls in (LL/HL/LH/HH), which is what
    the reference decoder would be told (tcd.go:301-318) -- the REF DWT layout has no other band geometry.

Tables are numpy structured arrays byte-identical to j2k_image_t / j2k_tilecomp_t / j2k_cblk_t.
"""
import ctypes as C
import os

import numpy as np

from . import (dc_shift_forward, decompose53, decompose97, encode_blocks, fwd_ict, fwd_rct)

CBLK_DT = np.dtype([("data_off", "<u8"), ("data_len", "<u4"), ("tilecomp", "<u4"), ("x0", "<u2"), ("y0", "<u2"),
                    ("w", "<u2"), ("h", "<u2"), ("band", "u1"), ("level", "u1"), ("num_bps", "u1"),
                    ("num_passes", "u1"), ("step", "<f4"), ("len_cleanup", "<u4"), ("rsv", "<u4")], align=True)
TILECOMP_DT = np.dtype([("comp", "<u4"), ("x0", "<u4"), ("y0", "<u4"), ("x1", "<u4"), ("y1", "<u4"),
                        ("coeff_off", "<u8")], align=True)
assert CBLK_DT.itemsize == 40 and TILECOMP_DT.itemsize == 32


def synth_image(width, height, ncomp, prec, seed):
    """three low-frequency cosines per channel at 60 % of range + N(0, 2 % of range), clipped."""
    rng = np.random.default_rng(seed)
    maxv = (1 << prec) - 1
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    out = np.empty((ncomp, height, width), np.int32)
    for c in range(ncomp):
        acc = np.zeros((height, width), np.float32)
        for _ in range(3):
            fx, fy = rng.uniform(0.5, 4.0, 2) * 2 * np.pi / max(width, height)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.cos(fx * xx + fy * yy + ph).astype(np.float32)
        img = (0.5 + 0.1 * acc) * maxv * 0.6 / 0.8 + rng.normal(0, 0.02 * maxv, (height, width)).astype(np.float32)
        out[c] = np.clip(np.rint(img), 0, maxv).astype(np.int32)
    return out


def synth_image_fast(width, height, ncomp, prec, seed):
    """the same kind of content as synth_image (three low-frequency cosines per channel + 2 % noise), about 5x cheaper:
    each cosine is built from two outer products (cos(a + b) = cos a cos b - sin a sin b).  Used where many large frames
    are needed (bench.py, full-size tests); NOT bit-identical to synth_image for the same seed."""
    rng = np.random.default_rng(seed)
    maxv = (1 << prec) - 1
    x = np.arange(width, dtype=np.float32)
    y = np.arange(height, dtype=np.float32)
    out = np.empty((ncomp, height, width), np.int32)
    for c in range(ncomp):
        acc = np.zeros((height, width), np.float32)
        for _ in range(3):
            fx, fy = rng.uniform(0.5, 4.0, 2) * 2 * np.pi / max(width, height)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.outer(np.cos(fy * y), np.cos(fx * x + ph)).astype(np.float32)
            acc -= np.outer(np.sin(fy * y), np.sin(fx * x + ph)).astype(np.float32)
        acc *= np.float32(0.1 * maxv * 0.6 / 0.8)
        acc += np.float32(0.5 * maxv * 0.6 / 0.8)
        acc += rng.standard_normal((height, width), dtype=np.float32) * np.float32(0.02 * maxv)
        np.rint(acc, out=acc)
        out[c] = np.clip(acc, 0, maxv).astype(np.int32)
    return out


def forward_tile(samples, prec, reversible, nlevels, quality=1.0, sgnd=0, mct=1):
    """samples: int32 [ncomp, h, w] -> list of int32 coefficient planes (dense-prefix layout), flat"""
    ncomp, h, w = samples.shape
    comps = [samples[c].reshape(-1).astype(np.int32) for c in range(ncomp)]
    if not sgnd:
        comps = [dc_shift_forward(c, prec) for c in comps]                    # encoder.go:218-220
    if ncomp >= 3 and mct:
        if reversible:
            r, g, b = fwd_rct(comps[0], comps[1], comps[2])                    # encoder.go:224-225
        else:
            fr, fg, fb = fwd_ict(*(c.astype(np.float64) for c in comps[:3]))   # encoder.go:227-245
            r, g, b = (np.where(f >= 0, np.trunc(f + 0.5), np.trunc(f - 0.5)).astype(np.int32) for f in (fr, fg, fb))
        comps[0], comps[1], comps[2] = r, g, b
    planes = []
    for c in comps:
        if reversible:
            planes.append(decompose53(c, w, h, nlevels))                       # encoder.go:255-256
        else:
            f = decompose97(c.astype(np.float64), w, h, nlevels)               # encoder.go:258-263
            step = 1.0 / quality
            q = f / step
            planes.append(np.where(q >= 0, np.trunc(q + 0.5), np.trunc(q - 0.5)).astype(np.int32))  # :269-275
    return planes


def build_ref_job(samples, prec, tile_w=None, tile_h=None, nlevels=5, reversible=True, ht=False, cb=64,
                  sgnd=0, mct=1, quality=1.0, threads=None):
    """samples int32 [ncomp, H, W] -> dict with image header fields, tilecomps, cblks, blob and planes."""
    ncomp, H, W = samples.shape
    tile_w = tile_w or W
    tile_h = tile_h or H
    tcs, plane_list, blks, meta = [], [], [], []
    plane_off = 0
    for ty in range(0, H, tile_h):
        for tx in range(0, W, tile_w):
            x1, y1 = min(tx + tile_w, W), min(ty + tile_h, H)
            w, h = x1 - tx, y1 - ty
            planes = forward_tile(np.ascontiguousarray(samples[:, ty:y1, tx:x1]), prec, reversible, nlevels,
                                  quality, sgnd, mct)
            for c in range(ncomp):
                t_index = len(tcs)
                tcs.append((c, tx, ty, x1, y1, 0))
                plane_list.append(planes[c])
                hx, hy = (w + 1) // 2, (h + 1) // 2
                for by in range(0, h, cb):
                    for bx in range(0, w, cb):
                        bw, bh = min(cb, w - bx), min(cb, h - by)
                        band = (1 if bx >= hx else 0) + (2 if by >= hy else 0)
                        blks.append((plane_off, w, bx, by, bw, bh, band, 1 if ht else 0))
                        meta.append((t_index, bx, by, bw, bh, band))
                plane_off += w * h
    flat = np.concatenate(plane_list) if plane_list else np.zeros(0, np.int32)
    blob, offs, lens, nbps = encode_blocks(flat, blks, threads)
    cblks = np.zeros(len(blks), CBLK_DT)
    m = np.array(meta, np.int64).reshape(-1, 6)
    cblks["data_off"], cblks["data_len"], cblks["num_bps"] = offs, lens, nbps
    cblks["tilecomp"], cblks["x0"], cblks["y0"], cblks["w"], cblks["h"], cblks["band"] = (m[:, i] for i in range(6))
    cblks["step"] = 1.0
    tilecomps = np.array(tcs, TILECOMP_DT)
    return dict(width=W, height=H, ncomp=ncomp, prec=prec, sgnd=sgnd, mct=mct, reversible=int(bool(reversible)),
                nlevels=nlevels, ht=int(bool(ht)), tilecomps=tilecomps, cblks=cblks, blob=blob,
                planes=plane_list, samples=samples)


def as_ctypes(arr, ctype):
    """view a structured numpy table as a ctypes array of the byte-identical struct `ctype`"""
    assert arr.dtype.itemsize == C.sizeof(ctype)
    arr = np.ascontiguousarray(arr)
    n = len(arr)
    buf = (ctype * max(n, 1)).from_buffer_copy(arr.tobytes() if n else bytes(C.sizeof(ctype)))
    return buf


def build_iso_job(samples, prec, tile_w=None, tile_h=None, nlevels=5, mct=1, cb=64, ht_passes=1, ht_plane=0):
    """ISO-mode (J2KGPU_MODE_ISO) job: the same source image as a conformant HTJ2K codestream (lossless 5-3, RCT,
    HT cleanup-only blocks) plus the flat tables a tier-2 parser would hand to j2kgpu_decode.  Block placement
    (x0, y0) is in the tile-component's Mallat plane; num_bps = Mb - missing_msbs = 1 (HT cleanup carries every
    magnitude bit), num_passes = 1."""
    from . import codestream as cs
    ncomp, H, W = samples.shape
    tile_w, tile_h = tile_w or W, tile_h or H
    data, info = cs.write_htj2k(samples, prec, tile_w, tile_h, nlevels, mct=mct, cb=cb, ht_passes=ht_passes, ht_plane=ht_plane)
    ntx = cs.cdiv(W, tile_w)
    tcs, tc_index = [], {}
    for ty in range(cs.cdiv(H, tile_h)):
        for tx in range(ntx):
            for c in range(ncomp):
                tc_index[(ty * ntx + tx, c)] = len(tcs)
                tcs.append((c, tx * tile_w, ty * tile_h, min((tx + 1) * tile_w, W), min((ty + 1) * tile_h, H), 0))
    blks = info["blocks"]
    cblks = np.zeros(len(blks), CBLK_DT)
    blob = bytearray()
    for i, b in enumerate(blks):
        cblks[i] = (len(blob), len(b["data"]), tc_index[(b["tile"], b["comp"])], b["px"], b["py"], b["w"], b["h"],
                    b["band"], b["level"], 1 + ht_plane, max(b["passes"], 1), 1.0, b["lcup"] if b["passes"] > 1 else 0, 0)
        blob += b["data"]
    return dict(width=W, height=H, ncomp=ncomp, prec=prec, sgnd=0, mct=1 if (mct and ncomp >= 3) else 0, reversible=1,
                nlevels=nlevels, ht=1, mode=1, tilecomps=np.array(tcs, TILECOMP_DT), cblks=cblks,
                blob=np.frombuffer(bytes(blob), np.uint8).copy(), samples=samples, codestream=data,
                coef_bits=max(info["eps"]) + info["guard"] - 1)      # max Mb over bands (QCD exponent + guard bits - 1)


def build_iso_job_from_codestream(data, reduce=0):
    """ISO-mode job tables from any codestream datagen.codestream.parse_codestream understands (e.g. one written by
    OpenJPEG): classic EBCOT or HT blocks, any number of quality layers (the block's segments are concatenated and
    num_passes says how many coding passes they hold), reversible or irreversible (step = dequantisation step).
    reduce = Config.ReduceResolution (jpeg2000.go:205-207, decoder.go:289-295): the `reduce` finest resolutions are
    left out -- purely a matter of which blocks the host hands over: in the Mallat plane the remaining bands keep
    their positions, the image and tile bounds shrink by 2^reduce (ceil), nlevels drops by reduce."""
    from . import codestream as cs
    h = cs.parse_codestream(data)
    if reduce > h["nlevels"]:
        raise ValueError("reduce > number of decomposition levels")
    sc = 1 << reduce
    W, H, ncomp = cs.cdiv(h["width"], sc), cs.cdiv(h["height"], sc), h["ncomp"]
    ntx = cs.cdiv(h["width"], h["tile_w"])
    tcs, tc_index = [], {}
    for ty in range(cs.cdiv(h["height"], h["tile_h"])):
        for tx in range(ntx):
            for c in range(ncomp):
                tc_index[(ty * ntx + tx, c)] = len(tcs)
                tcs.append((c, cs.cdiv(tx * h["tile_w"], sc), cs.cdiv(ty * h["tile_h"], sc),
                            cs.cdiv(min((tx + 1) * h["tile_w"], h["width"]), sc), cs.cdiv(min((ty + 1) * h["tile_h"], h["height"]), sc), 0))
    blks = [b for b in h["blocks"] if b["res"] <= h["nlevels"] - reduce]
    cblks = np.zeros(len(blks), CBLK_DT)
    blob = bytearray()
    gain = {0: 0, 1: 1, 2: 1, 3: 2}
    for i, b in enumerate(blks):
        # Annex E.1: step = 2^(Rb - eps) * (1 + mu / 2^11), Rb = precision + band gain
        step = 1.0 if h["reversible"] else 2.0 ** (h["prec"] + gain[b["band"]] - b["expn"]) * (1.0 + b["mant"] / 2048.0)
        nb = b["num_bps"] if b["passes"] else 0
        cblks[i] = (len(blob), len(b["data"]) if nb else 0, tc_index[(b["tile"], b["comp"])], b["px"], b["py"], b["w"], b["h"],
                    b["band"], max(b["level"] - reduce, 0), nb, min(b["passes"], 255), step, b.get("lcup", 0), 0)
        blob += b["data"]
    return dict(width=W, height=H, ncomp=ncomp, prec=h["prec"], sgnd=h["sgnd"], mct=1 if (h["mct"] and ncomp >= 3) else 0,
                reversible=h["reversible"], nlevels=h["nlevels"] - reduce, ht=h["ht"], mode=1, tilecomps=np.array(tcs, TILECOMP_DT), cblks=cblks,
                blob=np.frombuffer(bytes(blob) + bytes(8), np.uint8).copy(), codestream=bytes(data), layers=h["layers"],
                coef_bits=max(e + h["guard"] - 1 for e, _ in h["qcd"]), cblk_style=0 if h["ht"] else h["cblk_style"] & 0x3F)
