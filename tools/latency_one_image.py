#!/usr/bin/env python3
"""Latency of ONE image through the plugin call (what a Go decoder sees per jpeg2000.Decode): j2kgpu_decode of a 4K RGB HTJ2K
frame with page-locked buffers, tables built inside the call; with the image in one piece (chunks=1) and pipelined by tile groups."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
j2k = load_package()
from datagen import jobs
import ctypes as C

def main():
    W, H = 3840, 2160
    s = jobs.synth_image(W, H, 3, 8, seed=5)
    job = jobs.build_iso_job(s, 8, 512, 512, 5)
    ctx = j2k.Context(0)
    img = j2k.make_image(W, H, 3, 8, nlevels=5, ht=1, mode=j2k.MODE_ISO, coef_bits=job["coef_bits"])
    tcs, cbs = jobs.as_ctypes(job["tilecomps"], j2k.TileComp), jobs.as_ctypes(job["cblks"], j2k.CBlk)
    blob = ctx.host_alloc(job["blob"].size); blob[:] = job["blob"]
    out = ctx.host_alloc(W * H * 4)
    cs = ctx.host_alloc(len(job["codestream"])); cs[:] = np.frombuffer(job["codestream"], np.uint8)
    res = {}
    def call_tables():
        ctx._check(j2k.lib().j2kgpu_decode(ctx._h, C.byref(img), tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(j2k.u8p), blob.size,
                                           out.ctypes.data_as(j2k.u8p), W * 4))
    def call_cs():
        ctx._check(j2k.lib().j2kgpu_decode_codestream(ctx._h, cs.ctypes.data_as(j2k.u8p), cs.size, 0, out.ctypes.data_as(j2k.u8p), W * 4))
    for name, fn, opt in (("tables, one chunk", call_tables, "1"), ("tables, tile groups", call_tables, ""),
                          ("codestream (tier-2 inside), one chunk", call_cs, "1"), ("codestream (tier-2 inside), tile groups", call_cs, "")):
        ctx.set_option("chunks", opt)
        for _ in range(5): fn()
        t0 = time.perf_counter()
        n = 30
        for _ in range(n): fn()
        res[name] = round(1e3 * (time.perf_counter() - t0) / n, 3)
        assert np.array_equal(out.reshape(H, W, 4)[:, :, :3], np.moveaxis(s, 0, 2))
    print(json.dumps({"workload": "one 3840x2160 RGB lossless HTJ2K frame, 512x512 tiles, page-locked buffers", "ms_per_call": res}))

if __name__ == "__main__":
    main()
