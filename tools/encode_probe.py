#!/usr/bin/env python
"""Forward path alone (for launch lists / ncu): one 4K RGB lossless frame and one 1080p lossy frame through
j2kgpu_encode_tile with device pointers, per-stage device times from CUDA events on the context's stream."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from datagen import jobs  # noqa: E402


def main():
    import torch
    j2k = load_package()
    if os.environ.get("J2K_PROBE_LIB"):                              # A/B builds of the library
        j2k.LIB_PATH = os.environ["J2K_PROBE_LIB"]
        print("library:", j2k.LIB_PATH, flush=True)
    ctx = j2k.Context(0)
    L = j2k.lib()
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    # DefaultOptions() (jpeg2000.go:305-320): lossy, Quality 75, 6 resolutions, CodeBlockSize {6, 6} = 256 x 256 blocks
    w, h = 3840, 2160
    rgb = jobs.synth_image_fast(w, h, 3, 8, seed=13)
    pix = np.full((h, w, 4), 255, np.uint8)
    pix[:, :, :3] = np.moveaxis(rgb, 0, 2)
    p = j2k.EncodeParams(width=w, height=h, ncomp=3, pix_bits=8, lossless=0, num_resolutions=6, cb_x=6, cb_y=6, quality=75)
    t0 = time.perf_counter()
    a = ctx.encode_tile(p, pix.reshape(-1))
    t1 = time.perf_counter()
    a = ctx.encode_tile(p, pix.reshape(-1))
    t2 = time.perf_counter()
    b = O.encode_tile(p, pix.reshape(-1), threads=os.cpu_count() or 1)
    t3 = time.perf_counter()
    print("DefaultOptions (lossy q=75, 256x256 blocks) 4K: GPU %.1f ms (first call %.1f), %d blocks, %d bytes; CPU checker %d threads %.1f ms; equal %s" %
          ((t2 - t1) * 1e3, (t1 - t0) * 1e3, len(a[1]), len(a[0]), os.cpu_count() or 1, (t3 - t2) * 1e3,
           all(np.array_equal(x, y) for x, y in zip(a, b))), flush=True)
    for (w, h, lossless, q) in ((3840, 2160, 1, 0), (1920, 1080, 0, 75)):
        rgb = jobs.synth_image_fast(w, h, 3, 8, seed=11)
        pix = np.full((h, w, 4), 255, np.uint8)
        pix[:, :, :3] = np.moveaxis(rgb, 0, 2)
        p = j2k.EncodeParams(width=w, height=h, ncomp=3, pix_bits=8, lossless=lossless, num_resolutions=6, cb_x=4, cb_y=4, quality=q,
                             flags=j2k.ENC_DEVICE_PTRS)
        n = int(L.j2kgpu_encode_block_count(C.byref(p)))
        d_pix = torch.from_numpy(pix.reshape(-1)).cuda()
        d_out = torch.zeros(w * h * 6, dtype=torch.uint8, device="cuda")
        d_planes = torch.zeros(3 * w * h, dtype=torch.int32, device="cuda")
        got = C.c_uint64(0)
        torch.cuda.synchronize()
        whole = lambda: L.j2kgpu_encode_tile(ctx._h, C.byref(p), d_pix.data_ptr(), w * 4, d_out.data_ptr(), d_out.numel(),  # noqa: E731
                                             C.byref(got), None, None, 0)
        for name, fn, opt in (("preprocess (pixels -> planes, DWT, quantiser)",
                               lambda: L.j2kgpu_encode_preprocess(ctx._h, C.byref(p), d_pix.data_ptr(), w * 4, d_planes.data_ptr()), 0),
                              ("whole forward path, flag-byte tier-1 coder", whole, 1),
                              ("whole forward path, row-mask tier-1 coder", whole, 0)):
            ctx.set_option("enc_bytes", opt)
            for _ in range(2):
                assert fn() == 0
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                assert fn() == 0
            dt = (time.perf_counter() - t0) / reps
            print("%dx%d %s: %s %.3f ms (%.1f Mpixel/s), %d blocks, %d bytes" %
                  (w, h, "lossless" if lossless else "lossy q=%d" % q, name, dt * 1e3, w * h / 1e6 / dt, n, got.value), flush=True)
    # several encoders at once, one context (and stream) per caller as the Go binding keeps them: blocks of different images
    # share the machine, so the calls overlap where one call alone leaves SMs idle behind its longest chains
    import threading
    w, h = 3840, 2160
    rgb = jobs.synth_image_fast(w, h, 3, 8, seed=12)
    pix = np.full((h, w, 4), 255, np.uint8)
    pix[:, :, :3] = np.moveaxis(rgb, 0, 2)
    d_pix = torch.from_numpy(pix.reshape(-1)).cuda()
    torch.cuda.synchronize()
    for nthreads, cb, lossless, q, label in ((1, 4, 1, 0, "4K lossless"), (2, 4, 1, 0, "4K lossless"), (4, 4, 1, 0, "4K lossless"),
                                             (8, 4, 1, 0, "4K lossless"), (1, 6, 0, 75, "4K DefaultOptions"), (4, 6, 0, 75, "4K DefaultOptions"),
                                             (8, 6, 0, 75, "4K DefaultOptions"), (16, 6, 0, 75, "4K DefaultOptions")):
        ctxs = [j2k.Context(0) for _ in range(nthreads)]
        outs = [torch.zeros(w * h * 6, dtype=torch.uint8, device="cuda") for _ in range(nthreads)]
        torch.cuda.synchronize()
        reps = 0

        def work(i):
            p = j2k.EncodeParams(width=w, height=h, ncomp=3, pix_bits=8, lossless=lossless, num_resolutions=6, cb_x=cb, cb_y=cb, quality=q,
                                 flags=j2k.ENC_DEVICE_PTRS)
            got = C.c_uint64(0)
            for _ in range(reps + 1):
                assert L.j2kgpu_encode_tile(ctxs[i]._h, C.byref(p), d_pix.data_ptr(), w * 4, outs[i].data_ptr(), outs[i].numel(), C.byref(got), None, None, 0) == 0

        reps = 0
        for i in range(nthreads):                                       # warm every context: its pool allocates on the first call
            work(i)
        reps = 4 if cb == 4 else 2
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        print("%d contexts at once, %s: %.1f Mpixel/s in total (%.2f ms per frame and context)" %
              (nthreads, label, nthreads * (reps + 1) * w * h / 1e6 / dt, dt / (reps + 1) * 1e3), flush=True)
        for c in ctxs:
            c.close()
    ctx.close()


if __name__ == "__main__":
    main()
