#!/usr/bin/env python3
"""bench.py -- decoded Mpixels/s of the B200 JPEG 2000 tile-component decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--no-extra]

Headline workload = BASELINE configs[1] as a CONFORMANT HTJ2K codestream (ISO/IEC 15444-15; J2KGPU_MODE_ISO):
3840x2160 RGB 8-bit lossless 5-3, 5 decomposition levels (6 resolutions), 512x512 tiles, RCT, 64x64 code blocks,
HT cleanup-only blocks, F DISTINCT frames per step.  One "step" = one pass of the whole hot path (HT block decode ->
inverse 5-3 DWT -> inverse RCT -> DC shift -> clamp -> RGBA pack) over that batch.

  value      device-resident: inputs in HBM when the timed region starts, CUDA events on the launching stream, barrier +
             synchronize on both sides, max over ranks, EXACTLY --steps steps (`sustained` repeats it over >= 200 steps).
  e2e        the plugin call j2kgpu_decode_batch with pinned HOST buffers: table validation / flattening / upload, H2D of the
             compressed bytes, kernels and D2H of the pixels all inside the clock (`e2e.prebuilt_job` = the same batch through a
             job whose tables were uploaded beforehand, j2kgpu_job_run_host).
  guard      before anything is timed every frame of every timed workload is decoded once and compared: lossless frames with
             their source image, frame 0 additionally with the CPU checker (oracle/) and, for the headline, with OpenJPEG.
  roofline   the HBM-bound kernel of the path (IDWT levels 1+0 + RCT + DC + clamp + RGBA pack): ALGORITHMIC bytes per launch
             (4*W*H*C + W*H*bpp per frame -- the reference's int32 coefficients, whatever the plane type in HBM)
             over its CUDA-event time, against MEASURED_PEAKS.json; `moved_*` = the bytes the kernel really has to move.
  side       the same geometry in REF semantics: the reference's own (non-conformant, one row in four) HT coder and its
             EBCOT coder (ref_ht, ebcot_ref); cfg4 (8192^2 16-bit grey, 1024^2 tiles) decoded as ONE image whose tiles are
             sharded over the ranks into one shared host buffer (cfg4_tile_sharded).
The batch working set (F x (50-100 MB coefficients + 33 MB pixels)) is larger than the 126 MB L2: no L2 flush needed.

--impl reference times the CPU implementation of the same workload (the C checker of oracle/ with all host threads; the Go
reference cannot run: no Go toolchain, and its own decodeTile is a placeholder -- SURVEY.md F1/F5).
"""
import argparse
import ctypes as C
import json
import mmap
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, NCOMP, PREC, TILE, LEVELS = 3840, 2160, 3, 8, 512, 5
METRIC = "decoded_mpixels_per_s"
UNIT = "Mpixel/s"
# dram__bytes_read.sum + dram__bytes_write.sum of the roofline kernel over its algorithmic bytes, from the ncu --set full
# captures committed under profiles/ (an OFFLINE capture of this same command; the bench cannot run ncu on itself)
# (keyed by bytes per coefficient in HBM; the ratio is to the bytes that variant has to move)
NCU_TRAFFIC = {4: (1.069, "profiles/r1_ncu_idwt53_wide_int32_final.txt"), 2: (1.053, "profiles/r2_ncu_final_headline_kernels.txt")}


def workload_name(frames):
    return ("cfg2 as a conformant HTJ2K codestream: %dx%d RGB 8-bit lossless 5-3, %d levels, %dx%d tiles, RCT, 64x64 blocks, "
            "HT cleanup-only code blocks (ISO/IEC 15444-15), batch of %d distinct frames" % (W, H, LEVELS, TILE, TILE, frames))


# ---- input generation (CPU, before CUDA is touched: worker processes are forked) ---------------------------------------------
def _make_frame(spec):
    kind, seed, keep_cs = spec
    from datagen import jobs
    s = jobs.synth_image_fast(W, H, NCOMP, PREC, seed=seed)
    if kind == "iso":
        j = jobs.build_iso_job(s, PREC, TILE, TILE, LEVELS)
        if not keep_cs:
            j.pop("codestream", None)
    else:
        j = jobs.build_ref_job(s, PREC, TILE, TILE, nlevels=LEVELS, reversible=True, ht=(kind == "ref_ht"), threads=2)
        j.pop("planes", None)
    j["samples"] = s.astype(np.uint8)
    return j


def _make_cfg4_tile(spec):
    seed, tx, ty, tile, prec, nlevels = spec
    from datagen import jobs
    s = jobs.synth_image_fast(tile, tile, 1, prec, seed=seed)
    j = jobs.build_iso_job(s, prec, None, None, nlevels)
    j.pop("codestream", None)
    return j


def build_many(fn, specs, workers):
    if workers <= 1 or len(specs) <= 1:
        return [fn(s) for s in specs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(workers, len(specs))) as pool:
        return pool.map(fn, specs, chunksize=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm       # samples taken under load
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pixels(job, threads):
    """frame through the CPU checker: ISO jobs -> oracle/iso_path.c, REF jobs -> orc_decode_image; returns (pixels, seconds)"""
    import oracle_lib as O
    from datagen import jobs
    t0 = time.perf_counter()
    if job.get("mode") == 1:
        out = O.iso_decode_job(job, threads=threads)
    else:
        img = O.Image()
        img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
        for c in range(job["ncomp"]):
            img.prec[c], img.sgnd[c] = job["prec"], job["sgnd"]
        img.mct, img.reversible, img.nlevels, img.ht = job["mct"], job["reversible"], job["nlevels"], job["ht"]
        out = O.decode_image(img, jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk),
                             job["blob"], job["width"] * 4, job["width"] * job["height"] * 4, threads=threads)
    return out, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the CPU implementation of the headline workload on the host cores (all threads), one frame of the
    batch per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    job = _make_frame(("iso", 2002, False))
    warm = min(args.warmup, 1)
    for _ in range(warm):
        oracle_pixels(job, threads)
    times = []
    for _ in range(max(1, args.steps)):
        out, dt = oracle_pixels(job, threads)
        times.append(dt)
    pix = out.reshape(H, W, 4)
    assert all(np.array_equal(pix[:, :, c], job["samples"][c]) for c in range(3)), "the CPU arm decodes its input incorrectly"
    tot = sum(times)
    val = (W * H / 1e6) * len(times) / tot
    sample = "1 frame of the workload per step (of %d in the GPU arm's batch), all host threads" % args.frames
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": warm, "ms_per_step": round(1e3 * tot / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.frames), "mode": "ISO", "sample": sample},
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "C statement of the same ISO-mode path (oracle/iso_path.c, bit-identical to OpenJPEG 2.5.4); "
                                     "the Go reference cannot run here (no Go toolchain) and its decodeTile is a placeholder"},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def build_items(j2k, jobs, frames, mode):
    """pinned host buffers + batch items for a list of frame jobs"""
    import torch
    keep, items, host_out = [], [], []
    for j in frames:
        bpp = j2k.fmt_bpp(j["ncomp"], j["prec"])
        stride = j["width"] * bpp
        tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
        hb = torch.from_numpy(np.ascontiguousarray(j["blob"])).pin_memory()
        ho = torch.empty(stride * j["height"], dtype=torch.uint8).pin_memory()
        keep += [tcs, cbs, hb]
        host_out.append(ho)
        img = j2k.make_image(j["width"], j["height"], j["ncomp"], j["prec"], nlevels=j["nlevels"], ht=j["ht"], mode=mode,
                             coef_bits=j.get("coef_bits", 0))
        items.append(j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), C.cast(hb.data_ptr(), j2k.u8p), hb.numel(),
                                   C.cast(ho.data_ptr(), j2k.u8p), stride, 0, 0))
    return items, host_out, keep


def measure(args, ctx, j2k, jobs, frames, mode, stream, barrier, steps, e2e_steps, long_steps, rank0_checks):
    """device-resident and end-to-end timing of one workload; returns a dict of raw measurements"""
    import torch
    stride = W * 4
    F = len(frames)
    items, host_out, keep = build_items(j2k, jobs, frames, mode)
    job = j2k.Job(ctx, items)
    d_blob = torch.cat([torch.from_numpy(np.ascontiguousarray(j["blob"])) for j in frames] + [torch.zeros(64, dtype=torch.uint8)]).cuda()
    d_out = torch.empty(job.out_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    # ---- correctness guard on the exact inputs that are timed: every frame against its source image (all three
    # workloads are lossless where the coder decodes every sample), frame 0 against the CPU checker / OpenJPEG ----
    job.run(d_blob.data_ptr(), d_out.data_ptr())
    stream.synchronize()
    lossless = not (mode == 0 and frames[0]["ht"])               # the reference's HT coder decodes one row in four (ht.go:677)
    firsts = []
    for i, fr in enumerate(frames):
        off = job.out_offset(i)
        pix = d_out[off: off + stride * H].cpu().numpy().reshape(H, W, 4)
        if lossless:
            for c in range(3):
                assert np.array_equal(pix[:, :, c], fr["samples"][c]), "bench input %d decodes incorrectly" % i
        if i == 0:
            firsts = pix.copy()
    guard = {"frames_equal_source": F if lossless else 0}
    if rank0_checks:
        want, cpu_s = oracle_pixels(frames[0], os.cpu_count() or 1)
        assert np.array_equal(firsts.reshape(-1), want), "frame 0 differs from the CPU checker"
        guard["frame0_equals_cpu_checker"] = True
        guard["cpu_s"] = cpu_s
        if frames[0].get("codestream"):
            try:
                import io
                from PIL import Image
                t0 = time.perf_counter()
                im = Image.open(io.BytesIO(frames[0]["codestream"]))
                im.load()
                guard["openjpeg_s"] = time.perf_counter() - t0
                assert np.array_equal(np.array(im), firsts[:, :, :3]), "frame 0 differs from OpenJPEG's decode"
                guard["frame0_equals_openjpeg"] = True
            except ImportError:
                pass
    for _ in range(args.warmup):
        job.run(d_blob.data_ptr(), d_out.data_ptr())

    def timed(n):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        l0 = ctx.launches
        ev[0].record(stream)
        for _ in range(n):
            job.run(d_blob.data_ptr(), d_out.data_ptr())
        ev[1].record(stream)
        barrier()
        return ev[0].elapsed_time(ev[1]), ctx.launches - l0

    ms_total, launches = timed(steps)
    ms_long, _ = timed(long_steps) if long_steps else (0.0, 0)

    def time_fn(fn, reps, pre=None):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            if pre:
                pre()
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    reps = max(5, min(steps, 20))
    ent_ms = time_fn(lambda: job.run_entropy(d_blob.data_ptr()), reps)
    dwt_ms = time_fn(lambda: job.run_dwt_mct(d_out.data_ptr()), reps)
    nl = frames[0]["nlevels"]
    # the roofline kernel alone: the coarser levels refill the ping-pong buffers first
    last_ms = time_fn(lambda: job.run_level(0, d_out.data_ptr()), reps, pre=lambda: [job.run_level(l) for l in range(nl - 1, 0, -1)])
    # ---- end to end: the plugin call (tables built inside), and the prebuilt job beside it ----
    for _ in range(2):
        ctx.decode_batch(items)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.decode_batch(items)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(host_out[0].numpy().reshape(H, W, 4), firsts), "plugin call differs from the device path"
    for _ in range(2):
        job.run_host()
    barrier()
    t0 = time.perf_counter()
    pre_steps = max(1, min(e2e_steps, 10))
    for _ in range(pre_steps):
        job.run_host()
    barrier()
    pre_s = time.perf_counter() - t0
    # ---- the codestream front door: raw codestream bytes in (page-locked), tier-2 on host threads inside the library ----
    cs_s, cs_steps, cs_bytes = 0.0, 0, 0
    if all(fr.get("codestream") for fr in frames):
        cs_host = [torch.from_numpy(np.frombuffer(fr["codestream"], np.uint8).copy()).pin_memory() for fr in frames]
        cs_arrs = [t.numpy() for t in cs_host]
        outs = [t.numpy() for t in host_out]
        for t in host_out:
            t.zero_()
        for _ in range(2):
            ctx.decode_codestreams(cs_arrs, outs=outs)
        assert np.array_equal(host_out[0].numpy().reshape(H, W, 4), firsts), "codestream front door differs from the device path"
        barrier()
        t0 = time.perf_counter()
        cs_steps = max(1, min(e2e_steps, 10))
        for _ in range(cs_steps):
            ctx.decode_codestreams(cs_arrs, outs=outs)
        barrier()
        cs_s = time.perf_counter() - t0
        cs_bytes = int(sum(a.size for a in cs_arrs))
    res = dict(cs_s=cs_s, cs_steps=cs_steps, cs_bytes=cs_bytes, ms_total=ms_total, steps=steps, launches=int(launches), ms_long=ms_long, long_steps=long_steps, ent_ms=ent_ms,
               dwt_ms=dwt_ms, last_ms=last_ms, e2e_s=e2e_s, e2e_steps=e2e_steps, pre_s=pre_s, pre_steps=pre_steps,
               h2d=int(d_blob.numel()) - 64, d2h=stride * H * F, n_blocks=sum(len(j["cblks"]) for j in frames), F=F,
               fused_levels=job.fused_levels, coef_bytes=job.coef_bytes, plan=job.plan, guard=guard)
    job.close()
    return res


def bind_to_gpu_numa_node(local):
    """pin this rank (and therefore its pinned host buffers: first touch) to the CPUs of the NUMA node its GPU hangs off,
    so that 8 ranks do not push their H2D / D2H traffic across the socket interconnect; returns a note for the JSON line"""
    try:
        import torch
        props = torch.cuda.get_device_properties(local)
        bus = getattr(props, "pci_bus_id", None)
        dom = getattr(props, "pci_domain_id", 0)
        dev = getattr(props, "pci_device_id", 0)
        if bus is None:
            return "numa: unknown"
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "numa: rank bound to node %d (%d cpus)" % (node, len(allowed))
        return "numa: node %d has no allowed cpus" % node
    except Exception as e:                       # best effort: never fail the bench over affinity
        return "numa: not bound (%s)" % str(e)[:60]


def run_cfg4_tile_sharded(ctx, j2k, jobs, shard, tile_jobs, my_tiles, world, rank, barrier, reduce_max, steps):
    """BASELINE configs[3]: ONE 8192x8192 16-bit grey image, 1024x1024 tiles, its 64 tiles sharded over the ranks (LPT on
    compressed bytes, go-jpeg2000_b200/shard.py); every rank decodes its tiles and copies ONLY their pixel rectangles
    (J2KGPU_ITEM_TILES_ONLY) into ONE host image shared by all ranks (a page-locked /dev/shm mapping); rank 0 checks the
    assembled image against the sources.  No collective on the data path."""
    import torch
    G, T, prec, nl = 8192, 1024, 16, 5
    ntx = G // T
    stride = G * 2
    path = "/dev/shm/j2kgpu_cfg4_%s.pix" % os.environ.get("MASTER_PORT", "0")
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(stride * G)
    barrier()
    fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, stride * G)
    host = np.frombuffer(mm, np.uint8)
    ctx.host_register(host)
    # this rank's tiles go in as several items (groups of tiles) of ONE batch call: the library pipelines items, so the copy-in
    # of a group overlaps the decode of the previous one and the copy-out of the one before
    ngroups = max(1, min(8, len(my_tiles)))
    groups = [my_tiles[g * len(my_tiles) // ngroups:(g + 1) * len(my_tiles) // ngroups] for g in range(ngroups)]
    cbits = max(tile_jobs[t]["coef_bits"] for t in my_tiles) if my_tiles else 0
    img = j2k.make_image(G, G, 1, prec, nlevels=nl, ht=1, mode=1, coef_bits=cbits)
    items, keep = [], []
    for grp in groups:
        tcs, cbs, blobs, boff = [], [], [], 0
        for t in grp:
            j = tile_jobs[t]
            tx, ty = t % ntx, t // ntx
            tc = j["tilecomps"].copy()
            tc["x0"] += tx * T; tc["x1"] += tx * T; tc["y0"] += ty * T; tc["y1"] += ty * T
            cb = j["cblks"].copy()
            cb["tilecomp"] += sum(len(a) for a in tcs)
            cb["data_off"] += boff
            tcs.append(tc)
            cbs.append(cb)
            blobs.append(j["blob"])
            boff += j["blob"].size
        tc_arr = np.concatenate(tcs) if tcs else np.zeros(0, jobs.TILECOMP_DT)
        cb_arr = np.concatenate(cbs) if cbs else np.zeros(0, jobs.CBLK_DT)
        blob = torch.from_numpy(np.concatenate(blobs) if blobs else np.zeros(8, np.uint8)).pin_memory()
        c_tcs, c_cbs = jobs.as_ctypes(tc_arr, j2k.TileComp), jobs.as_ctypes(cb_arr, j2k.CBlk)
        keep += [c_tcs, c_cbs, blob]
        items.append(j2k.BatchItem(img, c_tcs, len(tc_arr), c_cbs, len(cb_arr), C.cast(blob.data_ptr(), j2k.u8p), blob.numel(),
                                   C.cast(host.ctypes.data, j2k.u8p), stride, j2k.ITEM_TILES_ONLY, 0))
    for _ in range(2):
        ctx.decode_batch(items)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.decode_batch(items)
    barrier()
    (dt,) = reduce_max(time.perf_counter() - t0)
    ok = True
    if rank == 0:                                            # the assembled image, every tile, against its source
        img16 = host.reshape(G, G, 2)
        val = (img16[:, :, 0].astype(np.uint16) << 8) | img16[:, :, 1]
        for t, j in enumerate(tile_jobs):
            tx, ty = t % ntx, t // ntx
            ok &= bool(np.array_equal(val[ty * T:(ty + 1) * T, tx * T:(tx + 1) * T], j["samples"][0].astype(np.uint16)))
        del img16, val
    barrier()
    ctx.host_unregister(host)
    del host, items                                          # every view of the mapping has to go before it can be closed
    try:
        mm.close()
    except BufferError:                                      # a stray view keeps it alive until exit: harmless
        pass
    os.close(fd)
    barrier()
    if rank == 0:
        os.unlink(path)
    return dict(value=round(G * G / 1e6 * steps / dt, 1), unit=UNIT, ms_per_image=round(1e3 * dt / steps, 3), steps=steps,
                tiles_per_rank=[len(p) for p in shard.shard_units([tile_jobs[t]["blob"].size for t in range(len(tile_jobs))], world)],
                assembled_image_equals_source=ok, items_per_rank=ngroups,
                api="j2kgpu_decode_batch, J2KGPU_ITEM_TILES_ONLY, one shared page-locked host image; a rank's tiles go in as up to 8 items so that copy-in, decode and copy-out overlap",
                workload="cfg4: 8192x8192 grey 16-bit lossless HTJ2K, 1024x1024 tiles, ONE image, tiles sharded over %d rank(s), "
                         "end to end (H2D + kernels + D2H of the owned tile rectangles)" % world)


def run_forward_path(ctx, j2k, stream, steps):
    """side object: the forward path (encoder.go:79-281, 597-743) on one 4K RGB frame, lossless, 6 resolutions,
    64 x 64 blocks -- device-resident (pixels and tile bytes in HBM), end to end from page-locked host pixels, and the CPU
    checker (oracle/orc_enc.c, all host threads) on the same frame; the three outputs are compared"""
    import ctypes as C
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from datagen import jobs
    w, h = W, H
    rgb = jobs.synth_image_fast(w, h, 3, 8, seed=4242)
    pix = np.full((h, w, 4), 255, np.uint8)
    pix[:, :, :3] = np.moveaxis(rgb, 0, 2)
    pix = pix.reshape(-1)
    kw = dict(width=w, height=h, ncomp=3, pix_bits=8, lossless=1, num_resolutions=6, cb_x=4, cb_y=4)
    p_dev, p_host = j2k.EncodeParams(flags=j2k.ENC_DEVICE_PTRS, **kw), j2k.EncodeParams(**kw)
    L = j2k.lib()
    n = int(L.j2kgpu_encode_block_count(C.byref(p_host)))
    lens, bps = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
    cap = w * h * 3 * 2
    got = C.c_uint64(0)
    d_pix = torch.from_numpy(pix).cuda()
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    h_pix = torch.from_numpy(pix).pin_memory()
    h_out = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def call(params, src, dst):
        rc = L.j2kgpu_encode_tile(ctx._h, C.byref(params), src.data_ptr(), w * 4, dst.data_ptr(), cap, C.byref(got),
                                  lens.ctypes.data_as(C.POINTER(C.c_uint32)), bps.ctypes.data_as(j2k.u8p), n)
        if rc:
            raise RuntimeError("j2kgpu_encode_tile rc=%d" % rc)

    def timed(params, src, dst):
        for _ in range(2):
            call(params, src, dst)
        n0 = ctx.launches
        t0 = time.perf_counter()
        for _ in range(steps):
            call(params, src, dst)
        return (time.perf_counter() - t0) / steps, (ctx.launches - n0) // steps

    dev_s, launches = timed(p_dev, d_pix, d_out)
    dev_bytes = d_out[: got.value].cpu().numpy().copy()
    e2e_s, _ = timed(p_host, h_pix, h_out)
    host_bytes = h_out[: got.value].numpy().copy()
    t0 = time.perf_counter()
    want, wlens, wbps = O.encode_tile(p_host, pix, threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(dev_bytes, want) and np.array_equal(host_bytes, want) and np.array_equal(lens, wlens) and np.array_equal(bps, wbps))
    mp = w * h / 1e6
    return dict(workload="one 3840x2160 RGB 8-bit frame, lossless 5-3 + RCT, 6 resolutions, 64x64 blocks, the reference encoder's EBCOT / MQ coder",
                value=round(mp / dev_s, 1), unit=UNIT, ms_per_frame=round(dev_s * 1e3, 3), code_blocks=n, tile_bytes=int(got.value),
                gpu_launches=int(launches), timing="wall clock around the blocking call (it synchronises before it returns)",
                e2e=dict(value=round(mp / e2e_s, 1), unit=UNIT, ms_per_frame=round(e2e_s * 1e3, 3), h2d_bytes_per_step=int(pix.size),
                         d2h_bytes_per_step=int(got.value) + 5 * n, api="j2kgpu_encode_tile, page-locked host pixels in, tile bytes out"),
                cpu_baseline=dict(value=round(mp / cpu_s, 2), unit=UNIT, cores=os.cpu_count() or 1, kind="port",
                                  sample="the same frame, oracle/orc_enc.c on all host threads"),
                bytes_equal_checker=same)


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    F = args.frames
    # ---- which frames are this rank's: the batch of world * F frames is sharded with the package's planner (equal costs:
    # the frames do not exist yet; every rank derives the same plan without talking to the others) ----
    from __graft_entry__ import load_package
    j2k = load_package()
    from go_jpeg2000_b200 import shard
    mine = shard.shard_units([W * H] * (world * F), world, rank)
    assert len(mine) == F
    workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    t_gen = time.perf_counter()
    specs = [("iso", 2002 + g, True) for i, g in enumerate(mine)]      # every frame keeps its codestream (front-door e2e)
    n_ref, n_eb = (min(F, 4), min(F, 2)) if not args.no_extra else (0, 0)
    specs += [("ref_ht", 1002 + mine[i], False) for i in range(n_ref)] + [("ebcot", 3002 + mine[i], False) for i in range(n_eb)]
    built = build_many(_make_frame, specs, workers)
    iso_frames, ref_frames, eb_frames = built[:F], built[F:F + n_ref], built[F + n_ref:]
    tile_jobs = []
    if not args.no_extra:
        tile_jobs = build_many(_make_cfg4_tile, [(4004 + t, t % 8, t // 8, 1024, 16, 5) for t in range(64)], workers)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    from datagen import jobs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "numa: single rank, not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = j2k.Context(local)
    # an explicit torch stream: the library launches on it and the CUDA events are recorded on it
    # (torch's legacy default stream has handle 0, which j2kgpu_set_stream reads as "use the ctx's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        if world == 1:
            return vals
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    peak, peak_src = peaks()

    def summarize(m):
        Fm = m["F"]
        mpix_step = W * H * Fm * world / 1e6
        alg = (4 * W * H * NCOMP + W * H * 4) * Fm                  # SURVEY.md 8(d), per launch of the roofline kernel
        moved = (m["coef_bytes"] * W * H * NCOMP + W * H * 4) * Fm   # what this variant's kernel has to move (int16 planes: less)
        ms_total, ms_long, e2e_s, pre_s, cs_s = reduce_max(m["ms_total"], m["ms_long"], m["e2e_s"], m["pre_s"], m["cs_s"])
        ach, stg = alg / (m["last_ms"] / 1e3) / 1e9, alg / (m["dwt_ms"] / 1e3) / 1e9
        last_kernel = ("k_idwt53_wide (16 columns per lane)" if m["plan"] & 4 else
                       "k_idwt53_fused (4 columns per lane)" if m["plan"] & 1 else "k_idwt53_stream")
        ratio, src = NCU_TRAFFIC.get(m["coef_bytes"], (None, None))
        out = dict(value=round(mpix_step * m["steps"] / (ms_total / 1e3), 1), ms_per_step=round(ms_total / m["steps"], 4),
                   frames_per_gpu_per_step=Fm,
                   e2e=dict(value=round(mpix_step * m["e2e_steps"] / e2e_s, 1), unit=UNIT, h2d_bytes_per_step=m["h2d"],
                            d2h_bytes_per_step=m["d2h"], ms_per_step=round(1e3 * e2e_s / m["e2e_steps"], 3), steps=m["e2e_steps"],
                            api="j2kgpu_decode_batch (tables validated, flattened and uploaded inside the call; pinned host buffers)",
                            prebuilt_job=dict(value=round(mpix_step * m["pre_steps"] / pre_s, 1), unit=UNIT,
                                              ms_per_step=round(1e3 * pre_s / m["pre_steps"], 3), api="j2kgpu_job_run_host")),
                   gpu_launches=m["launches"], code_blocks_per_step=m["n_blocks"] * world,
                   plan=dict(idwt_levels_in_last_kernel=m["fused_levels"], coef_plane_bytes_per_sample=m["coef_bytes"], last_kernel=last_kernel),
                   stages_ms=dict(entropy=round(m["ent_ms"], 4), dwt_mct_pack=round(m["dwt_ms"], 4), last_level_fused=round(m["last_ms"], 4)),
                   # achieved = the bytes this variant of the kernel has to move (every coefficient once at the plane's element
                   # size, every pixel once) / its measured duration.  SURVEY 8(d) counts coefficients at the reference's int32
                   # (4 B): with int16 planes that figure exceeds what the hardware moves, so it is reported beside, not as, `achieved`
                   roofline=dict(bound="hbm", achieved=round(moved / (m["last_ms"] / 1e3) / 1e9, 1), peak=peak, unit="GB/s",
                                 frac=round(moved / (m["last_ms"] / 1e3) / 1e9 / peak, 4),
                                 kernel=last_kernel + ": IDWT levels 1+0 + RCT + DC shift + clamp + RGBA pack" if m["fused_levels"] == 2 else
                                        last_kernel + ": last IDWT level + RCT + DC shift + clamp + RGBA pack",
                                 peak_source=peak_src, algorithmic_bytes_per_launch=moved, duration_ms=round(m["last_ms"], 4),
                                 coef_plane_bytes_per_sample=m["coef_bytes"],
                                 survey_8d_bytes_per_launch=alg, survey_8d_gbs=round(ach, 1), survey_8d_frac=round(ach / peak, 4),
                                 dwt_mct_stage_gbs=round(moved / (m["dwt_ms"] / 1e3) / 1e9, 1),
                                 dwt_mct_stage_frac=round(moved / (m["dwt_ms"] / 1e3) / 1e9 / peak, 4),
                                 dwt_mct_stage_survey_8d_frac=round(stg / peak, 4),
                                 traffic=int(moved * ratio) if ratio else None,
                                 traffic_source=("offline ncu --set full capture of this command (dram__bytes_read.sum + dram__bytes_write.sum "
                                                 "of the kernel, as a ratio to its bytes), " + src) if ratio else
                                                "no ncu capture of this variant is wired in: see profiles/"),
                   guard=m["guard"])
        if m["cs_steps"]:
            out["e2e"]["codestream_front_door"] = dict(value=round(mpix_step * m["cs_steps"] / cs_s, 1), unit=UNIT,
                                                       ms_per_step=round(1e3 * cs_s / m["cs_steps"], 3), h2d_bytes_per_step=m["cs_bytes"],
                                                       api="j2kgpu_decode_codestreams (raw codestream bytes in: main header, tile-part index "
                                                           "and tier-2 on host threads inside the call, overlapped with the device)")
        if m["long_steps"]:
            out["sustained"] = dict(steps=m["long_steps"], value=round(mpix_step * m["long_steps"] / (ms_long / 1e3), 1),
                                    ms_per_step=round(ms_long / m["long_steps"], 4))
        return out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e2e_steps = max(args.steps, 20) if not args.quick else args.steps
    long_steps = 0 if args.quick else max(200, args.steps)
    m_iso = measure(args, ctx, j2k, jobs, iso_frames, 1, stream, barrier, args.steps, e2e_steps, long_steps, rank == 0)
    clocks = sampler.stop() if rank == 0 else None
    main = summarize(m_iso)

    extra = {}
    if not args.no_extra:
        s_side = max(3, min(args.steps, 10))
        m_ref = measure(args, ctx, j2k, jobs, ref_frames, 0, stream, barrier, s_side, s_side, 0, rank == 0)
        extra["ref_ht"] = summarize(m_ref)
        extra["ref_ht"]["workload"] = ("cfg2 geometry, REF semantics, the reference's own HT coder (ht.go: not ISO/IEC 15444-15, decodes one "
                                       "sample row in four, parity unpinned) -- kept OUT of the headline; %d distinct frames" % len(ref_frames))
        m_eb = measure(args, ctx, j2k, jobs, eb_frames, 0, stream, barrier, 3, 3, 0, rank == 0)
        extra["ebcot_ref"] = summarize(m_eb)
        extra["ebcot_ref"]["workload"] = "cfg2 geometry, REF semantics, the reference's EBCOT/MQ coder (t1.go), %d distinct frames" % len(eb_frames)
        costs = [j["blob"].size for j in tile_jobs]
        if rank == 0 and world == 1:
            extra["forward_path"] = run_forward_path(ctx, j2k, stream, 5)
        extra["cfg4_tile_sharded"] = run_cfg4_tile_sharded(ctx, j2k, jobs, shard, tile_jobs, shard.shard_units(costs, world, rank),
                                                           world, rank, barrier, reduce_max, 5)

    if rank == 0:
        g = m_iso["guard"]
        cpu_threads = os.cpu_count() or 1
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(F), "mode": "ISO", "frames_per_gpu_per_step": F, "distinct_frames": F,
                       "code_blocks_per_step": main["code_blocks_per_step"], "l2": "working set > L2 (no flush needed)",
                       "parallelism": "frames sharded across GPUs (go-jpeg2000_b200/shard.py), no collective", "host": numa_note,
                       "input_generation_s": round(t_gen, 1)},
            "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "stages_ms": main["stages_ms"], "plan": main["plan"],
            "roofline": main["roofline"], "sustained": main.get("sustained"), "guard": {k: v for k, v in g.items() if not k.endswith("_s")},
            "cpu_baseline": {"value": round((W * H / 1e6) / g["cpu_s"], 2), "unit": UNIT, "cores": cpu_threads, "kind": "port",
                             "sample": "frame 0 of the batch, whole path, all host threads",
                             "note": "C statement of the same ISO-mode path (oracle/iso_path.c, bit-identical to OpenJPEG); Go is absent"},
            "clocks": clocks,
        }
        if "openjpeg_s" in g:
            line["openjpeg_cpu"] = dict(value=round(W * H / 1e6 / g["openjpeg_s"], 2), unit=UNIT,
                                        note="OpenJPEG 2.5.4 via Pillow decoding frame 0's codestream (context; same pixels)")
        line.update(extra)
        print(json.dumps(line))
    ctx.set_stream(0)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16, help="frames per GPU per step (the metric is quoted on a batch)")
    ap.add_argument("--no-extra", action="store_true", help="skip the REF-mode and cfg4 side measurements")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no sustained loop, e2e steps = --steps")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
