#!/usr/bin/env python3
"""bench.py -- decoded Mpixels/s of the B200 JPEG 2000 tile-component decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--coder ht|ebcot] [--frames F]

One "step" = one pass of the whole hot path (block entropy decode -> inverse DWT -> inverse RCT -> DC shift ->
clamp -> RGBA pack) over one batch of F frames of BASELINE configs[1]: 3840x2160 RGB 8-bit lossless,
5 decomposition levels (6 resolutions), 512x512 tiles, RCT, 64x64 code blocks, block bitstreams produced by the
reference encoder restated in datagen/ (REF semantics).  Inputs are resident in HBM when the
timed region starts (`value`); `e2e` times the same batch through the host-buffer C-ABI call with pinned host
buffers, H2D and D2H inside the timed region.  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  The batch working set (F x (99.5 MB coefficients + 33 MB pixels))
is larger than the 126 MB L2, so no explicit L2 flush is needed between steps.

Extra objects on the JSON line: roofline (the fused kernel "IDWT levels 1+0 + RCT + DC shift + clamp + RGBA pack",
algorithmic bytes 4*W*H*C + W*H*bpp per frame over its CUDA-event time, against MEASURED_PEAKS.json), cpu_baseline
(the C oracle, a restatement of the reference's Go stage functions, all host threads, bounded sample), clocks, and --
unless --no-extra -- the same geometry as a conformant HTJ2K codestream in ISO mode (iso_htj2k, cross-checked with
OpenJPEG) and with the reference's EBCOT coder (ebcot_ref).

--impl reference times the CPU implementation alone (oracle port; the Go reference cannot run: no Go toolchain).
"""
import argparse
import ctypes as C
import json
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
NCU_TRAFFIC_RATIO = {4: 1.069}     # measured DRAM bytes / algorithmic bytes of the fused kernel (ncu capture, see profiles/)
METRIC = "decoded_mpixels_per_s"
UNIT = "Mpixel/s"


def workload_name(coder, frames):
    return ("cfg2: %dx%d RGB 8-bit lossless 5-3, %d levels, %dx%d tiles, RCT, 64x64 blocks, %s block coder "
            "(REF semantics), batch of %d frames" % (W, H, LEVELS, TILE, TILE,
                                                    "reference HT" if coder == "ht" else "EBCOT/MQ", frames))


def build_frame(coder, seed, threads):
    from datagen import jobs
    s = jobs.synth_image(W, H, NCOMP, PREC, seed=seed)
    return jobs.build_ref_job(s, PREC, TILE, TILE, nlevels=LEVELS, reversible=True, ht=(coder == "ht"), threads=threads)


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
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_image(O, job):
    img = O.Image()
    img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
    for c in range(job["ncomp"]):
        img.prec[c], img.sgnd[c] = job["prec"], job["sgnd"]
    img.mct, img.reversible, img.nlevels, img.ht = job["mct"], job["reversible"], job["nlevels"], job["ht"]
    return img


def cpu_decode_time(job, threads, reps):
    """seconds per frame of the oracle's whole path (orc_decode_image) with `threads` host threads"""
    import oracle_lib as O
    from datagen import jobs
    img = oracle_image(O, job)
    tcs, cbs = jobs.as_ctypes(job["tilecomps"], O.TileComp), jobs.as_ctypes(job["cblks"], O.CBlk)
    blob = np.ascontiguousarray(job["blob"])
    out = np.zeros(W * H * 4, np.uint8)
    fn = O.lib().orc_decode_image
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = fn(C.byref(img), tcs, len(tcs), cbs, len(cbs), blob.ctypes.data_as(O.u8p), C.c_uint64(blob.size),
                out.ctypes.data_as(O.u8p), C.c_uint64(W * 4), threads)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    return best, out


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port, all threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    job = build_frame(args.coder, 1002, threads)
    for _ in range(min(args.warmup, 1)):
        cpu_decode_time(job, threads, 1)
    times = []
    for _ in range(max(1, args.steps)):
        t, _ = cpu_decode_time(job, threads, 1)
        times.append(t)
    tot = sum(times)
    val = (W * H / 1e6) * len(times) / tot
    sample = "1 frame of the workload per step (of %d in the GPU arm's batch)" % args.frames
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": round(1e3 * tot / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.coder, args.frames), "mode": "REF", "sample": sample},
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "C restatement of the reference's Go stage functions (oracle/), not Go: no Go toolchain on the box"},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def build_items(j2k, jobs, frames, mode):
    """pinned host buffers + batch items for a list of frame jobs"""
    import torch
    stride = W * 4
    keep, items, host_out = [], [], []
    for j in frames:
        tcs, cbs = jobs.as_ctypes(j["tilecomps"], j2k.TileComp), jobs.as_ctypes(j["cblks"], j2k.CBlk)
        hb = torch.from_numpy(np.ascontiguousarray(j["blob"])).pin_memory()
        ho = torch.empty(stride * H, dtype=torch.uint8).pin_memory()
        keep += [tcs, cbs, hb]
        host_out.append(ho)
        img = j2k.make_image(W, H, NCOMP, PREC, nlevels=LEVELS, ht=j["ht"], mode=mode, coef_bits=j.get("coef_bits", 0))
        items.append(j2k.BatchItem(img, tcs, len(tcs), cbs, len(cbs), C.cast(hb.data_ptr(), j2k.u8p), hb.numel(),
                                   C.cast(ho.data_ptr(), j2k.u8p), stride))
    return items, host_out, keep


def measure(args, ctx, j2k, jobs, frames, mode, stream, world, barrier, steps, check_lossless):
    """device-resident and end-to-end timing of one workload; returns a dict of raw measurements"""
    import torch
    stride = W * 4
    F = len(frames)
    items, host_out, keep = build_items(j2k, jobs, frames, mode)
    job = j2k.Job(ctx, items)
    d_blob = torch.cat([torch.from_numpy(np.ascontiguousarray(j["blob"])) for j in frames] + [torch.zeros(64, dtype=torch.uint8)]).cuda()
    d_out = torch.empty(job.out_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    # correctness guard on the exact bench inputs
    job.run(d_blob.data_ptr(), d_out.data_ptr())
    stream.synchronize()
    first = d_out[: stride * H].cpu().numpy().reshape(H, W, 4)
    if check_lossless:
        src = frames[0]["samples"]
        for c in range(3):
            assert np.array_equal(first[:, :, c], src[c].astype(np.uint8)), "bench inputs decode incorrectly"
    for _ in range(args.warmup):
        job.run(d_blob.data_ptr(), d_out.data_ptr())
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    l0 = ctx.launches
    ev[0].record(stream)
    for _ in range(steps):
        job.run(d_blob.data_ptr(), d_out.data_ptr())
    ev[1].record(stream)
    barrier()
    ms_total = ev[0].elapsed_time(ev[1])
    launches = ctx.launches - l0

    def time_fn(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    reps = max(3, min(steps, 10))
    ent_ms = time_fn(lambda: job.run_entropy(d_blob.data_ptr()), reps)
    dwt_ms = time_fn(lambda: job.run_dwt_mct(d_out.data_ptr()), reps)
    last_ts = []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):                      # the dominant kernel alone: coarser levels refill the ping-pong buffers first
        for lvl in range(LEVELS - 1, 0, -1):
            job.run_level(lvl)
        a.record(stream)
        job.run_level(0, d_out.data_ptr())
        b.record(stream)
        b.synchronize()
        last_ts.append(a.elapsed_time(b))
    last_ms = float(np.mean(last_ts))
    # end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    for _ in range(2):
        job.run_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(steps, 5))
    for _ in range(e2e_steps):
        job.run_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(host_out[0].numpy().reshape(H, W, 4), first), "host path differs from device path"
    res = dict(ms_total=ms_total, steps=steps, launches=int(launches), ent_ms=ent_ms, dwt_ms=dwt_ms, last_ms=last_ms,
               e2e_s=e2e_s, e2e_steps=e2e_steps, h2d=int(d_blob.numel()) - 64, d2h=stride * H * F,
               n_blocks=sum(len(j["cblks"]) for j in frames), F=F, fused_levels=job.fused_levels, coef_bytes=job.coef_bytes,
               plan=job.plan)
    job.close()
    return res


def bind_to_gpu_numa_node(local):
    """pin this rank (and therefore its pinned host buffers: first touch) to the CPUs of the NUMA node its GPU hangs off,
    so that 8 ranks do not push their H2D / D2H traffic across the socket interconnect; returns a note for the JSON line"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from datagen import jobs
    from __graft_entry__ import load_package
    j2k = load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    F = args.frames
    mpix_step = W * H * F * world / 1e6
    alg_bytes = (4 * W * H * NCOMP + W * H * 4) * F               # SURVEY.md 8(d), per launch of the fused kernel
    peak, peak_src = peaks()

    def summarize(m):
        ms_total, e2e_s = reduce_max(m["ms_total"], m["e2e_s"])
        # bytes this variant's fused kernel has to move: its coefficient planes (int32, or int16 when the plan says so) + pixels
        vb = (m["coef_bytes"] * W * H * NCOMP + W * H * 4) * F
        ach = vb / (m["last_ms"] / 1e3) / 1e9
        stg = vb / (m["dwt_ms"] / 1e3) / 1e9
        return dict(value=round(mpix_step * m["steps"] / (ms_total / 1e3), 1), ms_per_step=round(ms_total / m["steps"], 4),
                    e2e=dict(value=round(mpix_step * m["e2e_steps"] / e2e_s, 1), unit=UNIT, h2d_bytes_per_step=m["h2d"],
                             d2h_bytes_per_step=m["d2h"], ms_per_step=round(1e3 * e2e_s / m["e2e_steps"], 3),
                             api="j2kgpu_job_run_host (pinned host buffers)"),
                    gpu_launches=m["launches"], code_blocks_per_step=m["n_blocks"] * world,
                    plan=dict(idwt_levels_in_last_kernel=m["fused_levels"], coef_plane_bytes_per_sample=m["coef_bytes"],
                              last_kernel="k_idwt53_wide (16 columns per lane)" if m["plan"] & 4 else
                                          "k_idwt53_fused (4 columns per lane)" if m["plan"] & 1 else "k_idwt53_stream"),
                    stages_ms=dict(entropy=round(m["ent_ms"], 4), dwt_mct_pack=round(m["dwt_ms"], 4),
                                   last_level_fused=round(m["last_ms"], 4)),
                    roofline=dict(bound="hbm", achieved=round(ach, 1), peak=peak, unit="GB/s", frac=round(ach / peak, 4),
                                  dwt_mct_stage_gbs=round(stg, 1), dwt_mct_stage_frac=round(stg / peak, 4),
                                  bytes_per_launch=vb))

    sampler = ClockSampler(local)
    # ---- headline: REF semantics (bit-identical to the reference's stage functions) -----------------------------------
    base = [build_frame(args.coder, 1002 + 17 * rank + i, threads) for i in range(min(2, F))]
    frames = [base[i % len(base)] for i in range(F)]
    if rank == 0:
        sampler.start()
    m_ref = measure(args, ctx, j2k, jobs, frames, 0, stream, world, barrier, args.steps, args.coder == "ebcot")
    clocks = sampler.stop() if rank == 0 else None
    main = summarize(m_ref)

    extra = {}
    if not args.no_extra:
        # ---- real HTJ2K (ISO/IEC 15444-15), same geometry: conformant codestream, cross-checked with OpenJPEG ----
        iso_base = [jobs.build_iso_job(jobs.synth_image(W, H, NCOMP, PREC, seed=2002 + 17 * rank + i), PREC, TILE, TILE, LEVELS)
                    for i in range(min(2, F))]
        iso_frames = [iso_base[i % len(iso_base)] for i in range(F)]
        m_iso = measure(args, ctx, j2k, jobs, iso_frames, 1, stream, world, barrier, args.steps, True)
        extra["iso_htj2k"] = summarize(m_iso)
        extra["iso_htj2k"]["workload"] = ("configs[1] as a conformant HTJ2K codestream (lossless 5-3, RCT, HT cleanup-only blocks), "
                                          "J2KGPU_MODE_ISO, %d frames; pixels == source image == OpenJPEG decode" % F)
        if rank == 0:
            try:
                import io
                from PIL import Image
                t0 = time.perf_counter()
                im = Image.open(io.BytesIO(iso_base[0]["codestream"]))
                im.load()
                dt = time.perf_counter() - t0
                extra["iso_htj2k"]["openjpeg_cpu"] = dict(value=round(W * H / 1e6 / dt, 2), unit=UNIT,
                                                          note="OpenJPEG 2.5.4 via Pillow, same codestream, 1 frame, context only")
            except Exception as e:  # pragma: no cover
                extra["iso_htj2k"]["openjpeg_cpu"] = dict(error=str(e)[:100])
        # ---- classic EBCOT/MQ, REF semantics ----------------------------------------------------------------
        if args.coder != "ebcot":
            eb = [build_frame("ebcot", 3002 + 17 * rank, threads)]
            eb_frames = [eb[0]] * F
            m_eb = measure(args, ctx, j2k, jobs, eb_frames, 0, stream, world, barrier, max(2, min(args.steps, 3)), True)
            extra["ebcot_ref"] = summarize(m_eb)
            extra["ebcot_ref"]["workload"] = workload_name("ebcot", F)

    if rank == 0:
        cpu_threads = os.cpu_count() or 1
        cpu_t, _ = cpu_decode_time(frames[0], cpu_threads, 1 if args.coder == "ebcot" else 2)
        cpu_val = (W * H / 1e6) / cpu_t
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.coder, F), "mode": "REF", "frames_per_gpu_per_step": F,
                       "code_blocks_per_step": main["code_blocks_per_step"], "l2": "working set > L2 (no flush needed)",
                       "parallelism": "frames sharded across GPUs, no collective", "host": numa_note},
            "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "stages_ms": main["stages_ms"], "plan": main["plan"],
            "roofline": dict(main["roofline"], kernel=main["plan"]["last_kernel"] +
                             (": IDWT levels 1+0 + RCT + DC + clamp + RGBA pack" if main["plan"]["idwt_levels_in_last_kernel"] == 2
                              else ": last IDWT level + RCT + DC + clamp + RGBA pack"),
                             peak_source=peak_src, algorithmic_bytes_per_launch=alg_bytes,
                             # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, from the ncu --set full capture
                             # in profiles/ (the bench's own 16-frame launch: 1.763 GB read + 0.507 GB written against 2.123 GB algorithmic)
                             traffic=int(alg_bytes * NCU_TRAFFIC_RATIO.get(main["plan"]["coef_plane_bytes_per_sample"], 1.0)),
                             traffic_source="profiles/r1_ncu_idwt53_wide_int32_final.txt (16-frame launch: 2.270 GB measured / 2.123 GB algorithmic = 1.069)"),
            "cpu_baseline": {"value": round(cpu_val, 2), "unit": UNIT, "cores": cpu_threads, "kind": "port",
                             "sample": "1 frame of the batch, whole path, all host threads",
                             "note": "C restatement of the reference's Go stage functions (oracle/); Go itself is absent"},
            "clocks": clocks,
        }
        line.update(extra)
        print(json.dumps(line))
    ctx.set_stream(0)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--coder", default="ht", choices=["ht", "ebcot"])
    ap.add_argument("--frames", type=int, default=16, help="frames per GPU per step (the metric is quoted on a batch)")
    ap.add_argument("--no-extra", action="store_true", help="skip the ISO HTJ2K and EBCOT side measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
