#!/usr/bin/env python3
"""Fuzz the whole front door (tier-2 + kernels) with mutated codestreams on the CPU emulation build (tools/emu): every mutant
must decode or return an error -- never fault, never hang.  python tools/fuzz_codestream.py [seed] [n]   (run `make -C tools/emu` first)"""
import io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
j2k = load_package()
j2k.LIB_PATH = os.environ.get("J2K_EMU_LIB") or os.path.join(ROOT, "tools", "emu", "_build", "libj2kgpu_emu.so")
j2k._lib = None
from datagen import jobs
from PIL import Image


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rng = np.random.default_rng(seed)
    s = jobs.synth_image(96, 80, 3, 8, seed=4)
    buf = io.BytesIO()
    Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8)).save(buf, format="JPEG2000", no_jp2=True, num_resolutions=3, mct=1,
                                                               quality_layers=[10, 1], tile_size=(64, 64), precinct_size=(32, 32))
    streams = [buf.getvalue(), jobs.build_iso_job(s, 8, 64, 64, 2, ht_passes=3, ht_plane=1)["codestream"],
               jobs.build_iso_job(s, 8, 64, 64, 2)["codestream"]]
    ctx = j2k.Context(0)
    ok = err = 0
    t0 = time.time()
    for k in range(n):
        b = bytearray(streams[k % len(streams)])
        for _ in range(int(rng.integers(1, 6))):
            i = int(rng.integers(40, len(b)))                  # past SIZ: the image stays 96 x 80 x 3
            b[i] = int(rng.integers(0, 256))
        if rng.random() < 0.2:
            b = b[: int(rng.integers(60, len(b)))]
        try:
            ctx.decode_codestream(bytes(b))
            ok += 1
        except j2k.J2KError:
            err += 1
    print("fuzz_codestream seed %d: %d mutants, %d decoded, %d refused, %.1f s" % (seed, n, ok, err, time.time() - t0))


if __name__ == "__main__":
    main()
