#!/usr/bin/env python3
"""Development aid: fuzz the HT block decoders of the CPU emulation build (tools/emu) against the oracle.
    python tools/fuzz_ht.py ref [seed]     reference "HT" coder: arbitrary bytes (0xFF / 0x00 heavy, plausible scup)
    python tools/fuzz_ht.py iso [seed]     ISO cleanup decoder: encoder streams with damaged MagSgn / VLC / MEL bytes
Needs `make -C tools/emu` (run_emu.py builds it).  Not part of the tests, the bench or the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from datagen import iso_ht_encode
from __graft_entry__ import load_package

mod = load_package()
mod.LIB_PATH = os.path.join(ROOT, "tools", "emu", "_build", "libj2kgpu_emu.so")
mod._lib = None
ctx = mod.Context(0)
kind = sys.argv[1] if len(sys.argv) > 1 else "ref"
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0


def fuzz_ref():
    global bad
    for rnd in range(12):
        blocks = []
        for _ in range(500):
            n = int(rng.integers(2, 1500))
            s = rng.integers(0, 256, n).astype(np.uint8)
            r = rng.random()
            if r < 0.3:
                s[rng.random(n) < 0.3] = 0xFF
            elif r < 0.5:
                s[rng.random(n) < 0.7] = 0x00
            elif r < 0.6:
                s[:] = 0xFF
            if rng.random() < 0.9:
                scup = int(rng.integers(2, min(n, 4095) + 1))
                s[-1] = scup & 0xFF
                s[-2] = (s[-2] & 0xF0) | (scup >> 8)
            blocks.append((s.tobytes(), int(rng.integers(1, 65)), int(rng.integers(1, 65)), 0, 0))
        for i, ((s, w, h, _, _), got) in enumerate(zip(blocks, ctx.ht_decode_blocks(blocks))):
            if not np.array_equal(got, O.ht_decode(s, w, h)):
                bad += 1
                print("MISMATCH", rnd, i, w, h, len(s))


def fuzz_iso():
    global bad
    for rnd in range(8):
        blocks = []
        for _ in range(400):
            w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
            nb = int(rng.integers(1, 20))
            d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int64).astype(np.int32)
            d[rng.random(w * h) < rng.uniform(0, 0.95)] = 0
            s = np.frombuffer(iso_ht_encode(d, w, h), np.uint8).copy()
            if s.size < 2:
                continue
            scup = (int(s[-1]) << 4) + (int(s[-2]) & 0x0F)
            L = s.size - scup
            r = rng.random()
            if r < 0.5 and L > 0:                           # damage the MagSgn segment
                for _ in range(int(rng.integers(1, 6))):
                    s[int(rng.integers(0, L))] = rng.choice([0xFF, 0x7F, 0x00, int(rng.integers(0, 256))])
            elif r < 0.8 and s.size > 2:                    # damage anywhere (VLC / MEL included)
                for _ in range(int(rng.integers(1, 4))):
                    s[int(rng.integers(0, s.size - 2))] = int(rng.integers(0, 256))
            blocks.append((s.tobytes(), w, h, int(rng.integers(1, 4)), 0))
        for i, ((s, w, h, nbp, _), got) in enumerate(zip(blocks, ctx.ht_decode_blocks(blocks, mode=1))):
            if not np.array_equal(got, O.iso_ht_decode(s, w, h, nbp)[0]):
                bad += 1
                print("MISMATCH", rnd, i, w, h, len(s))


(fuzz_iso if kind == "iso" else fuzz_ref)()
print("done, mismatches =", bad)
sys.exit(1 if bad else 0)
