"""How often does CUDA's pow() change a pixel of the four math.Pow colour conversions (colorspace.go:250-427) against the
libm's pow() of the CPU checker?  Prints one JSON line per conversion: samples, pixels that differ, largest difference."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O  # noqa: E402

j2k = importlib.import_module("go-jpeg2000_b200")
ctx = j2k.Context(0)
NAMES = {7: "CIELab", 8: "CIEJab", 9: "e-sRGB", 10: "ROMM-RGB"}
for prec in (8, 12, 16):
    for cs in (7, 8, 9, 10):
        w, h = 2048, 1024
        rng = np.random.default_rng(prec * 100 + cs)
        half = 1 << (prec - 1)
        comps = [rng.integers(-half, half, w * h).astype(np.int32) for _ in range(3)]   # in range after the DC shift
        img = j2k.make_image(w, h, 3, prec, sgnd=0, mct=0, reversible=1, colorspace=cs)
        got = np.asarray(ctx.mct_dc_pack(img, comps, apply_tail=True), np.uint8).astype(np.int32)
        after = O.colour_convert(O.decoder_tail(comps, 0, 1, [prec] * 3, [0] * 3), prec, cs)
        want = np.asarray(O.create_image(after, w, h, prec)[0], np.uint8).astype(np.int32)
        if prec > 8:
            got, want = (got[0::2] << 8) | got[1::2], (want[0::2] << 8) | want[1::2]
        d = np.abs(got - want)
        print(json.dumps({"conversion": NAMES[cs], "precision": prec, "samples": int(d.size), "differ": int((d > 0).sum()),
                          "max_abs_diff": int(d.max()), "distinct_output_values": int(np.unique(got).size)}))
