#!/usr/bin/env python3
"""Writes tests/golden/iso_styles.npz: small codestreams WRITTEN by OpenJPEG 2.5.4 (through its C API, datagen/opj_direct.py)
with the code-block styles BYPASS / RESET / TERMALL / VCAUSAL / PREDTERM / SEGSYM, together with the pixels OpenJPEG itself decodes from them.
The tests decode the stored bytes and must reproduce the stored pixels exactly -- independent of libopenjp2 at test time.

    python tests/golden/make_golden_styles.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from datagen import jobs, opj_direct  # noqa: E402

CASES = {
    "reset": (96, 80, 3, dict(mode=0x02, num_resolutions=3)),
    "vcausal": (96, 80, 3, dict(mode=0x08, num_resolutions=3)),
    "segsym": (96, 80, 1, dict(mode=0x20, num_resolutions=3, cblk=(32, 32))),
    "all_four_layers_tiles": (160, 128, 3, dict(mode=0x3A, num_resolutions=4, tile=(64, 64), rates=[20, 5, 1])),
    "all_four_lossy_97": (128, 128, 3, dict(mode=0x3A, num_resolutions=5, irreversible=True, rates=[30, 10])),
    "termall": (96, 80, 3, dict(mode=0x04, num_resolutions=3)),
    "bypass": (96, 80, 3, dict(mode=0x01, num_resolutions=3)),
    "all_six_layers_tiles": (160, 128, 3, dict(mode=0x3F, num_resolutions=4, tile=(64, 64), rates=[20, 5, 1])),
    "bypass_termall_lossy_97": (128, 128, 3, dict(mode=0x05, num_resolutions=5, irreversible=True, rates=[30, 10])),
}


def main():
    out = {}
    for name, (w, h, nc, kw) in CASES.items():
        s = jobs.synth_image(w, h, nc, 8, seed=sum(map(ord, name)))
        data = opj_direct.encode(s, **kw)
        dec = np.array(Image.open(io.BytesIO(data)))
        out[name + "_j2k"] = np.frombuffer(data, np.uint8)
        out[name + "_pix"] = dec
        print(name, len(data), "bytes", dec.shape, dec.dtype)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "iso_styles.npz"), **out)


if __name__ == "__main__":
    main()
