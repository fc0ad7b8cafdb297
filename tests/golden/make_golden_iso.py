#!/usr/bin/env python3
"""Writes tests/golden/iso_openjpeg.npz: small codestreams WRITTEN by OpenJPEG 2.5.4 (through Pillow) together with the
pixels OpenJPEG itself decodes from them.  The ISO-mode tests decode the stored bytes (harness tier-2 + GPU kernels, or
tier-2 + CPU oracle) and must reproduce the stored pixels exactly -- independent of Pillow being installed.

    python tests/golden/make_golden_iso.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from datagen import jobs  # noqa: E402

CASES = {
    "rgb_lossless": (128, 96, 3, 8, dict(irreversible=False, num_resolutions=4, mct=1)),
    "rgb_lossless_layers_tiles": (160, 128, 3, 8, dict(irreversible=False, num_resolutions=3, mct=1, tile_size=(64, 64),
                                                       quality_mode="rates", quality_layers=[20, 5, 1])),
    "rgb_truncated_53": (128, 128, 3, 8, dict(irreversible=False, num_resolutions=5, mct=1, quality_mode="rates", quality_layers=[16])),
    "rgb_lossy_97": (128, 128, 3, 8, dict(irreversible=True, num_resolutions=5, mct=1, quality_mode="rates", quality_layers=[30, 10])),
    "gray_lossy_97_odd": (117, 83, 1, 8, dict(irreversible=True, num_resolutions=4, quality_mode="dB", quality_layers=[36])),
    "gray16_lossless": (96, 64, 1, 16, dict(irreversible=False, num_resolutions=3)),
}


def main():
    out = {}
    for name, (w, h, nc, prec, kw) in CASES.items():
        s = jobs.synth_image(w, h, nc, prec, seed=sum(map(ord, name)))
        if prec == 16:
            im = Image.fromarray(s[0].astype(np.uint16))
        else:
            im = Image.fromarray(np.moveaxis(s, 0, 2).astype(np.uint8) if nc == 3 else s[0].astype(np.uint8))
        buf = io.BytesIO()
        im.save(buf, format="JPEG2000", no_jp2=True, **kw)
        data = buf.getvalue()
        dec = np.array(Image.open(io.BytesIO(data)))
        out[name + "_j2k"] = np.frombuffer(data, np.uint8)
        out[name + "_pix"] = dec
        print(name, len(data), "bytes", dec.shape, dec.dtype)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "iso_openjpeg.npz"), **out)


if __name__ == "__main__":
    main()
