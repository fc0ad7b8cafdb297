"""Shared by the forward-path tests: deterministic Go-image byte buffers and the option sets that are exercised."""
import numpy as np


def go_image(w, h, ncomp, bits, seed):
    """bytes of an image.Gray / Gray16 / RGBA / RGBA64 / NRGBA / NRGBA64 (16-bit samples big-endian): smooth + noise"""
    rng = np.random.default_rng(seed)
    ch = 1 if ncomp == 1 else 4
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(xx / 9.0 + seed) + np.cos(yy / 7.0) + 2) / 4
    a = np.stack([base * (0.5 + 0.12 * c) + rng.random((h, w)) * 0.08 for c in range(ch)], axis=2)
    m = (1 << bits) - 1
    v = np.clip(a * m, 0, m).astype(np.uint32)
    if bits == 8:
        return v.astype(np.uint8).reshape(-1)
    out = np.zeros((h, w, ch, 2), np.uint8)
    out[..., 0] = v >> 8
    out[..., 1] = v & 255
    return out.reshape(-1)


# (width, height, ncomp, pix_bits, lossless, num_resolutions, cb_x, cb_y, quality, precision)
CASES = [
    (96, 64, 3, 8, 1, 6, 4, 4, 0, 0),          # the BASELINE shape in small: RGB 8-bit lossless, 6 resolutions, 64 x 64 blocks
    (100, 75, 3, 8, 0, 6, 4, 4, 75, 0),        # lossy: ICT, 9-7, Quality 75 (DefaultOptions)
    (67, 45, 1, 8, 1, 4, 3, 3, 0, 0),          # grey, odd sizes, 32 x 32 blocks
    (80, 60, 1, 16, 0, 3, 4, 2, 50, 12),       # Gray16 rescaled to 12 bits, 64 x 16 blocks
    (130, 70, 4, 16, 1, 1, 4, 4, 0, 0),        # NRGBA64; NumResolutions 1 -> 5 levels but one "band" (encoder.go:249-252, 601-604)
    (300, 280, 3, 8, 1, 0, 6, 6, 0, 0),        # DefaultOptions block size {6, 6} = 256 x 256 blocks; NumResolutions 0 -> 6
    (1, 1, 1, 8, 1, 6, 4, 4, 0, 0),            # one pixel
    (1, 37, 3, 8, 0, 4, 4, 4, 0, 0),           # one column; Quality 0 -> 100
    (41, 1, 1, 16, 1, 3, 4, 4, 0, 0),          # one row
    (2, 2, 3, 8, 1, 6, 0, 0, 0, 0),            # 4 x 4 blocks
    (257, 129, 3, 16, 0, 5, 4, 4, 20, 10),     # RGBA64 -> 10 bits, lossy
    (64, 64, 4, 8, 0, 6, 5, 3, 100, 0),        # NRGBA lossy (alpha through the 9-7 as well), 128 x 32 blocks
]
