"""CPU: ISO conformance pins for J2KGPU_MODE_ISO.

(1) OpenJPEG (2.5.4 through Pillow, 2.5.3 through OpenCV -- independent third-party decoders) decodes the HTJ2K
    codestreams written by datagen.codestream (HT cleanup encoder gen_iso_ht.c + tier-2 writer) to the source
    image, bit-exactly: the encoder and the writer are conformant to ISO/IEC 15444-15.
(2) The ISO HT oracle decoder (oracle/iso_ht.c) inverts the same encoder block by block.
Together: the ISO HT oracle is pinned against OpenJPEG transitively (same pattern as SURVEY.md 8c)."""
import io

import numpy as np
import pytest

import oracle_lib as O
from datagen import codestream as cs, iso_ht_encode, jobs

PIL_Image = pytest.importorskip("PIL.Image")


def opj_decode(data):
    im = PIL_Image.open(io.BytesIO(data))
    im.load()
    return np.array(im)


@pytest.mark.parametrize("w,h,ncomp,tw,th,nl", [
    (64, 64, 1, None, None, 0), (64, 64, 1, None, None, 1), (128, 128, 1, None, None, 2),
    (256, 256, 3, None, None, 5), (200, 150, 3, None, None, 3), (512, 512, 3, 256, 256, 5),
    (333, 211, 3, 128, 128, 4), (70, 3, 1, None, None, 2), (5, 90, 3, None, None, 3),
])
def test_openjpeg_decodes_our_htj2k_8bit(w, h, ncomp, tw, th, nl):
    s = jobs.synth_image(w, h, ncomp, 8, seed=w + h)
    data, info = cs.write_htj2k(s, 8, tw, th, nl)
    a = opj_decode(data)
    got = a[None] if ncomp == 1 else np.moveaxis(a, 2, 0)
    assert np.array_equal(got, s)


def test_openjpeg_decodes_our_htj2k_high_bit_depth():
    s = jobs.synth_image(320, 180, 1, 12, seed=3)
    data, _ = cs.write_htj2k(s, 12, None, None, 5)
    assert np.array_equal(opj_decode(data).astype(np.int64), s[0].astype(np.int64) << 4)   # Pillow scales 12 -> 16 bit
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED), s[0])
    s = jobs.synth_image(150, 100, 3, 16, seed=4)
    data, _ = cs.write_htj2k(s, 16, None, None, 4)
    b = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(b[:, :, ::-1].transpose(2, 0, 1), s)


def test_iso_ht_oracle_inverts_encoder_random_blocks():
    rng = np.random.default_rng(1)
    for t in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(1, 16))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.97)] = 0
        enc = iso_ht_encode(d, w, h)
        if not d.any():
            assert enc == b""
            continue
        out, rc = O.iso_ht_decode(enc, w, h, 1)
        assert rc == 0 and np.array_equal(out, d), t
        out3, rc = O.iso_ht_decode(enc, w, h, 3)                   # cleanup at bit-plane 2: magnitudes shifted
        assert rc == 0 and np.array_equal(out3, d * 4)


def test_iso_ht_oracle_rejects_malformed():
    out, rc = O.iso_ht_decode(b"", 8, 8, 1)
    assert rc < 0 and not out.any()
    out, rc = O.iso_ht_decode(b"\x00\x00", 8, 8, 1)              # Scup = 0
    assert rc < 0 and not out.any()
    rng = np.random.default_rng(2)
    for _ in range(200):                                          # garbage never faults
        n = int(rng.integers(2, 500))
        s = rng.integers(0, 256, n).astype(np.uint8)
        scup = int(rng.integers(2, min(n, 4079) + 1))
        s[-1], s[-2] = scup >> 4, (s[-2] & 0xF0) | (scup & 0xF)
        O.iso_ht_decode(s.tobytes(), int(rng.integers(1, 65)), int(rng.integers(1, 65)), 1)
