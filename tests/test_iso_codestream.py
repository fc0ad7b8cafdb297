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
        out3, rc = O.iso_ht_decode(enc, w, h, 3)                   # cleanup at bit-plane 2: magnitudes shifted, mid-point below
        assert rc == 0 and np.array_equal(out3, d * 4 + 2 * np.sign(d))


@pytest.mark.parametrize("w,h,ncomp,tw,nl,passes,P", [
    (200, 150, 1, None, 3, 1, 1), (200, 150, 1, None, 3, 2, 1), (200, 150, 1, None, 3, 3, 1), (256, 256, 3, None, 5, 3, 2),
    (333, 211, 3, 128, 4, 2, 3), (333, 211, 3, 128, 4, 3, 1), (70, 5, 1, None, 2, 3, 1), (5, 90, 3, None, 3, 2, 2),
])
def test_openjpeg_pins_sigprop_magref(w, h, ncomp, tw, nl, passes, P):
    """HT SigProp / MagRef (T.814 7.4, 7.5): OpenJPEG decodes the multi-pass codestreams of the extended writer, and the
    CPU checker (tier-2 parser -> oracle/iso_ht.c refinement passes -> oracle/iso_path.c) reproduces OpenJPEG's pixels
    bit for bit -- including the mid-point reconstruction of a cleanup pass that stops above bit-plane 0"""
    s = jobs.synth_image(w, h, ncomp, 8, seed=w + h)
    job = jobs.build_iso_job(s, 8, tw, tw, nl, ht_passes=passes, ht_plane=P)
    ref = opj_decode(job["codestream"]).reshape(h, w, ncomp)
    got = O.iso_decode_job(job).reshape(h, w, -1)[:, :, :ncomp]
    assert np.array_equal(got, ref)
    via_parser = O.iso_decode_job(jobs.build_iso_job_from_codestream(job["codestream"])).reshape(h, w, -1)[:, :, :ncomp]
    assert np.array_equal(via_parser, ref)
    err = np.abs(ref.astype(int) - np.moveaxis(s, 0, 2)).max()
    assert err <= (3 if (passes, P) == (3, 1) else 40)             # (3, 1): only isolated +-1 coefficients are lost


def test_iso_ht_refinement_checker_matches_writer_statement():
    """block level: the checker's SigProp / MagRef decode equals the reconstruction the writer states for its own stream"""
    from datagen import iso_ht_encode_passes
    rng = np.random.default_rng(5)
    n = 0
    for t in range(400):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        nb = int(rng.integers(2, 12))
        d = rng.integers(-(1 << nb) + 1, 1 << nb, w * h).astype(np.int32)
        d[rng.random(w * h) < rng.uniform(0, 0.97)] = 0
        sm = rng.random(w * h) < 0.3
        d[sm] = rng.integers(-3, 4, int(sm.sum()))
        P, npass = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        enc, lcup, recon = iso_ht_encode_passes(d, w, h, P, npass)
        if not enc:
            continue
        out, rc = O.iso_ht_decode_passes(enc, lcup, w, h, P + 1, npass)
        assert rc == 0 and np.array_equal(out, recon), t
        n += 1
    assert n > 300


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


# ---- classic EBCOT (ISO/IEC 15444-1 Annex C/D): tier-2 parser + oracle/iso_t1.c pinned by OpenJPEG -------------------
def opj_encode(s, **kw):
    a = np.moveaxis(s, 0, 2).astype(np.uint8) if s.shape[0] == 3 else s[0].astype(np.uint8)
    buf = io.BytesIO()
    PIL_Image.fromarray(a).save(buf, format="JPEG2000", no_jp2=True, **kw)
    return buf.getvalue()


def iso_inverse53(plane, w, h, levels):
    """ISO 15444-1 inverse 5-3 on a Mallat plane: per level rows first, then columns = Inverse2D53 of the transpose"""
    a = np.array(plane, np.int32).reshape(h, w).copy()
    dims = [(w, h)]
    for _ in range(levels - 1):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
    for (lw, lh) in (reversed(dims) if levels else []):
        sub = np.ascontiguousarray(a[:lh, :lw].T)
        a[:lh, :lw] = O.inv2d53(sub, lh, lw).reshape(lw, lh).T
    return a.reshape(-1)


def oracle_decode_reversible(data):
    """parser -> iso_t1_decode per block -> Mallat planes -> ISO inverse 5-3 -> inverse RCT -> DC shift -> clamp"""
    h = cs.parse_codestream(data)
    W, H, nc, nl = h["width"], h["height"], h["ncomp"], h["nlevels"]
    ntx = cs.cdiv(W, h["tile_w"])
    out = np.zeros((nc, H, W), np.int64)
    planes = {}
    for b in h["blocks"]:
        t = b["tile"]
        x0, y0 = (t % ntx) * h["tile_w"], (t // ntx) * h["tile_h"]
        x1, y1 = min(x0 + h["tile_w"], W), min(y0 + h["tile_h"], H)
        key = (t, b["comp"])
        if key not in planes:
            planes[key] = (np.zeros((y1 - y0, x1 - x0), np.int32), (x0, y0, x1, y1))
        if not b["passes"]:
            continue
        v = O.iso_t1_decode(b["data"], b["w"], b["h"], b["num_bps"], b["passes"], b["band"])
        v = np.sign(v) * (np.abs(v) >> 1)                                  # reversible reconstruction: m2 / 2, truncating
        planes[key][0][b["py"]:b["py"] + b["h"], b["px"]:b["px"] + b["w"]] = v.reshape(b["h"], b["w"])
    for (t, c), (pl, (x0, y0, x1, y1)) in planes.items():
        out[c, y0:y1, x0:x1] = iso_inverse53(pl.reshape(-1), x1 - x0, y1 - y0, nl).reshape(y1 - y0, x1 - x0)
    if h["mct"] and nc >= 3:
        y, u, v = out[0].copy(), out[1].copy(), out[2].copy()
        g = y - ((u + v) >> 2)
        out[0], out[1], out[2] = v + g, g, u + g
    out += 1 << (h["prec"] - 1)
    return np.clip(out, 0, (1 << h["prec"]) - 1), h


EBCOT_CASES = [
    (64, 64, 1, dict(num_resolutions=1)),
    (200, 150, 3, dict(num_resolutions=4, mct=1)),
    (512, 512, 3, dict(num_resolutions=6, mct=1)),                          # BASELINE configs[0] as a real codestream
    (333, 211, 3, dict(num_resolutions=5, mct=1, tile_size=(128, 128))),
    (300, 200, 1, dict(num_resolutions=3, codeblock_size=(32, 32))),
    (256, 128, 3, dict(num_resolutions=4, mct=1, quality_mode="rates", quality_layers=[30, 10, 1])),   # 3 layers, lossless in total
    (256, 192, 3, dict(num_resolutions=5, mct=1, quality_mode="rates", quality_layers=[25])),          # truncated passes (lossy 5-3)
    (320, 240, 3, dict(num_resolutions=4, mct=1, progression="RLCP", quality_mode="rates", quality_layers=[40, 12])),
]


@pytest.mark.parametrize("w,h,ncomp,kw", EBCOT_CASES)
def test_parser_and_iso_t1_oracle_match_openjpeg(w, h, ncomp, kw):
    """OpenJPEG writes the codestream AND decodes it; our parser + ISO T1 oracle + ISO 5-3 + RCT must give the same
    pixels (also when the stream is truncated: mid-point reconstruction m2 / 2 like OpenJPEG), and the source image
    when the stream is lossless."""
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data = opj_encode(s, irreversible=False, **kw)
    got, hdr = oracle_decode_reversible(data)
    ref = opj_decode(data)
    ref = ref[None] if ncomp == 1 else np.moveaxis(ref, 2, 0)
    assert np.array_equal(got, ref)
    lossless = "quality_layers" not in kw or kw["quality_layers"][-1] <= 1
    assert np.array_equal(got, s) == lossless
    if "quality_layers" in kw:
        assert hdr["layers"] == len(kw["quality_layers"])


def test_parser_reads_our_own_writer():
    s = jobs.synth_image(200, 150, 3, 8, seed=3)
    data, info = cs.write_htj2k(s, 8, 128, 128, 3)
    h = cs.parse_codestream(data)
    key = lambda e: (e["tile"], e["comp"], e["res"], e["band"], e["py"], e["px"])
    a, b = sorted(h["blocks"], key=key), sorted(info["blocks"], key=key)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x["data"] == y["data"] and (x["px"], x["py"], x["w"], x["h"]) == (y["px"], y["py"], y["w"], y["h"])


# ---- irreversible path (9-7, ICT): float32 model with OpenJPEG's constants and operation order ------------------------
F32 = np.float32
K97, IK97 = F32(1.230174105), F32(1.625732422) * F32(0.5)
AL97, BE97, GA97, DE97 = F32(-1.586134342), F32(-0.052980118), F32(0.882911075), F32(0.443506852)


def inv97_1d_f32(x, axis):
    """one line direction of the inverse 9-7, band order in -> interleaved out: low * K, high * (1.625732422 / 2), then
    x += (l + r) * c for -delta, -gamma, -beta, -alpha with mirrored ends (opj_v8dwt_decode)"""
    x = np.moveaxis(x, axis, 0).astype(F32)
    n = x.shape[0]
    if n < 2:
        return np.moveaxis(x, 0, axis)
    nl = (n + 1) // 2
    out = np.empty_like(x)
    out[0::2] = x[:nl] * K97
    out[1::2] = x[nl:] * IK97
    for par, c in ((0, -DE97), (1, -GA97), (0, -BE97), (1, -AL97)):
        idx = np.arange(par, n, 2)
        l = np.where(idx - 1 >= 0, idx - 1, idx + 1)
        r = np.where(idx + 1 < n, idx + 1, idx - 1)
        out[idx] = out[idx] + (out[l] + out[r]) * c
    return np.moveaxis(out, 0, axis)


def oracle_decode_irreversible(data):
    h = cs.parse_codestream(data)
    W, H, nc, nl = h["width"], h["height"], h["ncomp"], h["nlevels"]
    ntx = cs.cdiv(W, h["tile_w"])
    out = np.zeros((nc, H, W), F32)
    planes, gain = {}, {0: 0, 1: 1, 2: 1, 3: 2}
    for b in h["blocks"]:
        t = b["tile"]
        x0, y0 = (t % ntx) * h["tile_w"], (t // ntx) * h["tile_h"]
        x1, y1 = min(x0 + h["tile_w"], W), min(y0 + h["tile_h"], H)
        key = (t, b["comp"])
        if key not in planes:
            planes[key] = (np.zeros((y1 - y0, x1 - x0), F32), (x0, y0, x1, y1))
        if not b["passes"]:
            continue
        v = O.iso_t1_decode(b["data"], b["w"], b["h"], b["num_bps"], b["passes"], b["band"])
        step = F32(2.0 ** (h["prec"] + gain[b["band"]] - b["expn"]) * (1.0 + b["mant"] / 2048.0))      # Annex E.1
        planes[key][0][b["py"]:b["py"] + b["h"], b["px"]:b["px"] + b["w"]] = (v.astype(F32) * (F32(0.5) * step)).reshape(b["h"], b["w"])
    for (t, c), (pl, (x0, y0, x1, y1)) in planes.items():
        a = pl.copy()
        dims = [(x1 - x0, y1 - y0)]
        for _ in range(nl - 1):
            dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
        for (lw, lh) in (reversed(dims) if nl else []):
            a[:lh, :lw] = inv97_1d_f32(inv97_1d_f32(a[:lh, :lw], 1), 0)             # rows first, then columns
        out[c, y0:y1, x0:x1] = a
    if h["mct"] and nc >= 3:
        y, u, v = out[0].copy(), out[1].copy(), out[2].copy()
        out[0], out[1], out[2] = y + v * F32(1.402), (y - u * F32(0.34413)) - v * F32(0.71414), y + u * F32(1.772)
    q = np.rint(out).astype(np.int64) + (1 << (h["prec"] - 1))
    return np.clip(q, 0, (1 << h["prec"]) - 1), h


LOSSY_CASES = [
    (256, 256, 1, dict(num_resolutions=4, quality_mode="rates", quality_layers=[10])),
    (256, 256, 3, dict(num_resolutions=6, mct=1, quality_mode="rates", quality_layers=[20])),
    (200, 150, 3, dict(num_resolutions=4, mct=1, quality_mode="rates", quality_layers=[40, 20, 8])),
    (333, 211, 3, dict(num_resolutions=5, mct=1, tile_size=(128, 128), quality_mode="dB", quality_layers=[38])),
]


@pytest.mark.parametrize("w,h,ncomp,kw", LOSSY_CASES)
def test_irreversible_model_matches_openjpeg(w, h, ncomp, kw):
    """lossy 9-7 codestreams written by OpenJPEG: parser + ISO T1 oracle + dequantisation + float32 9-7 + ICT with
    OpenJPEG's constants reproduce OpenJPEG's decode bit for bit (north_star allows 1 LSB; none is needed)"""
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data = opj_encode(s, irreversible=True, **kw)
    got, hdr = oracle_decode_irreversible(data)
    ref = opj_decode(data)
    ref = ref[None] if ncomp == 1 else np.moveaxis(ref, 2, 0)
    assert np.array_equal(got, ref)
    assert 10 * np.log10(255.0 ** 2 / np.mean((got - s) ** 2.0)) > 30            # and it is a sensible image


# ---- committed golden vectors: codestreams written by OpenJPEG + the pixels OpenJPEG decodes from them ----------------
import os

GOLD_ISO = os.path.join(os.path.dirname(__file__), "golden", "iso_openjpeg.npz")
GOLD_NAMES = ["rgb_lossless", "rgb_lossless_layers_tiles", "rgb_truncated_53", "rgb_lossy_97", "gray_lossy_97_odd", "gray16_lossless"]


@pytest.mark.parametrize("name", GOLD_NAMES)
def test_iso_oracle_reproduces_openjpeg_golden_vectors(name):
    """tests/golden/iso_openjpeg.npz (made by tests/golden/make_golden_iso.py): the parser + ISO oracle decode the stored
    bytes to the stored pixels, bit for bit"""
    g = np.load(GOLD_ISO)
    data, pix = g[name + "_j2k"].tobytes(), g[name + "_pix"]
    hdr = cs.parse_codestream(data)
    got, _ = (oracle_decode_reversible if hdr["reversible"] else oracle_decode_irreversible)(data)
    ref = pix[None] if pix.ndim == 2 else np.moveaxis(pix, 2, 0)
    assert np.array_equal(got, ref.astype(np.int64))


# ---- lossy HTJ2K (9-7 + ICT + dead-zone quantiser + HT cleanup blocks), written by datagen.codestream ------------------
@pytest.mark.parametrize("w,h,ncomp,tw,nl,step", [(128, 128, 1, None, 3, 2.0), (256, 192, 3, None, 5, 1.0), (333, 211, 3, 128, 4, 4.0)])
def test_openjpeg_decodes_our_lossy_htj2k_and_model_matches(w, h, ncomp, tw, nl, step):
    """OpenJPEG decodes the lossy HTJ2K codestreams of our writer to a sensible image (the writer is conformant), and
    the float32 model (parser + ISO HT oracle + mid-point dequantisation (|q| + 1/2) * step + OpenJPEG-order 9-7 / ICT)
    reproduces OpenJPEG's pixels bit for bit"""
    s = jobs.synth_image(w, h, ncomp, 8, seed=w)
    data, _ = cs.write_htj2k(s, 8, tw, tw, nl, lossy_step=step)
    ref = opj_decode(data)
    ref = ref[None] if ncomp == 1 else np.moveaxis(ref, 2, 0)
    assert 10 * np.log10(255.0 ** 2 / np.mean((ref.astype(np.float64) - s) ** 2)) > 33
    hdr = cs.parse_codestream(data)
    assert hdr["ht"] == 1 and hdr["reversible"] == 0
    W, H, nc = hdr["width"], hdr["height"], hdr["ncomp"]
    ntx = cs.cdiv(W, hdr["tile_w"])
    out = np.zeros((nc, H, W), F32)
    planes, gain = {}, {0: 0, 1: 1, 2: 1, 3: 2}
    for b in hdr["blocks"]:
        t = b["tile"]
        x0, y0 = (t % ntx) * hdr["tile_w"], (t // ntx) * hdr["tile_h"]
        x1, y1 = min(x0 + hdr["tile_w"], W), min(y0 + hdr["tile_h"], H)
        key = (t, b["comp"])
        if key not in planes:
            planes[key] = (np.zeros((y1 - y0, x1 - x0), F32), (x0, y0, x1, y1))
        if not b["passes"]:
            continue
        q, rc = O.iso_ht_decode(b["data"], b["w"], b["h"], b["num_bps"])
        assert rc == 0
        q = q.astype(np.int64)
        stp = F32(2.0 ** (hdr["prec"] + gain[b["band"]] - b["expn"]) * (1.0 + b["mant"] / 2048.0))
        f = (np.sign(q) * (2 * np.abs(q) + (q != 0))).astype(F32) * (F32(0.5) * stp)
        planes[key][0][b["py"]:b["py"] + b["h"], b["px"]:b["px"] + b["w"]] = f.reshape(b["h"], b["w"])
    for (t, c), (pl, (x0, y0, x1, y1)) in planes.items():
        a = pl.copy()
        dims = [(x1 - x0, y1 - y0)]
        for _ in range(hdr["nlevels"] - 1):
            dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
        for (lw, lh) in reversed(dims):
            a[:lh, :lw] = inv97_1d_f32(inv97_1d_f32(a[:lh, :lw], 1), 0)
        out[c, y0:y1, x0:x1] = a
    if hdr["mct"] and nc >= 3:
        y, u, v = out[0].copy(), out[1].copy(), out[2].copy()
        out[0], out[1], out[2] = y + v * F32(1.402), (y - u * F32(0.34413)) - v * F32(0.71414), y + u * F32(1.772)
    got = np.clip(np.rint(out).astype(np.int64) + 128, 0, 255)
    assert np.array_equal(got, ref)
