"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference path).

Test infrastructure only -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Builds the library with oracle/Makefile on first use.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# encoder-side helpers come from the input generator (datagen/), re-exported for the tests' convenience
from datagen import (mq_encode, t1_encode, ht_encode, fwd53, fwd97, fwd2d53, fwd2d97, decompose53, decompose97,  # noqa: E402,F401
                     quantize, fwd_rct, fwd_ict, dc_shift_forward, encode_blocks)
_lib = None

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class CBlk(C.Structure):  # == j2k_cblk_t
    _fields_ = [("data_off", C.c_uint64), ("data_len", C.c_uint32), ("tilecomp", C.c_uint32),
                ("x0", C.c_uint16), ("y0", C.c_uint16), ("w", C.c_uint16), ("h", C.c_uint16),
                ("band", C.c_uint8), ("level", C.c_uint8), ("num_bps", C.c_uint8), ("num_passes", C.c_uint8),
                ("step", C.c_float), ("len_cleanup", C.c_uint32), ("rsv", C.c_uint32)]


class TileComp(C.Structure):  # == j2k_tilecomp_t
    _fields_ = [("comp", C.c_uint32), ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32),
                ("y1", C.c_uint32), ("coeff_off", C.c_uint64)]


class Image(C.Structure):  # == j2k_image_t
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("ncomp", C.c_uint16),
                ("prec", C.c_uint8 * 4), ("sgnd", C.c_uint8 * 4), ("mct", C.c_uint8),
                ("reversible", C.c_uint8), ("nlevels", C.c_uint8), ("ht", C.c_uint8),
                ("mode", C.c_uint8), ("out_fmt", C.c_uint8), ("coef_bits", C.c_uint8), ("colorspace", C.c_uint8),
                ("cblk_style", C.c_uint8), ("rsv", C.c_uint8)]


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_create_image.restype = C.c_int
        L.orc_decode_image.restype = C.c_int
        L.orc_t1_zc_lut.restype = u8p
        L.iso_ht_decode.restype = C.c_int
        L.iso_ht_decode_passes.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)




def mq_decode(data, ctxs):
    ctxs = np.ascontiguousarray(ctxs, np.uint8)
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(len(ctxs), np.uint8)
    lib().orc_mq_decode(_p(buf, u8p), len(data), _p(ctxs, u8p), len(ctxs), _p(out, u8p))
    return out




def t1_decode(data, w, h, num_bps, band):
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(w * h, np.int32)
    lib().orc_t1_decode(_p(buf, u8p), len(data), w, h, num_bps, band, _p(out, i32p))
    return out




def ht_decode(data, w, h):
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(w * h, np.int32)
    lib().orc_ht_decode(_p(buf, u8p), len(data), w, h, _p(out, i32p))
    return out


def iso_t1_decode(data, w, h, num_bps, num_passes, band, style=0):
    """ISO EBCOT block decoder -> int32 [w*h] at twice scale with the mid-point (see oracle/iso_t1.c)"""
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(w * h, np.int32)
    rc = lib().iso_t1_decode_style(_p(buf, u8p), len(data), w, h, num_bps, num_passes, band, style, _p(out, i32p))
    assert rc == 0
    return out


def iso_ht_decode(data, w, h, num_bps=1):
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(w * h, np.int32)
    rc = lib().iso_ht_decode(_p(buf, u8p), len(data), w, h, num_bps, _p(out, i32p))
    return out, rc


def iso_ht_decode_passes(data, lcup, w, h, num_bps, num_passes):
    """one HT set: cleanup segment (lcup bytes) + refinement segment -> (sign * Q in quarter units, rc)"""
    buf = np.frombuffer(bytes(data), np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(w * h, np.int32)
    rc = lib().iso_ht_decode_passes(_p(buf, u8p), lcup, len(data) - lcup, w, h, num_bps, num_passes, _p(out, i32p))
    return out, rc


def _inplace(fn, arr, *args):
    fn(arr.ctypes.data_as(C.c_void_p), *args)
    return arr


def inv53(d): d = np.array(d, np.int32); return _inplace(lib().orc_inv53, d, len(d))
def inv97(d): d = np.array(d, np.float64); return _inplace(lib().orc_inv97, d, len(d))
def inv2d53(d, w, h): d = np.array(d, np.int32).reshape(-1); return _inplace(lib().orc_inv2d53, d, w, h)
def inv2d97(d, w, h): d = np.array(d, np.float64).reshape(-1); return _inplace(lib().orc_inv2d97, d, w, h)
def reconstruct53(d, w, h, L): d = np.array(d, np.int32).reshape(-1); return _inplace(lib().orc_reconstruct53, d, w, h, L)
def reconstruct97(d, w, h, L): d = np.array(d, np.float64).reshape(-1); return _inplace(lib().orc_reconstruct97, d, w, h, L)


def apply_inverse_dwt(d, w, h, levels, reversible):
    d = np.array(d, np.int32).reshape(-1)
    return _inplace(lib().orc_apply_inverse_dwt, d, w, h, levels, int(reversible))




def _three(fn, a, b, c, dt):
    a, b, c = (np.array(x, dt).reshape(-1) for x in (a, b, c))
    fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
    return a, b, c


def inv_rct(y, u, v): return _three(lib().orc_inv_rct, y, u, v, np.int32)
def inv_ict(y, cb, cr): return _three(lib().orc_inv_ict, y, cb, cr, np.float64)


def dc_shift_inverse(d, prec):
    d = np.array(d, np.int32).reshape(-1)
    lib().orc_dc_shift_inverse(_p(d, i32p), C.c_size_t(d.size), prec)
    return d




def decoder_tail(comps, mct, reversible, prec, sgnd):
    comps = [np.array(c, np.int32).reshape(-1) for c in comps]
    arr = (i32p * len(comps))(*[_p(c, i32p) for c in comps])
    pr = (C.c_uint8 * 4)(*(list(prec) + [0] * (4 - len(prec))))
    sg = (C.c_uint8 * 4)(*(list(sgnd) + [0] * (4 - len(sgnd))))
    lib().orc_decoder_tail(arr, len(comps), C.c_size_t(comps[0].size), int(mct), int(reversible), pr, sg)
    return comps


def colour_convert(comps, prec, cs):
    """colorspace.go conversions of the YCbCr family (cs 1 = BT.709 matrix, 2 = BT.601), in place on copies"""
    comps = [np.array(c, np.int32).reshape(-1) for c in comps]
    arr = (i32p * len(comps))(*[_p(c, i32p) for c in comps])
    lib().orc_colour_convert(arr, len(comps), C.c_size_t(comps[0].size), int(prec), int(cs))
    return comps


def create_image(comps, w, h, prec):
    comps = [np.ascontiguousarray(c, np.int32).reshape(-1) for c in comps]
    arr = (i32p * max(len(comps), 1))(*[_p(c, i32p) for c in comps])
    pix = np.zeros(w * h * 8, np.uint8)
    bpp = lib().orc_create_image(arr, w, h, len(comps), prec, _p(pix, u8p))
    if bpp < 0:
        raise ValueError("unsupported number of components: %d" % len(comps))
    return pix[: w * h * bpp].copy(), bpp


def decode_image(img, tcs, cbs, blob, out_stride, out_size, threads=1):
    """img: Image; tcs: ctypes array of TileComp; cbs: ctypes array of CBlk; blob: np.uint8."""
    out = np.zeros(out_size, np.uint8)
    blob = np.ascontiguousarray(blob, np.uint8)
    rc = lib().orc_decode_image(C.byref(img), tcs, len(tcs), cbs, len(cbs), _p(blob, u8p),
                                C.c_uint64(blob.size), _p(out, u8p), C.c_uint64(out_stride), threads)
    if rc != 0:
        raise RuntimeError("orc_decode_image rc=%d" % rc)
    return out


def iso_decode_job(job, threads=4, out=None):
    """the ISO-mode whole path (oracle/iso_path.c) on a job dict of datagen.jobs (build_iso_job / build_iso_job_from_codestream)
    -> packed pixels (uint8, width * bpp per row)"""
    from datagen import jobs as J
    img = Image()
    img.width, img.height, img.ncomp = job["width"], job["height"], job["ncomp"]
    for c in range(job["ncomp"]):
        img.prec[c], img.sgnd[c] = job["prec"], job["sgnd"]
    img.mct, img.reversible, img.nlevels, img.ht, img.mode = job["mct"], job["reversible"], job["nlevels"], job["ht"], 1
    img.cblk_style = job.get("cblk_style", 0)
    img.colorspace = job.get("colorspace", 0)
    bpp = (1 if job["prec"] <= 8 else 2) if job["ncomp"] == 1 else (4 if job["prec"] <= 8 else 8)
    stride = job["width"] * bpp
    if out is None:
        out = np.zeros(stride * job["height"], np.uint8)
    tcs, cbs = J.as_ctypes(job["tilecomps"], TileComp), J.as_ctypes(job["cblks"], CBlk)
    blob = np.ascontiguousarray(job["blob"], np.uint8)
    L = lib()
    L.iso_decode_image.restype = C.c_int
    rc = L.iso_decode_image(C.byref(img), tcs, len(tcs), cbs, len(cbs), _p(blob, u8p), C.c_uint64(blob.size),
                            _p(out, u8p), C.c_uint64(stride), threads)
    if rc != 0:
        raise RuntimeError("iso_decode_image rc=%d" % rc)
    return out


# ---- forward path checker (oracle/orc_enc.c) ---------------------------------------------------------------------------
class EncodeParams(C.Structure):  # == j2k_encode_t / orc_encode_t
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("ncomp", C.c_uint16), ("pix_bits", C.c_uint8),
                ("precision", C.c_uint8), ("lossless", C.c_uint8), ("num_resolutions", C.c_uint8), ("cb_x", C.c_uint8),
                ("cb_y", C.c_uint8), ("quality", C.c_int32), ("flags", C.c_uint32), ("rsv", C.c_uint32 * 2)]


def _enc_params(p):
    q = EncodeParams()
    for name, _ in EncodeParams._fields_[:-1]:
        setattr(q, name, getattr(p, name))
    return q


def encode_block_count(p):
    lib().orc_encode_block_count.restype = C.c_uint32
    return int(lib().orc_encode_block_count(C.byref(_enc_params(p))))


def encode_preprocess(p, pix, stride=None):
    """encoder.extractImageData + preprocess -> componentData (ncomp, height, width) int32"""
    pix = np.ascontiguousarray(pix, np.uint8).reshape(-1)
    stride = stride or p.width * (1 if p.ncomp == 1 else 4) * (p.pix_bits // 8)
    planes = np.zeros((p.ncomp, p.height, p.width), np.int32)
    rc = lib().orc_encode_preprocess(C.byref(_enc_params(p)), _p(pix, u8p), C.c_uint64(stride), _p(planes, i32p))
    if rc != 0:
        raise ValueError("bad encoder options")
    return planes


def encode_tile(p, pix, stride=None, threads=4):
    """... + encodeTile's tileData -> (bytes, bytes per block, bit planes per block)"""
    pix = np.ascontiguousarray(pix, np.uint8).reshape(-1)
    stride = stride or p.width * (1 if p.ncomp == 1 else 4) * (p.pix_bits // 8)
    n = encode_block_count(p)
    lens, bps = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint8)
    cap = p.width * p.height * p.ncomp * 4 + 65536
    out = np.zeros(cap, np.uint8)
    lib().orc_encode_tile.restype = C.c_int64
    got = lib().orc_encode_tile(C.byref(_enc_params(p)), _p(pix, u8p), C.c_uint64(stride), _p(out, u8p), C.c_uint64(cap),
                                lens.ctypes.data_as(C.POINTER(C.c_uint32)), _p(bps, u8p), threads)
    if got < 0:
        raise ValueError("orc_encode_tile rc=%d" % got)
    return out[:got].copy(), lens[:n], bps[:n]
