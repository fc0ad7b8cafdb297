"""datagen -- synthetic-input generator (reference ENCODER side restated in C, see datagen.h).

Manufactures what the reference encoder would hand its decoder: coefficient planes (DC shift ->
RCT/ICT -> dense-prefix multi-level DWT -> quantiser, encoder.go:216-281) and per-block bitstreams
(T1.Encode / HTEncoder.Encode), plus the flat job tables of include/j2kgpu.h.  Used by tests/ and by
bench.py to build inputs.  It is not the oracle and not the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class GenBlk(C.Structure):
    _fields_ = [("plane_off", C.c_uint64), ("stride", C.c_uint32), ("x0", C.c_uint16), ("y0", C.c_uint16),
                ("w", C.c_uint16), ("h", C.c_uint16), ("band", C.c_uint8), ("ht", C.c_uint8),
                ("r0", C.c_uint8), ("r1", C.c_uint8)]


def build(force=False):
    so = os.path.join(HERE, "libdatagen.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "libdatagen.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.gen_mq_encode.restype = C.c_int
        _lib.gen_t1_encode.restype = C.c_int
        _lib.gen_ht_encode.restype = C.c_int
        _lib.gen_encode_blocks.restype = C.c_int64
        _lib.gen_iso_ht_encode.restype = C.c_int
        _lib.gen_iso_ht_encode_passes.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def mq_encode(ctxs, bits):
    ctxs = np.ascontiguousarray(ctxs, np.uint8)
    bits = np.ascontiguousarray(bits, np.uint8)
    out = np.zeros(2 * len(bits) + 64, np.uint8)
    n = lib().gen_mq_encode(_p(ctxs, u8p), _p(bits, u8p), len(bits), _p(out, u8p), len(out))
    assert n >= 0
    return out[:n].tobytes()


def t1_encode(coeffs, w, h, band):
    """T1.SetData + T1.Encode -> (bytes, num_bps); b'' for an all-zero block (Go nil)."""
    c = np.ascontiguousarray(coeffs, np.int32).reshape(-1)
    assert c.size == w * h
    out = np.zeros(w * h * 4 + 16384, np.uint8)
    nb = C.c_int(0)
    n = lib().gen_t1_encode(_p(c, i32p), w, h, band, _p(out, u8p), len(out), C.byref(nb))
    assert n >= 0
    return out[:n].tobytes(), nb.value


def ht_encode(coeffs, w, h, band=0):
    c = np.ascontiguousarray(coeffs, np.int32).reshape(-1)
    assert c.size == w * h
    out = np.zeros(max(w * h * 2, 64) * 2 + 64, np.uint8)
    n = lib().gen_ht_encode(_p(c, i32p), w, h, band, _p(out, u8p), len(out))
    if n < 0:
        raise OverflowError("reference HT encoder would index out of range")
    return out[:n].tobytes()


def iso_ht_encode(coeffs, w, h):
    """conformant HT cleanup-pass encoder (T.814); b'' for an all-zero block"""
    c = np.ascontiguousarray(coeffs, np.int32).reshape(-1)
    assert c.size == w * h
    out = np.zeros(w * h * 5 + 8192, np.uint8)
    n = lib().gen_iso_ht_encode(_p(c, i32p), w, h, _p(out, u8p), len(out))
    if n < 0:
        raise ValueError("HT encode failed (magnitude too large?)")
    return out[:n].tobytes()


def iso_ht_encode_passes(coeffs, w, h, P, npasses):
    """one HT set (T.814): cleanup at bit-plane P, SigProp (npasses >= 2) and MagRef (npasses == 3) at bit-plane P - 1
    -> (bytes = cleanup segment + refinement segment, Lcup, expected reconstruction in quarter units, signed)"""
    c = np.ascontiguousarray(coeffs, np.int32).reshape(-1)
    assert c.size == w * h
    out = np.zeros(w * h * 6 + 8192, np.uint8)
    recon = np.zeros(w * h, np.int32)
    lcup = C.c_int(0)
    n = lib().gen_iso_ht_encode_passes(_p(c, i32p), w, h, P, npasses, _p(out, u8p), len(out), C.byref(lcup), _p(recon, i32p))
    if n < 0:
        raise ValueError("HT encode failed")
    return out[:n].tobytes(), lcup.value, recon


def _inplace(fn, arr, *args):
    fn(arr.ctypes.data_as(C.c_void_p), *args)
    return arr


def fwd53(d): d = np.array(d, np.int32); return _inplace(lib().gen_fwd53, d, len(d))
def fwd97(d): d = np.array(d, np.float64); return _inplace(lib().gen_fwd97, d, len(d))
def fwd2d53(d, w, h): d = np.array(d, np.int32).reshape(-1); return _inplace(lib().gen_fwd2d53, d, w, h)
def fwd2d97(d, w, h): d = np.array(d, np.float64).reshape(-1); return _inplace(lib().gen_fwd2d97, d, w, h)
def decompose53(d, w, h, L): d = np.array(d, np.int32).reshape(-1); return _inplace(lib().gen_decompose53, d, w, h, L)
def decompose97(d, w, h, L): d = np.array(d, np.float64).reshape(-1); return _inplace(lib().gen_decompose97, d, w, h, L)


def quantize(d, step):
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(d.size, np.int32)
    lib().gen_quantize(_p(d, f64p), C.c_double(step), _p(out, i32p), C.c_size_t(d.size))
    return out


def _three(fn, a, b, c, dt):
    a, b, c = (np.array(x, dt).reshape(-1) for x in (a, b, c))
    fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
    return a, b, c


def fwd_rct(r, g, b): return _three(lib().gen_fwd_rct, r, g, b, np.int32)
def fwd_ict(r, g, b): return _three(lib().gen_fwd_ict, r, g, b, np.float64)


def dc_shift_forward(d, prec):
    d = np.array(d, np.int32).reshape(-1)
    lib().gen_dc_shift_forward(_p(d, i32p), C.c_size_t(d.size), prec)
    return d


def encode_blocks(planes, blks, threads=None):
    """planes: flat int32 array; blks: list of (plane_off, stride, x0, y0, w, h, band, ht).
    -> (blob uint8 array, offs uint64, lens uint32, nbps uint8)"""
    planes = np.ascontiguousarray(planes, np.int32).reshape(-1)
    n = len(blks)
    arr = (GenBlk * max(n, 1))()
    cap = 0
    for i, (po, st, x0, y0, w, h, band, ht) in enumerate(blks):
        arr[i] = GenBlk(po, st, x0, y0, w, h, band, ht, 0, 0)
        cap += w * h * 4 + 16384 if not ht else max(w * h * 2, 64) * 2 + 64
    out = np.zeros(cap + 16, np.uint8)
    offs = np.zeros(max(n, 1), np.uint64)
    lens = np.zeros(max(n, 1), np.uint32)
    nbps = np.zeros(max(n, 1), np.uint8)
    if threads is None:
        threads = os.cpu_count() or 1
    tot = lib().gen_encode_blocks(_p(planes, i32p), arr, n, _p(out, u8p), C.c_uint64(out.size),
                                  offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  lens.ctypes.data_as(C.POINTER(C.c_uint32)), _p(nbps, u8p), threads)
    if tot < 0:
        raise RuntimeError("gen_encode_blocks failed")
    return out[:tot].copy(), offs[:n], lens[:n], nbps[:n]
