"""TEST HARNESS: OpenJPEG's encoder through its C API (ctypes on the libopenjp2 that Pillow bundles), for the encoder
parameters Pillow does not pass through -- the code-block style byte of COD (`mode`: 1 BYPASS, 2 RESET, 4 TERMALL,
8 VSC, 16 PREDTERM, 32 SEGSYM) and component sub-sampling.  The streams it writes are decoded by OpenJPEG itself (Pillow)
in the tests, which is what pins the product's handling of those styles.

The layout of opj_cparameters_t is not assumed: the offset of `numresolution, cblockw_init, cblockh_init, mode,
irreversible` is found by looking for the defaults (6, 64, 64, 0, 0) that opj_set_default_encoder_parameters writes.
"""
import ctypes as C
import glob
import os
import tempfile

import numpy as np

_lib = None


def lib():
    global _lib
    if _lib is None:
        import PIL
        cands = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libopenjp2*"))
        if not cands:
            raise OSError("libopenjp2 not found next to Pillow")
        L = C.CDLL(cands[0])
        L.opj_create_compress.restype = C.c_void_p
        L.opj_create_compress.argtypes = [C.c_int]
        L.opj_image_create.restype = C.c_void_p
        L.opj_image_create.argtypes = [C.c_uint32, C.c_void_p, C.c_int]
        L.opj_stream_create_default_file_stream.restype = C.c_void_p
        L.opj_stream_create_default_file_stream.argtypes = [C.c_char_p, C.c_int]
        L.opj_setup_encoder.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.opj_start_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.opj_encode.argtypes = [C.c_void_p, C.c_void_p]
        L.opj_end_compress.argtypes = [C.c_void_p, C.c_void_p]
        L.opj_stream_destroy.argtypes = [C.c_void_p]
        L.opj_destroy_codec.argtypes = [C.c_void_p]
        L.opj_image_destroy.argtypes = [C.c_void_p]
        L.opj_set_default_encoder_parameters.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class CmptParm(C.Structure):            # opj_image_cmptparm_t (2.5.x keeps the deprecated bpp member)
    _fields_ = [(n, C.c_uint32) for n in ("dx", "dy", "w", "h", "x0", "y0", "prec", "bpp", "sgnd")]


class Comp(C.Structure):                # opj_image_comp_t
    _fields_ = [(n, C.c_uint32) for n in ("dx", "dy", "w", "h", "x0", "y0", "prec", "bpp", "sgnd", "resno_decoded", "factor")] + \
               [("data", C.POINTER(C.c_int32)), ("alpha", C.c_uint16)]


class Image(C.Structure):               # opj_image_t
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32), ("numcomps", C.c_uint32),
                ("color_space", C.c_int), ("comps", C.POINTER(Comp)), ("icc", C.c_void_p), ("icc_len", C.c_uint32)]


def _params_offset(buf):
    a = np.frombuffer(buf, np.int32)
    for i in range(len(a) - 5):
        if a[i] == 6 and a[i + 1] == 64 and a[i + 2] == 64 and a[i + 3] == 0 and a[i + 4] == 0:
            return i
    raise RuntimeError("opj_cparameters_t: defaults not found")


def encode(samples, prec=8, mode=0, irreversible=False, num_resolutions=6, mct=None, cblk=(64, 64), tile=None, rates=None,
           sub=None):
    """samples: (ncomp, h, w) integers, or a list of 2-D planes when sub = [(dx, dy), ...] gives the components different
    sampling grids.  Returns the raw codestream (J2K) bytes OpenJPEG writes."""
    L = lib()
    planes = [np.ascontiguousarray(p, np.int32) for p in samples]
    nc = len(planes)
    sub = sub or [(1, 1)] * nc
    W = max(p.shape[1] * s[0] for p, s in zip(planes, sub))
    H = max(p.shape[0] * s[1] for p, s in zip(planes, sub))
    raw = (C.c_uint8 * 32768)()
    L.opj_set_default_encoder_parameters(raw)
    P = np.frombuffer(raw, np.int32)
    o = _params_offset(raw)
    P[o] = num_resolutions
    P[o + 1], P[o + 2] = cblk
    P[o + 3] = mode
    P[o + 4] = 1 if irreversible else 0
    # tcp_numlayers sits 200 floats (tcp_rates[100], tcp_distoratio[100]) + itself before numresolution
    rates = list(rates or [0.0])
    P[o - 201] = len(rates)
    np.frombuffer(raw, np.float32)[o - 200:o - 200 + len(rates)] = rates
    P[5] = 1                                                # cp_disto_alloc
    if tile:
        P[0] = 1                                            # tile_size_on, cp_tx0, cp_ty0, cp_tdx, cp_tdy
        P[3], P[4] = tile
    # tcp_mct is a char near the end of the struct; OpenJPEG derives it in opj_setup_encoder from the component count
    # when it is left at its default of 0 only for RGB colour space images -- set by image colour space below
    cp = (CmptParm * nc)()
    for i, (p, s) in enumerate(zip(planes, sub)):
        cp[i].dx, cp[i].dy = s
        cp[i].w, cp[i].h = p.shape[1], p.shape[0]
        cp[i].prec = cp[i].bpp = prec
    use_mct = (nc == 3 and all(s == (1, 1) for s in sub)) if mct is None else bool(mct)
    img = L.opj_image_create(nc, cp, 1 if nc >= 3 else 2)   # OPJ_CLRSPC_SRGB / GRAY
    im = C.cast(img, C.POINTER(Image)).contents
    im.x0 = im.y0 = 0
    im.x1, im.y1 = W, H
    for i, p in enumerate(planes):
        C.memmove(im.comps[i].data, p.ctypes.data, p.nbytes)
    codec = L.opj_create_compress(0)                        # OPJ_CODEC_J2K
    _set_mct(raw, use_mct)
    fd, path = tempfile.mkstemp(suffix=".j2k")
    os.close(fd)
    try:
        if not L.opj_setup_encoder(codec, raw, img):
            raise RuntimeError("opj_setup_encoder failed")
        st = L.opj_stream_create_default_file_stream(path.encode(), 0)
        ok = L.opj_start_compress(codec, img, st) and L.opj_encode(codec, st) and L.opj_end_compress(codec, st)
        L.opj_stream_destroy(st)
        if not ok:
            raise RuntimeError("OpenJPEG failed to encode")
        with open(path, "rb") as f:
            return f.read()
    finally:
        L.opj_destroy_codec(codec)
        L.opj_image_destroy(img)
        os.unlink(path)


def _set_mct(raw, on):
    """tcp_mct: a char 18 int-sized slots ... after cod_format; located from the default subsampling_dx, subsampling_dy,
    decod_format, cod_format = (1, 1, -1, -1) and the documented member order that follows them (jpwl_* block of 118
    ints, cp_cinema, max_comp_size, cp_rsiz, tp_on, tp_flag, tcp_mct).  tests/test_iso_styles.py checks the COD byte."""
    a = np.frombuffer(raw, np.int32)
    for i in range(len(a) - 4):
        if a[i] == 1 and a[i + 1] == 1 and a[i + 2] == -1 and a[i + 3] == -1:
            raw[(i + 4 + 118 + 3) * 4 + 2] = 1 if on else 0
            return
    raise RuntimeError("opj_cparameters_t: tail not found")
