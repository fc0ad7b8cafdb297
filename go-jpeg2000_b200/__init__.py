"""go-jpeg2000_b200 -- B200 (sm_100a) JPEG 2000 tile-component decode path behind a C ABI.

The product is csrc/ -> libj2kgpu.so (include/j2kgpu.h).  This Python package is the thin host
mirror used by tests and bench.py: ctypes structs identical to the C ones, and functions named after
the reference functions they stand in for (reference = mrjoshuak/go-jpeg2000):

    t1_decode / ht_decode          entropy.T1.Decode (t1.go:1261) / entropy.HTDecoder.Decode (ht.go:93)
    reconstruct_multilevel53/97    dwt.ReconstructMultiLevel53/97 (dwt.go:534, 561)
    apply_inverse_dwt              tcd.TileDecoder.ApplyInverseDWT (tcd.go:416)
    inverse_rct / inverse_ict      mct.InverseRCT / InverseICT (mct.go:56, 43)
    dc_level_shift_inverse         mct.DCLevelShiftInverse (mct.go:113)
    create_image                   decoder.createImage (decoder.go:417)
    decode_tiles                   decoder.decodeTiles (decoder.go:282) -- the whole path

There is no CPU fallback: if libj2kgpu.so is missing or no CUDA device is present, import / Context()
raise.  (The directory name carries a hyphen, so load it with importlib -- see load_package() in
__graft_entry__.py -- under the module name go_jpeg2000_b200.)
"""
import ctypes as C
import os
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libj2kgpu.so")

MODE_REF, MODE_ISO = 0, 1
PLAN_FUSED, PLAN_FAST_EPILOGUE, PLAN_WIDE, PLAN_COEF16 = 1, 2, 4, 8
BAND_LL, BAND_HL, BAND_LH, BAND_HH = 0, 1, 2, 3
FMT_AUTO, FMT_GRAY8, FMT_GRAY16, FMT_RGBA8, FMT_RGBA64 = 0, 1, 2, 3, 4
E_ARG, E_RANGE, E_UNSUPPORTED, E_CUDA, E_NOMEM, E_NODEVICE, E_INTERNAL = -1, -2, -3, -4, -5, -6, -7
CS_NONE, CS_YCC709, CS_YCC601, CS_PHOTOYCC, CS_CMY, CS_CMYK, CS_YCCK = 0, 1, 2, 3, 4, 5, 6          # J2KGPU_CS_*

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class Image(C.Structure):          # j2k_image_t
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("ncomp", C.c_uint16),
                ("prec", C.c_uint8 * 4), ("sgnd", C.c_uint8 * 4), ("mct", C.c_uint8),
                ("reversible", C.c_uint8), ("nlevels", C.c_uint8), ("ht", C.c_uint8),
                ("mode", C.c_uint8), ("out_fmt", C.c_uint8), ("coef_bits", C.c_uint8), ("colorspace", C.c_uint8),
                ("cblk_style", C.c_uint8), ("rsv", C.c_uint8)]


class TileComp(C.Structure):       # j2k_tilecomp_t
    _fields_ = [("comp", C.c_uint32), ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32),
                ("y1", C.c_uint32), ("coeff_off", C.c_uint64)]


class CBlk(C.Structure):           # j2k_cblk_t
    _fields_ = [("data_off", C.c_uint64), ("data_len", C.c_uint32), ("tilecomp", C.c_uint32),
                ("x0", C.c_uint16), ("y0", C.c_uint16), ("w", C.c_uint16), ("h", C.c_uint16),
                ("band", C.c_uint8), ("level", C.c_uint8), ("num_bps", C.c_uint8), ("num_passes", C.c_uint8),
                ("step", C.c_float), ("len_cleanup", C.c_uint32), ("rsv", C.c_uint32)]


class BatchItem(C.Structure):      # j2k_batch_item_t
    _fields_ = [("image", Image),
                ("tilecomps", C.POINTER(TileComp)), ("n_tilecomps", C.c_uint32),
                ("cblks", C.POINTER(CBlk)), ("n_cblks", C.c_uint32),
                ("blob", u8p), ("blob_len", C.c_uint64),
                ("out_pix", u8p), ("out_stride", C.c_uint64), ("flags", C.c_uint32), ("rsv", C.c_uint32)]


class EncodeParams(C.Structure):
    """j2k_encode_t: the reference encoder's Options as its forward path reads them (encoder.go:197-281, 597-673)"""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("ncomp", C.c_uint16), ("pix_bits", C.c_uint8),
                ("precision", C.c_uint8), ("lossless", C.c_uint8), ("num_resolutions", C.c_uint8), ("cb_x", C.c_uint8),
                ("cb_y", C.c_uint8), ("quality", C.c_int32), ("flags", C.c_uint32), ("rsv", C.c_uint32 * 2)]


ENC_DEVICE_PTRS = 1


ITEM_TILES_ONLY = 1                # J2KGPU_ITEM_TILES_ONLY


class BlkJob(C.Structure):         # j2k_blkjob_t
    _fields_ = [("data_off", C.c_uint64), ("data_len", C.c_uint32), ("out_off", C.c_uint32),
                ("w", C.c_uint16), ("h", C.c_uint16), ("band", C.c_uint8), ("num_bps", C.c_uint8),
                ("rsv0", C.c_uint8), ("rsv1", C.c_uint8), ("len_cleanup", C.c_uint32), ("rsv2", C.c_uint32)]


EXPORTS = [
    "j2kgpu_abi_version", "j2kgpu_create", "j2kgpu_destroy", "j2kgpu_strerror", "j2kgpu_last_error",
    "j2kgpu_set_stream", "j2kgpu_set_option", "j2kgpu_launch_count", "j2kgpu_decode", "j2kgpu_decode_batch",
    "j2kgpu_job_create", "j2kgpu_job_destroy", "j2kgpu_job_blob_bytes", "j2kgpu_job_out_bytes",
    "j2kgpu_job_out_offset", "j2kgpu_job_run", "j2kgpu_job_run_entropy", "j2kgpu_job_run_dwt_mct",
    "j2kgpu_job_run_level", "j2kgpu_job_fused_levels", "j2kgpu_job_coef_bytes", "j2kgpu_job_plan",
    "j2kgpu_job_run_host", "j2kgpu_sync", "j2kgpu_t1_decode_blocks", "j2kgpu_ht_decode_blocks",
    "j2kgpu_idwt53", "j2kgpu_idwt97", "j2kgpu_apply_inverse_dwt", "j2kgpu_inverse_rct",
    "j2kgpu_inverse_ict", "j2kgpu_dc_level_shift_inverse", "j2kgpu_mct_dc_pack",
    "j2kgpu_host_alloc", "j2kgpu_host_free", "j2kgpu_host_register", "j2kgpu_host_unregister",
    "j2kgpu_parse_codestream", "j2kgpu_parsed_free", "j2kgpu_parsed_error", "j2kgpu_parsed_item", "j2kgpu_parsed_info",
    "j2kgpu_decode_codestream", "j2kgpu_decode_codestreams",
    "j2kgpu_encode_block_count", "j2kgpu_encode_preprocess", "j2kgpu_encode_tile",
]

_lib = None


class J2KError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("j2kgpu error %d: %s" % (code, detail))


def lib():
    """Load libj2kgpu.so.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libj2kgpu.so not built: run __graft_entry__.build() "
                              "(make -C go-jpeg2000_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.j2kgpu_strerror.restype = C.c_char_p
        L.j2kgpu_last_error.restype = C.c_char_p
        L.j2kgpu_last_error.argtypes = [C.c_void_p]
        L.j2kgpu_launch_count.restype = C.c_uint64
        L.j2kgpu_launch_count.argtypes = [C.c_void_p]
        for name in ("j2kgpu_job_blob_bytes", "j2kgpu_job_out_bytes"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        L.j2kgpu_job_out_offset.restype = C.c_uint64
        L.j2kgpu_job_out_offset.argtypes = [C.c_void_p, C.c_uint32]
        L.j2kgpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.j2kgpu_destroy.argtypes = [C.c_void_p]
        L.j2kgpu_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.j2kgpu_sync.argtypes = [C.c_void_p]
        L.j2kgpu_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.j2kgpu_decode.argtypes = [C.c_void_p, C.POINTER(Image), C.POINTER(TileComp), C.c_uint32,
                                    C.POINTER(CBlk), C.c_uint32, u8p, C.c_uint64, u8p, C.c_uint64]
        L.j2kgpu_decode_batch.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(BatchItem)]
        L.j2kgpu_job_create.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(BatchItem), C.POINTER(C.c_void_p)]
        L.j2kgpu_job_destroy.argtypes = [C.c_void_p]
        L.j2kgpu_job_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.j2kgpu_job_run_entropy.argtypes = [C.c_void_p, C.c_void_p]
        L.j2kgpu_job_run_dwt_mct.argtypes = [C.c_void_p, C.c_void_p]
        L.j2kgpu_job_run_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.j2kgpu_job_run_host.argtypes = [C.c_void_p, C.POINTER(BatchItem)]
        L.j2kgpu_job_fused_levels.argtypes = [C.c_void_p]
        L.j2kgpu_job_coef_bytes.argtypes = [C.c_void_p]
        L.j2kgpu_job_plan.argtypes = [C.c_void_p]
        for name in ("j2kgpu_t1_decode_blocks", "j2kgpu_ht_decode_blocks"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.POINTER(BlkJob), C.c_uint32, u8p, C.c_uint64,
                                         i32p, C.c_uint64]
        L.j2kgpu_idwt53.argtypes = [C.c_void_p, C.c_int, i32p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.j2kgpu_idwt97.argtypes = [C.c_void_p, C.c_int, f64p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.j2kgpu_apply_inverse_dwt.argtypes = [C.c_void_p, C.c_int, i32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.j2kgpu_inverse_rct.argtypes = [C.c_void_p, i32p, i32p, i32p, C.c_uint64]
        L.j2kgpu_inverse_ict.argtypes = [C.c_void_p, f64p, f64p, f64p, C.c_uint64]
        L.j2kgpu_dc_level_shift_inverse.argtypes = [C.c_void_p, i32p, C.c_uint64, C.c_int]
        L.j2kgpu_mct_dc_pack.argtypes = [C.c_void_p, C.POINTER(Image), C.POINTER(i32p), C.c_int, u8p, C.c_uint64]
        L.j2kgpu_encode_block_count.argtypes = [C.POINTER(EncodeParams)]
        L.j2kgpu_encode_block_count.restype = C.c_uint32
        L.j2kgpu_encode_preprocess.argtypes = [C.c_void_p, C.POINTER(EncodeParams), C.c_void_p, C.c_uint64, C.c_void_p]
        L.j2kgpu_encode_tile.argtypes = [C.c_void_p, C.POINTER(EncodeParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), u8p, C.c_uint32]
        L.j2kgpu_parse_codestream.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        L.j2kgpu_parsed_free.argtypes = [C.c_void_p]
        L.j2kgpu_parsed_free.restype = None
        L.j2kgpu_parsed_error.argtypes = [C.c_void_p]
        L.j2kgpu_parsed_error.restype = C.c_char_p
        L.j2kgpu_parsed_item.argtypes = [C.c_void_p, C.POINTER(BatchItem)]
        L.j2kgpu_parsed_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.j2kgpu_decode_codestream.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint32, u8p, C.c_uint64]
        L.j2kgpu_decode_codestreams.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(u8p), C.POINTER(C.c_uint64), C.c_uint32,
                                                C.POINTER(u8p), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def fmt_bpp(ncomp, prec):
    """bytes per output pixel, as decoder.createImage picks the Go image type (decoder.go:427-523)"""
    return (1 if prec <= 8 else 2) if ncomp == 1 else (4 if prec <= 8 else 8)


def make_image(width, height, ncomp, prec, sgnd=0, mct=1, reversible=1, nlevels=5, ht=0, mode=MODE_REF, coef_bits=0,
               colorspace=0, cblk_style=0):
    im = Image()
    im.width, im.height, im.ncomp = width, height, ncomp
    precs = list(prec) if isinstance(prec, (list, tuple)) else [prec] * ncomp
    sg = list(sgnd) if isinstance(sgnd, (list, tuple)) else [sgnd] * ncomp
    for c in range(min(ncomp, 4)):
        im.prec[c] = precs[c]
        im.sgnd[c] = sg[c]
    im.mct, im.reversible, im.nlevels, im.ht, im.mode, im.out_fmt = mct, reversible, nlevels, ht, mode, FMT_AUTO
    im.coef_bits = coef_bits
    im.colorspace = colorspace
    im.cblk_style = cblk_style
    return im


def _p(a, t):
    return a.ctypes.data_as(t)


class Context:
    """One j2kgpu_ctx (one GPU, one stream, calls serialised) -- the object a Go decoder would hold."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._jobs = weakref.WeakSet()                           # live Job objects: a job must not outlive its context
        rc = lib().j2kgpu_create(device, C.byref(self._h))
        if rc != 0:
            raise J2KError(rc, lib().j2kgpu_strerror(rc).decode())

    def close(self):
        if self._h:
            for job in list(self._jobs):
                job.close()
            lib().j2kgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise J2KError(rc, lib().j2kgpu_last_error(self._h).decode() or lib().j2kgpu_strerror(rc).decode())

    @property
    def launches(self):
        return int(lib().j2kgpu_launch_count(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._check(lib().j2kgpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._check(lib().j2kgpu_sync(self._h))

    def set_option(self, name, value):
        """A/B switch / test hook (j2kgpu_set_option): e.g. set_option("no_fuse", "1"), set_option("chunks", "1,1,2")"""
        self._check(lib().j2kgpu_set_option(self._h, name.encode(), str(value).encode()))

    def options(self, **kw):
        """context manager: set the given options, restore the defaults ("0" / "") afterwards"""
        ctx = self

        class _Scope:
            def __enter__(self_):
                for k, v in kw.items():
                    ctx.set_option(k, v)

            def __exit__(self_, *a):
                for k in kw:
                    ctx.set_option(k, "" if k == "chunks" else ("-1" if k == "host_alpha" else "0"))
        return _Scope()

    # ---- page-locked host memory (j2kgpu_host_*) ----------------------------------------------
    def host_alloc(self, nbytes):
        """a uint8 numpy array over `nbytes` of page-locked memory; release it with host_free(array)"""
        p = C.c_void_p()
        self._check(lib().j2kgpu_host_alloc(self._h, C.c_uint64(nbytes), C.byref(p)))
        a = np.ctypeslib.as_array(C.cast(p, u8p), shape=(max(int(nbytes), 1),))[:nbytes]
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[a.ctypes.data] = p
        return a

    def host_free(self, a):
        p = self._pinned.pop(a.ctypes.data)
        self._check(lib().j2kgpu_host_free(self._h, p))

    def host_register(self, a):
        """page-lock a contiguous numpy array the caller owns (until host_unregister)"""
        self._check(lib().j2kgpu_host_register(self._h, C.c_void_p(a.ctypes.data), C.c_uint64(a.nbytes)))

    def host_unregister(self, a):
        self._check(lib().j2kgpu_host_unregister(self._h, C.c_void_p(a.ctypes.data)))

    # ---- entropy stage ------------------------------------------------------------------------
    def _decode_blocks(self, fn, blocks, mode, style=0):
        """blocks: list of (bytes, w, h, num_bps, band[, num_passes[, len_cleanup]]) -> list of int32 arrays (w*h each);
        style: code-block style bits (ISO mode, classic blocks)"""
        n = len(blocks)
        jobs = (BlkJob * max(n, 1))()
        blob = bytearray()
        off = 0
        for i, blk in enumerate(blocks):
            data, w, h, nbps, band = blk[:5]
            jobs[i] = BlkJob(len(blob), len(data), off, w, h, band, nbps, blk[5] if len(blk) > 5 else 0, style,
                             blk[6] if len(blk) > 6 else 0, 0)
            blob += bytes(data)
            off += w * h
        blob_np = np.frombuffer(bytes(blob), np.uint8) if blob else np.zeros(1, np.uint8)
        out = np.zeros(max(off, 1), np.int32)
        self._check(fn(self._h, mode, jobs, n, _p(blob_np, u8p), len(blob), _p(out, i32p), off))
        res, o = [], 0
        for blk in blocks:
            w, h = blk[1], blk[2]
            res.append(out[o:o + w * h].copy())
            o += w * h
        return res

    def t1_decode_blocks(self, blocks, mode=MODE_REF, style=0):
        return self._decode_blocks(lib().j2kgpu_t1_decode_blocks, blocks, mode, style)

    def ht_decode_blocks(self, blocks, mode=MODE_REF):
        return self._decode_blocks(lib().j2kgpu_ht_decode_blocks, blocks, mode)

    def t1_decode(self, data, w, h, num_bps, band, mode=MODE_REF):
        """entropy.T1.Decode(data, numBPS, bandType) on a fresh NewT1(w, h)  (t1.go:1261)"""
        return self.t1_decode_blocks([(data, w, h, num_bps, band)], mode)[0]

    def ht_decode(self, data, w, h, num_bitplanes=0, band=0, mode=MODE_REF):
        """entropy.HTDecoder.Decode(data, numBitplanes, bandType) on a fresh NewHTDecoder(w, h)  (ht.go:93)"""
        return self.ht_decode_blocks([(data, w, h, num_bitplanes, band)], mode)[0]

    # ---- DWT stage ----------------------------------------------------------------------------
    def reconstruct_multilevel53(self, data, width, height, levels, mode=MODE_REF):
        d = np.array(data, np.int32).reshape(-1)
        assert d.size == width * height
        self._check(lib().j2kgpu_idwt53(self._h, mode, _p(d, i32p), width, height, levels))
        return d

    def reconstruct_multilevel97(self, data, width, height, levels, mode=MODE_REF):
        d = np.array(data, np.float64).reshape(-1)
        assert d.size == width * height
        self._check(lib().j2kgpu_idwt97(self._h, mode, _p(d, f64p), width, height, levels))
        return d

    def apply_inverse_dwt(self, data, width, height, levels, reversible, mode=MODE_REF):
        d = np.array(data, np.int32).reshape(-1)
        assert d.size == width * height
        self._check(lib().j2kgpu_apply_inverse_dwt(self._h, mode, _p(d, i32p), width, height, levels, int(reversible)))
        return d

    # ---- MCT / DC / pack ------------------------------------------------------------------------
    def inverse_rct(self, y, u, v):
        y, u, v = (np.array(a, np.int32).reshape(-1) for a in (y, u, v))
        self._check(lib().j2kgpu_inverse_rct(self._h, _p(y, i32p), _p(u, i32p), _p(v, i32p), y.size))
        return y, u, v

    def inverse_ict(self, y, cb, cr):
        y, cb, cr = (np.array(a, np.float64).reshape(-1) for a in (y, cb, cr))
        self._check(lib().j2kgpu_inverse_ict(self._h, _p(y, f64p), _p(cb, f64p), _p(cr, f64p), y.size))
        return y, cb, cr

    def dc_level_shift_inverse(self, data, precision):
        d = np.array(data, np.int32).reshape(-1)
        self._check(lib().j2kgpu_dc_level_shift_inverse(self._h, _p(d, i32p), d.size, precision))
        return d

    def mct_dc_pack(self, img, comps, apply_tail=True):
        comps = [np.ascontiguousarray(c, np.int32).reshape(-1) for c in comps]
        bpp = fmt_bpp(img.ncomp, img.prec[0])
        stride = img.width * bpp
        pix = np.zeros(max(stride * img.height, 1), np.uint8)
        arr = (i32p * max(len(comps), 1))(*[_p(c, i32p) for c in comps])
        self._check(lib().j2kgpu_mct_dc_pack(self._h, C.byref(img), arr, int(apply_tail), _p(pix, u8p), stride))
        return pix[: stride * img.height]

    def create_image(self, comps, width, height, prec):
        """decoder.createImage (decoder.go:417): planar components -> Pix; raises for a bad component count"""
        img = make_image(width, height, len(comps), prec, sgnd=1, mct=0)
        return self.mct_dc_pack(img, comps, apply_tail=False)

    # ---- forward path (encoder.go:79-281, 597-743) --------------------------------------------------
    def encode_preprocess(self, params, pix, stride=None):
        """extractImageData + preprocess: Go image bytes -> encoder.componentData (ncomp x height x width int32)"""
        pix = np.ascontiguousarray(pix, np.uint8).reshape(-1)
        bpp = (1 if params.ncomp == 1 else 4) * (params.pix_bits // 8)
        stride = stride or params.width * bpp
        planes = np.zeros((params.ncomp, params.height, params.width), np.int32)
        self._check(lib().j2kgpu_encode_preprocess(self._h, C.byref(params), pix.ctypes.data, stride, planes.ctypes.data))
        return planes

    def encode_tile(self, params, pix, stride=None):
        """... + encodeTile up to createTileHeader -> (tileData bytes, bytes per block, bit planes per block)"""
        pix = np.ascontiguousarray(pix, np.uint8).reshape(-1)
        bpp = (1 if params.ncomp == 1 else 4) * (params.pix_bits // 8)
        stride = stride or params.width * bpp
        n = int(lib().j2kgpu_encode_block_count(C.byref(params)))
        lens, bps = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint8)
        cap = params.width * params.height * params.ncomp * 4 + 65536
        out = np.zeros(cap, np.uint8)
        got = C.c_uint64(0)
        self._check(lib().j2kgpu_encode_tile(self._h, C.byref(params), pix.ctypes.data, stride, out.ctypes.data, cap, C.byref(got),
                                             lens.ctypes.data_as(C.POINTER(C.c_uint32)), _p(bps, u8p), n))
        return out[: got.value].copy(), lens[:n], bps[:n]

    # ---- whole path -------------------------------------------------------------------------------
    def decode_tiles(self, img, tilecomps, cblks, blob, out_stride=None):
        """decoder.decodeTiles (decoder.go:282): job tables + compressed blob -> packed pixels (host buffers)"""
        bpp = fmt_bpp(img.ncomp, img.prec[0])
        stride = out_stride or img.width * bpp
        blob = np.ascontiguousarray(blob, np.uint8)
        out = np.zeros(max(stride * img.height, 1), np.uint8)
        self._check(lib().j2kgpu_decode(self._h, C.byref(img), tilecomps, len(tilecomps), cblks, len(cblks),
                                        _p(blob, u8p), blob.size, _p(out, u8p), stride))
        return out[: stride * img.height]

    def decode_batch(self, items):
        arr = (BatchItem * len(items))(*items)
        self._check(lib().j2kgpu_decode_batch(self._h, len(items), arr))

    def decode_codestreams(self, streams, reduce=0, outs=None):
        """raw codestreams -> pixels (j2kgpu_decode_codestreams): tier-2 of the frames runs on host threads inside the
        library, overlapped with the copy / kernel / copy pipeline.  -> list of uint8 arrays (stride = width * bpp);
        `outs` (e.g. page-locked arrays from host_alloc) is filled instead when given."""
        n = len(streams)
        bufs = [s if isinstance(s, np.ndarray) else np.frombuffer(bytes(s), np.uint8) for s in streams]
        dims = [siz_of(b, reduce) for b in bufs]
        strides = [w * fmt_bpp(nc, pr) for (w, h, nc, pr) in dims]
        if outs is None:
            outs = [np.zeros(max(st * d[1], 1), np.uint8) for st, d in zip(strides, dims)]
        cs = (u8p * n)(*[_p(b, u8p) for b in bufs])
        lens = (C.c_uint64 * n)(*[b.size for b in bufs])
        po = (u8p * n)(*[_p(o, u8p) for o in outs])
        so = (C.c_uint64 * n)(*strides)
        self._check(lib().j2kgpu_decode_codestreams(self._h, n, cs, lens, reduce, po, so))
        return [o[: st * d[1]] for o, st, d in zip(outs, strides, dims)]

    def decode_codestream(self, data, reduce=0):
        """jpeg2000.Decode of one raw codestream, device path (j2kgpu_decode_codestream)"""
        return self.decode_codestreams([data], reduce)[0]


class Parsed:
    """j2kgpu_parse_codestream: the codestream front door's host half (main header, tile-part index, tier-2) -- what the
    Go side's buildGPUJob does in the reference layout (codestream.Parser.ReadHeader parser.go:44, ReadTilePartHeader
    parser.go:894, the missing tier-2 of decoder.go:375).  Needs no device.  Keeps `data` alive: the tables point into it."""
    INFO = ("layers", "tiles", "tile_parts", "packets", "progression", "plt_packets", "tlm_tile_parts", "zero_copy")

    def __init__(self, data, reduce=0, threads=0):
        self.data = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        self._h = C.c_void_p()
        buf = self.data if self.data.size else np.zeros(1, np.uint8)
        rc = lib().j2kgpu_parse_codestream(_p(buf, u8p), self.data.size, reduce, threads, C.byref(self._h))
        if rc != 0:
            msg = lib().j2kgpu_parsed_error(self._h).decode() if self._h else ""
            self.close()
            raise J2KError(rc, msg)
        self.item = BatchItem()
        lib().j2kgpu_parsed_item(self._h, C.byref(self.item))
        info = (C.c_uint32 * 8)()
        lib().j2kgpu_parsed_info(self._h, info)
        self.info = dict(zip(self.INFO, (int(v) for v in info)))
        self.image = self.item.image

    def tables(self):
        """(tilecomps, cblks, blob) as numpy copies (structured arrays with the C field names)"""
        it = self.item
        tcs = np.ctypeslib.as_array(it.tilecomps, shape=(it.n_tilecomps,)).copy() if it.n_tilecomps else np.zeros(0, TileComp)
        cbs = np.ctypeslib.as_array(it.cblks, shape=(it.n_cblks,)).copy() if it.n_cblks else np.zeros(0, CBlk)
        blob = np.ctypeslib.as_array(it.blob, shape=(int(it.blob_len),)).copy() if it.blob_len else np.zeros(0, np.uint8)
        return tcs, cbs, blob

    def close(self):
        if self._h:
            lib().j2kgpu_parsed_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def siz_of(data, reduce=0):
    """(width, height, ncomp, precision) from the SIZ marker segment, which directly follows SOC (A.5.1); reduce as in Parsed"""
    d = bytes(data[:64])
    if d[4:8] == b"jP  ":                              # a JP2 file: the codestream sits in its jp2c box
        i = bytes(data).find(b"jp2c\xff\x4f\xff\x51")
        d = bytes(data[i + 4:i + 68]) if i >= 0 else b""
    if len(d) < 43 or d[:4] != b"\xff\x4f\xff\x51":
        raise J2KError(E_ARG, "not a codestream (no SOC + SIZ)")
    be = lambda o, n: int.from_bytes(d[o:o + n], "big")
    sc = 1 << reduce
    return -(-be(8, 4) // sc), -(-be(12, 4) // sc), be(40, 2), (d[42] & 0x7F) + 1


class Job:
    """A validated, uploaded batch (j2kgpu_job): run it with device pointers, repeatedly."""

    def __init__(self, ctx, items):
        self.ctx = ctx
        self._items = (BatchItem * len(items))(*items)
        self._h = C.c_void_p()
        ctx._check(lib().j2kgpu_job_create(ctx._h, len(items), self._items, C.byref(self._h)))
        ctx._jobs.add(self)
        self.blob_bytes = int(lib().j2kgpu_job_blob_bytes(self._h))
        self.out_bytes = int(lib().j2kgpu_job_out_bytes(self._h))
        self.n = len(items)
        self.fused_levels = int(lib().j2kgpu_job_fused_levels(self._h))
        self.coef_bytes = int(lib().j2kgpu_job_coef_bytes(self._h))
        self.plan = int(lib().j2kgpu_job_plan(self._h))

    def out_offset(self, i):
        return int(lib().j2kgpu_job_out_offset(self._h, i))

    def run(self, d_blob_ptr, d_out_ptr):
        self.ctx._check(lib().j2kgpu_job_run(self._h, C.c_void_p(d_blob_ptr), C.c_void_p(d_out_ptr)))

    def run_entropy(self, d_blob_ptr):
        self.ctx._check(lib().j2kgpu_job_run_entropy(self._h, C.c_void_p(d_blob_ptr)))

    def run_dwt_mct(self, d_out_ptr):
        self.ctx._check(lib().j2kgpu_job_run_dwt_mct(self._h, C.c_void_p(d_out_ptr)))

    def run_level(self, lvl, d_out_ptr=0):
        self.ctx._check(lib().j2kgpu_job_run_level(self._h, lvl, C.c_void_p(d_out_ptr)))

    def run_host(self):
        self.ctx._check(lib().j2kgpu_job_run_host(self._h, self._items))

    def close(self):
        if self._h:
            lib().j2kgpu_job_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
