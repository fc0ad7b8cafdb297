/*
 * j2kgpu.h -- C ABI of the B200 (sm_100a) JPEG 2000 tile-component decode path.
 *
 * This is the drop-in boundary for mrjoshuak/go-jpeg2000 (reference, pure Go):
 * the reference has NO FFI of its own, so each entry point below
 * replaces one Go-internal function and cites it.  A cgo binding
 * (go/j2kgpu_cgo.go, shown in INTEGRATION.md) swaps the body of
 * decoder.decodeTiles (decoder.go:311-359) for one call to j2kgpu_decode().
 *
 * Rules of the boundary (cgo-safe):
 *  - POD only; tables hold offsets, never pointers; host byte order.
 *  - The caller owns every host buffer.  The library copies what it needs before
 *    returning and keeps no host pointer after the call.
 *  - Calls are blocking unless the name ends in _async.  A ctx serialises its
 *    calls (one decode at a time); use one ctx per goroutine / GPU.
 *  - Errors: 0 = OK, negative = J2KGPU_E_*; j2kgpu_last_error(ctx) holds the
 *    sticky detail (CUDA error string, offending index).  Nothing aborts or
 *    panics on malformed input (reference fuzz contract, fuzz_test.go:10-63).
 *  - There is no CPU fallback: without a CUDA device j2kgpu_create() fails.
 */
#ifndef J2KGPU_H
#define J2KGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* colour conversions of the reference that are built (colorspace.go), float64 with the reference's evaluation order and its
 * clampToInt32 rounding.  A job with a conversion decodes its last IDWT level with the general tiled kernel. */
#define J2KGPU_CS_NONE   0   /* sRGB, greyscale, unspecified: no conversion                                          */
#define J2KGPU_CS_YCC709 1   /* ITU-R BT.709 matrix: ColorSpaceSYCC, EYCC, YPbPr60, YPbPr50 (colorspace.go:90-114, 429-482) */
#define J2KGPU_CS_YCC601 2   /* ITU-R BT.601 matrix: ColorSpaceYCbCr2, YCbCr3 (colorspace.go:116-140)                 */
#define J2KGPU_CS_PHOTOYCC 3 /* ColorSpacePhotoYCC (colorspace.go:142-168)                                           */
#define J2KGPU_CS_CMY    4   /* ColorSpaceCMY: maxVal - v (colorspace.go:170-189)                                    */
#define J2KGPU_CS_CMYK   5   /* ColorSpaceCMYK, 4 components; the 4th stays and becomes alpha (colorspace.go:191-217) */
#define J2KGPU_CS_YCCK   6   /* ColorSpaceYCCK, 4 components (colorspace.go:219-250)                                 */
/* The four conversions that go through math.Pow.  Everything but the power function itself is evaluated in the reference's
 * order in float64; CUDA's pow() (<= 2 ulp) stands where Go has math.Pow, so a result can differ from the reference's by
 * 1 LSB when the real value lies within ~1e-13 of a rounding boundary (tolerance in the tests: 1 LSB; measured: 0). */
#define J2KGPU_CS_CIELAB 7   /* ColorSpaceCIELab (colorspace.go:250-292)                                             */
#define J2KGPU_CS_CIEJAB 8   /* ColorSpaceCIEJab: the reference treats J as L* (colorspace.go:319-359), same arithmetic */
#define J2KGPU_CS_ESRGB  9   /* ColorSpaceESRGB (colorspace.go:363-389)                                              */
#define J2KGPU_CS_ROMM   10  /* ColorSpaceROMMRGB (colorspace.go:393-427)                                            */

#define J2KGPU_ABI_VERSION 4

/* ---- status codes ---------------------------------------------------------- */
enum {
    J2KGPU_OK            = 0,
    J2KGPU_E_ARG         = -1,   /* null pointer, zero size, inconsistent table        */
    J2KGPU_E_RANGE       = -2,   /* block outside its tile-component / blob, bad offset  */
    J2KGPU_E_UNSUPPORTED = -3,   /* component count (decoder.go:585), block > 64x64, ... */
    J2KGPU_E_CUDA        = -4,   /* CUDA runtime error, see j2kgpu_last_error()          */
    J2KGPU_E_NOMEM       = -5,
    J2KGPU_E_NODEVICE    = -6,
    J2KGPU_E_INTERNAL    = -7    /* an internal bound was exceeded (reported, never silent)   */
};

/* conformance mode */
enum {
    J2KGPU_MODE_REF = 0,   /* bit-exact to the reference's stage functions (parity contract) */
    J2KGPU_MODE_ISO = 1    /* ISO/IEC 15444-1/-15 semantics (what real codestreams need)     */
};

/* band type, internal/entropy/t1.go:125-130 */
enum { J2KGPU_BAND_LL = 0, J2KGPU_BAND_HL = 1, J2KGPU_BAND_LH = 2, J2KGPU_BAND_HH = 3 };

/* output pixel layout == the Go image types chosen by createImage (decoder.go:417-588) */
enum {
    J2KGPU_FMT_AUTO    = 0,   /* pick from ncomp/prec exactly like createImage          */
    J2KGPU_FMT_GRAY8   = 1,   /* image.Gray    1 B/px                                   */
    J2KGPU_FMT_GRAY16  = 2,   /* image.Gray16  2 B/px big-endian                        */
    J2KGPU_FMT_RGBA8   = 3,   /* image.RGBA    4 B/px, A = 255 for 3 components         */
    J2KGPU_FMT_RGBA64  = 4    /* image.RGBA64  8 B/px big-endian, A = 65535             */
};

/* ---- job description (flattened codestream.Header + tcd model) ------------ */

/* one image: codestream.Header fields the path reads (header.go:8-48) */
typedef struct {
    uint32_t width, height;     /* image area: Xsiz-XOsiz, Ysiz-YOsiz (decoder.go:286-287)   */
    uint16_t ncomp;             /* 1, 3 or 4 (decoder.go:427,585)                           */
    uint8_t  prec[4];           /* ComponentInfo[c].Precision()                              */
    uint8_t  sgnd[4];           /* ComponentInfo[c].IsSigned()                               */
    uint8_t  mct;               /* COD MultipleComponentXf (decoder.go:322)                  */
    uint8_t  reversible;        /* CodingStyle.IsReversible(): 1 = 5-3/RCT, 0 = 9-7/ICT      */
    uint8_t  nlevels;           /* NumDecompositions                                         */
    uint8_t  ht;                /* Header.IsHTJ2K() (header.go:241-257)                      */
    uint8_t  mode;              /* J2KGPU_MODE_*                                             */
    uint8_t  out_fmt;           /* J2KGPU_FMT_*                                              */
    uint8_t  coef_bits;         /* ISO mode: upper bound on the magnitude bits of any coefficient, i.e. max over
                                 * bands of Mb = guard bits + exponent - 1 (QCD/QCC); 0 = unknown.  When <= 14 the
                                 * library keeps the coefficient planes as int16 in HBM (half the DWT read traffic);
                                 * a block whose decoded magnitudes exceed the bound is malformed and decodes to
                                 * zero.  REF mode ignores it (EBCOT bounds come from num_bps, the reference HT
                                 * coder has no bound).                                                       */
    uint8_t  colorspace;        /* J2KGPU_CS_*: conversion to sRGB after the DC shift, the step of decoder.go:350-356
                                 * (getColorConversion, colorspace.go:54-88); 0 = none (sRGB, grey, unknown)    */
    uint8_t  cblk_style;        /* ISO mode, classic (non-HT) blocks: code-block style bits of COD SPcod (Table A.19),
                                 * all six decoded.  With J2KGPU_CBLK_BYPASS or _TERMALL a block consists of several
                                 * codeword segments (with TERMALL one per coding pass; with BYPASS alone passes 0..9,
                                 * then per bit-plane one raw segment for significance + refinement and one MQ segment
                                 * for the cleanup pass): the block's data_len bytes are the segments back to back, and
                                 * they are FOLLOWED in the blob by one little-endian uint32 byte count per segment its
                                 * num_passes touch (not counted in data_len; B.10.7.2 signals these lengths in the
                                 * packet headers).  REF: set 0 */
    uint8_t  rsv;               /* set 0                                                      */
} j2k_image_t;
#define J2KGPU_CBLK_BYPASS   0x01u
#define J2KGPU_CBLK_RESET    0x02u
#define J2KGPU_CBLK_TERMALL  0x04u
#define J2KGPU_CBLK_VCAUSAL  0x08u
#define J2KGPU_CBLK_PREDTERM 0x10u
#define J2KGPU_CBLK_SEGSYM   0x20u

/* one tile-component (tcd.TileComponent, tcd.go:272-283): bounds in image
 * coordinates relative to the image origin; the coefficient plane is
 * (x1-x0) x (y1-y0) int32, row-major. */
typedef struct {
    uint32_t comp;
    uint32_t x0, y0, x1, y1;
    uint64_t coeff_off;         /* reserved (device arena offset); set 0                    */
} j2k_tilecomp_t;

/* one code block (tcd.CodeBlock, tcd.go:18-41) */
typedef struct {
    uint64_t data_off;          /* into blob                                                 */
    uint32_t data_len;          /* 0 => block not coded => all-zero (tcd.go:394-396)         */
    uint32_t tilecomp;          /* index into the tile-component table                       */
    uint16_t x0, y0, w, h;      /* placement inside the tile-component plane; w,h <= 64      */
    uint8_t  band;              /* J2KGPU_BAND_*                                             */
    uint8_t  level;             /* decomposition level of the band (ISO mode)                */
    uint8_t  num_bps;           /* CodeBlock.TotalBitPlanes: bit length of max|x|            */
    uint8_t  num_passes;        /* ISO mode: coding passes present (0 = all); REF: ignored.  HT blocks: 1 = cleanup,
                                 * 2 = + SigProp, 3 = + SigProp + MagRef (one HT set, T.814 clause 7)           */
    float    step;              /* ISO mode dequantisation step; REF: ignored                */
    uint32_t len_cleanup;       /* ISO mode, HT blocks with num_passes > 1 (ABI v4): the data is the cleanup segment
                                 * (Lcup = len_cleanup bytes) followed by the refinement segment (SigProp bytes
                                 * forward, MagRef bytes backward; data_len - len_cleanup bytes).  0 = data_len  */
    uint32_t rsv;               /* set 0                                                     */
} j2k_cblk_t;

/* one image of a batch (cfg5: many frames, one call) */
typedef struct {
    j2k_image_t           image;
    const j2k_tilecomp_t *tilecomps;  uint32_t n_tilecomps;
    const j2k_cblk_t     *cblks;      uint32_t n_cblks;
    const uint8_t        *blob;       uint64_t blob_len;
    uint8_t              *out_pix;    uint64_t out_stride;   /* bytes per output row          */
    uint32_t              flags;      /* J2KGPU_ITEM_* (ABI v4)                               */
    uint32_t              rsv;        /* set 0                                                 */
} j2k_batch_item_t;

/* j2k_batch_item_t.flags */
/* The item carries only SOME tiles of its image (large-image tile sharding, SURVEY.md 8e: the tile loop of
 * decoder.go:315-319 split over several contexts / GPUs).  Only the pixel rectangles those tiles cover are written to
 * out_pix (decoder.go:398-410 copies tile by tile); the rest of the buffer is not touched, so that contexts given
 * disjoint tile subsets of ONE image fill one shared host buffer.  Without the flag a call writes the whole image and
 * pixels no tile covers hold what the reference's zero-initialised planes decode to (decoder.go:305-309); with the flag
 * those pixels are left undefined in the device buffer as well (device-resident runs of such an item).  Several items of
 * one batch call may name the same out_pix with disjoint tile subsets: the call then pipelines the groups of tiles. */
#define J2KGPU_ITEM_TILES_ONLY 1u

/* stage-level block job (entropy stage in isolation) */
typedef struct {
    uint64_t data_off;  uint32_t data_len;
    uint32_t out_off;           /* element offset into the int32 output array               */
    uint16_t w, h;
    uint8_t  band, num_bps;
    uint8_t  rsv0;              /* ISO mode: number of coding passes to decode (0 = all; HT: 1..3); else 0 */
    uint8_t  rsv1;              /* ISO mode, classic blocks: code-block style (j2k_image_t.cblk_style); else 0 */
    uint32_t len_cleanup;       /* ISO mode, HT, passes > 1: Lcup (see j2k_cblk_t.len_cleanup); 0 = data_len */
    uint32_t rsv2;              /* set 0 */
} j2k_blkjob_t;

typedef struct j2kgpu_ctx j2kgpu_ctx;
typedef struct j2kgpu_job j2kgpu_job;

/* ---- context ---------------------------------------------------------------- */
int         j2kgpu_abi_version(void);
int         j2kgpu_create(int device, j2kgpu_ctx **out);
void        j2kgpu_destroy(j2kgpu_ctx *ctx);
const char *j2kgpu_strerror(int code);
const char *j2kgpu_last_error(const j2kgpu_ctx *ctx);
/* Use an externally owned CUDA stream (cudaStream_t, e.g. torch's current stream)
 * for all work of this ctx; NULL restores the ctx's own stream. */
int         j2kgpu_set_stream(j2kgpu_ctx *ctx, void *cuda_stream);
/* A/B switches and test hooks (no_fuse, no_wide, no_fast_epi, coef32, no_preclear, wide_sp, debug_plan, chunks).  Their
 * initial values come from the environment variables J2KGPU_<NAME>, read once in j2kgpu_create; nothing else in the
 * library reads the environment. */
int         j2kgpu_set_option(j2kgpu_ctx *ctx, const char *name, const char *value);
/* kernels launched by this ctx since creation (bench.py's gpu_launches) */
uint64_t    j2kgpu_launch_count(const j2kgpu_ctx *ctx);

/* ---- whole path: replaces decoder.decodeTiles (decoder.go:282-360) ---------- *
 * blocks -> coefficient planes (DecodeCodeBlock tcd.go:393) -> ApplyInverseDWT
 * (tcd.go:416) -> inverse MCT + DC shift (decoder.go:321-348) -> createImage
 * (decoder.go:417).  Host buffers in, host pixels out (H2D/D2H inside).       */
int j2kgpu_decode(j2kgpu_ctx *ctx, const j2k_image_t *img,
                  const j2k_tilecomp_t *tilecomps, uint32_t n_tilecomps,
                  const j2k_cblk_t *cblks, uint32_t n_cblks,
                  const uint8_t *blob, uint64_t blob_len,
                  uint8_t *out_pix, uint64_t out_stride);
/* many images in one call; all items must share ncomp/prec/sgnd/mct/reversible/
 * nlevels/ht/mode (geometry may differ). */
int j2kgpu_decode_batch(j2kgpu_ctx *ctx, uint32_t n_img, const j2k_batch_item_t *items);

/* ---- codestream front door (ABI v4): tier-2 inside the library ----------------- *
 * The reference keeps codestream parsing in Go (internal/codestream) but has no working tier-2: decodeTile is a
 * placeholder (decoder.go:375-380), internal/tcd/t2.go a toy.  These entry points do what the Go-side `buildGPUJob`
 * of INTEGRATION.md has to do -- main header (parser.go:44-124), tile-part index (parser.go:894-982, Psot / TLM), packet
 * headers (tag trees, passes, Lblock, lengths; SOP / EPH; PLT cross-check; all five progression orders, maximal or
 * user-defined precincts, quality layers, classic and HT blocks) -- and fill the tables above in J2KGPU_MODE_ISO, tiles parsed concurrently on host threads.
 * Raw codestreams, or JP2 files: the contiguous codestream box is located here and the enumerated colour space of the JP2
 * header's colour specification box selects j2k_image_t.colorspace (decoder.getColorSpace decoder.go:135-178, applied as in
 * decoder.go:350-356); every other box -- palette, channel definition, resolution -- stays with decoder.readJP2
 * (decoder.go:206-253).  reduce =
 * Config.ReduceResolution (jpeg2000.go:205-207).  Unsupported features (sub-sampling, POC/PPM/PPT, COC with differing components,
 * HT blocks with classic style bits) return J2KGPU_E_UNSUPPORTED; all six classic code-block styles are decoded
 * (j2k_image_t.cblk_style). */
typedef struct j2kgpu_parsed j2kgpu_parsed;
/* *out is always set (free it with j2kgpu_parsed_free); on failure j2kgpu_parsed_error(*out) says why.  No CUDA involved. */
int         j2kgpu_parse_codestream(const uint8_t *cs, uint64_t len, uint32_t reduce, uint32_t threads, j2kgpu_parsed **out);
void        j2kgpu_parsed_free(j2kgpu_parsed *p);
const char *j2kgpu_parsed_error(const j2kgpu_parsed *p);
/* item <- image header, tables and blob of the parsed codestream (pointers into p, and into cs when every block's bytes
 * are contiguous there: cs must then outlive the decode call); out_pix / out_stride are left for the caller */
int         j2kgpu_parsed_item(const j2kgpu_parsed *p, j2k_batch_item_t *item);
/* info: layers, tiles, tile-parts, packets, progression order, packets checked against PLT, tile-parts listed by TLM, zero-copy */
int         j2kgpu_parsed_info(const j2kgpu_parsed *p, uint32_t info[8]);
/* parse + decode; for n codestreams the tier-2 of later frames runs on host threads while the device decodes earlier ones */
int         j2kgpu_decode_codestream(j2kgpu_ctx *ctx, const uint8_t *cs, uint64_t len, uint32_t reduce, uint8_t *out_pix, uint64_t out_stride);
int         j2kgpu_decode_codestreams(j2kgpu_ctx *ctx, uint32_t n, const uint8_t *const *cs, const uint64_t *lens, uint32_t reduce,
                                      uint8_t *const *out_pix, const uint64_t *out_stride);

/* ---- device-resident form (inputs and outputs stay in HBM) ------------------ *
 * A job is a batch whose tables have been validated, flattened and uploaded.
 * j2kgpu_job_run launches the whole path on the ctx stream with DEVICE pointers
 * and returns without synchronising (CUDA-graph friendly).  d_blob holds the
 * items' blobs back to back in item order; d_out holds the items' pixel buffers
 * back to back (item i at byte offset j2kgpu_job_out_offset(job, i)).
 * Lifetime: a job belongs to its context and must be destroyed before it
 * (j2kgpu_job_destroy after j2kgpu_destroy of the owning context is undefined). */
int      j2kgpu_job_create(j2kgpu_ctx *ctx, uint32_t n_img, const j2k_batch_item_t *items, j2kgpu_job **out);
void     j2kgpu_job_destroy(j2kgpu_job *job);
uint64_t j2kgpu_job_blob_bytes(const j2kgpu_job *job);
uint64_t j2kgpu_job_out_bytes(const j2kgpu_job *job);
uint64_t j2kgpu_job_out_offset(const j2kgpu_job *job, uint32_t item);
int      j2kgpu_job_run(j2kgpu_job *job, const void *d_blob, void *d_out);
/* stage subsets of the same job, for per-stage timing: entropy only / DWT+MCT+pack only */
int      j2kgpu_job_run_entropy(j2kgpu_job *job, const void *d_blob);
int      j2kgpu_job_run_dwt_mct(j2kgpu_job *job, void *d_out);
/* one inverse-DWT level of the job (lvl = nlevels-1 .. 0; level 0 is the fused IDWT+MCT+DC+pack kernel and
 * needs d_out); levels must be run coarse to fine after the entropy stage -- used to time a single kernel */
int      j2kgpu_job_run_level(j2kgpu_job *job, int lvl, void *d_out);
/* how the job was planned: 2 when inverse-DWT levels 1 and 0 run as one fused kernel (then run_level(1) is a
 * no-op and run_level(0) runs both), else 1; bytes per element of the coefficient planes in HBM (2 or 4) */
int      j2kgpu_job_fused_levels(const j2kgpu_job *job);
int      j2kgpu_job_coef_bytes(const j2kgpu_job *job);
/* the kernels the job will use, as J2KGPU_PLAN_* bits (for tests and bench reports) */
enum { J2KGPU_PLAN_FUSED = 1,          /* IDWT levels 1+0 + MCT + DC + pack in one kernel (idwt_fused.cu)          */
       J2KGPU_PLAN_FAST_EPILOGUE = 2,  /* ... with the fixed 3 x 8-bit RCT -> RGBA8 epilogue                        */
       J2KGPU_PLAN_WIDE = 4,           /* ... in the 16-columns-per-lane variant (idwt_wide.cu)                     */
       J2KGPU_PLAN_COEF16 = 8 };       /* int16 coefficient planes                                                  */
int      j2kgpu_job_plan(const j2kgpu_job *job);
/* host-buffer run of a prepared job: H2D, kernels and D2H pipelined over chunks of the batch on three streams;
 * blocking.  Give pinned host buffers for the copies to overlap the kernels. */
int      j2kgpu_job_run_host(j2kgpu_job *job, const j2k_batch_item_t *items);
int      j2kgpu_sync(j2kgpu_ctx *ctx);

/* ---- page-locked host memory (ABI v3) --------------------------------------- *
 * The host-buffer entry points copy asynchronously, overlapped with the kernels,
 * only from / to page-locked memory; pageable memory (a Go slice, malloc) still
 * works but every copy is staged by the driver and blocks.  Either allocate the
 * blob / pixel buffers here, or page-lock memory the caller already owns (the
 * Pix of an image.RGBA, kept alive and unmoved for the duration: runtime.Pinner)
 * and release it before the memory is freed.  The reference has no counterpart:
 * its buffers are plain Go slices (decoder.go:296-305, 417-588).                */
int      j2kgpu_host_alloc(j2kgpu_ctx *ctx, uint64_t bytes, void **out);
int      j2kgpu_host_free(j2kgpu_ctx *ctx, void *p);
int      j2kgpu_host_register(j2kgpu_ctx *ctx, void *p, uint64_t bytes);
int      j2kgpu_host_unregister(j2kgpu_ctx *ctx, void *p);

/* ---- per-stage entry points (differential tests against the oracle) --------- */
/* entropy.T1.Decode (t1.go:1261) over n blocks; out = concatenated w*h int32 */
int j2kgpu_t1_decode_blocks(j2kgpu_ctx *ctx, int mode, const j2k_blkjob_t *jobs, uint32_t n,
                            const uint8_t *blob, uint64_t blob_len, int32_t *out, uint64_t out_len);
/* entropy.HTDecoder.Decode (ht.go:93) over n blocks */
int j2kgpu_ht_decode_blocks(j2kgpu_ctx *ctx, int mode, const j2k_blkjob_t *jobs, uint32_t n,
                            const uint8_t *blob, uint64_t blob_len, int32_t *out, uint64_t out_len);
/* dwt.ReconstructMultiLevel53 (dwt.go:534): in place on a host buffer */
int j2kgpu_idwt53(j2kgpu_ctx *ctx, int mode, int32_t *data, uint32_t width, uint32_t height, uint32_t levels);
/* dwt.ReconstructMultiLevel97 (dwt.go:561): float64, in place */
int j2kgpu_idwt97(j2kgpu_ctx *ctx, int mode, double *data, uint32_t width, uint32_t height, uint32_t levels);
/* tcd.TileDecoder.ApplyInverseDWT (tcd.go:416): int32 plane in place, 9-7 through float64 */
int j2kgpu_apply_inverse_dwt(j2kgpu_ctx *ctx, int mode, int32_t *data, uint32_t width, uint32_t height,
                             uint32_t levels, int reversible);
/* mct.InverseRCT (mct.go:56) / mct.InverseICT (mct.go:43) / mct.DCLevelShiftInverse (mct.go:113) */
int j2kgpu_inverse_rct(j2kgpu_ctx *ctx, int32_t *y, int32_t *u, int32_t *v, uint64_t n);
int j2kgpu_inverse_ict(j2kgpu_ctx *ctx, double *y, double *cb, double *cr, uint64_t n);
int j2kgpu_dc_level_shift_inverse(j2kgpu_ctx *ctx, int32_t *data, uint64_t n, int precision);
/* decoder.go:321-348 + createImage (decoder.go:417): planar int32 components -> Pix.
 * apply_tail != 0 runs inverse MCT + DC shift first; 0 packs the planes as they are. */
int j2kgpu_mct_dc_pack(j2kgpu_ctx *ctx, const j2k_image_t *img, const int32_t *const *comps,
                       int apply_tail, uint8_t *out_pix, uint64_t out_stride);

/* ---- forward path ------------------------------------------------------------------------------ *
 * What encoder.encode does between extractImageData and createTileHeader (encoder.go:79-281, 597-743), byte for byte:
 * Go image bytes -> component planes (Options.Precision rescale, encoder.go:197-211) -> DC shift, ForwardRCT / ForwardICT
 * (rounded half away from zero), DecomposeMultiLevel53 / 97 (rows then columns, dense-prefix levels), v / stepSize +- 0.5
 * -> encodeTile's code-block list (component, resolution, band, block row, block column; blocks cut from the top-left
 * corner of the component plane whatever the band, encoder.go:763-796, kept as written) -> T1.SetData + T1.Encode per
 * block -> the blocks' bytes appended in list order (`tileData`, the argument of createTileHeader).  Marker segments,
 * the tile-part header and the JP2 boxes stay in Go. */
typedef struct {
    uint32_t width, height;     /* img.Bounds().Dx() / Dy()                                                        */
    uint16_t ncomp;             /* 1: image.Gray / Gray16; 3: image.RGBA / RGBA64 (alpha ignored, encoder.go:109,126);
                                 * 4: image.NRGBA / NRGBA64 (alpha is component 3)                                  */
    uint8_t  pix_bits;          /* 8 or 16: sample size of the Go image type (16-bit samples big-endian); pixels are
                                 * ncomp == 1 ? 1 : 4 samples wide                                                  */
    uint8_t  precision;         /* Options.Precision (0 = keep)                                                     */
    uint8_t  lossless;          /* Options.Lossless: 5-3 + RCT, else 9-7 + ICT + Quality                            */
    uint8_t  num_resolutions;   /* Options.NumResolutions; 0 -> 6 (and <= 1 -> 5 decomposition levels, encoder.go:249-252) */
    uint8_t  cb_x, cb_y;        /* Options.CodeBlockSize: blocks of 1 << (v + 2) samples (encoder.go:606-607), v <= 6 */
    int32_t  quality;           /* Options.Quality; <= 0 -> 100 (encoder.go:265-268)                                */
    uint32_t flags;             /* J2KGPU_ENC_*                                                                     */
    uint32_t rsv[2];            /* set 0                                                                            */
} j2k_encode_t;
#define J2KGPU_ENC_DEVICE_PTRS 1u   /* pix and out (planes) are device pointers; blk_len / blk_bps stay host arrays   */
/* number of entries of encodeTile's job list for these options (0: invalid options) */
uint32_t j2kgpu_encode_block_count(const j2k_encode_t *p);
/* extractImageData + preprocess: planes receives ncomp planes of width x height int32 (encoder.componentData) */
int j2kgpu_encode_preprocess(j2kgpu_ctx *ctx, const j2k_encode_t *p, const uint8_t *pix, uint64_t pix_stride, int32_t *planes);
/* ... + encodeTile up to createTileHeader: out receives *out_len bytes (J2KGPU_E_ARG with *out_len = the size needed when
 * out_cap is too small).  blk_len / blk_bps (optional, n_blk >= the block count): bytes per block (0 = the reference's
 * nil: all-zero block) and the bit-plane count T1.Encode derived (t1_fast5.go:23-27), which the reference leaves out of
 * its codestream although its decoder needs it (t1.go:1261). */
int j2kgpu_encode_tile(j2kgpu_ctx *ctx, const j2k_encode_t *p, const uint8_t *pix, uint64_t pix_stride, uint8_t *out,
                       uint64_t out_cap, uint64_t *out_len, uint32_t *blk_len, uint8_t *blk_bps, uint32_t n_blk);

#ifdef __cplusplus
}
#endif
#endif /* J2KGPU_H */
