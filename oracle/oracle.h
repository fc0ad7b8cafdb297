/*
 * oracle.h -- CPU restatement of the go-jpeg2000 tile-component decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the timed CPU baseline.  The CUDA path never calls into it.
 *
 * Every function restates a function of the reference
 * (mrjoshuak/go-jpeg2000, read-only at /root/reference) and cites the
 * file:line it follows.  The reference is pure Go and there is no Go
 * toolchain in this image, so the reference itself cannot be run here:
 * the restatement is pinned by replaying the reference's own exact
 * round-trip tests (tests/test_oracle_*.py).
 * Parity status:  MQ / T1 / DWT / RCT / ICT / DC / pack = pinned by those
 * properties;  HT (ht.go) = "parity unpinned" (the reference tests assert
 * no decoded value, ht_test.go:57-75) -- the code is the only contract.
 *
 * Go semantics honoured throughout: int32/uint32 arithmetic wraps, shifts
 * by >= width give 0 (arithmetic >> gives the sign fill), float64 -> int32
 * conversion truncates toward zero, float64 expressions are evaluated
 * without FMA contraction (build with -ffp-contract=off).
 */
#ifndef J2K_ORACLE_H
#define J2K_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* band types, internal/entropy/t1.go:125-130 */
enum { ORC_BAND_LL = 0, ORC_BAND_HL = 1, ORC_BAND_LH = 2, ORC_BAND_HH = 3 };

/* ---- MQ coder (internal/entropy/mqc.go) -------------------------------- */
/* Decode n decisions with the given context sequence.  mqc.go:370-497 */
void orc_mq_decode(const uint8_t *data, int len, const uint8_t *ctxs, int n, uint8_t *bits_out);

/* ---- EBCOT tier-1 (internal/entropy/t1.go, t1_luts.go, t1_fast5.go) ---- */
/* T1.Decode on a fresh T1 (t1.go:1261-1292).  out has w*h entries. */
void orc_t1_decode(const uint8_t *data, int len, int w, int h, int num_bps, int band, int32_t *out);
/* the ZC context table (t1_luts.go:35-110), 4*256 entries, for table tests */
const uint8_t *orc_t1_zc_lut(void);

/* ---- HT block coder (internal/entropy/ht.go, ht_luts.go) --------------- */
/* HTDecoder.Decode on a fresh (zeroed) decoder, ht.go:93-150. */
void orc_ht_decode(const uint8_t *data, int len, int w, int h, int32_t *out);

/* ---- ISO-mode checkers (not restatements of the reference; see the file headers) ------ */
/* ISO/IEC 15444-15 HT block decoder, one HT set: data = cleanup segment (lcup bytes) + refinement segment (lref bytes),
 * num_passes 1..3 (cleanup, + SigProp, + MagRef), cleanup bit-plane P = num_bps - 1.  out = sign * Q in quarter units
 * (integer LSB = bit 2, mid-point bit below the last decoded bit-plane, as OpenJPEG reconstructs); 0 or <0 if malformed */
int  iso_ht_decode_passes(const uint8_t *data, int lcup, int lref, int w, int h, int num_bps, int num_passes, int32_t *out);
/* cleanup only, reversible value: out = sign * (Q >> 2) */
int  iso_ht_decode(const uint8_t *data, int len, int w, int h, int num_bps, int32_t *out);
/* ISO/IEC 15444-1 Annex C/D code-block decoder (iso_t1.c): num_bps magnitude bit-planes, the first num_passes coding
 * passes; out = sign * (2 * magnitude + mid-point of the last decoded bit-plane); band 0 LL, 1 HL, 2 LH, 3 HH */
int  iso_t1_decode(const uint8_t *data, int len, int w, int h, int num_bps, int num_passes, int band, int32_t *out);
int  iso_t1_num_segments(int style, int num_passes);
int  iso_t1_decode_style(const uint8_t *data, int len, int w, int h, int num_bps, int num_passes, int band, int style, int32_t *out);

/* ---- DWT (internal/dwt/dwt.go) ------------------------------------------ */
void orc_inv53(int32_t *d, int n);                      /* dwt.go:122-147 */
void orc_inv97(double *d, int n);                       /* dwt.go:213-262 */
void orc_inv2d53(int32_t *d, int w, int h);             /* dwt.go:410-429 */
void orc_inv2d97(double *d, int w, int h);              /* dwt.go:454-473 */
void orc_reconstruct53(int32_t *d, int w, int h, int levels); /* dwt.go:534-548 */
void orc_reconstruct97(double *d, int w, int h, int levels);  /* dwt.go:561-573 */
void orc_dequantize(const int32_t *in, double step, double *out, size_t n); /* dwt.go:514-520 */
/* TileDecoder.ApplyInverseDWT, internal/tcd/tcd.go:416-437 */
void orc_apply_inverse_dwt(int32_t *d, int w, int h, int levels, int reversible);

/* ---- MCT / DC shift (internal/mct/mct.go) ------------------------------- */
void orc_inv_rct(int32_t *y, int32_t *u, int32_t *v, size_t n);   /* mct.go:56-66  */
void orc_inv_ict(double *y, double *cb, double *cr, size_t n);    /* mct.go:43-53  */
void orc_dc_shift_inverse(int32_t *d, size_t n, int prec);        /* mct.go:113-118 */

/* ---- decoder tail (decoder.go) ------------------------------------------- */
/* decoder.go:321-348: inverse MCT (RCT, or ICT through float64 with the
 * truncating int32(v+0.5)) then DC shift of every unsigned component. */
void orc_decoder_tail(int32_t *const *comps, int ncomp, size_t n, int mct, int reversible,
                      const uint8_t *prec, const uint8_t *sgnd);
/* decoder.createImage, decoder.go:417-588.  pix layout = Go image.Gray /
 * Gray16 / RGBA / RGBA64 Pix (16-bit big-endian).  Returns bytes per pixel,
 * or -1 for an unsupported component count (decoder.go:585-586). */
int  orc_create_image(const int32_t *const *comps, int w, int h, int ncomp, int prec, uint8_t *pix);

/* ---- whole-path driver used as the CPU baseline --------------------------- */
/* Go float64 -> int32 conversion as the reference runs it on amd64 (CVTTSD2SL): truncation toward zero, and the
 * "integer indefinite" value 0x80000000 for NaN and for anything outside the int32 range (the Go spec leaves the
 * out-of-range result implementation-specific; REF parity is defined against amd64).  Written out instead of a C
 * cast, which would be undefined behaviour out of range. */
static inline int32_t orc_f64_to_i32(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int32_t)0x80000000u;
    return (int32_t)v;
}

typedef struct {
    uint64_t data_off; uint32_t data_len;
    uint32_t tilecomp; uint16_t x0, y0, w, h;
    uint8_t band, level, num_bps, num_passes;
    float step;
    uint32_t len_cleanup, rsv;
} orc_cblk_t;                 /* same layout as j2k_cblk_t (include/j2kgpu.h) */
typedef struct {
    uint32_t comp, x0, y0, x1, y1;
    uint64_t coeff_off;
} orc_tilecomp_t;             /* same layout as j2k_tilecomp_t */
typedef struct {
    uint32_t width, height;
    uint16_t ncomp; uint8_t prec[4]; uint8_t sgnd[4];
    uint8_t mct, reversible, nlevels, ht, mode, out_fmt;
    uint8_t coef_bits, colorspace, cblk_style, rsv;
} orc_image_t;                /* same layout as j2k_image_t */

/* REF-mode whole path with `threads` host threads: per block T1.Decode /
 * HTDecoder.Decode into the tile-component plane, ApplyInverseDWT per
 * tile-component, copy into image planes (decoder.go:398-410), decoder tail,
 * createImage.  Returns 0, or negative on bad arguments. */
/* colour conversion to sRGB, YCbCr family (colorspace.go:90-140, 429-452): cs 1 = BT.709 matrix, 2 = BT.601 matrix */
void orc_colour_convert(int32_t *const *comps, int ncomp, size_t n, int prec, int cs);

int orc_decode_image(const orc_image_t *img, const orc_tilecomp_t *tcs, uint32_t n_tc,
                     const orc_cblk_t *cbs, uint32_t n_cb, const uint8_t *blob, uint64_t blob_len,
                     uint8_t *out_pix, uint64_t out_stride, int threads);

/* ISO-mode (J2KGPU_MODE_ISO) whole path, see iso_path.c: same tables, OpenJPEG's arithmetic */
int iso_decode_image(const orc_image_t *img, const orc_tilecomp_t *tcs, uint32_t n_tc,
                     const orc_cblk_t *cbs, uint32_t n_cb, const uint8_t *blob, uint64_t blob_len,
                     uint8_t *out_pix, uint64_t out_stride, int threads);

/* forward path (orc_enc.c): encoder.extractImageData + preprocess + encodeTile's tileData (encoder.go:79-281, 597-743) */
typedef struct {
    uint32_t width, height; uint16_t ncomp; uint8_t pix_bits, precision, lossless, num_resolutions, cb_x, cb_y;
    int32_t quality; uint32_t flags; uint32_t rsv[2];
} orc_encode_t;               /* same layout as j2k_encode_t */
uint32_t orc_encode_block_count(const orc_encode_t *p);
int orc_encode_preprocess(const orc_encode_t *p, const uint8_t *pix, uint64_t stride, int32_t *planes);
int64_t orc_encode_tile(const orc_encode_t *p, const uint8_t *pix, uint64_t stride, uint8_t *out, uint64_t cap,
                        uint32_t *blk_len, uint8_t *blk_bps, int threads);

#ifdef __cplusplus
}
#endif
#endif
