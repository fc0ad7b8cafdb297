/*
 * datagen.h -- synthetic-input generator: restatement of the reference ENCODER side
 * (mrjoshuak/go-jpeg2000 encoder.go preprocess + entropy encoders), used to manufacture code-block
 * bitstreams and coefficient planes for tests and for bench.py inputs ("synthetic codestreams generated
 * by the reference encoder", BASELINE.json).  It is neither the product (which only decodes, on the GPU)
 * nor the oracle (oracle/ = decoder-side checker).  Every function cites the reference file:line.
 */
#ifndef J2K_DATAGEN_H
#define J2K_DATAGEN_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
enum { GEN_BAND_LL = 0, GEN_BAND_HL = 1, GEN_BAND_LH = 2, GEN_BAND_HH = 3 };
/* MQ encoder over n (ctx,bit) decisions; returns bytes written or -1.  mqc.go:185-349 */
int  gen_mq_encode(const uint8_t *ctxs, const uint8_t *bits, int n, uint8_t *out, int cap);
/* T1.SetData + T1.Encode; returns byte count (0 == nil for an all-zero block), -1 if cap too small;
 * *num_bps = bit length of max|x| (t1_fast5.go:13-28) */
int  gen_t1_encode(const int32_t *coeffs, int w, int h, int band, uint8_t *out, int cap, int *num_bps);
/* HTEncoder.Encode ht.go:942-1045; -1 where the reference would index out of range */
int  gen_ht_encode(const int32_t *coeffs, int w, int h, int band, uint8_t *out, int cap);
/* ISO/IEC 15444-15 HT cleanup-pass encoder (conformant; see gen_iso_ht.c) */
int  gen_iso_ht_encode(const int32_t *coeffs, int w, int h, uint8_t *out, int cap);
/* one HT set: cleanup at bit-plane P, + SigProp (npasses >= 2) + MagRef (npasses == 3) at bit-plane P - 1; returns
 * Lcup + Lref (0: nothing significant at bit-plane P), *lcup = Lcup; see gen_iso_ht.c */
int  gen_iso_ht_encode_passes(const int32_t *coeffs, int w, int h, int P, int npasses, uint8_t *out, int cap,
                              int *lcup, int32_t *recon);
void gen_fwd53(int32_t *d, int n);                              /* dwt.go:73-118  */
void gen_fwd97(double *d, int n);                               /* dwt.go:161-210 */
void gen_fwd2d53(int32_t *d, int w, int h);                     /* dwt.go:356-407 */
void gen_fwd2d97(double *d, int w, int h);                      /* dwt.go:432-451 */
void gen_decompose53(int32_t *d, int w, int h, int levels);     /* dwt.go:524-531 */
void gen_decompose97(double *d, int w, int h, int levels);      /* dwt.go:551-558 */
void gen_quantize(const double *in, double step, int32_t *out, size_t n);   /* dwt.go:500-511 */
void gen_fwd_rct(int32_t *r, int32_t *g, int32_t *b, size_t n); /* mct.go:28-38  */
void gen_fwd_ict(double *r, double *g, double *b, size_t n);    /* mct.go:14-24  */
void gen_dc_shift_forward(int32_t *d, size_t n, int prec);      /* mct.go:96-101 */
/* Encode every block of a job in parallel (the shape of the reference encoder's goroutine pool,
 * encoder.go:690-742): block i covers plane[y0..y0+h) x [x0..x0+w) of a stride-`stride` int32 plane that
 * starts at planes[plane_off]; bytes are appended to out (capacity cap) in block order and offs / lens / nbps
 * receive the per-block side information.  Returns total bytes or -1. */
typedef struct { uint64_t plane_off; uint32_t stride; uint16_t x0, y0, w, h; uint8_t band, ht, r0, r1; } gen_blk_t;
int64_t gen_encode_blocks(const int32_t *planes, const gen_blk_t *blks, uint32_t n, uint8_t *out, uint64_t cap,
                          uint64_t *offs, uint32_t *lens, uint8_t *nbps, int threads);
#ifdef __cplusplus
}
#endif
#endif
