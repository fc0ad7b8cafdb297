/*
 * orc_tail.c -- restatement of internal/mct/mct.go (RCT, ICT, DC shift) and of
 * the tail of the reference decoder: decoder.go:321-348 (inverse MCT + DC
 * shift) and decoder.createImage decoder.go:417-599 (clamp, precision scaling,
 * packing into the Go image Pix layouts).  Also the threaded whole-path driver
 * used as the CPU baseline.  Oracle / test infrastructure only.
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))
#define WMUL(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))

void orc_inv_rct(int32_t *y, int32_t *u, int32_t *v, size_t n)      /* mct.go:56-66 */
{
    for (size_t i = 0; i < n; i++) {
        int32_t g = WSUB(y[i], WADD(u[i], v[i]) >> 2);
        int32_t r = WADD(v[i], g), b = WADD(u[i], g);
        y[i] = r; u[i] = g; v[i] = b;
    }
}

void orc_inv_ict(double *y, double *cb, double *cr, size_t n)       /* mct.go:43-53 */
{
    for (size_t i = 0; i < n; i++) {
        double r = y[i] + 1.402 * cr[i];
        double g = y[i] - 0.34413 * cb[i] - 0.71414 * cr[i];
        double b = y[i] + 1.772 * cb[i];
        y[i] = r; cb[i] = g; cr[i] = b;
    }
}

void orc_dc_shift_inverse(int32_t *d, size_t n, int prec)           /* mct.go:113-118 */
{
    int32_t s = (int32_t)((uint32_t)1 << (prec - 1));
    for (size_t i = 0; i < n; i++) d[i] = WADD(d[i], s);
}

/* decoder.go:321-348 */
void orc_decoder_tail(int32_t *const *comps, int ncomp, size_t n, int mct, int reversible,
                      const uint8_t *prec, const uint8_t *sgnd)
{
    if (mct != 0 && ncomp >= 3) {
        if (reversible) {
            orc_inv_rct(comps[0], comps[1], comps[2], n);
        } else {
            double *f[3];
            for (int c = 0; c < 3; c++) {
                f[c] = malloc(sizeof(double) * (n ? n : 1));
                for (size_t i = 0; i < n; i++) f[c][i] = (double)comps[c][i];
            }
            orc_inv_ict(f[0], f[1], f[2], n);
            for (int c = 0; c < 3; c++) {
                for (size_t i = 0; i < n; i++) comps[c][i] = orc_f64_to_i32(f[c][i] + 0.5);  /* truncating */
                free(f[c]);
            }
        }
    }
    for (int c = 0; c < ncomp; c++)
        if (!sgnd[c]) orc_dc_shift_inverse(comps[c], n, prec[c]);
}

/* clampToInt32 colorspace.go:483-491 */
static inline int32_t cs_clamp_round(double v, double maxv)
{
    if (v < 0.0) return 0;
    if (v > maxv) return (int32_t)maxv;
    return orc_f64_to_i32(v + 0.5);
}

/* labInverseF colorspace.go:294-301; the constants are Go's exact untyped constants rounded to float64 (6/29, 4/29,
 * 3 * (6/29)^2 = 108/841: the C constant expressions below round to the same doubles, tests/test_oracle_tail.py) */
static inline double lab_inverse_f(double t)
{
    if (t > 6.0 / 29.0) return t * t * t;
    return (3 * (6.0 / 29.0) * (6.0 / 29.0)) * (t - 4.0 / 29.0);
}

/* srgbGamma colorspace.go:303-309.  pow() is the libm's where Go has math.Pow: the only place this restatement can
 * differ from the reference, in the last bit of the power before the rounding to an integer ("parity unpinned at 1 LSB"
 * for conversions 7-10; there is no Go here to run). */
static inline double srgb_gamma(double lin)
{
    if (lin <= 0.0031308) return 12.92 * lin;
    return 1.055 * pow(lin, 1.0 / 2.4) - 0.055;
}

static inline double clampf(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   /* colorspace.go:494-502 */

/* XYZ -> linear sRGB -> gamma -> integer: the common tail of convertCIELabToRGB / convertCIEJabToRGB (no clamp before the
 * gamma, colorspace.go:277-289) and convertROMMRGBToRGB (clamped to [0, 1], :416-424) */
static inline void xyz_to_srgb(double x, double y, double z, int clamp01, double maxv, int32_t out[3])
{
    double rl = 3.2404542 * x - 1.5371385 * y - 0.4985314 * z;
    double gl = -0.9692660 * x + 1.8760108 * y + 0.0415560 * z;
    double bl = 0.0556434 * x - 0.2040259 * y + 1.0572252 * z;
    if (clamp01) { rl = clampf(rl, 0, 1); gl = clampf(gl, 0, 1); bl = clampf(bl, 0, 1); }
    out[0] = cs_clamp_round(srgb_gamma(rl) * maxv, maxv);
    out[1] = cs_clamp_round(srgb_gamma(gl) * maxv, maxv);
    out[2] = cs_clamp_round(srgb_gamma(bl) * maxv, maxv);
}

/* decoder.go:350-356 -> getColorConversion (colorspace.go:54-88).  cs 1: convertSYCCToRGB (:90-114), convertYPbPr709ToRGB
 * (:429-452), convertEYCCToRGB (:454-482) -- the BT.709 matrix; cs 2: convertYCbCr601ToRGB (:116-140); cs 3:
 * convertPhotoYCCToRGB (:142-168); cs 4: convertCMYToRGB (:170-189); cs 5: convertCMYKToRGB (:191-217); cs 6:
 * convertYCCKToRGB (:219-250); cs 7 / 8: convertCIELabToRGB (:250-292) / convertCIEJabToRGB (:319-359, the same arithmetic);
 * cs 9: convertESRGBToRGB (:363-389); cs 10: convertROMMRGBToRGB (:393-427). */
void orc_colour_convert(int32_t *const *comps, int ncomp, size_t n, int prec, int cs)
{
    if (cs == 0 || ncomp < 3) return;
    const int32_t maxi = (int32_t)(((uint32_t)1 << prec) - 1);
    const double maxv = (double)maxi, half = (double)(int32_t)((uint32_t)1 << (prec - 1));
    if (cs == 4) {
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) comps[c][i] = (int32_t)((uint32_t)maxi - (uint32_t)comps[c][i]);
        return;
    }
    if (cs == 5) {
        if (ncomp < 4) return;
        for (size_t i = 0; i < n; i++) {
            const double c = (double)comps[0][i] / maxv, m = (double)comps[1][i] / maxv, y = (double)comps[2][i] / maxv,
                         k = (double)comps[3][i] / maxv;
            const double r = (1 - c) * (1 - k) * maxv, g = (1 - m) * (1 - k) * maxv, b = (1 - y) * (1 - k) * maxv;
            comps[0][i] = cs_clamp_round(r, maxv); comps[1][i] = cs_clamp_round(g, maxv); comps[2][i] = cs_clamp_round(b, maxv);
        }
        return;
    }
    if (cs == 7 || cs == 8) {
        for (size_t i = 0; i < n; i++) {
            const double L = (double)comps[0][i] / maxv * 100.0;
            const double a = (double)comps[1][i] / maxv * 255.0 - 128.0, b = (double)comps[2][i] / maxv * 255.0 - 128.0;
            const double fy = (L + 16.0) / 116.0, fx = a / 500.0 + fy, fz = fy - b / 200.0;
            int32_t o[3];
            xyz_to_srgb(0.96422 * lab_inverse_f(fx), 1.0 * lab_inverse_f(fy), 0.82521 * lab_inverse_f(fz), 0, maxv, o);
            comps[0][i] = o[0]; comps[1][i] = o[1]; comps[2][i] = o[2];
        }
        return;
    }
    if (cs == 9) {
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) {
                const double v = (double)comps[c][i] / maxv * 1.25 - 0.25;
                comps[c][i] = cs_clamp_round(srgb_gamma(clampf(v, 0, 1)) * maxv, maxv);
            }
        return;
    }
    if (cs == 10) {
        for (size_t i = 0; i < n; i++) {
            const double rr = pow((double)comps[0][i] / maxv, 1.8), gr = pow((double)comps[1][i] / maxv, 1.8),
                         br = pow((double)comps[2][i] / maxv, 1.8);
            const double x = 0.7977 * rr + 0.1352 * gr + 0.0313 * br;
            const double y = 0.2880 * rr + 0.7119 * gr + 0.0001 * br;
            const double z = 0.0000 * rr + 0.0000 * gr + 0.8249 * br;
            int32_t o[3];
            xyz_to_srgb(x, y, z, 1, maxv, o);
            comps[0][i] = o[0]; comps[1][i] = o[1]; comps[2][i] = o[2];
        }
        return;
    }
    if (cs == 3 || cs == 6) {
        if (cs == 6 && ncomp < 4) return;
        const double scale = maxv / 255.0;
        for (size_t i = 0; i < n; i++) {
            const double y = (double)comps[0][i] / scale, c1 = (double)comps[1][i] / scale - 156.0,
                         c2 = (double)comps[2][i] / scale - 156.0;
            double r = y + 1.3584 * c2;
            double g = y - 0.4302 * c1 - 0.7915 * c2;
            double b = y + 2.2179 * c1;
            if (cs == 6) {
                const double k = (double)comps[3][i] / maxv;
                r = r * scale * (1 - k); g = g * scale * (1 - k); b = b * scale * (1 - k);
            } else {
                r = r * scale; g = g * scale; b = b * scale;
            }
            comps[0][i] = cs_clamp_round(r, maxv); comps[1][i] = cs_clamp_round(g, maxv); comps[2][i] = cs_clamp_round(b, maxv);
        }
        return;
    }
    const double kr = cs == 1 ? 1.5748 : 1.402, kgb = cs == 1 ? 0.1873 : 0.344136, kgr = cs == 1 ? 0.4681 : 0.714136,
                 kb = cs == 1 ? 1.8556 : 1.772;
    for (size_t i = 0; i < n; i++) {
        const double y = (double)comps[0][i], cb = (double)comps[1][i] - half, cr = (double)comps[2][i] - half;
        const double r = y + kr * cr;
        const double g = y - kgb * cb - kgr * cr;
        const double b = y + kb * cb;
        comps[0][i] = cs_clamp_round(r, maxv);
        comps[1][i] = cs_clamp_round(g, maxv);
        comps[2][i] = cs_clamp_round(b, maxv);
    }
}

static inline int32_t clamp_i32(int32_t v, int32_t lo, int32_t hi)   /* decoder.go:591-599 */
{
    return v < lo ? lo : (v > hi ? hi : v);
}

/* decoder.createImage decoder.go:417-588.  `v*65535/maxVal` is evaluated in
 * int32 like the reference (the product wraps for 16-bit samples). */
int orc_create_image(const int32_t *const *comps, int w, int h, int ncomp, int prec, uint8_t *pix)
{
    int32_t maxv = (int32_t)(((uint64_t)1 << prec) - 1);
    size_t n = (size_t)w * h;
    if (ncomp != 1 && ncomp != 3 && ncomp != 4) return -1;         /* decoder.go:585-586 */
    int nch = ncomp == 1 ? 1 : 4;
    if (prec <= 8) {
        for (size_t i = 0; i < n; i++) {
            for (int c = 0; c < nch; c++) {
                int32_t v;
                if (c < ncomp) {
                    v = clamp_i32(comps[c][i], 0, maxv);
                    if (prec != 8) v = WMUL(v, 255) / maxv;
                } else v = 255;                                      /* opaque alpha for 3 components */
                pix[i * nch + c] = (uint8_t)v;
            }
        }
        return nch;
    }
    for (size_t i = 0; i < n; i++) {
        for (int c = 0; c < nch; c++) {
            int32_t v;
            if (c < ncomp) {
                v = clamp_i32(comps[c][i], 0, maxv);
                v = WMUL(v, 65535) / maxv;
            } else v = 65535;
            uint16_t u = (uint16_t)v;
            pix[(i * nch + c) * 2] = (uint8_t)(u >> 8);              /* Go image.Gray16/RGBA64: big-endian */
            pix[(i * nch + c) * 2 + 1] = (uint8_t)u;
        }
    }
    return nch * 2;
}

/* ---------------- whole-path driver (CPU baseline) ---------------------------- */
typedef struct {
    const orc_image_t *img; const orc_tilecomp_t *tcs; uint32_t n_tc;
    const orc_cblk_t *cbs; uint32_t n_cb; const uint8_t *blob;
    int32_t **planes;            /* per tile-component coefficient planes */
    volatile uint32_t next;      /* work counter */
    int stage;
    /* stage 2: image tail over row chunks */
    int32_t *comps[4]; uint8_t *out_pix; uint64_t out_stride; int bpp; int rc;
} job_t;

#define TAIL_ROWS 16

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (j->stage == 0) {                       /* block entropy decode */
            if (i >= j->n_cb) break;
            const orc_cblk_t *cb = &j->cbs[i];
            const orc_tilecomp_t *tc = &j->tcs[cb->tilecomp];
            int tw = (int)(tc->x1 - tc->x0);
            int32_t *tmp = malloc(sizeof(int32_t) * (size_t)cb->w * cb->h + 4);
            if (cb->data_len == 0) memset(tmp, 0, sizeof(int32_t) * (size_t)cb->w * cb->h);  /* tcd.go:394-396 */
            else if (j->img->ht) orc_ht_decode(j->blob + cb->data_off, (int)cb->data_len, cb->w, cb->h, tmp);
            else orc_t1_decode(j->blob + cb->data_off, (int)cb->data_len, cb->w, cb->h, cb->num_bps, cb->band, tmp);
            int32_t *plane = j->planes[cb->tilecomp];
            for (int y = 0; y < cb->h; y++)
                memcpy(plane + (size_t)(cb->y0 + y) * tw + cb->x0, tmp + (size_t)y * cb->w, sizeof(int32_t) * cb->w);
            free(tmp);
        } else if (j->stage == 2) {                /* decoder tail + createImage on a chunk of rows (elementwise) */
            size_t W = j->img->width, H = j->img->height;
            size_t y0 = (size_t)i * TAIL_ROWS;
            if (y0 >= H) break;
            size_t rows = H - y0 < TAIL_ROWS ? H - y0 : TAIL_ROWS;
            int32_t *c[4] = {0, 0, 0, 0};
            for (int k = 0; k < j->img->ncomp; k++) c[k] = j->comps[k] + y0 * W;
            orc_decoder_tail(c, j->img->ncomp, rows * W, j->img->mct, j->img->reversible, j->img->prec, j->img->sgnd);
            orc_colour_convert(c, j->img->ncomp, rows * W, j->img->prec[0], j->img->colorspace);   /* decoder.go:350-356 */
            if (j->out_stride == W * (size_t)j->bpp) {
                if (orc_create_image((const int32_t *const *)c, (int)W, (int)rows, j->img->ncomp, j->img->prec[0],
                                     j->out_pix + y0 * j->out_stride) < 0) j->rc = -3;
            } else {
                uint8_t *tmp = malloc(rows * W * (size_t)j->bpp + 1);
                if (orc_create_image((const int32_t *const *)c, (int)W, (int)rows, j->img->ncomp, j->img->prec[0], tmp) < 0) j->rc = -3;
                else for (size_t y = 0; y < rows; y++) memcpy(j->out_pix + (y0 + y) * j->out_stride, tmp + y * W * j->bpp, W * (size_t)j->bpp);
                free(tmp);
            }
        } else {                                   /* inverse DWT per tile-component */
            if (i >= j->n_tc) break;
            const orc_tilecomp_t *tc = &j->tcs[i];
            orc_apply_inverse_dwt(j->planes[i], (int)(tc->x1 - tc->x0), (int)(tc->y1 - tc->y0),
                                  j->img->nlevels, j->img->reversible);
        }
    }
    return NULL;
}

static void run_stage(job_t *j, int stage, int threads)
{
    j->stage = stage; j->next = 0;
    if (threads <= 1) { worker(j); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

int orc_decode_image(const orc_image_t *img, const orc_tilecomp_t *tcs, uint32_t n_tc,
                     const orc_cblk_t *cbs, uint32_t n_cb, const uint8_t *blob, uint64_t blob_len,
                     uint8_t *out_pix, uint64_t out_stride, int threads)
{
    if (!img || !tcs || !out_pix || img->ncomp == 0 || img->ncomp > 4) return -1;
    for (uint32_t i = 0; i < n_cb; i++) {
        if (cbs[i].tilecomp >= n_tc) return -1;
        if (cbs[i].data_off + cbs[i].data_len > blob_len) return -2;
    }
    job_t j; memset(&j, 0, sizeof j);
    j.img = img; j.tcs = tcs; j.n_tc = n_tc; j.cbs = cbs; j.n_cb = n_cb; j.blob = blob;
    j.planes = calloc(n_tc ? n_tc : 1, sizeof(int32_t *));
    for (uint32_t i = 0; i < n_tc; i++) {
        size_t n = (size_t)(tcs[i].x1 - tcs[i].x0) * (tcs[i].y1 - tcs[i].y0);
        j.planes[i] = calloc(n ? n : 1, sizeof(int32_t));           /* tcd.go:283 zero Data */
    }
    run_stage(&j, 0, threads);
    run_stage(&j, 1, threads);

    size_t W = img->width, H = img->height, n = W * H;
    int32_t *comps[4] = {0, 0, 0, 0};
    for (int c = 0; c < img->ncomp; c++) comps[c] = calloc(n ? n : 1, sizeof(int32_t));
    for (uint32_t i = 0; i < n_tc; i++) {                           /* decoder.go:398-410 */
        const orc_tilecomp_t *tc = &tcs[i];
        if (tc->comp >= img->ncomp) continue;
        size_t tw = tc->x1 - tc->x0;
        for (uint32_t y = tc->y0; y < tc->y1 && y < H; y++)
            for (uint32_t x = tc->x0; x < tc->x1 && x < W; x++)
                comps[tc->comp][(size_t)y * W + x] = j.planes[i][(size_t)(y - tc->y0) * tw + (x - tc->x0)];
    }
    int bpp = img->ncomp == 1 ? (img->prec[0] <= 8 ? 1 : 2) : (img->prec[0] <= 8 ? 4 : 8);
    if (img->ncomp != 1 && img->ncomp != 3 && img->ncomp != 4) j.rc = -3;          /* decoder.go:585-586 */
    else {
        for (int c = 0; c < img->ncomp; c++) j.comps[c] = comps[c];
        j.out_pix = out_pix; j.out_stride = out_stride; j.bpp = bpp;
        run_stage(&j, 2, threads);
    }
    int rc = j.rc;
    (void)n;
    for (int c = 0; c < img->ncomp; c++) free(comps[c]);
    for (uint32_t i = 0; i < n_tc; i++) free(j.planes[i]);
    free(j.planes);
    return rc;
}
