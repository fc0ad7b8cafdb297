/*
 * iso_path.c -- CPU statement of the whole tile-component decode path in J2KGPU_MODE_ISO: the checker of the
 * full-size ISO-mode GPU tests and the CPU arm of bench.py for the conformant HTJ2K workload.
 * Oracle / test infrastructure only (see oracle.h): nothing of the product calls it.
 *
 * The reference (mrjoshuak/go-jpeg2000) has no working counterpart of this mode: its decodeTile is a placeholder
 * (decoder.go:375-380) and its HT coder is not ISO/IEC 15444-15.  What is restated here is the
 * published algorithm, in the arithmetic of OpenJPEG 2.5.4 -- the independent decoder that pins it: the numpy models in
 * tests/test_iso_codestream.py reproduce OpenJPEG's pixels bit for bit on codestreams OpenJPEG wrote, and
 * tests/test_oracle_iso_path.py checks this C code against those models and against OpenJPEG itself.
 *
 *   blocks      iso_ht_decode_passes (HT cleanup / SigProp / MagRef) or iso_t1_decode (Annex C/D), placed in the
 *               tile-component's Mallat plane; reversible: sign * floor(magnitude), irreversible: float32
 *               value * step with the mid-point of the last decoded bit-plane (dequantisation, Annex E)
 *   inverse DWT per level, coarsest first: every row, then every column (ISO order), 5-3 in int32 (Annex F.3.8.2:
 *               the same lifting steps as dwt.go:122-147), 9-7 in float32 with OpenJPEG's constants and
 *               operation order (low * K, high * 1.625732422 / 2, then delta, gamma, beta, alpha)
 *   tail        inverse RCT (mct.go:56-66 arithmetic) or inverse ICT in float32 + round to nearest even; DC shift;
 *               clamp; pack as decoder.createImage lays pixels out (decoder.go:417-588) without its int32 overflow
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const orc_image_t *img; const orc_tilecomp_t *tcs; uint32_t n_tc; const orc_cblk_t *cbs; uint32_t n_cb;
    const uint8_t *blob; int32_t **planes;        /* per tile-component: w * h int32 (float32 bits when irreversible) */
    uint8_t *out; uint64_t out_stride;
    int threads, tid, phase;
} iso_work_t;

static void inv53_lines(int32_t *p, int n, int count, size_t line_stride, size_t elem_stride, int32_t *tmp)
{
    for (int i = 0; i < count; i++) {
        int32_t *l = p + (size_t)i * line_stride;
        for (int k = 0; k < n; k++) tmp[k] = l[(size_t)k * elem_stride];
        orc_inv53(tmp, n);                                    /* interleave L | H, then the two lifting steps */
        for (int k = 0; k < n; k++) l[(size_t)k * elem_stride] = tmp[k];
    }
}

/* one line of the inverse 9-7 in float32: band order in (nl low-pass, then high-pass), interleaved out */
static void inv97f_line(float *x, int n, float *t)
{
    if (n < 2) return;
    static const float K = 1.230174105f, AL = -1.586134342f, BE = -0.052980118f, GA = 0.882911075f, DE = 0.443506852f;
    const float IK = 1.625732422f * 0.5f;
    const int nl = (n + 1) / 2;
    for (int i = 0; i < nl; i++) t[2 * i] = x[i] * K;
    for (int i = 0; i < n - nl; i++) t[2 * i + 1] = x[nl + i] * IK;
    const float cs[4] = {-DE, -GA, -BE, -AL};
    for (int s = 0; s < 4; s++) {
        const float c = cs[s];
        for (int i = s & 1; i < n; i += 2) {
            const int l = i - 1 >= 0 ? i - 1 : i + 1, r = i + 1 < n ? i + 1 : i - 1;
            const float sum = t[l] + t[r];
            const float prod = sum * c;
            t[i] = t[i] + prod;
        }
    }
    memcpy(x, t, sizeof(float) * (size_t)n);
}

static void inv97f_lines(float *p, int n, int count, size_t line_stride, size_t elem_stride, float *a, float *b)
{
    for (int i = 0; i < count; i++) {
        float *l = p + (size_t)i * line_stride;
        for (int k = 0; k < n; k++) a[k] = l[(size_t)k * elem_stride];
        inv97f_line(a, n, b);
        for (int k = 0; k < n; k++) l[(size_t)k * elem_stride] = a[k];
    }
}

static uint32_t pack_value(int32_t v, int prec)
{
    const int32_t maxv = (int32_t)((1u << prec) - 1u);
    if (v < 0) v = 0;
    if (v > maxv) v = maxv;
    if (prec <= 8) return (uint32_t)(prec != 8 ? (v * 255) / maxv : v) & 0xFF;
    return (uint32_t)(((uint64_t)(uint32_t)v * 65535u) / (uint32_t)maxv) & 0xFFFF;
}

static void *iso_worker(void *arg)
{
    iso_work_t *j = arg;
    const orc_image_t *im = j->img;
    const int irrev = !im->reversible;
    if (j->phase == 0) {                                      /* ---- blocks ---- */
        int32_t *buf = malloc(sizeof(int32_t) * 64 * 64);
        for (uint32_t i = (uint32_t)j->tid; i < j->n_cb; i += (uint32_t)j->threads) {
            const orc_cblk_t *cb = &j->cbs[i];
            if (cb->w == 0 || cb->h == 0) continue;
            const orc_tilecomp_t *tc = &j->tcs[cb->tilecomp];
            const int pw = (int)(tc->x1 - tc->x0);
            int32_t *dst = j->planes[cb->tilecomp] + (size_t)cb->y0 * pw + cb->x0;
            memset(buf, 0, sizeof(int32_t) * (size_t)cb->w * cb->h);
            int shift;                                        /* fractional bits of buf below the integer LSB */
            if (cb->data_len == 0 || cb->num_bps == 0) shift = 0;
            else if (im->ht) {
                const uint32_t lcup = (cb->len_cleanup && cb->len_cleanup <= cb->data_len) ? cb->len_cleanup : cb->data_len;
                int np = cb->num_passes ? cb->num_passes : 1;
                if (np > 3) np = 3;
                iso_ht_decode_passes(j->blob + cb->data_off, (int)lcup, (int)(cb->data_len - lcup), cb->w, cb->h, cb->num_bps, np, buf);
                shift = 2;
            } else {
                const int all = 3 * cb->num_bps - 2;
                iso_t1_decode_style(j->blob + cb->data_off, (int)cb->data_len, cb->w, cb->h, cb->num_bps,
                                    cb->num_passes && cb->num_passes < all ? cb->num_passes : all, cb->band, im->cblk_style, buf);
                shift = 1;
            }
            const float sc = cb->step * (shift == 2 ? 0.25f : (shift == 1 ? 0.5f : 1.0f));
            for (int y = 0; y < cb->h; y++)
                for (int x = 0; x < cb->w; x++) {
                    const int32_t v = buf[y * cb->w + x];
                    int32_t o;
                    if (irrev) { const float f = (float)v * sc; memcpy(&o, &f, 4); }
                    else { const uint32_t m = (uint32_t)(v < 0 ? -(int64_t)v : v) >> shift; o = (int32_t)(v < 0 ? 0u - m : m); }
                    dst[(size_t)y * pw + x] = o;
                }
        }
        free(buf);
    } else if (j->phase == 1) {                               /* ---- inverse DWT per tile-component ---- */
        for (uint32_t t = (uint32_t)j->tid; t < j->n_tc; t += (uint32_t)j->threads) {
            const int w = (int)(j->tcs[t].x1 - j->tcs[t].x0), h = (int)(j->tcs[t].y1 - j->tcs[t].y0);
            const int m = (w > h ? w : h) + 8;
            void *ta = malloc(sizeof(float) * (size_t)m), *tb = malloc(sizeof(float) * (size_t)m);
            for (int lvl = im->nlevels - 1; lvl >= 0; lvl--) {
                const int lw = (w + (1 << lvl) - 1) >> lvl, lh = (h + (1 << lvl) - 1) >> lvl;
                if (irrev) {
                    inv97f_lines((float *)j->planes[t], lw, lh, (size_t)w, 1, ta, tb);      /* rows */
                    inv97f_lines((float *)j->planes[t], lh, lw, 1, (size_t)w, ta, tb);      /* columns */
                } else {
                    inv53_lines(j->planes[t], lw, lh, (size_t)w, 1, ta);
                    inv53_lines(j->planes[t], lh, lw, 1, (size_t)w, ta);
                }
            }
            free(ta); free(tb);
        }
    } else {                                                  /* ---- tail: tiles -> pixels (one tile-component set per step) ---- */
        const int nc = im->ncomp, prec = im->prec[0];
        const int bpp = nc == 1 ? (prec <= 8 ? 1 : 2) : (prec <= 8 ? 4 : 8);
        for (uint32_t t = (uint32_t)j->tid; t < j->n_tc; t += (uint32_t)j->threads) {
            const orc_tilecomp_t *tc = &j->tcs[t];
            if (tc->comp != 0) continue;
            const int32_t *pl[4] = {j->planes[t], 0, 0, 0};
            for (uint32_t u = 0; u < j->n_tc; u++)
                if (j->tcs[u].x0 == tc->x0 && j->tcs[u].y0 == tc->y0 && j->tcs[u].x1 == tc->x1 && j->tcs[u].y1 == tc->y1 && j->tcs[u].comp < 4)
                    pl[j->tcs[u].comp] = j->planes[u];
            const int w = (int)(tc->x1 - tc->x0), h = (int)(tc->y1 - tc->y0);
            for (int y = 0; y < h && tc->y0 + (uint32_t)y < im->height; y++)
                for (int x = 0; x < w && tc->x0 + (uint32_t)x < im->width; x++) {
                    int32_t v[4] = {0, 0, 0, 0};
                    const size_t k = (size_t)y * w + x;
                    if (irrev) {
                        float f[4] = {0, 0, 0, 0};
                        for (int c = 0; c < nc; c++) if (pl[c]) memcpy(&f[c], &pl[c][k], 4);
                        if (im->mct && nc >= 3) {
                            const float yy = f[0], u = f[1], ww = f[2];
                            const float a = ww * 1.402f, r = yy + a;
                            const float b1 = u * 0.34413f, b2 = ww * 0.71414f, g0 = yy - b1, g = g0 - b2;
                            const float c1 = u * 1.772f, b = yy + c1;
                            f[0] = r; f[1] = g; f[2] = b;
                        }
                        for (int c = 0; c < nc; c++) v[c] = (int32_t)lrintf(f[c]);
                    } else {
                        for (int c = 0; c < nc; c++) if (pl[c]) v[c] = pl[c][k];
                        if (im->mct && nc >= 3) {
                            const uint32_t yy = (uint32_t)v[0], u = (uint32_t)v[1], ww = (uint32_t)v[2];
                            const uint32_t g = yy - (uint32_t)((int32_t)(u + ww) >> 2);
                            v[0] = (int32_t)(ww + g); v[1] = (int32_t)g; v[2] = (int32_t)(u + g);
                        }
                    }
                    for (int c = 0; c < nc; c++) if (!im->sgnd[c]) v[c] = (int32_t)((uint32_t)v[c] + (1u << (im->prec[c] - 1)));
                    if (im->colorspace && nc >= 3) {                                   /* decoder.go:350-356 */
                        int32_t *cp[4] = {&v[0], &v[1], &v[2], &v[3]};
                        orc_colour_convert(cp, nc, 1, im->prec[0], im->colorspace);
                    }
                    uint8_t *o = j->out + (size_t)(tc->y0 + (uint32_t)y) * j->out_stride + (size_t)(tc->x0 + (uint32_t)x) * (size_t)bpp;
                    if (bpp == 1) o[0] = (uint8_t)pack_value(v[0], prec);
                    else if (bpp == 2) { const uint32_t p = pack_value(v[0], prec); o[0] = (uint8_t)(p >> 8); o[1] = (uint8_t)p; }
                    else if (bpp == 4) {
                        for (int c = 0; c < 3; c++) o[c] = (uint8_t)pack_value(v[c], prec);
                        o[3] = nc == 4 ? (uint8_t)pack_value(v[3], prec) : 255;
                    } else {
                        for (int c = 0; c < 4; c++) {
                            const uint32_t p = (c < 3 || nc == 4) ? pack_value(v[c], prec) : 65535u;
                            o[2 * c] = (uint8_t)(p >> 8); o[2 * c + 1] = (uint8_t)p;
                        }
                    }
                }
        }
    }
    return 0;
}

/* ISO-mode whole path with `threads` host threads; same tables as j2kgpu_decode (include/j2kgpu.h).  Pixels no tile
 * covers are left as the caller passed them.  Returns 0, or negative on bad arguments. */
int iso_decode_image(const orc_image_t *img, const orc_tilecomp_t *tcs, uint32_t n_tc,
                     const orc_cblk_t *cbs, uint32_t n_cb, const uint8_t *blob, uint64_t blob_len,
                     uint8_t *out_pix, uint64_t out_stride, int threads)
{
    if (!img || !out_pix || (n_tc && !tcs) || (n_cb && !cbs)) return -1;
    if (img->ncomp != 1 && img->ncomp != 3 && img->ncomp != 4) return -3;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (uint32_t i = 0; i < n_cb; i++) {
        const orc_cblk_t *cb = &cbs[i];
        if (cb->tilecomp >= n_tc) return -2;
        const orc_tilecomp_t *tc = &tcs[cb->tilecomp];
        if (cb->x0 + cb->w > tc->x1 - tc->x0 || cb->y0 + cb->h > tc->y1 - tc->y0 || cb->w > 64 || cb->h > 64) return -2;
        if (cb->data_off > blob_len || cb->data_len > blob_len - cb->data_off) return -2;
        if (!img->ht && (img->cblk_style & 0x05) && cb->data_len &&           /* the segment-length table behind the code bytes */
            4ull * (uint64_t)iso_t1_num_segments(img->cblk_style, cb->num_passes ? cb->num_passes : 3 * cb->num_bps - 2) >
                blob_len - cb->data_off - cb->data_len) return -2;
    }
    int32_t **planes = calloc(n_tc ? n_tc : 1, sizeof(int32_t *));
    for (uint32_t t = 0; t < n_tc; t++)
        planes[t] = calloc((size_t)(tcs[t].x1 - tcs[t].x0) * (tcs[t].y1 - tcs[t].y0), sizeof(int32_t));
    pthread_t th[256];
    iso_work_t wk[256];
    for (int phase = 0; phase < 3; phase++) {
        for (int t = 0; t < threads; t++) {
            wk[t] = (iso_work_t){img, tcs, n_tc, cbs, n_cb, blob, planes, out_pix, out_stride, threads, t, phase};
            if (t) pthread_create(&th[t], 0, iso_worker, &wk[t]);
        }
        iso_worker(&wk[0]);
        for (int t = 1; t < threads; t++) pthread_join(th[t], 0);
    }
    for (uint32_t t = 0; t < n_tc; t++) free(planes[t]);
    free(planes);
    return 0;
}
