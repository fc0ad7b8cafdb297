/*
 * gen_fwd.c -- restatement of the reference ENCODER-side transforms: Forward53/Forward97
 * (internal/dwt/dwt.go:73-118, 161-210), Forward2D53/97 (dwt.go:356-407, 432-451),
 * DecomposeMultiLevel53/97 (dwt.go:524-531, 551-558; dense-prefix layout), Quantize (dwt.go:500-511),
 * ForwardRCT/ForwardICT (internal/mct/mct.go:14-38) and DCLevelShiftForward (mct.go:96-101).
 * Part of datagen/ (synthetic-input generator).  float64 without FMA (-ffp-contract=off), int32 wraps.
 */
#include "datagen.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))
#define WMUL(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))

static const double kAlpha = -1.586134342059924;   /* dwt.go:150-157 */
static const double kBeta  = -0.052980118572961;
static const double kGamma = 0.882911075530934;
static const double kDelta = 0.443506852043971;
static const double kK     = 1.230174104914001;
static const double kKInv  = 0.812893066115961;

/* deinterleave dwt.go:265-284, 309-326 */
static void deinterleave_i(int32_t *d, int n, int32_t *tmp)
{
    int half = (n + 1) / 2;
    for (int i = 0, j = 0; i < n; i += 2, j++) tmp[j] = d[i];
    for (int i = 1, j = half; i < n; i += 2, j++) tmp[j] = d[i];
    memcpy(d, tmp, sizeof(int32_t) * (size_t)n);
}
static void deinterleave_f(double *d, int n, double *tmp)
{
    int half = (n + 1) / 2;
    for (int i = 0, j = 0; i < n; i += 2, j++) tmp[j] = d[i];
    for (int i = 1, j = half; i < n; i += 2, j++) tmp[j] = d[i];
    memcpy(d, tmp, sizeof(double) * (size_t)n);
}
static void fwd53_t(int32_t *d, int n, int32_t *tmp)     /* dwt.go:73-118 */
{
    if (n < 2) return;
    for (int i = 1; i < n - 1; i += 2) d[i] = WSUB(d[i], WADD(d[i - 1], d[i + 1]) >> 1);
    if ((n & 1) == 0) d[n - 1] = WSUB(d[n - 1], d[n - 2]);
    d[0] = WADD(d[0], WADD(WADD(d[1], d[1]), 2) >> 2);
    for (int i = 2; i < n - 1; i += 2) d[i] = WADD(d[i], WADD(WADD(d[i - 1], d[i + 1]), 2) >> 2);
    if (n & 1) d[n - 1] = WADD(d[n - 1], WADD(WADD(d[n - 2], d[n - 2]), 2) >> 2);
    deinterleave_i(d, n, tmp);
}

static void fwd97_t(double *d, int n, double *tmp)       /* dwt.go:161-210 */
{
    if (n < 2) return;
    for (int i = 1; i < n - 1; i += 2) d[i] += kAlpha * (d[i - 1] + d[i + 1]);
    if ((n & 1) == 0) d[n - 1] += (2 * kAlpha) * d[n - 2];
    d[0] += (2 * kBeta) * d[1];
    for (int i = 2; i < n - 1; i += 2) d[i] += kBeta * (d[i - 1] + d[i + 1]);
    if (n & 1) d[n - 1] += (2 * kBeta) * d[n - 2];
    for (int i = 1; i < n - 1; i += 2) d[i] += kGamma * (d[i - 1] + d[i + 1]);
    if ((n & 1) == 0) d[n - 1] += (2 * kGamma) * d[n - 2];
    d[0] += (2 * kDelta) * d[1];
    for (int i = 2; i < n - 1; i += 2) d[i] += kDelta * (d[i - 1] + d[i + 1]);
    if (n & 1) d[n - 1] += (2 * kDelta) * d[n - 2];
    for (int i = 0; i < n; i += 2) d[i] *= kKInv;
    for (int i = 1; i < n; i += 2) d[i] *= kK;
    deinterleave_f(d, n, tmp);
}

void gen_fwd53(int32_t *d, int n) { if (n < 2) return; int32_t *t = malloc(sizeof(int32_t) * (size_t)n); fwd53_t(d, n, t); free(t); }
void gen_fwd97(double *d, int n) { if (n < 2) return; double *t = malloc(sizeof(double) * (size_t)n); fwd97_t(d, n, t); free(t); }

void gen_fwd2d53(int32_t *d, int w, int h)               /* rows then columns, dwt.go:356-407 */
{
    int m = w > h ? w : h;
    int32_t *tmp = malloc(sizeof(int32_t) * (size_t)m), *col = malloc(sizeof(int32_t) * (size_t)h);
    for (int y = 0; y < h; y++) fwd53_t(d + (size_t)y * w, w, tmp);
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) col[y] = d[(size_t)y * w + x];
        fwd53_t(col, h, tmp);
        for (int y = 0; y < h; y++) d[(size_t)y * w + x] = col[y];
    }
    free(tmp); free(col);
}

void gen_fwd2d97(double *d, int w, int h)                /* dwt.go:432-451 */
{
    int m = w > h ? w : h;
    double *tmp = malloc(sizeof(double) * (size_t)m), *col = malloc(sizeof(double) * (size_t)h);
    for (int y = 0; y < h; y++) fwd97_t(d + (size_t)y * w, w, tmp);
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) col[y] = d[(size_t)y * w + x];
        fwd97_t(col, h, tmp);
        for (int y = 0; y < h; y++) d[(size_t)y * w + x] = col[y];
    }
    free(tmp); free(col);
}

void gen_decompose53(int32_t *d, int w, int h, int levels)     /* dwt.go:524-531 */
{
    for (int l = 0; l < levels; l++) { gen_fwd2d53(d, w, h); w = (w + 1) / 2; h = (h + 1) / 2; }
}

void gen_decompose97(double *d, int w, int h, int levels)      /* dwt.go:551-558 */
{
    for (int l = 0; l < levels; l++) { gen_fwd2d97(d, w, h); w = (w + 1) / 2; h = (h + 1) / 2; }
}

void gen_quantize(const double *in, double step, int32_t *out, size_t n)   /* dwt.go:500-511 */
{
    double inv = 1.0 / step;
    for (size_t i = 0; i < n; i++)
        out[i] = in[i] >= 0 ? (int32_t)floor(in[i] * inv + 0.5) : (int32_t)ceil(in[i] * inv - 0.5);
}

void gen_fwd_rct(int32_t *r, int32_t *g, int32_t *b, size_t n)      /* mct.go:28-38 */
{
    for (size_t i = 0; i < n; i++) {
        int32_t y = WADD(WADD(r[i], WMUL(2, g[i])), b[i]) >> 2;
        int32_t u = WSUB(b[i], g[i]), v = WSUB(r[i], g[i]);
        r[i] = y; g[i] = u; b[i] = v;
    }
}

void gen_fwd_ict(double *r, double *g, double *b, size_t n)         /* mct.go:14-24 */
{
    for (size_t i = 0; i < n; i++) {
        double y  = 0.299 * r[i] + 0.587 * g[i] + 0.114 * b[i];
        double cb = -0.16875 * r[i] - 0.33126 * g[i] + 0.5 * b[i];
        double cr = 0.5 * r[i] - 0.41869 * g[i] - 0.08131 * b[i];
        r[i] = y; g[i] = cb; b[i] = cr;
    }
}

void gen_dc_shift_forward(int32_t *d, size_t n, int prec)           /* mct.go:96-101 */
{
    int32_t s = (int32_t)((uint32_t)1 << (prec - 1));
    for (size_t i = 0; i < n; i++) d[i] = WSUB(d[i], s);
}

