/*
 * orc_dwt.c -- restatement of internal/dwt/dwt.go (5-3 int32 and 9-7 float64
 * lifting, 2-D, multi-level) and of TileDecoder.ApplyInverseDWT
 * (internal/tcd/tcd.go:416-437).  Oracle / test infrastructure only.
 *
 * Kept from the reference: even-origin signals only; 2-D inverse = all
 * columns, then all rows (dwt.go:410-429); multi-level works on the DENSE
 * PREFIX of the buffer -- level l transforms data[0 : w_l*h_l] as a
 * w_l x h_l image of stride w_l (dwt.go:524-548), which is not the Mallat
 * layout; float64 math is mul-then-add without FMA (build with
 * -ffp-contract=off); int32 math wraps.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))

static const double kAlpha = -1.586134342059924;   /* dwt.go:150-157 */
static const double kBeta  = -0.052980118572961;
static const double kGamma = 0.882911075530934;
static const double kDelta = 0.443506852043971;
static const double kK     = 1.230174104914001;
static const double kKInv  = 0.812893066115961;

/* interleave dwt.go:287-306, 329-346 */
static void interleave_i(int32_t *d, int n, int32_t *tmp)
{
    int half = (n + 1) / 2;
    memcpy(tmp, d, sizeof(int32_t) * (size_t)n);
    for (int i = 0, j = 0; j < half; i += 2, j++) d[i] = tmp[j];
    for (int i = 1, j = half; j < n; i += 2, j++) d[i] = tmp[j];
}
static void interleave_f(double *d, int n, double *tmp)
{
    int half = (n + 1) / 2;
    memcpy(tmp, d, sizeof(double) * (size_t)n);
    for (int i = 0, j = 0; j < half; i += 2, j++) d[i] = tmp[j];
    for (int i = 1, j = half; j < n; i += 2, j++) d[i] = tmp[j];
}

static void inv53_t(int32_t *d, int n, int32_t *tmp)     /* dwt.go:122-147 */
{
    if (n < 2) return;
    interleave_i(d, n, tmp);
    d[0] = WSUB(d[0], WADD(WADD(d[1], d[1]), 2) >> 2);
    for (int i = 2; i < n - 1; i += 2) d[i] = WSUB(d[i], WADD(WADD(d[i - 1], d[i + 1]), 2) >> 2);
    if (n & 1) d[n - 1] = WSUB(d[n - 1], WADD(WADD(d[n - 2], d[n - 2]), 2) >> 2);
    for (int i = 1; i < n - 1; i += 2) d[i] = WADD(d[i], WADD(d[i - 1], d[i + 1]) >> 1);
    if ((n & 1) == 0) d[n - 1] = WADD(d[n - 1], d[n - 2]);
}

static void inv97_t(double *d, int n, double *tmp)       /* dwt.go:213-262 */
{
    if (n < 2) return;
    interleave_f(d, n, tmp);
    for (int i = 0; i < n; i += 2) d[i] *= kK;
    for (int i = 1; i < n; i += 2) d[i] *= kKInv;
    d[0] -= (2 * kDelta) * d[1];
    for (int i = 2; i < n - 1; i += 2) d[i] -= kDelta * (d[i - 1] + d[i + 1]);
    if (n & 1) d[n - 1] -= (2 * kDelta) * d[n - 2];
    for (int i = 1; i < n - 1; i += 2) d[i] -= kGamma * (d[i - 1] + d[i + 1]);
    if ((n & 1) == 0) d[n - 1] -= (2 * kGamma) * d[n - 2];
    d[0] -= (2 * kBeta) * d[1];
    for (int i = 2; i < n - 1; i += 2) d[i] -= kBeta * (d[i - 1] + d[i + 1]);
    if (n & 1) d[n - 1] -= (2 * kBeta) * d[n - 2];
    for (int i = 1; i < n - 1; i += 2) d[i] -= kAlpha * (d[i - 1] + d[i + 1]);
    if ((n & 1) == 0) d[n - 1] -= (2 * kAlpha) * d[n - 2];
}

void orc_inv53(int32_t *d, int n) { if (n < 2) return; int32_t *t = malloc(sizeof(int32_t) * (size_t)n); inv53_t(d, n, t); free(t); }
void orc_inv97(double *d, int n) { if (n < 2) return; double *t = malloc(sizeof(double) * (size_t)n); inv97_t(d, n, t); free(t); }

void orc_inv2d53(int32_t *d, int w, int h)               /* columns then rows, dwt.go:410-429 */
{
    int m = w > h ? w : h;
    int32_t *tmp = malloc(sizeof(int32_t) * (size_t)m), *col = malloc(sizeof(int32_t) * (size_t)h);
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) col[y] = d[(size_t)y * w + x];
        inv53_t(col, h, tmp);
        for (int y = 0; y < h; y++) d[(size_t)y * w + x] = col[y];
    }
    for (int y = 0; y < h; y++) inv53_t(d + (size_t)y * w, w, tmp);
    free(tmp); free(col);
}

void orc_inv2d97(double *d, int w, int h)                /* dwt.go:454-473 */
{
    int m = w > h ? w : h;
    double *tmp = malloc(sizeof(double) * (size_t)m), *col = malloc(sizeof(double) * (size_t)h);
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) col[y] = d[(size_t)y * w + x];
        inv97_t(col, h, tmp);
        for (int y = 0; y < h; y++) d[(size_t)y * w + x] = col[y];
    }
    for (int y = 0; y < h; y++) inv97_t(d + (size_t)y * w, w, tmp);
    free(tmp); free(col);
}

void orc_reconstruct53(int32_t *d, int w, int h, int levels)   /* dwt.go:534-548 */
{
    if (levels <= 0) return;
    int *ws = malloc(sizeof(int) * (size_t)levels), *hs = malloc(sizeof(int) * (size_t)levels);
    for (int l = 0; l < levels; l++) { ws[l] = w; hs[l] = h; w = (w + 1) / 2; h = (h + 1) / 2; }
    for (int l = levels - 1; l >= 0; l--) orc_inv2d53(d, ws[l], hs[l]);
    free(ws); free(hs);
}

void orc_reconstruct97(double *d, int w, int h, int levels)    /* dwt.go:561-573 */
{
    if (levels <= 0) return;
    int *ws = malloc(sizeof(int) * (size_t)levels), *hs = malloc(sizeof(int) * (size_t)levels);
    for (int l = 0; l < levels; l++) { ws[l] = w; hs[l] = h; w = (w + 1) / 2; h = (h + 1) / 2; }
    for (int l = levels - 1; l >= 0; l--) orc_inv2d97(d, ws[l], hs[l]);
    free(ws); free(hs);
}

void orc_dequantize(const int32_t *in, double step, double *out, size_t n) /* dwt.go:514-520 */
{
    for (size_t i = 0; i < n; i++) out[i] = (double)in[i] * step;
}

/* ApplyInverseDWT tcd.go:416-437: 9-7 goes int32 -> float64 -> IDWT -> int32(v + 0.5),
 * where the Go conversion truncates toward zero (so -1.7 -> -1). */
void orc_apply_inverse_dwt(int32_t *d, int w, int h, int levels, int reversible)
{
    if (reversible) { orc_reconstruct53(d, w, h, levels); return; }
    size_t n = (size_t)w * h;
    double *f = malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) f[i] = (double)d[i];
    orc_reconstruct97(f, w, h, levels);
    for (size_t i = 0; i < n; i++) d[i] = orc_f64_to_i32(f[i] + 0.5);
    free(f);
}
