/*
 * orc_t1.c -- restatement of the reference EBCOT tier-1 block coder
 * (internal/entropy/t1.go, t1_luts.go, t1_fast5.go).  Oracle / test
 * infrastructure only.
 *
 * Reference behaviour kept on purpose:
 *  - significance-propagation and magnitude-refinement passes scan in RASTER
 *    order (t1.go:1298-1299, 1334-1335); only the cleanup pass scans 4-row
 *    stripes column by column (t1.go:1353-1354);
 *  - every bit-plane runs all three passes, including the first;
 *  - run-length mode only when the full 4-row column exists (t1.go:1196);
 *  - the ZC table of t1_luts.go:35-110 and the sign rule of t1.go:387-460
 *    (vc is not negated when hc != 0);
 *  - all MQ contexts start in state 0 (UNI in 92).
 */
#include "oracle.h"
#include "orc_mq.h"
#include <stdlib.h>
#include <string.h>

/* flag bits, t1.go:72-91 (the four directional bits are an encoder-side
 * shortcut only, t1_fast5.go:91; they never change a coded decision) */
enum { F_SIG = 1, F_VISIT = 2, F_REFINE = 4, F_NEG = 8 };

static uint8_t g_zc_lut[4 * 256];
static int g_zc_ready;

/* t1_luts.go:35-110 */
static void zc_lut_init(void)
{
    if (g_zc_ready) return;
    for (int band = 0; band < 4; band++) {
        for (int p = 0; p < 256; p++) {
            int hcount = (p & 1) + ((p >> 1) & 1);
            int vcount = ((p >> 2) & 1) + ((p >> 3) & 1);
            int dcount = ((p >> 4) & 1) + ((p >> 5) & 1) + ((p >> 6) & 1) + ((p >> 7) & 1);
            int ctx;
            if (band == ORC_BAND_HH) {
                int hv = hcount + vcount;
                if (hv >= 3) ctx = 8;
                else if (hv == 2) ctx = dcount >= 2 ? 7 : (dcount >= 1 ? 6 : 5);
                else if (hv == 1) ctx = dcount >= 2 ? 4 : 3;
                else ctx = dcount >= 2 ? 2 : (dcount >= 1 ? 1 : 0);
            } else {
                if (band == ORC_BAND_HL) { int t = hcount; hcount = vcount; vcount = t; }
                if (hcount == 2) ctx = 8;
                else if (hcount == 1) ctx = vcount >= 1 ? 7 : (dcount >= 1 ? 6 : 5);
                else if (vcount == 2) ctx = 4;
                else if (vcount == 1) ctx = dcount >= 1 ? 3 : 2;
                else ctx = dcount >= 2 ? 1 : 0;
            }
            g_zc_lut[band * 256 + p] = (uint8_t)ctx;
        }
    }
    g_zc_ready = 1;
}

const uint8_t *orc_t1_zc_lut(void) { zc_lut_init(); return g_zc_lut; }

typedef struct {
    int w, h, stride, band;
    uint8_t *flags;      /* (w+2)*(h+2), 1-sample border, t1.go:138 */
    int32_t *mag;        /* w*h magnitudes */
} t1_state;

#define FIDX(t, x, y) (((y) + 1) * (t)->stride + (x) + 1)

/* getZCContext t1.go:349-384 */
static int zc_context(const t1_state *t, int i)
{
    const uint8_t *f = t->flags;
    int s = t->stride, p = 0;
    if (f[i - 1] & F_SIG) p |= 0x01;
    if (f[i + 1] & F_SIG) p |= 0x02;
    if (f[i - s] & F_SIG) p |= 0x04;
    if (f[i + s] & F_SIG) p |= 0x08;
    if (f[i - s - 1] & F_SIG) p |= 0x10;
    if (f[i - s + 1] & F_SIG) p |= 0x20;
    if (f[i + s - 1] & F_SIG) p |= 0x40;
    if (f[i + s + 1] & F_SIG) p |= 0x80;
    return g_zc_lut[t->band * 256 + p];
}

/* hasSignificantNeighbor t1.go:1087-1092 */
static int has_sig_neighbor(const t1_state *t, int i)
{
    const uint8_t *f = t->flags;
    int s = t->stride;
    return ((f[i - 1] | f[i + 1] | f[i - s] | f[i + s] |
             f[i - s - 1] | f[i - s + 1] | f[i + s - 1] | f[i + s + 1]) & F_SIG) != 0;
}

/* getSCContext t1.go:387-460 */
static int sc_context(const t1_state *t, int i, int *pred)
{
    const uint8_t *f = t->flags;
    int s = t->stride, hc = 0, vc = 0;
    if (f[i - 1] & F_SIG) hc += (f[i - 1] & F_NEG) ? -1 : 1;
    if (f[i + 1] & F_SIG) hc += (f[i + 1] & F_NEG) ? -1 : 1;
    if (f[i - s] & F_SIG) vc += (f[i - s] & F_NEG) ? -1 : 1;
    if (f[i + s] & F_SIG) vc += (f[i + s] & F_NEG) ? -1 : 1;
    *pred = 0;
    if (hc < 0) { *pred = 1; hc = -hc; }
    if (hc == 0 && vc < 0) { *pred = 1; vc = -vc; }
    int ctx = ORC_CTX_SC0;
    if (hc == 1) ctx = ORC_CTX_SC0 + (vc == 1 ? 4 : (vc == 0 ? 2 : 1));
    else if (hc == 0) ctx = ORC_CTX_SC0 + (vc == 1 ? 1 : 0);
    else if (hc == 2) ctx = ORC_CTX_SC0 + 3;
    return ctx;
}

/* getMRContext t1.go:463-479 */
static int mr_context(const t1_state *t, int i)
{
    if (t->flags[i] & F_REFINE) return ORC_CTX_MAG0 + 2;
    return has_sig_neighbor(t, i) ? ORC_CTX_MAG0 + 1 : ORC_CTX_MAG0;
}

/* canUseRunLength t1.go:1195-1208 */
static int can_run_length(const t1_state *t, int x, int y)
{
    if (y + 4 > t->h) return 0;
    for (int yy = y; yy < y + 4; yy++) {
        int i = FIDX(t, x, yy);
        if (t->flags[i] & (F_SIG | F_VISIT)) return 0;
        if (has_sig_neighbor(t, i)) return 0;
    }
    return 1;
}

/* ============================ decoder ======================================== */

static void dec_sign(t1_state *t, orc_mqdec *mq, int i)            /* t1.go:1322-1328 */
{
    int pred, ctx = sc_context(t, i, &pred);
    if (orc_mqdec_decode(mq, ctx) ^ pred) t->flags[i] |= F_NEG;
}

static void dec_new_sig(t1_state *t, orc_mqdec *mq, int x, int y, int32_t bit)
{
    int i = FIDX(t, x, y);
    t->mag[y * t->w + x] = bit;
    dec_sign(t, mq, i);
    t->flags[i] |= F_SIG;
}

void orc_t1_decode(const uint8_t *data, int len, int w, int h, int num_bps, int band, int32_t *out)
{
    zc_lut_init();
    t1_state t;
    t.w = w; t.h = h; t.stride = w + 2; t.band = band;
    t.flags = (uint8_t *)calloc((size_t)(w + 2) * (h + 2), 1);
    t.mag = out;                                  /* decode magnitudes in place, sign applied last */
    memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    orc_mqdec mq;
    orc_mqdec_init(&mq, data, len);               /* t1.go:1264 */

    for (int bp = num_bps - 1; bp >= 0; bp--) {   /* t1.go:1275-1279 */
        int32_t bit = (int32_t)((uint32_t)1 << bp);
        /* significance propagation, raster order t1.go:1295-1319 */
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                int i = FIDX(&t, x, y);
                if (t.flags[i] & F_SIG) continue;
                if (!has_sig_neighbor(&t, i)) continue;
                if (orc_mqdec_decode(&mq, zc_context(&t, i))) dec_new_sig(&t, &mq, x, y, bit);
                t.flags[i] |= F_VISIT;
            }
        }
        /* magnitude refinement, raster order t1.go:1331-1347 */
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                int i = FIDX(&t, x, y);
                if (!(t.flags[i] & F_SIG) || (t.flags[i] & F_VISIT)) continue;
                if (orc_mqdec_decode(&mq, mr_context(&t, i))) t.mag[y * w + x] |= bit;
                t.flags[i] |= F_REFINE;
            }
        }
        /* cleanup, 4-row stripes column by column t1.go:1350-1381 */
        for (int y = 0; y < h; y += 4) {
            for (int x = 0; x < w; x++) {
                if (can_run_length(&t, x, y)) {   /* decodeRunLength t1.go:1384-1410 */
                    if (!orc_mqdec_decode(&mq, ORC_CTX_RL)) continue;
                    int pos = orc_mqdec_decode(&mq, ORC_CTX_UNI) << 1;
                    pos |= orc_mqdec_decode(&mq, ORC_CTX_UNI);
                    dec_new_sig(&t, &mq, x, y + pos, bit);
                    for (int k = pos + 1; k < 4 && y + k < h; k++) {
                        int i = FIDX(&t, x, y + k);
                        if (orc_mqdec_decode(&mq, zc_context(&t, i))) dec_new_sig(&t, &mq, x, y + k, bit);
                    }
                    continue;
                }
                for (int yy = y; yy < y + 4 && yy < h; yy++) {
                    int i = FIDX(&t, x, yy);
                    if (t.flags[i] & F_VISIT) { t.flags[i] &= (uint8_t)~F_VISIT; continue; }
                    if (t.flags[i] & F_SIG) continue;
                    if (orc_mqdec_decode(&mq, zc_context(&t, i))) dec_new_sig(&t, &mq, x, yy, bit);
                }
            }
        }
    }
    /* apply signs t1.go:1282-1289 (Go int32 negation wraps) */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (t.flags[FIDX(&t, x, y)] & F_NEG)
                out[y * w + x] = (int32_t)(0u - (uint32_t)out[y * w + x]);
    free(t.flags);
}

/* ============================ encoder ======================================== */
/* SetData t1.go:292-304 + EncodeFast5 t1_fast5.go:10-899 (same decisions as
 * EncodeSafe t1.go:923-947; the directional-flag shortcut only skips samples
 * that have no significant neighbour). */

static void enc_sign(t1_state *t, orc_mqenc *mq, int i)            /* t1.go:482-555 */
{
    int pred, ctx = sc_context(t, i, &pred);
    int sign = (t->flags[i] & F_NEG) ? 1 : 0;
    orc_mqenc_encode(mq, ctx, sign ^ pred);
}

int orc_t1_encode(const int32_t *coeffs, int w, int h, int band, uint8_t *out, int cap, int *num_bps)
{
    zc_lut_init();
    t1_state t;
    t.w = w; t.h = h; t.stride = w + 2; t.band = band;
    t.flags = (uint8_t *)calloc((size_t)(w + 2) * (h + 2), 1);
    t.mag = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * h);
    int32_t maxv = 0;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int32_t v = coeffs[y * w + x];
            if (v < 0) { v = (int32_t)(0u - (uint32_t)v); t.flags[FIDX(&t, x, y)] |= F_NEG; }
            t.mag[y * w + x] = v;
            if (v > maxv) maxv = v;
        }
    }
    int nbps = 0;
    for (int32_t m = maxv; m > 0; m >>= 1) nbps++;                  /* t1_fast5.go:23-27 */
    if (num_bps) *num_bps = nbps;
    if (maxv == 0) { free(t.flags); free(t.mag); return 0; }        /* t1_fast5.go:20-22 -> nil */

    int tmpcap = w * h * 4 + 16384;
    uint8_t *buf = (uint8_t *)malloc((size_t)tmpcap);
    orc_mqenc mq;
    orc_mqenc_init(&mq, buf, tmpcap);

    for (int bp = nbps - 1; bp >= 0; bp--) {
        int32_t bit = (int32_t)((uint32_t)1 << bp);
        /* SPP t1.go:558-639 */
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                int i = FIDX(&t, x, y);
                if (t.flags[i] & F_SIG) continue;
                if (!has_sig_neighbor(&t, i)) continue;
                int sig = (t.mag[y * w + x] & bit) != 0;
                orc_mqenc_encode(&mq, zc_context(&t, i), sig);
                if (sig) { enc_sign(&t, &mq, i); t.flags[i] |= F_SIG; }
                t.flags[i] |= F_VISIT;
            }
        }
        /* MRP t1.go:642-683 */
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                int i = FIDX(&t, x, y);
                if (!(t.flags[i] & F_SIG) || (t.flags[i] & F_VISIT)) continue;
                orc_mqenc_encode(&mq, mr_context(&t, i), (t.mag[y * w + x] & bit) != 0);
                t.flags[i] |= F_REFINE;
            }
        }
        /* cleanup t1.go:686-770, run length t1.go:816-914 */
        for (int y = 0; y < h; y += 4) {
            for (int x = 0; x < w; x++) {
                if (can_run_length(&t, x, y)) {
                    int first = -1;
                    for (int k = 0; k < 4; k++)
                        if (t.mag[(y + k) * w + x] & bit) { first = k; break; }
                    if (first < 0) { orc_mqenc_encode(&mq, ORC_CTX_RL, 0); continue; }
                    orc_mqenc_encode(&mq, ORC_CTX_RL, 1);
                    orc_mqenc_encode(&mq, ORC_CTX_UNI, (first >> 1) & 1);
                    orc_mqenc_encode(&mq, ORC_CTX_UNI, first & 1);
                    int i = FIDX(&t, x, y + first);
                    enc_sign(&t, &mq, i);
                    t.flags[i] |= F_SIG;
                    for (int k = first + 1; k < 4 && y + k < h; k++) {
                        i = FIDX(&t, x, y + k);
                        int sig = (t.mag[(y + k) * w + x] & bit) != 0;
                        orc_mqenc_encode(&mq, zc_context(&t, i), sig);
                        if (sig) { enc_sign(&t, &mq, i); t.flags[i] |= F_SIG; }
                    }
                    continue;
                }
                for (int yy = y; yy < y + 4 && yy < h; yy++) {
                    int i = FIDX(&t, x, yy);
                    if (t.flags[i] & F_VISIT) { t.flags[i] &= (uint8_t)~F_VISIT; continue; }
                    if (t.flags[i] & F_SIG) continue;
                    int sig = (t.mag[yy * w + x] & bit) != 0;
                    orc_mqenc_encode(&mq, zc_context(&t, i), sig);
                    if (sig) { enc_sign(&t, &mq, i); t.flags[i] |= F_SIG; }
                }
            }
        }
    }
    const uint8_t *start;
    int n = orc_mqenc_flush(&mq, &start);                            /* t1_fast5.go:878-898 */
    if (n > cap) n = -1;
    if (n > 0) memcpy(out, start, (size_t)n);
    free(buf); free(t.flags); free(t.mag);
    return n;
}
