/*
 * orc_enc.c -- CPU checker of the FORWARD path: restatement of encoder.extractImageData
 * (encoder.go:79-214), encoder.preprocess (encoder.go:216-281) and encoder.encodeTile up to createTileHeader
 * (encoder.go:597-743, extractCodeBlockData :763-796).  Test infrastructure, like the rest of oracle/: used by tests/,
 * smoke() and bench.py's CPU legs only.  The transforms and the tier-1 / MQ encoders are the restatements the input
 * generator already holds (datagen/gen_fwd.c, gen_t1.c, gen_mq.c: dwt.go:73-210, 356-451, 524-558, mct.go:14-38, 96-101,
 * t1.go:292-304, t1_fast5.go:10-899, mqc.go:185-341), compiled into this library as they are; they are pinned by the
 * reference's own test strategy for the encoder, the exact round trip (decode(encode(x)) == x through orc_t1.c, which the
 * reference's golden vectors pin).
 */
#include "oracle.h"
#include "../datagen/datagen.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int w, h, nc, pix_bits, prec, lossless, levels, num_res, cbw, cbh; double step; } enc_geom;

static int geometry(const orc_encode_t *p, enc_geom *g)
{
    if (!p->width || !p->height || (p->ncomp != 1 && p->ncomp != 3 && p->ncomp != 4) || (p->pix_bits != 8 && p->pix_bits != 16)) return -1;
    g->w = (int)p->width; g->h = (int)p->height; g->nc = p->ncomp; g->pix_bits = p->pix_bits;
    g->prec = (p->precision > 0 && p->precision <= 16 && p->precision != p->pix_bits) ? p->precision : p->pix_bits;   /* encoder.go:197 */
    g->lossless = p->lossless != 0;
    g->levels = (int)p->num_resolutions - 1;                       /* encoder.go:249-252 */
    if (g->levels <= 0) g->levels = 5;
    g->num_res = p->num_resolutions ? p->num_resolutions : 6;      /* encoder.go:601-604 */
    g->cbw = 1 << (p->cb_x + 2); g->cbh = 1 << (p->cb_y + 2);      /* encoder.go:606-607 */
    g->step = 1.0 / (double)(p->quality > 0 ? p->quality : 100);   /* encoder.go:265-269 */
    return 0;
}

/* extractImageData + preprocess -> planes[nc][h][w] */
int orc_encode_preprocess(const orc_encode_t *p, const uint8_t *pix, uint64_t stride, int32_t *planes)
{
    enc_geom g;
    if (geometry(p, &g)) return -1;
    const size_t n = (size_t)g.w * g.h;
    for (int y = 0; y < g.h; y++) {                                /* encoder.go:84-181 */
        const uint8_t *row = pix + (size_t)y * stride;
        for (int x = 0; x < g.w; x++)
            for (int c = 0; c < g.nc; c++) {
                int32_t v;
                if (g.nc == 1) v = g.pix_bits == 8 ? row[x] : (row[2 * x] << 8 | row[2 * x + 1]);
                else v = g.pix_bits == 8 ? row[4 * x + c] : (row[8 * x + 2 * c] << 8 | row[8 * x + 2 * c + 1]);
                planes[c * n + (size_t)y * g.w + x] = v;
            }
    }
    if (g.prec != g.pix_bits) {                                    /* encoder.go:197-211 */
        const int32_t src_max = (int32_t)((1u << g.pix_bits) - 1), dst_max = (int32_t)((1u << g.prec) - 1);
        for (size_t i = 0; i < n * g.nc; i++) planes[i] = planes[i] * dst_max / src_max;
    }
    for (int c = 0; c < g.nc; c++) gen_dc_shift_forward(planes + c * n, n, g.prec);      /* encoder.go:218-220 */
    double *f = g.lossless ? NULL : malloc(sizeof(double) * n * 3);
    if (g.nc >= 3) {
        if (g.lossless) gen_fwd_rct(planes, planes + n, planes + 2 * n, n);
        else {                                                     /* encoder.go:227-246 */
            for (size_t i = 0; i < 3 * n; i++) f[i] = (double)planes[i];
            gen_fwd_ict(f, f + n, f + 2 * n, n);
            for (size_t i = 0; i < 3 * n; i++) planes[i] = f[i] >= 0 ? (int32_t)(f[i] + 0.5) : (int32_t)(f[i] - 0.5);
        }
    }
    for (int c = 0; c < g.nc; c++) {                               /* encoder.go:254-278 */
        int32_t *d = planes + c * n;
        if (g.lossless) gen_decompose53(d, g.w, g.h, g.levels);
        else {
            for (size_t i = 0; i < n; i++) f[i] = (double)d[i];
            gen_decompose97(f, g.w, g.h, g.levels);
            for (size_t i = 0; i < n; i++) d[i] = f[i] >= 0 ? (int32_t)(f[i] / g.step + 0.5) : (int32_t)(f[i] / g.step - 0.5);
        }
    }
    free(f);
    return 0;
}

typedef struct { int c, sx, sy, w, h, band; } enc_blk;

/* encodeTile's job list, encoder.go:615-673 */
static uint32_t block_list(const enc_geom *g, enc_blk *out)
{
    uint32_t n = 0;
    for (int c = 0; c < g->nc; c++)
        for (int r = 0; r < g->num_res; r++) {
            const int nb = r == 0 ? 1 : 3;
            for (int b = 0; b < nb; b++) {
                const int band = r == 0 ? GEN_BAND_LL : (b == 0 ? GEN_BAND_HL : (b == 1 ? GEN_BAND_LH : GEN_BAND_HH));
                const int64_t scale = (int64_t)1 << (g->num_res - 1 - r);
                int bw = (int)((g->w + scale - 1) / scale), bh = (int)((g->h + scale - 1) / scale);
                if (r > 0) { bw = (bw + 1) / 2; bh = (bh + 1) / 2; }
                for (int cby = 0; cby * g->cbh < bh; cby++)
                    for (int cbx = 0; cbx * g->cbw < bw; cbx++, n++) {
                        if (!out) continue;
                        enc_blk *e = &out[n];
                        e->c = c; e->band = band; e->sx = cbx * g->cbw; e->sy = cby * g->cbh;
                        e->w = g->cbw; e->h = g->cbh;
                        if (e->sx + e->w > bw) e->w = bw - e->sx;
                        if (e->sy + e->h > bh) e->h = bh - e->sy;
                    }
            }
        }
    return n;
}

uint32_t orc_encode_block_count(const orc_encode_t *p)
{
    enc_geom g;
    return geometry(p, &g) ? 0 : block_list(&g, NULL);
}

typedef struct {
    const enc_geom *g; const enc_blk *blks; uint32_t n; const int32_t *planes;
    uint8_t **bufs; int *blen; uint8_t *bps; volatile uint32_t next;
} enc_pool;

static void *enc_worker(void *arg)
{
    enc_pool *p = (enc_pool *)arg;
    const enc_geom *g = p->g;
    int32_t *tmp = malloc(sizeof(int32_t) * (size_t)g->cbw * g->cbh);
    for (;;) {
        const uint32_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        const enc_blk *b = &p->blks[i];
        const int32_t *plane = p->planes + (size_t)b->c * g->w * g->h;
        for (int y = 0; y < b->h; y++)                             /* extractCodeBlockData, encoder.go:778-793 */
            for (int x = 0; x < b->w; x++) {
                const int sx = b->sx + x, sy = b->sy + y;
                tmp[y * b->w + x] = (sx < g->w && sy < g->h) ? plane[(size_t)sy * g->w + sx] : 0;
            }
        const int cap = b->w * b->h * 8 + 16384;
        uint8_t *buf = malloc((size_t)cap);
        int nb = 0;
        p->blen[i] = gen_t1_encode(tmp, b->w, b->h, b->band, buf, cap, &nb);
        p->bufs[i] = buf; p->bps[i] = (uint8_t)nb;
    }
    free(tmp);
    return NULL;
}

/* extractImageData + preprocess + encodeTile's tileData; returns the byte count, -1: bad options, -2: out too small */
int64_t orc_encode_tile(const orc_encode_t *p, const uint8_t *pix, uint64_t stride, uint8_t *out, uint64_t cap,
                        uint32_t *blk_len, uint8_t *blk_bps, int threads)
{
    enc_geom g;
    if (geometry(p, &g)) return -1;
    int32_t *planes = malloc(sizeof(int32_t) * (size_t)g.w * g.h * g.nc);
    orc_encode_preprocess(p, pix, stride, planes);
    const uint32_t n = block_list(&g, NULL);
    enc_blk *blks = malloc(sizeof(enc_blk) * (n ? n : 1));
    block_list(&g, blks);
    enc_pool pool; memset(&pool, 0, sizeof pool);
    pool.g = &g; pool.blks = blks; pool.n = n; pool.planes = planes;
    pool.bufs = calloc(n ? n : 1, sizeof(uint8_t *)); pool.blen = calloc(n ? n : 1, sizeof(int)); pool.bps = calloc(n ? n : 1, 1);
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, enc_worker, &pool);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    int64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {                             /* results in list order, encoder.go:726-740 */
        const int len = pool.blen[i] > 0 ? pool.blen[i] : 0;
        if (total >= 0 && (pool.blen[i] < 0 || (uint64_t)total + (uint64_t)len > cap)) total = -2;
        if (total >= 0) { if (len) memcpy(out + total, pool.bufs[i], (size_t)len); total += len; }
        if (blk_len) blk_len[i] = (uint32_t)len;
        if (blk_bps) blk_bps[i] = pool.bps[i];
        free(pool.bufs[i]);
    }
    free(pool.bufs); free(pool.blen); free(pool.bps); free(blks); free(planes);
    return total;
}
