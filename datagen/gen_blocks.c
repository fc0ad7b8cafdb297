/*
 * gen_blocks.c -- parallel block-encode driver for datagen/ (the shape of the reference encoder's
 * code-block worker pool, encoder.go:690-742): every block is encoded independently with
 * gen_t1_encode / gen_ht_encode and the bitstreams are concatenated in block order.
 */
#include "datagen.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const int32_t *planes; const gen_blk_t *blks; uint32_t n;
    uint8_t **bufs; int *blen; uint8_t *nbps;
    volatile uint32_t next;
} pool_t;

static void *worker(void *arg)
{
    pool_t *p = (pool_t *)arg;
    int32_t tmp[64 * 64];
    for (;;) {
        uint32_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        const gen_blk_t *b = &p->blks[i];
        if (b->w == 0 || b->h == 0 || b->w > 64 || b->h > 64) { p->blen[i] = -1; continue; }
        const int32_t *src = p->planes + b->plane_off + (size_t)b->y0 * b->stride + b->x0;
        for (int y = 0; y < b->h; y++) memcpy(tmp + y * b->w, src + (size_t)y * b->stride, sizeof(int32_t) * b->w);
        int cap = b->w * b->h * 4 + 16384;
        uint8_t *buf = malloc((size_t)cap);
        int nb = 0, len;
        if (b->ht) len = gen_ht_encode(tmp, b->w, b->h, b->band, buf, cap);
        else len = gen_t1_encode(tmp, b->w, b->h, b->band, buf, cap, &nb);
        p->bufs[i] = buf; p->blen[i] = len; p->nbps[i] = (uint8_t)nb;
    }
    return NULL;
}

int64_t gen_encode_blocks(const int32_t *planes, const gen_blk_t *blks, uint32_t n, uint8_t *out, uint64_t cap,
                          uint64_t *offs, uint32_t *lens, uint8_t *nbps, int threads)
{
    pool_t p; memset(&p, 0, sizeof p);
    p.planes = planes; p.blks = blks; p.n = n; p.nbps = nbps;
    p.bufs = calloc(n ? n : 1, sizeof(uint8_t *));
    p.blen = calloc(n ? n : 1, sizeof(int));
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &p);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    int64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (total >= 0 && (p.blen[i] < 0 || (uint64_t)total + (uint64_t)p.blen[i] > cap)) total = -1;
        if (total >= 0) {
            offs[i] = (uint64_t)total; lens[i] = (uint32_t)p.blen[i];
            if (p.blen[i] > 0) memcpy(out + total, p.bufs[i], (size_t)p.blen[i]);
            total += p.blen[i];
        }
        free(p.bufs[i]);
    }
    free(p.bufs); free(p.blen);
    return total;
}
