/*
 * gen_mq.c -- restatement of the reference MQ ENCODER (internal/entropy/mqc.go:169-349).
 * Part of datagen/: the synthetic-input generator (reference encoder side).  Not the oracle,
 * not the product: it only manufactures code-block bitstreams for tests and bench inputs.
 */
#include "datagen.h"
#include "gen_mq.h"
#include <string.h>

uint32_t gen_mq_qe[94];
uint8_t  gen_mq_nmps[94];
uint8_t  gen_mq_nlps[94];

/* ISO/IEC 15444-1 Table C.2: Qe, NMPS, NLPS, SWITCH */
static const struct { uint16_t qe; uint8_t nmps, nlps, sw; } k_iso_c2[47] = {
    {0x5601, 1, 1, 1},  {0x3401, 2, 6, 0},  {0x1801, 3, 9, 0},  {0x0AC1, 4, 12, 0},
    {0x0521, 5, 29, 0}, {0x0221, 38, 33, 0}, {0x5601, 7, 6, 1},  {0x5401, 8, 14, 0},
    {0x4801, 9, 14, 0}, {0x3801, 10, 14, 0}, {0x3001, 11, 17, 0}, {0x2401, 12, 18, 0},
    {0x1C01, 13, 20, 0}, {0x1601, 29, 21, 0}, {0x5601, 15, 14, 1}, {0x5401, 16, 14, 0},
    {0x5101, 17, 15, 0}, {0x4801, 18, 16, 0}, {0x3801, 19, 17, 0}, {0x3401, 20, 18, 0},
    {0x3001, 21, 19, 0}, {0x2801, 22, 19, 0}, {0x2401, 23, 20, 0}, {0x2201, 24, 21, 0},
    {0x1C01, 25, 22, 0}, {0x1801, 26, 23, 0}, {0x1601, 27, 24, 0}, {0x1401, 28, 25, 0},
    {0x1201, 29, 26, 0}, {0x1101, 30, 27, 0}, {0x0AC1, 31, 28, 0}, {0x09C1, 32, 29, 0},
    {0x08A1, 33, 30, 0}, {0x0521, 34, 31, 0}, {0x0441, 35, 32, 0}, {0x02A1, 36, 33, 0},
    {0x0221, 37, 34, 0}, {0x0141, 38, 35, 0}, {0x0111, 39, 36, 0}, {0x0085, 40, 37, 0},
    {0x0049, 41, 38, 0}, {0x0025, 42, 39, 0}, {0x0015, 43, 40, 0}, {0x0009, 44, 41, 0},
    {0x0005, 45, 42, 0}, {0x0001, 45, 43, 0}, {0x5601, 46, 46, 0},
};

static int g_tables_ready;

void gen_mq_tables_init(void)
{
    if (g_tables_ready) return;
    for (int i = 0; i < 47; i++) {
        for (int m = 0; m < 2; m++) {
            int s = 2 * i + m;
            gen_mq_qe[s]   = k_iso_c2[i].qe;
            gen_mq_nmps[s] = (uint8_t)(2 * k_iso_c2[i].nmps + m);
            gen_mq_nlps[s] = (uint8_t)(2 * k_iso_c2[i].nlps + (m ^ k_iso_c2[i].sw));
        }
    }
    g_tables_ready = 1;
}

/* ---- encoder: NewMQEncoder mqc.go:185-201 ---------------------------------- */
void gen_mqenc_init(gen_mqenc *e, uint8_t *buf, int cap)
{
    gen_mq_tables_init();
    e->A = 0x8000; e->C = 0; e->CT = 12;
    e->buf = buf; e->cap = cap; e->bp = 0; e->overflow = 0;
    if (cap > 0) buf[0] = 0;                 /* the dummy byte "bp[-1]" */
    else e->overflow = 1;
    memset(e->ctx, 0, sizeof e->ctx);        /* every context starts in state 0 ... */
    e->ctx[GEN_CTX_UNI] = 92;                /* ... except UNI (mqc.go:194-199)      */
}

static void enc_put(gen_mqenc *e, uint8_t b)
{
    e->bp++;
    if (e->bp >= e->cap) { e->overflow = 1; e->bp = e->cap - 1; return; }
    e->buf[e->bp] = b;
}

/* byteOut mqc.go:270-310 */
static void enc_byte_out(gen_mqenc *e)
{
    if (e->overflow) return;
    if (e->buf[e->bp] == 0xFF) {
        enc_put(e, (uint8_t)(e->C >> 20));
        e->C &= 0xFFFFF; e->CT = 7;
    } else if ((e->C & 0x8000000u) == 0) {
        enc_put(e, (uint8_t)(e->C >> 19));
        e->C &= 0x7FFFF; e->CT = 8;
    } else {
        e->buf[e->bp]++;
        if (e->buf[e->bp] == 0xFF) {
            e->C &= 0x7FFFFFFu;
            enc_put(e, (uint8_t)(e->C >> 20));
            e->C &= 0xFFFFF; e->CT = 7;
        } else {
            enc_put(e, (uint8_t)(e->C >> 19));
            e->C &= 0x7FFFF; e->CT = 8;
        }
    }
}

/* renormEnc mqc.go:258-267 */
static void enc_renorm(gen_mqenc *e)
{
    while ((e->A & 0x8000) == 0) {
        e->A <<= 1; e->C <<= 1; e->CT--;
        if (e->CT == 0) enc_byte_out(e);
    }
}

/* Encode mqc.go:224-255 */
void gen_mqenc_encode(gen_mqenc *e, int ctx, int d)
{
    uint8_t s = e->ctx[ctx];
    uint32_t qe = gen_mq_qe[s];
    int mps = s & 1;
    e->A -= qe;
    if ((d & 1) == mps) {
        if ((e->A & 0x8000) == 0) {
            if (e->A < qe) e->A = qe; else e->C += qe;
            e->ctx[ctx] = gen_mq_nmps[s];
            enc_renorm(e);
        } else {
            e->C += qe;
        }
    } else {
        if (e->A < qe) e->C += qe; else e->A = qe;
        e->ctx[ctx] = gen_mq_nlps[s];
        enc_renorm(e);
    }
}

/* Flush mqc.go:313-341 (setbits, two byteOuts, drop trailing 0xFF and the dummy byte) */
int gen_mqenc_flush(gen_mqenc *e, const uint8_t **start)
{
    uint32_t tempC = e->C + e->A;
    e->C |= 0xFFFF;
    if (e->C >= tempC) e->C -= 0x8000;
    e->C <<= e->CT; enc_byte_out(e);
    e->C <<= e->CT; enc_byte_out(e);
    if (e->overflow) { *start = NULL; return -1; }
    int end = e->bp + 1;
    if (end > 0 && e->buf[end - 1] == 0xFF) end--;
    if (end > 1) { *start = e->buf + 1; return end - 1; }
    *start = NULL;
    return 0;
}

/* ---- flat entry point ------------------------------------------------------------ */
int gen_mq_encode(const uint8_t *ctxs, const uint8_t *bits, int n, uint8_t *out, int cap)
{
    int tmpcap = 2 * n + 64;
    uint8_t stackbuf[4096];
    uint8_t *buf = stackbuf;
    uint8_t *heap = NULL;
    if (tmpcap > (int)sizeof stackbuf) { heap = (uint8_t *)__builtin_malloc((size_t)tmpcap); buf = heap; }
    gen_mqenc e;
    gen_mqenc_init(&e, buf, tmpcap);
    for (int i = 0; i < n; i++) gen_mqenc_encode(&e, ctxs[i], bits[i]);
    const uint8_t *start;
    int len = gen_mqenc_flush(&e, &start);
    if (len > cap) len = -1;
    if (len > 0) memcpy(out, start, (size_t)len);
    if (heap) __builtin_free(heap);
    return len;
}

