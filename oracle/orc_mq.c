/*
 * orc_mq.c -- restatement of the reference MQ arithmetic coder
 * (internal/entropy/mqc.go).  Oracle / test infrastructure only.
 *
 * The 94-entry state table of mqc.go:21-116 is the ISO/IEC 15444-1 Table C.2
 * 47-state machine expanded to (2*state + mps); it is rebuilt here from the
 * 47 standard rows.  Entry 46 is the self-looping uniform state (-> 92/93).
 */
#include "orc_mq.h"
#include "oracle.h"
#include <string.h>

uint32_t orc_mq_qe[94];
uint8_t  orc_mq_nmps[94];
uint8_t  orc_mq_nlps[94];

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

void orc_mq_tables_init(void)
{
    if (g_tables_ready) return;
    for (int i = 0; i < 47; i++) {
        for (int m = 0; m < 2; m++) {
            int s = 2 * i + m;
            orc_mq_qe[s]   = k_iso_c2[i].qe;
            orc_mq_nmps[s] = (uint8_t)(2 * k_iso_c2[i].nmps + m);
            orc_mq_nlps[s] = (uint8_t)(2 * k_iso_c2[i].nlps + (m ^ k_iso_c2[i].sw));
        }
    }
    g_tables_ready = 1;
}

/* ---- decoder: NewMQDecoder mqc.go:370-399 ---------------------------------- */
static void dec_byte_in(orc_mqdec *d);

void orc_mqdec_init(orc_mqdec *d, const uint8_t *data, int len)
{
    orc_mq_tables_init();
    d->A = 0x8000; d->C = 0; d->CT = 0;
    d->data = data; d->len = len; d->bp = -1;
    memset(d->ctx, 0, sizeof d->ctx);
    d->ctx[ORC_CTX_UNI] = 92;
    if (len == 0) d->C = 0xFFu << 16;
    else { d->bp = 0; d->C = (uint32_t)data[0] << 16; }
    dec_byte_in(d);
    d->C <<= 7;
    d->CT -= 7;
    d->A = 0x8000;
}

/* byteIn mqc.go:402-439 */
static void dec_byte_in(orc_mqdec *d)
{
    if (d->bp < 0) d->bp = 0;
    if (d->bp >= d->len) { d->C += 0xFF00; d->CT = 8; return; }
    uint8_t next = (d->bp + 1 < d->len) ? d->data[d->bp + 1] : 0xFF;
    if (d->data[d->bp] == 0xFF) {
        if (next > 0x8F) { d->C += 0xFF00; d->CT = 8; }          /* marker: do not advance */
        else { d->bp++; d->C += (uint32_t)next << 9; d->CT = 7; }
    } else {
        d->bp++; d->C += (uint32_t)next << 8; d->CT = 8;
    }
}

/* renormDec mqc.go:488-497 */
static void dec_renorm(orc_mqdec *d)
{
    while ((d->A & 0x8000) == 0) {
        if (d->CT == 0) dec_byte_in(d);
        d->A <<= 1; d->C <<= 1; d->CT--;
    }
}

/* Decode mqc.go:443-485 */
int orc_mqdec_decode(orc_mqdec *d, int ctx)
{
    uint8_t s = d->ctx[ctx];
    uint32_t qe = orc_mq_qe[s];
    int mps = s & 1, dec;
    d->A -= qe;
    if ((d->C >> 16) < qe) {
        if (d->A < qe) { d->A = qe; dec = mps;     d->ctx[ctx] = orc_mq_nmps[s]; }
        else           { d->A = qe; dec = 1 - mps; d->ctx[ctx] = orc_mq_nlps[s]; }
        dec_renorm(d);
        return dec;
    }
    d->C -= qe << 16;
    if ((d->A & 0x8000) == 0) {
        if (d->A < qe) { dec = 1 - mps; d->ctx[ctx] = orc_mq_nlps[s]; }
        else           { dec = mps;     d->ctx[ctx] = orc_mq_nmps[s]; }
        dec_renorm(d);
        return dec;
    }
    return mps;
}

/* ---- flat test entry point ---------------------------------------------------- */
void orc_mq_decode(const uint8_t *data, int len, const uint8_t *ctxs, int n, uint8_t *bits_out)
{
    orc_mqdec d;
    orc_mqdec_init(&d, data, len);
    for (int i = 0; i < n; i++) bits_out[i] = (uint8_t)orc_mqdec_decode(&d, ctxs[i]);
}
