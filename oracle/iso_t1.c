/* iso_t1.c -- TEST INFRASTRUCTURE (CPU checker, never on the product path).
 *
 * ISO/IEC 15444-1 Annex C + D code-block decoder (MQ arithmetic decoder, significance propagation / magnitude
 * refinement / cleanup passes in stripe order, standard context tables and initial states, pass-count truncation)
 * for J2KGPU_MODE_ISO.  The reference's entropy/t1.go is not conformant (raster-order passes, its own
 * context tables, all contexts starting in state 0) and is restated in orc_t1.c; this file follows the published
 * algorithm instead and is pinned by OpenJPEG: tests/test_iso_codestream.py decodes codestreams written by OpenJPEG
 * (through Pillow) with this decoder + the ISO inverse transform and compares with OpenJPEG's own output.
 * Code-block styles (COD SPcod, Table A.19): RESET (0x02: contexts back to Table D.7 at every pass boundary), VCAUSAL
 * (0x08: the row below a stripe counts as insignificant, D.4.3... D.7) and SEGSYM (0x20: four UNIFORM symbols after each
 * cleanup pass, D.5) are decoded; PREDTERM (0x10) needs nothing from a decoder; BYPASS (0x01, D.6: from the fifth
 * bit-plane on the significance and refinement passes are raw bits) and TERMALL (0x04: every pass its own terminated
 * codeword) split the block into several codeword segments: their byte counts follow the block's data_len bytes as
 * little-endian 32-bit words (seglens; the job-table convention of include/j2kgpu.h).  Pinned by streams OpenJPEG wrote
 * with those styles (datagen/opj_direct.py) and decodes itself.
 *
 * Output convention (what the CUDA kernel reproduces): out[y*w+x] = sign * m2, where m2 is the magnitude at TWICE
 * scale with the mid-point of the last decoded bit-plane added: m2 = 2 * (decoded magnitude bits) + (1 << p_last),
 * p_last = lowest bit-plane at which the sample was coded.  Reversible reconstruction = m2 / 2 (truncating);
 * irreversible = m2 * 0.5 * step.  A fully decoded block has p_last = 0 for every significant sample, so m2 / 2 is the
 * exact integer.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* Table C.2: Qe, NMPS, NLPS, SWITCH */
static const uint16_t QE[47] = {
    0x5601, 0x3401, 0x1801, 0x0AC1, 0x0521, 0x0221, 0x5601, 0x5401, 0x4801, 0x3801, 0x3001, 0x2401,
    0x1C01, 0x1601, 0x5601, 0x5401, 0x5101, 0x4801, 0x3801, 0x3401, 0x3001, 0x2801, 0x2401, 0x2201,
    0x1C01, 0x1801, 0x1601, 0x1401, 0x1201, 0x1101, 0x0AC1, 0x09C1, 0x08A1, 0x0521, 0x0441, 0x02A1,
    0x0221, 0x0141, 0x0111, 0x0085, 0x0049, 0x0025, 0x0015, 0x0009, 0x0005, 0x0001, 0x5601};
static const uint8_t NMPS[47] = {1, 2, 3, 4, 5, 38, 7, 8, 9, 10, 11, 12, 13, 29, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
                                 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 45, 46};
static const uint8_t NLPS[47] = {1, 6, 9, 12, 29, 33, 6, 14, 14, 14, 17, 18, 20, 21, 14, 14, 15, 16, 17, 18, 19, 19, 20, 21,
                                 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 46};
static const uint8_t SW[47] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1};

typedef struct {
    const uint8_t *d; int len, bp;
    uint32_t A, C; int CT;
    uint8_t idx[19], mps[19];
} mq_t;

static uint32_t mq_byte(const mq_t *m, int p) { return p < m->len ? m->d[p] : 0xFFu; }   /* past the end: 0xFF fill */

static void mq_bytein(mq_t *m)
{
    if (mq_byte(m, m->bp) == 0xFF) {
        if (mq_byte(m, m->bp + 1) > 0x8F) { m->C += 0xFF00; m->CT = 8; }
        else { m->bp++; m->C += mq_byte(m, m->bp) << 9; m->CT = 7; }
    } else { m->bp++; m->C += mq_byte(m, m->bp) << 8; m->CT = 8; }
}

static void mq_reset_contexts(mq_t *m)                       /* Table D.7: ZC context 0, run-length, uniform */
{
    memset(m->idx, 0, sizeof m->idx); memset(m->mps, 0, sizeof m->mps);
    m->idx[0] = 4; m->idx[17] = 3; m->idx[18] = 46;
}

static void mq_init(mq_t *m, const uint8_t *d, int len)
{
    memset(m, 0, sizeof *m);
    m->d = d; m->len = len; m->bp = 0;
    mq_reset_contexts(m);
    m->C = mq_byte(m, 0) << 16;
    mq_bytein(m);
    m->C <<= 7; m->CT -= 7; m->A = 0x8000;
}

/* raw (bypass) bit, D.6: bytes are read MSB first, the byte after an 0xFF carries 7 bits; past the end: 0xFF */
static int raw_decode(mq_t *m)
{
    if (m->CT == 0) {
        if (m->C == 0xFF) {
            if (mq_byte(m, m->bp) > 0x8F) { m->C = 0xFF; m->CT = 8; }
            else { m->C = mq_byte(m, m->bp); m->bp++; m->CT = 7; }
        } else { m->C = mq_byte(m, m->bp); m->bp++; m->CT = 8; }
    }
    m->CT--;
    return (int)((m->C >> m->CT) & 1);
}

/* a new codeword segment: the contexts survive (unless RESET), the arithmetic / raw state starts over */
static void seg_start(mq_t *m, const uint8_t *d, int len, int raw)
{
    m->d = d; m->len = len; m->bp = 0;
    if (raw) { m->C = 0; m->CT = 0; return; }
    m->C = mq_byte(m, 0) << 16;
    m->CT = 0;
    mq_bytein(m);
    m->C <<= 7; m->CT -= 7; m->A = 0x8000;
}

static int mq_decode(mq_t *m, int cx)
{
    const uint32_t qe = QE[m->idx[cx]];
    int d;
    m->A -= qe;
    if ((m->C >> 16) < qe) {                                 /* LPS exchange (C.3.2) */
        if (m->A < qe) { d = m->mps[cx]; m->idx[cx] = NMPS[m->idx[cx]]; }
        else { d = 1 - m->mps[cx]; if (SW[m->idx[cx]]) m->mps[cx] ^= 1; m->idx[cx] = NLPS[m->idx[cx]]; }
        m->A = qe;
    } else {
        m->C -= qe << 16;
        if (m->A & 0x8000) return m->mps[cx];
        if (m->A < qe) { d = 1 - m->mps[cx]; if (SW[m->idx[cx]]) m->mps[cx] ^= 1; m->idx[cx] = NLPS[m->idx[cx]]; }
        else { d = m->mps[cx]; m->idx[cx] = NMPS[m->idx[cx]]; }
    }
    do {
        if (m->CT == 0) mq_bytein(m);
        m->A <<= 1; m->C <<= 1; m->CT--;
    } while (!(m->A & 0x8000));
    return d;
}

/* Table D.1: zero-coding context from the counts of significant horizontal, vertical, diagonal neighbours */
static int zc_ctx(int band, int h, int v, int d)
{
    if (band == 3) {                                          /* HH */
        const int hv = h + v;
        if (d >= 3) return 8;
        if (d == 2) return hv >= 1 ? 7 : 6;
        if (d == 1) return hv >= 2 ? 5 : (hv == 1 ? 4 : 3);
        return hv >= 2 ? 2 : (hv == 1 ? 1 : 0);
    }
    if (band == 1) { const int t = h; h = v; v = t; }          /* HL: roles of h and v swapped */
    if (h == 2) return 8;
    if (h == 1) return v >= 1 ? 7 : (d >= 1 ? 6 : 5);
    if (v == 2) return 4;
    if (v == 1) return 3;
    return d >= 2 ? 2 : (d == 1 ? 1 : 0);
}

typedef struct {
    int w, h, sw;                 /* sw = w + 2 */
    int raw;                      /* the current pass reads raw bits (selective bypass) */
    uint8_t *sig, *neg, *pi, *ref;
    int vsc;                      /* vertically causal context formation */
    int32_t *mag;                 /* magnitude bits decoded so far */
    int8_t *plast;                /* lowest bit-plane at which the sample was coded */
} blk_t;

#define AT(a, x, y) ((a)[((y) + 1) * b->sw + (x) + 1])
/* the row below sample (x, y): with the vertically causal style a stripe never looks into the next one */
#define BELOW(a, xx, y) ((b->vsc && ((y) & 3) == 3) ? 0 : AT(a, xx, (y) + 1))

static int sign_decode(mq_t *m, const blk_t *b, int x, int y)      /* Table D.2 / D.3 */
{
    if (b->raw) return raw_decode(m);                               /* D.6: the sign bit itself, no prediction */
    int hc = 0, vc = 0;
    if (AT(b->sig, x - 1, y)) hc += AT(b->neg, x - 1, y) ? -1 : 1;
    if (AT(b->sig, x + 1, y)) hc += AT(b->neg, x + 1, y) ? -1 : 1;
    if (AT(b->sig, x, y - 1)) vc += AT(b->neg, x, y - 1) ? -1 : 1;
    if (BELOW(b->sig, x, y)) vc += AT(b->neg, x, y + 1) ? -1 : 1;
    hc = hc > 1 ? 1 : (hc < -1 ? -1 : hc);
    vc = vc > 1 ? 1 : (vc < -1 ? -1 : vc);
    int flip = 0;
    if (hc < 0 || (hc == 0 && vc < 0)) { flip = 1; hc = -hc; vc = -vc; }
    int cx;
    if (hc == 1) cx = vc == 1 ? 13 : (vc == 0 ? 12 : 11);
    else cx = vc == 1 ? 10 : 9;                                 /* hc == 0: vc is 1 or 0 after the flip */
    return mq_decode(m, cx) ^ flip;
}

static int zc_of(const blk_t *b, int band, int x, int y)
{
    const int h = AT(b->sig, x - 1, y) + AT(b->sig, x + 1, y);
    const int v = AT(b->sig, x, y - 1) + BELOW(b->sig, x, y);
    const int d = AT(b->sig, x - 1, y - 1) + AT(b->sig, x + 1, y - 1) + BELOW(b->sig, x - 1, y) + BELOW(b->sig, x + 1, y);
    return zc_ctx(band, h, v, d);
}

static int any_neighbour(const blk_t *b, int x, int y)
{
    return AT(b->sig, x - 1, y) | AT(b->sig, x + 1, y) | AT(b->sig, x, y - 1) | BELOW(b->sig, x, y) |
           AT(b->sig, x - 1, y - 1) | AT(b->sig, x + 1, y - 1) | BELOW(b->sig, x - 1, y) | BELOW(b->sig, x + 1, y);
}

static void become_sig(mq_t *m, blk_t *b, int x, int y, int bp)
{
    const int s = sign_decode(m, b, x, y);
    AT(b->sig, x, y) = 1; AT(b->neg, x, y) = (uint8_t)s;
    b->mag[y * b->w + x] |= 1 << bp;
    b->plast[y * b->w + x] = (int8_t)bp;
}

/* codeword segments that num_passes coding passes touch (1 unless the style has BYPASS or TERMALL) */
int iso_t1_num_segments(int style, int num_passes)
{
    if (num_passes <= 0 || !(style & 0x05)) return 1;
    const int i = num_passes - 1;
    if (style & 0x04) return i + 1;
    if (i < 10) return 1;
    return 1 + 2 * ((i - 10) / 3) + ((i - 10) % 3 == 2 ? 1 : 0) + 1;
}

int iso_t1_decode(const uint8_t *data, int len, int w, int h, int num_bps, int num_passes, int band, int32_t *out)
{
    return iso_t1_decode_style(data, len, w, h, num_bps, num_passes, band, 0, out);
}

int iso_t1_decode_style(const uint8_t *data, int len, int w, int h, int num_bps, int num_passes, int band, int style, int32_t *out)
{

    memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    if (w < 1 || h < 1 || w > 1024 || h > 1024 || num_bps < 0 || num_bps > 30) return -1;
    if (num_bps == 0 || num_passes <= 0) return 0;
    const int max_passes = 3 * num_bps - 2;
    if (num_passes > max_passes) num_passes = max_passes;
    blk_t blk, *b = &blk;
    b->w = w; b->h = h; b->sw = w + 2; b->vsc = (style & 0x08) != 0;
    const size_t fl = (size_t)(w + 2) * (h + 2);
    b->sig = calloc(4 * fl, 1); b->neg = b->sig + fl; b->pi = b->neg + fl; b->ref = b->pi + fl;
    b->mag = calloc((size_t)w * h, sizeof(int32_t));
    b->plast = calloc((size_t)w * h, 1);
    mq_t mq;
    mq_init(&mq, data, len);
    int bp = num_bps - 1, type = 2;                              /* the first pass is a cleanup pass */
    const int segmented = style & (0x01 | 0x04);
    const uint8_t *seglens = data + len;                         /* segmented styles: the table behind the code bytes */
    int seg = 0, seg_pos = 0, seg_left = 0;
    b->raw = 0;
    for (int pass = 0; pass < num_passes; pass++) {
        if (pass && (style & 0x02)) mq_reset_contexts(&mq);
        if (segmented) {
            const int raw = (style & 0x01) && pass >= 10 && type != 2;
            if (seg_left == 0) {
                const int sl = (int)((uint32_t)seglens[4 * seg] | (uint32_t)seglens[4 * seg + 1] << 8 | (uint32_t)seglens[4 * seg + 2] << 16 |
                                     (uint32_t)seglens[4 * seg + 3] << 24);
                const int avail = sl < 0 || sl > len - seg_pos ? len - seg_pos : sl;
                seg_start(&mq, data + seg_pos, avail, raw);
                seg_pos += avail;
                seg_left = (style & 0x04) ? 1 : (seg == 0 ? 10 : ((seg & 1) ? 2 : 1));
                seg++;
            }
            seg_left--;
            b->raw = raw;
        }
        for (int y0 = 0; y0 < h; y0 += 4)
            for (int x = 0; x < w; x++) {
                const int rows = y0 + 4 <= h ? 4 : h - y0;
                if (type == 0) {                                                       /* significance propagation, D.3.1 */
                    for (int k = 0; k < rows; k++) {
                        const int y = y0 + k;
                        if (AT(b->sig, x, y) || !any_neighbour(b, x, y)) continue;
                        if (b->raw ? raw_decode(&mq) : mq_decode(&mq, zc_of(b, band, x, y))) become_sig(&mq, b, x, y, bp);
                        AT(b->pi, x, y) = 1;
                    }
                } else if (type == 1) {                                                /* magnitude refinement, D.3.3 */
                    for (int k = 0; k < rows; k++) {
                        const int y = y0 + k;
                        if (!AT(b->sig, x, y) || AT(b->pi, x, y)) continue;
                        const int cx = AT(b->ref, x, y) ? 16 : (any_neighbour(b, x, y) ? 15 : 14);
                        if (b->raw ? raw_decode(&mq) : mq_decode(&mq, cx)) b->mag[y * w + x] |= 1 << bp;
                        AT(b->ref, x, y) = 1;
                        b->plast[y * w + x] = (int8_t)bp;
                    }
                } else {                                                               /* cleanup, D.3.4 */
                    int k = 0;
                    if (rows == 4) {
                        int quiet = 1;
                        for (int j = 0; j < 4; j++)
                            if (AT(b->sig, x, y0 + j) || AT(b->pi, x, y0 + j) || any_neighbour(b, x, y0 + j)) quiet = 0;
                        if (quiet) {
                            if (!mq_decode(&mq, 17)) continue;                          /* run of four zeros */
                            k = mq_decode(&mq, 18) << 1;
                            k |= mq_decode(&mq, 18);
                            become_sig(&mq, b, x, y0 + k, bp);
                            k++;
                        }
                    }
                    for (; k < rows; k++) {
                        const int y = y0 + k;
                        if (AT(b->sig, x, y) || AT(b->pi, x, y)) continue;
                        if (mq_decode(&mq, zc_of(b, band, x, y))) become_sig(&mq, b, x, y, bp);
                    }
                }
            }
        if (type == 2 && (style & 0x20))                                               /* segmentation symbol 1010 (D.5) */
            for (int i = 0; i < 4; i++) mq_decode(&mq, 18);
        if (type == 2) {                                                               /* end of the bit-plane */
            memset(b->pi, 0, fl);
            bp--;
        }
        type = (type + 1) % 3;
    }
    for (int i = 0; i < w * h; i++) {
        const int x = i % w, y = i / w;
        if (!AT(b->sig, x, y)) continue;
        const int32_t m2 = (int32_t)(((uint32_t)b->mag[i] << 1) | (1u << b->plast[i]));
        out[i] = AT(b->neg, x, y) ? -m2 : m2;
    }
    free(b->sig); free(b->mag); free(b->plast);
    return 0;
}
