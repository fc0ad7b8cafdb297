/*
 * iso_ht.c -- ISO/IEC 15444-15 (ITU-T T.814) HT cleanup-pass block decoder, CPU checker for
 * J2KGPU_MODE_ISO.  Oracle / test infrastructure only.
 *
 * This is NOT a restatement of the reference's ht.go (which is not conformant); it is
 * written from the published algorithm (T.814 clause 7, as also implemented by OpenJPH / OpenJPEG
 * ht_dec.c): three byte streams inside the cleanup segment -- MagSgn growing forward from byte 0, MEL
 * growing forward from Lcup-Scup, VLC growing backward from Lcup-2 -- 2x2 quads scanned in pairs,
 * CxtVLC tables (ht_vlc_tables.inc), U-VLC, exponent predictor kappa from the previous quad row.
 * Pinned by: OpenJPEG 2.5.4 decodes the streams of datagen/gen_iso_ht.c to the source image
 * (tests/test_iso_codestream.py), and this decoder inverts the same streams.
 *
 * plus the SigProp and MagRef passes of the same HT set (clause 7.4 / 7.5), pinned the same way by OpenJPEG
 * decoding multi-pass streams of the extended writer.
 *
 * Reconstruction follows OpenJPEG (the pin): the cleanup pass at bit-plane P = num_bps - 1 (num_bps = Mb - missing
 * MSBs) gives magnitude index mu; a mid-point bit sits below the last decoded bit-plane.  iso_ht_decode_passes
 * returns sign * Q in QUARTER units (integer LSB = bit 2), so that the refinement of bit-plane P - 1 = -1 and its
 * mid-point are still integers: cleanup only Q = (2 mu + 1) << (P + 1); MagRef'ed Q = mu << (P + 2) | bit << (P + 1)
 * | 1 << P; newly significant in SigProp Q = 3 << P.  Reversible value = sign * (Q >> 2); irreversible =
 * sign * Q * step / 4.  iso_ht_decode = cleanup only, reversible value.
 * Returns 0, or a negative value for a malformed segment (out is then all zero).
 */
#include "oracle.h"
#include <string.h>
#include <stdlib.h>

#include "ht_vlc_tables.inc"
static const uint16_t k_tbl0[1024] = HT_VLC_TBL0_INIT;
static const uint16_t k_tbl1[1024] = HT_VLC_TBL1_INIT;
static const uint8_t k_mel_exp[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5};

/* ---- MEL: forward, MSB first, a byte after 0xFF carries 7 bits ---- */
typedef struct {
    const uint8_t *d; int pos, left;      /* left = bytes still readable (Scup-1 in total) */
    uint32_t tmp; int bits; int unstuff;
    int k, zeros; int one_after;
} mel_t;

static int mel_bit(mel_t *m)
{
    if (m->bits == 0) {
        uint32_t b = 0xFF;
        if (m->left > 0) {
            b = m->d[m->pos++];
            m->left--;
            if (m->left == 0) b |= 0x0F;   /* last byte is shared with the VLC stream */
        }
        m->bits = m->unstuff ? 7 : 8;
        m->tmp = m->unstuff ? (b & 0x7F) : b;
        m->unstuff = (b == 0xFF);
    }
    m->bits--;
    return (int)((m->tmp >> m->bits) & 1);
}

static int mel_event(mel_t *m)
{
    if (m->zeros == 0 && !m->one_after) {
        int e = k_mel_exp[m->k];
        if (mel_bit(m)) {                    /* "1": a full run of 2^e zeros */
            m->zeros = 1 << e;
            if (m->k < 12) m->k++;
        } else {                             /* "0" + e bits: that many zeros, then a one */
            int r = 0;
            for (int i = 0; i < e; i++) r = (r << 1) | mel_bit(m);
            m->zeros = r; m->one_after = 1;
            if (m->k > 0) m->k--;
        }
    }
    if (m->zeros > 0) { m->zeros--; return 0; }
    m->one_after = 0;
    return 1;
}

/* ---- VLC: backward, LSB first; a byte whose low 7 bits are all 1 after a byte > 0x8F carries 7 bits ---- */
typedef struct { const uint8_t *d; int pos, left; uint64_t tmp; int bits; int unstuff; } vlc_t;

static void vlc_fill(vlc_t *v)
{
    while (v->bits <= 32) {
        uint32_t b = 0;
        if (v->left > 0) { b = v->d[v->pos--]; v->left--; }
        int nb = (v->unstuff && (b & 0x7F) == 0x7F) ? 7 : 8;
        v->tmp |= (uint64_t)b << v->bits;
        v->bits += nb;
        v->unstuff = b > 0x8F;
    }
}
static uint32_t vlc_peek(vlc_t *v) { vlc_fill(v); return (uint32_t)v->tmp; }
static void vlc_skip(vlc_t *v, int n) { v->tmp >>= n; v->bits -= n; }

/* ---- MagSgn: forward, LSB first; a byte after 0xFF carries 7 bits; exhausted stream feeds 0xFF ---- */
typedef struct { const uint8_t *d; int pos, left; uint64_t tmp; int bits; int unstuff; } ms_t;

static void ms_fill(ms_t *s)
{
    while (s->bits <= 32) {
        uint32_t b = 0xFF;
        if (s->left > 0) { b = s->d[s->pos++]; s->left--; }
        int nb = s->unstuff ? 7 : 8;
        s->tmp |= (uint64_t)b << s->bits;
        s->bits += nb;
        s->unstuff = (b == 0xFF);
    }
}
static uint32_t ms_get(ms_t *s, int n)
{
    ms_fill(s);
    uint32_t v = (uint32_t)(s->tmp & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1)));
    s->tmp >>= n; s->bits -= n;
    return v;
}

/* U-VLC prefix table: prefix_len | suffix_len << 2 | base << 5 (T.814 Table 3) */
static const uint8_t k_uvlc[8] = {
    3 | (5 << 2) | (5 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5),
    3 | (1 << 2) | (3 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5)};

/* returns bits consumed; u[i] = u_q (NOT yet + kappa).  mode: bit0 = u_off of quad 0, bit1 = of quad 1;
 * mode 4 = initial row, both set and the MEL event was 1 (both u > 2). */
static int uvlc_decode(uint32_t vlc, int mode, int initial, int u[2])
{
    int used = 0;
    u[0] = u[1] = 0;
    if (mode == 0) return 0;
    if (mode == 1 || mode == 2) {
        uint32_t t = k_uvlc[vlc & 7];
        int pl = t & 3, sl = (t >> 2) & 7;
        vlc >>= pl;
        u[mode - 1] = (int)((t >> 5) + (vlc & ((1u << sl) - 1)));
        return pl + sl;
    }
    if (mode == 3 && initial) {
        uint32_t t1 = k_uvlc[vlc & 7];
        int p1 = t1 & 3;
        vlc >>= p1; used += p1;
        if (p1 > 2) {                                  /* u0 > 2, so u1 is 1 or 2: one bit */
            u[1] = (int)(vlc & 1) + 1;
            vlc >>= 1; used++;
            int sl = (t1 >> 2) & 7;
            u[0] = (int)((t1 >> 5) + (vlc & ((1u << sl) - 1)));
            return used + sl;
        }
        uint32_t t2 = k_uvlc[vlc & 7];
        int p2 = t2 & 3;
        vlc >>= p2; used += p2;
        int s1 = (t1 >> 2) & 7, s2 = (t2 >> 2) & 7;
        u[0] = (int)((t1 >> 5) + (vlc & ((1u << s1) - 1)));
        vlc >>= s1;
        u[1] = (int)((t2 >> 5) + (vlc & ((1u << s2) - 1)));
        return used + s1 + s2;
    }
    /* mode 3 in a non-initial row, or mode 4: two full codes, prefixes first */
    uint32_t t1 = k_uvlc[vlc & 7];
    int p1 = t1 & 3;
    vlc >>= p1;
    uint32_t t2 = k_uvlc[vlc & 7];
    int p2 = t2 & 3;
    vlc >>= p2;
    int s1 = (t1 >> 2) & 7, s2 = (t2 >> 2) & 7;
    u[0] = (int)((t1 >> 5) + (vlc & ((1u << s1) - 1)));
    vlc >>= s1;
    u[1] = (int)((t2 >> 5) + (vlc & ((1u << s2) - 1)));
    if (mode == 4) { u[0] += 2; u[1] += 2; }
    return p1 + p2 + s1 + s2;
}

static inline int bitlen32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* cleanup pass: mu (>= 1 for every significant sample, 0 elsewhere) and sign per sample */
static int ht_cleanup(const uint8_t *data, int len, int w, int h, int P, uint32_t *out, uint8_t *sgn)
{
    memset(out, 0, sizeof(uint32_t) * (size_t)w * h);
    memset(sgn, 0, (size_t)w * h);
    if (len < 2) return -1;
    const int lcup = len;
    const int scup = ((int)data[lcup - 1] << 4) + (data[lcup - 2] & 0x0F);
    if (scup < 2 || scup > lcup || scup > 4079) return -2;

    mel_t mel; memset(&mel, 0, sizeof mel);
    mel.d = data; mel.pos = lcup - scup; mel.left = scup - 1;
    vlc_t vlc; memset(&vlc, 0, sizeof vlc);
    {
        uint32_t b = data[lcup - 2];
        vlc.d = data; vlc.pos = lcup - 3; vlc.left = scup - 2;
        vlc.tmp = b >> 4;
        vlc.bits = 4 - (((vlc.tmp & 7) == 7) ? 1 : 0);
        vlc.unstuff = (b | 0x0F) > 0x8F;
    }
    ms_t ms; memset(&ms, 0, sizeof ms);
    ms.d = data; ms.pos = 0; ms.left = lcup - scup;

    /* per column of the previous quad row's bottom sample row: significance and exponent; 2 columns of
     * zero padding on each side */
    const int nq = (w + 1) / 2;
    uint8_t *sg = calloc((size_t)2 * nq + 8, 1), *ex = calloc((size_t)2 * nq + 8, 1);
    uint8_t *nsg = calloc((size_t)2 * nq + 8, 1), *nex = calloc((size_t)2 * nq + 8, 1);
    uint8_t *SG = sg + 2, *EX = ex + 2, *NSG = nsg + 2, *NEX = nex + 2;
    int rc = 0;

    for (int y = 0; y < h && rc == 0; y += 2) {
        const int initial = (y == 0);
        const uint16_t *tbl = initial ? k_tbl0 : k_tbl1;
        int cw = 0;                                        /* context carried from the left quad */
        memset(nsg, 0, (size_t)2 * nq + 8); memset(nex, 0, (size_t)2 * nq + 8);
        for (int q = 0; q < nq; q += 2) {
            uint32_t qinf[2] = {0, 0};
            int U[2] = {0, 0};
            const int npair = (q + 1 < nq) ? 2 : 1;
            for (int i = 0; i < npair; i++) {
                const int qq = q + i;
                int c_q;
                if (initial) c_q = cw;
                else c_q = cw | (SG[2 * qq - 1] | SG[2 * qq]) | ((SG[2 * qq + 1] | SG[2 * qq + 2]) << 2);
                uint32_t e = tbl[(c_q << 7) | (vlc_peek(&vlc) & 0x7F)];
                if (c_q == 0 && !mel_event(&mel)) e = 0;   /* insignificant quad: no codeword was sent */
                vlc_skip(&vlc, e & 7);
                qinf[i] = e;
                const uint32_t rho = (e >> 4) & 0xF;
                if (initial) cw = (int)(((rho & 1) | ((rho >> 1) & 1)) | (((rho >> 2) & 1) << 1) | (((rho >> 3) & 1) << 2));
                else cw = (int)((((rho >> 2) & 1) | ((rho >> 3) & 1)) << 1);
            }
            int mode = (int)(((qinf[0] >> 3) & 1) | (((qinf[1] >> 3) & 1) << 1));
            if (initial && mode == 3 && mel_event(&mel)) mode = 4;
            int u[2];
            vlc_skip(&vlc, uvlc_decode(vlc_peek(&vlc), mode, initial, u));
            for (int i = 0; i < npair; i++) {
                const int qq = q + i;
                const uint32_t rho = (qinf[i] >> 4) & 0xF;
                int kappa = 1;
                if (!initial && (rho & (rho - 1))) {        /* gamma_q: more than one significant sample */
                    int E = EX[2 * qq - 1];
                    if (EX[2 * qq] > E) E = EX[2 * qq];
                    if (EX[2 * qq + 1] > E) E = EX[2 * qq + 1];
                    if (EX[2 * qq + 2] > E) E = EX[2 * qq + 2];
                    if (E - 1 > kappa) kappa = E - 1;
                }
                U[i] = u[i] + kappa;
                /* every magnitude must fit 31 bits in quarter units: (2 mu + 1) << (P + 1) with mu <= 2^U */
                if (U[i] + P > 28) rc = -3;
                if (U[i] > 31) U[i] = 31;
            }
            for (int i = 0; i < npair; i++) {
                const int qq = q + i;
                const uint32_t e = qinf[i], rho = (e >> 4) & 0xF;
                for (int n = 0; n < 4; n++) {
                    if (!((rho >> n) & 1)) continue;
                    const int x = 2 * qq + (n >> 1), yy = y + (n & 1);
                    if (x >= w || yy >= h) { rc = -4; continue; }      /* significance outside the block */
                    const int m = U[i] - (int)((e >> (12 + n)) & 1);
                    uint32_t v = ms_get(&ms, m);
                    const uint32_t sign = v & 1;
                    v |= ((e >> (8 + n)) & 1) << m;
                    v |= 1;
                    const uint32_t mu = (v >> 1) + 1;
                    out[yy * w + x] = mu;
                    sgn[yy * w + x] = (uint8_t)sign;
                    if (n & 1) { NSG[x] = 1; NEX[x] = (uint8_t)bitlen32(v); }
                }
            }
        }
        uint8_t *t;
        t = sg; sg = nsg; nsg = t; t = ex; ex = nex; nex = t;
        SG = sg + 2; EX = ex + 2; NSG = nsg + 2; NEX = nex + 2;
    }
    free(sg); free(ex); free(nsg); free(nex);
    if (rc) memset(out, 0, sizeof(uint32_t) * (size_t)w * h);
    return rc;
}

/* forward bit reader of the SigProp stream: LSB first, a byte after 0xFF carries 7 bits, zeros when exhausted */
typedef struct { const uint8_t *d; int pos, left; uint32_t tmp; int bits; int unstuff; } spp_t;
static int spp_bit(spp_t *s)
{
    if (s->bits == 0) {
        uint32_t b = 0;
        if (s->left > 0) { b = s->d[s->pos++]; s->left--; }
        s->bits = s->unstuff ? 7 : 8;
        s->tmp = b;
        s->unstuff = (b == 0xFF);
    }
    const int v = (int)(s->tmp & 1);
    s->tmp >>= 1; s->bits--;
    return v;
}
/* backward bit reader of the MagRef stream: LSB first, from the last byte of the refinement segment; a byte whose low
 * 7 bits are all ones after a byte > 0x8F carries 7 bits (the byte after the segment counts as > 0x8F); then zeros */
typedef struct { const uint8_t *d; int pos, left; uint32_t tmp; int bits; int unstuff; } mrp_t;
static int mrp_bit(mrp_t *s)
{
    if (s->bits == 0) {
        uint32_t b = 0;
        if (s->left > 0) { b = s->d[s->pos--]; s->left--; }
        s->bits = (s->unstuff && (b & 0x7F) == 0x7F) ? 7 : 8;
        s->tmp = b;
        s->unstuff = b > 0x8F;
    }
    const int v = (int)(s->tmp & 1);
    s->tmp >>= 1; s->bits--;
    return v;
}

int iso_ht_decode_passes(const uint8_t *data, int lcup, int lref, int w, int h, int num_bps, int num_passes, int32_t *out)
{
    const int n = w * h;
    memset(out, 0, sizeof(int32_t) * (size_t)n);
    if (num_bps < 1 || num_bps > 30 || num_passes < 1 || num_passes > 3 || lref < 0) return -1;
    uint32_t *mu = malloc(sizeof(uint32_t) * (size_t)n);
    uint8_t *sgn = malloc((size_t)n), *snew = calloc((size_t)n, 1);
    const int P = num_bps - 1;
    const int rc = ht_cleanup(data, lcup, w, h, P, mu, sgn);
    if (rc) { free(mu); free(sgn); free(snew); return rc; }
    if (num_passes > 1 && lref == 0) num_passes = 1;             /* no refinement bytes: cleanup only (as OpenJPEG) */
    if (num_passes >= 2) {
        spp_t sp = {data + lcup, 0, lref, 0, 0, 0};
        for (int y0 = 0; y0 < h; y0 += 4)
            for (int x0 = 0; x0 < w; x0 += 4) {
                int newi[16], nnew = 0;
                for (int x = x0; x < x0 + 4 && x < w; x++)
                    for (int y = y0; y < y0 + 4 && y < h; y++) {
                        const int i = y * w + x;
                        if (mu[i]) continue;
                        int mbr = 0;
                        for (int dy = -1; dy <= 1; dy++)
                            for (int dx = -1; dx <= 1; dx++) {
                                const int yy = y + dy, xx = x + dx;
                                if ((dy || dx) && yy >= 0 && yy < h && xx >= 0 && xx < w && (mu[yy * w + xx] || snew[yy * w + xx])) mbr = 1;
                            }
                        if (mbr && spp_bit(&sp)) { snew[i] = 1; newi[nnew++] = i; }
                    }
                for (int k = 0; k < nnew; k++) snew[newi[k]] = (uint8_t)(1 + 2 * spp_bit(&sp));   /* 1 positive, 3 negative */
            }
    }
    /* all arithmetic is uint32 (wraps on absurd inputs exactly like the device code) */
    mrp_t mr = {data + lcup, lref - 1, lref, 0, 0, 1};
    for (int y0 = 0; y0 < h; y0 += 4)
        for (int x = 0; x < w; x++)
            for (int y = y0; y < y0 + 4 && y < h; y++) {
                const int i = y * w + x;
                uint32_t q = 0, neg = 0;
                if (mu[i]) {
                    if (num_passes == 3) q = (mu[i] << (P + 2)) | ((uint32_t)mrp_bit(&mr) << (P + 1)) | (1u << P);
                    else q = (2u * mu[i] + 1u) << (P + 1);
                    neg = sgn[i];
                } else if (snew[i]) {
                    q = 3u << P;
                    neg = snew[i] & 2;
                }
                out[i] = (int32_t)(neg ? 0u - q : q);
            }
    free(mu); free(sgn); free(snew);
    return 0;
}

/* cleanup only, reversible value sign * (Q >> 2) (uint32 arithmetic) */
int iso_ht_decode(const uint8_t *data, int len, int w, int h, int num_bps, int32_t *out)
{
    const int rc = iso_ht_decode_passes(data, len, 0, w, h, num_bps, 1, out);
    for (int i = 0; i < w * h; i++) {
        const uint32_t u = (uint32_t)out[i];
        const uint32_t q = (out[i] < 0) ? 0u - u : u;
        out[i] = (int32_t)((out[i] < 0) ? 0u - (q >> 2) : (q >> 2));
    }
    return rc;
}
