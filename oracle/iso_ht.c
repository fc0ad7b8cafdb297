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
 * Output: out[y*w+x] = sign * (mu << (num_bps-1)) where mu is the decoded magnitude index at the
 * cleanup bit-plane (num_bps = Mb - missing_msbs, so num_bps = 1 for a lossless cleanup-only block).
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

int iso_ht_decode(const uint8_t *data, int len, int w, int h, int num_bps, int32_t *out)
{
    memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    if (len < 2 || num_bps < 1 || num_bps > 30) return -1;
    const int lcup = len;
    const int scup = ((int)data[lcup - 1] << 4) + (data[lcup - 2] & 0x0F);
    if (scup < 2 || scup > lcup || scup > 4079) return -2;
    const int shift = num_bps - 1;

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
                if (U[i] > 31) { rc = -3; U[i] = 31; }
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
                    const uint32_t mag = mu << shift;
                    out[yy * w + x] = (int32_t)(sign ? 0u - mag : mag);
                    if (n & 1) { NSG[x] = 1; NEX[x] = (uint8_t)bitlen32(v); }
                }
            }
        }
        uint8_t *t;
        t = sg; sg = nsg; nsg = t; t = ex; ex = nex; nex = t;
        SG = sg + 2; EX = ex + 2; NSG = nsg + 2; NEX = nex + 2;
    }
    free(sg); free(ex); free(nsg); free(nex);
    if (rc) memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    return rc;
}
