/*
 * gen_iso_ht.c -- ISO/IEC 15444-15 (ITU-T T.814) HT block ENCODER, cleanup pass only, all magnitude bits in
 * one HT set (the lossless / single-pass form every HTJ2K encoder emits by default).  Part of datagen/: it
 * manufactures conformant HT code-block bitstreams for ISO-mode tests and bench inputs, because OpenJPEG
 * (the ISO cross-check available in this image) decodes HTJ2K but cannot encode it.  Validated by letting
 * OpenJPEG 2.5.4 decode whole codestreams built from its output (tests/test_iso_codestream.py).
 *
 * Written from the published algorithm (T.814 clause 7 / annex C; structure as in OpenJPH's block encoder):
 * MagSgn stream forward with 0xFF bit-stuffing, MEL adaptive run-length stream forward, VLC stream backward
 * with its 0x8F/0x7F stuffing rule, CxtVLC encode tables derived by inverting the decode tables
 * (ht_vlc_tables.inc), U-VLC exponent bounds with the kappa predictor from the previous quad row.
 * MEL and VLC tails are not fused (allowed: fusing is an optional saving).
 */
#include "datagen.h"
#include <stdlib.h>
#include <string.h>

#include "ht_vlc_tables.inc"
static const uint16_t k_tbl0[1024] = HT_VLC_TBL0_INIT;
static const uint16_t k_tbl1[1024] = HT_VLC_TBL1_INIT;
static const uint8_t k_mel_exp[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5};

/* encode tables: [table][c_q][rho][emb] -> cwd << 8 | len << 4 | e_k ; 0 = no entry */
static uint16_t g_enc[2][8][16][16];
static int g_enc_ready;

static void build_enc_tables(void)
{
    if (g_enc_ready) return;
    for (int t = 0; t < 2; t++) {
        const uint16_t *tbl = t ? k_tbl1 : k_tbl0;
        for (int c = 0; c < 8; c++)
            for (int rho = 0; rho < 16; rho++)
                for (int emb = 0; emb < 16; emb++) {
                    g_enc[t][c][rho][emb] = 0;
                    if ((emb & rho) != emb || (rho == 0 && c == 0)) continue;
                    int best = -1, best_pop = -1, best_cwd = 0, best_len = 0, best_ek = 0;
                    for (int pat = 0; pat < 128; pat++) {
                        const uint16_t e = tbl[(c << 7) | pat];
                        const int len = e & 7;
                        if (!len || (pat >> len)) continue;            /* visit each codeword once (pat == cwd) */
                        if (((e >> 4) & 15) != rho) continue;
                        const int uoff = (e >> 3) & 1, ek = (e >> 12) & 15, e1 = (e >> 8) & 15;
                        if (emb) {
                            if (!uoff || (emb & ek) != e1) continue;
                            const int pop = __builtin_popcount((unsigned)ek);
                            if (pop >= best_pop) { best = pat; best_pop = pop; best_cwd = pat; best_len = len; best_ek = ek; }
                        } else {
                            if (uoff) continue;
                            if (best < 0) { best = pat; best_cwd = pat; best_len = len; best_ek = ek; }
                        }
                    }
                    if (best >= 0) g_enc[t][c][rho][emb] = (uint16_t)((best_cwd << 8) | (best_len << 4) | best_ek);
                }
    }
    g_enc_ready = 1;
}

/* ---- MagSgn writer: forward, LSB first, 7 bits after a 0xFF byte ---- */
typedef struct { uint8_t *buf; int cap, pos; uint32_t tmp; int used, maxb; int ovf; } msw_t;

static void ms_put(msw_t *s, uint32_t v, int n)
{
    while (n > 0) {
        int t = s->maxb - s->used;
        if (t > n) t = n;
        s->tmp |= (v & ((1u << t) - 1)) << s->used;
        s->used += t; v >>= t; n -= t;
        if (s->used >= s->maxb) {
            if (s->pos >= s->cap) { s->ovf = 1; return; }
            s->buf[s->pos++] = (uint8_t)s->tmp;
            s->maxb = (s->tmp == 0xFF) ? 7 : 8;
            s->tmp = 0; s->used = 0;
        }
    }
}
static void ms_finish(msw_t *s)
{
    if (s->used) {
        int t = s->maxb - s->used;                       /* pad with ones */
        s->tmp |= (0xFFu & ((1u << t) - 1)) << s->used;
        s->used += t;
        if (s->tmp != 0xFF) {
            if (s->pos >= s->cap) { s->ovf = 1; return; }
            s->buf[s->pos++] = (uint8_t)s->tmp;
        }
    } else if (s->maxb == 7) {
        s->pos--;                                        /* a trailing 0xFF is implied by the decoder */
    }
}

/* ---- MEL writer: forward, MSB first, 7 bits after a 0xFF byte ---- */
typedef struct { uint8_t *buf; int cap, pos; uint32_t tmp; int rem; int run, k, thr; int ovf; } melw_t;

static void mel_bit(melw_t *m, int b)
{
    m->tmp = (m->tmp << 1) | (uint32_t)b;
    if (--m->rem == 0) {
        if (m->pos >= m->cap) { m->ovf = 1; return; }
        m->buf[m->pos++] = (uint8_t)m->tmp;
        m->rem = (m->tmp == 0xFF) ? 7 : 8;
        m->tmp = 0;
    }
}
static void mel_encode(melw_t *m, int ev)
{
    if (!ev) {
        if (++m->run >= m->thr) {
            mel_bit(m, 1);
            m->run = 0;
            if (m->k < 12) m->k++;
            m->thr = 1 << k_mel_exp[m->k];
        }
    } else {
        mel_bit(m, 0);
        for (int t = k_mel_exp[m->k]; t > 0;) mel_bit(m, (m->run >> --t) & 1);
        m->run = 0;
        if (m->k > 0) m->k--;
        m->thr = 1 << k_mel_exp[m->k];
    }
}

/* ---- VLC writer: backward, LSB first ---- */
typedef struct { uint8_t *end; int cap, pos; uint32_t tmp; int used; int last8f; int ovf; } vlcw_t;

static void vlc_put(vlcw_t *v, uint32_t cwd, int n)
{
    while (n > 0) {
        int avail = 8 - v->last8f - v->used;
        int t = avail < n ? avail : n;
        v->tmp |= (cwd & ((1u << t) - 1)) << v->used;
        v->used += t; avail -= t; n -= t; cwd >>= t;
        if (avail == 0) {
            if (v->last8f && v->tmp != 0x7F) { v->last8f = 0; continue; }   /* one more bit fits after all */
            if (v->pos >= v->cap) { v->ovf = 1; return; }
            *(v->end - v->pos) = (uint8_t)v->tmp;
            v->pos++;
            v->last8f = v->tmp > 0x8F;
            v->tmp = 0; v->used = 0;
        }
    }
}

static void uvlc_prefix(int u, uint32_t *cw, int *len)
{
    if (u == 1) { *cw = 1; *len = 1; }
    else if (u == 2) { *cw = 2; *len = 2; }
    else if (u <= 4) { *cw = 4; *len = 3; }
    else { *cw = 0; *len = 3; }
}
static void uvlc_suffix(int u, uint32_t *cw, int *len)
{
    if (u <= 2) { *cw = 0; *len = 0; }
    else if (u <= 4) { *cw = (uint32_t)(u - 3); *len = 1; }
    else { *cw = (uint32_t)(u - 5); *len = 5; }
}

static inline int bitlen32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* Returns byte count (0 for an all-zero block: it is simply not included in the packet), -1 on overflow or if
 * a magnitude does not fit (|x| must be < 2^30). */
int gen_iso_ht_encode(const int32_t *coef, int w, int h, uint8_t *out, int cap)
{
    build_enc_tables();
    int any = 0;
    for (int i = 0; i < w * h; i++) {
        if (coef[i]) any = 1;
        if (coef[i] == INT32_MIN || (coef[i] < 0 ? -coef[i] : coef[i]) >= (1 << 30)) return -1;
    }
    if (!any) return 0;
    const int seg = w * h * 5 + 4096;
    uint8_t *msb = malloc((size_t)seg), *melb = malloc((size_t)seg), *vlcb = malloc((size_t)seg);
    msw_t ms = {msb, seg, 0, 0, 0, 8, 0};
    melw_t mel = {melb, seg, 0, 0, 8, 0, 0, 1, 0};
    vlcw_t vlc = {vlcb + seg - 1, seg, 1, 0xF, 4, 1, 0};
    vlcb[seg - 1] = 0xFF;                                   /* placeholder of the last byte (Scup MSBs) */

    const int nq = (w + 1) / 2;
    uint8_t *sg = calloc((size_t)2 * nq + 8, 1), *ex = calloc((size_t)2 * nq + 8, 1);
    uint8_t *nsg = calloc((size_t)2 * nq + 8, 1), *nex = calloc((size_t)2 * nq + 8, 1);
    uint8_t *SG = sg + 2, *EX = ex + 2, *NSG = nsg + 2, *NEX = nex + 2;
    int bad = 0;

    for (int y = 0; y < h; y += 2) {
        const int initial = (y == 0);
        int cw = 0;
        memset(nsg, 0, (size_t)2 * nq + 8); memset(nex, 0, (size_t)2 * nq + 8);
        for (int q = 0; q < nq; q += 2) {
            const int npair = (q + 1 < nq) ? 2 : 1;
            uint32_t v[2][4]; int e[2][4]; int rho[2] = {0, 0}, U[2] = {0, 0}, u[2] = {0, 0}, ek[2] = {0, 0};
            for (int i = 0; i < npair; i++) {
                const int qq = q + i;
                int emax = 0;
                for (int n = 0; n < 4; n++) {
                    const int x = 2 * qq + (n >> 1), yy = y + (n & 1);
                    v[i][n] = 0; e[i][n] = 0;
                    if (x >= w || yy >= h) continue;
                    const int32_t c = coef[yy * w + x];
                    if (!c) continue;
                    const uint32_t mu = (uint32_t)(c < 0 ? -c : c);
                    rho[i] |= 1 << n;
                    v[i][n] = 2 * (mu - 1) + (c < 0 ? 1u : 0u);
                    e[i][n] = bitlen32(2 * mu - 1);
                    if (e[i][n] > emax) emax = e[i][n];
                    if (n & 1) { NSG[x] = 1; NEX[x] = (uint8_t)e[i][n]; }
                }
                int c_q;
                if (initial) c_q = cw;
                else c_q = cw | (SG[2 * qq - 1] | SG[2 * qq]) | ((SG[2 * qq + 1] | SG[2 * qq + 2]) << 2);
                int kappa = 1;
                if (!initial && (rho[i] & (rho[i] - 1))) {
                    int E = EX[2 * qq - 1];
                    if (EX[2 * qq] > E) E = EX[2 * qq];
                    if (EX[2 * qq + 1] > E) E = EX[2 * qq + 1];
                    if (EX[2 * qq + 2] > E) E = EX[2 * qq + 2];
                    if (E - 1 > kappa) kappa = E - 1;
                }
                U[i] = emax > kappa ? emax : kappa;
                u[i] = U[i] - kappa;
                int emb = 0;
                if (u[i] > 0)
                    for (int n = 0; n < 4; n++)
                        if (((rho[i] >> n) & 1) && e[i][n] == emax) emb |= 1 << n;
                if (c_q == 0) mel_encode(&mel, rho[i] != 0);
                if (rho[i] != 0 || c_q != 0) {
                    const uint16_t t = g_enc[initial ? 0 : 1][c_q][rho[i]][emb];
                    if (!t) bad = 1;
                    vlc_put(&vlc, t >> 8, (t >> 4) & 7);
                    ek[i] = t & 15;
                }
                const int r = rho[i];
                if (initial) cw = ((r & 1) | ((r >> 1) & 1)) | (((r >> 2) & 1) << 1) | (((r >> 3) & 1) << 2);
                else cw = (((r >> 2) & 1) | ((r >> 3) & 1)) << 1;
            }
            /* U-VLC for the pair, in the order the decoder reads it */
            const int uo0 = u[0] > 0, uo1 = u[1] > 0;
            uint32_t pc, sc; int pl, sl;
            if (uo0 && uo1) {
                int a = u[0], b = u[1];
                if (initial) {
                    const int both = (a > 2 && b > 2);
                    mel_encode(&mel, both);
                    if (both) { a -= 2; b -= 2; }
                    else if (a > 2) {                       /* b is 1 or 2: prefix(a), one bit for b, suffix(a) */
                        uvlc_prefix(a, &pc, &pl); vlc_put(&vlc, pc, pl);
                        vlc_put(&vlc, (uint32_t)(b - 1), 1);
                        uvlc_suffix(a, &sc, &sl); vlc_put(&vlc, sc, sl);
                        a = b = 0;
                    }
                }
                if (a) {
                    uvlc_prefix(a, &pc, &pl); vlc_put(&vlc, pc, pl);
                    uvlc_prefix(b, &pc, &pl); vlc_put(&vlc, pc, pl);
                    uvlc_suffix(a, &sc, &sl); vlc_put(&vlc, sc, sl);
                    uvlc_suffix(b, &sc, &sl); vlc_put(&vlc, sc, sl);
                }
            } else if (uo0 || uo1) {
                const int a = uo0 ? u[0] : u[1];
                uvlc_prefix(a, &pc, &pl); vlc_put(&vlc, pc, pl);
                uvlc_suffix(a, &sc, &sl); vlc_put(&vlc, sc, sl);
            }
            if (u[0] > 36 || u[1] > 36) bad = 1;
            for (int i = 0; i < npair; i++)
                for (int n = 0; n < 4; n++)
                    if ((rho[i] >> n) & 1) {
                        const int m = U[i] - ((ek[i] >> n) & 1);
                        ms_put(&ms, v[i][n] & ((m >= 32) ? 0xFFFFFFFFu : ((1u << m) - 1)), m);
                    }
        }
        uint8_t *t;
        t = sg; sg = nsg; nsg = t; t = ex; ex = nex; nex = t;
        SG = sg + 2; EX = ex + 2; NSG = nsg + 2; NEX = nex + 2;
    }
    /* terminate: MagSgn padded with ones; MEL flushes a pending run; MEL and VLC tails written separately */
    ms_finish(&ms);
    if (mel.run > 0) mel_bit(&mel, 1);
    if (mel.rem != 8 || mel.pos == 0) {
        /* flush the partial MEL byte (a fresh, empty byte is still written so that the segment is never empty) */
        mel.tmp <<= mel.rem;
        if (mel.pos < mel.cap) mel.buf[mel.pos++] = (uint8_t)mel.tmp; else mel.ovf = 1;
    }
    if (vlc.used > 0) {                                      /* partial VLC byte (at least the Scup-nibble byte) */
        if (vlc.pos < vlc.cap) { *(vlc.end - vlc.pos) = (uint8_t)vlc.tmp; vlc.pos++; } else vlc.ovf = 1;
    }
    int ret = -1;
    if (!bad && !ms.ovf && !mel.ovf && !vlc.ovf) {
        const int scup = mel.pos + vlc.pos;
        const int total = ms.pos + scup;
        if (scup >= 2 && scup <= 4079 && total <= cap) {
            memcpy(out, msb, (size_t)ms.pos);
            memcpy(out + ms.pos, melb, (size_t)mel.pos);
            memcpy(out + ms.pos + mel.pos, vlc.end - vlc.pos + 1, (size_t)vlc.pos);
            out[total - 1] = (uint8_t)(scup >> 4);
            out[total - 2] = (uint8_t)((out[total - 2] & 0xF0) | (scup & 0x0F));
            ret = total;
        }
    }
    free(msb); free(melb); free(vlcb); free(sg); free(ex); free(nsg); free(nex);
    return ret;
}

/* ---- HT refinement passes (T.814 clause 7.4 SigProp, 7.5 MagRef) -------------------------------------------------
 * One HT set: the cleanup pass codes mu = |x| >> P (all bit-planes >= P), the SigProp and MagRef passes code
 * bit-plane P - 1.  SigProp: 4-row stripes, groups of 4 columns, column by column, top to bottom; a sample that is
 * insignificant after the cleanup pass and has a significant neighbour (cleanup significance of all 8 neighbours,
 * plus SigProp significance of the ones already visited) gets one bit; the sign bits of the samples that turned
 * significant follow after the group's (up to 16) significance bits.  Bits go LSB first into a forward-growing
 * stream with 7 bits after a 0xFF byte.  MagRef: one bit per cleanup-significant sample in stripe scan order, LSB
 * first into a stream that grows backward from the end of the refinement segment (the VLC stuffing rule, with the
 * byte after the end of the segment taken as 0xFF).  The refinement segment = SigProp bytes, then MagRef bytes.
 * recon (optional, w * h): what a conformant decoder reconstructs, in quarter units (integer LSB = bit 2) with the
 * mid-point bit below the last decoded bit-plane, signed: the writer's own statement of the expected result. */
int gen_iso_ht_encode_passes(const int32_t *coef, int w, int h, int P, int npasses, uint8_t *out, int cap,
                             int *lcup_out, int32_t *recon)
{
    if (npasses < 1 || npasses > 3 || P < 0 || P > 20 || (npasses > 1 && P < 1)) return -1;
    const int n = w * h;
    int32_t *mu = malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const int32_t a = coef[i] < 0 ? -coef[i] : coef[i];
        mu[i] = coef[i] < 0 ? -(a >> P) : (a >> P);
    }
    const int lcup = gen_iso_ht_encode(mu, w, h, out, cap);
    if (lcup_out) *lcup_out = lcup > 0 ? lcup : 0;
    if (recon) memset(recon, 0, sizeof(int32_t) * (size_t)n);
    if (lcup <= 0) { free(mu); return lcup; }
    if (recon)
        for (int i = 0; i < n; i++)
            if (mu[i]) {
                const int32_t m = mu[i] < 0 ? -mu[i] : mu[i];
                const int32_t q = (2 * m + 1) << (P + 1);
                recon[i] = mu[i] < 0 ? -q : q;
            }
    if (npasses == 1) { free(mu); return lcup; }

    const int seg = n / 2 + 64;
    uint8_t *sppb = malloc((size_t)seg), *mrpb = malloc((size_t)seg);
    uint8_t *sspp = calloc((size_t)n, 1);
    msw_t spp = {sppb, seg, 0, 0, 0, 8, 0};
    vlcw_t mrp = {mrpb + seg - 1, seg, 0, 0, 0, 1, 0};
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
                            if ((dy || dx) && yy >= 0 && yy < h && xx >= 0 && xx < w && (mu[yy * w + xx] || sspp[yy * w + xx])) mbr = 1;
                        }
                    if (!mbr) continue;
                    const int32_t a = coef[i] < 0 ? -coef[i] : coef[i];
                    const int bit = (a >> (P - 1)) & 1;
                    ms_put(&spp, (uint32_t)bit, 1);
                    if (bit) { sspp[i] = 1; newi[nnew++] = i; }
                }
            for (int k = 0; k < nnew; k++) {
                ms_put(&spp, coef[newi[k]] < 0 ? 1u : 0u, 1);
                if (recon) recon[newi[k]] = coef[newi[k]] < 0 ? -(3 << P) : (3 << P);
            }
        }
    if (spp.used) { spp.buf[spp.pos++] = (uint8_t)spp.tmp; }                 /* partial byte, zero padded */
    if (spp.pos == 0 || spp.buf[spp.pos - 1] == 0xFF) spp.buf[spp.pos++] = 0;  /* never empty, never ends in 0xFF */
    if (npasses == 3) {
        for (int y0 = 0; y0 < h; y0 += 4)
            for (int x = 0; x < w; x++)
                for (int y = y0; y < y0 + 4 && y < h; y++) {
                    const int i = y * w + x;
                    if (!mu[i]) continue;
                    const int32_t a = coef[i] < 0 ? -coef[i] : coef[i];
                    const int bit = (a >> (P - 1)) & 1;
                    vlc_put(&mrp, (uint32_t)bit, 1);
                    if (recon) {
                        const int32_t q = ((a >> P) << (P + 2)) | (bit << (P + 1)) | (1 << P);
                        recon[i] = coef[i] < 0 ? -q : q;
                    }
                }
        if (mrp.used > 0) { *(mrp.end - mrp.pos) = (uint8_t)mrp.tmp; mrp.pos++; }
    }
    int ret = -1;
    if (!spp.ovf && !mrp.ovf && spp.pos < seg - 2 && lcup + spp.pos + mrp.pos <= cap) {
        memcpy(out + lcup, sppb, (size_t)spp.pos);
        memcpy(out + lcup + spp.pos, mrp.end - mrp.pos + 1, (size_t)mrp.pos);
        ret = lcup + spp.pos + mrp.pos;
    }
    free(mu); free(sppb); free(mrpb); free(sspp);
    return ret;
}
