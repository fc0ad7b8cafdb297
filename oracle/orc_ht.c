/*
 * orc_ht.c -- restatement of the reference "HT" block coder
 * (internal/entropy/ht.go, ht_luts.go).  Oracle / test infrastructure only.
 *
 * PARITY UNPINNED: no reference test asserts a decoded value for this coder
 * (ht_test.go:57-75, tcd_htj2k_test.go:158-165), so the code below restates
 * ht.go statement by statement, including what is not ISO/IEC 15444-15:
 *   - only sample row y of each 4-row stripe is produced (ht.go:677,701);
 *   - "quads" are 1x4 row segments handled in pairs (ht.go:593);
 *   - the MEL stream is initialised (and can veto the block) but never read;
 *   - the VLC length field is read with mask 0x0F (ht.go:620);
 *   - the first-quad context is always 0 (sigma holds 4-bit rho, ht.go:602-606);
 *   - the encoder emits a zero-filled MEL segment of maxSize/4 bytes and a
 *     byte-reversed VLC segment (ht.go:978,1019,1036-1038), so encode->decode is
 *     not an identity.  Decoder parity is defined on a fresh (zeroed) decoder.
 * Go semantics: uint32 shifts by >= 32 give 0, uint32 "bits" counters wrap.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#include "ht_vlc_tables.inc"
static const uint16_t k_vlc_tbl0[1024] = HT_VLC_TBL0_INIT;
static const uint16_t k_vlc_tbl1[1024] = HT_VLC_TBL1_INIT;

static inline uint32_t shl32(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
static inline uint64_t shr64(uint64_t v, uint32_t n) { return n >= 64 ? 0u : v >> n; }
static inline uint64_t shl64(uint64_t v, uint32_t n) { return n >= 64 ? 0u : v << n; }

/* ------------------------------ readers --------------------------------------- */
typedef struct {            /* revBitstream ht.go:57-64 */
    const uint8_t *data; int len;
    int pos; uint64_t tmp; uint32_t bits; int size; int unstuff;
} rev_t;

typedef struct {            /* frwdBitstream ht.go:67-75 */
    const uint8_t *data; int len;
    int pos; uint64_t tmp; uint32_t bits; int unstuff; int size; uint32_t x;
} fwd_t;

/* initMEL ht.go:153-195 -- only its boolean result is observable */
static int mel_init_ok(const uint8_t *data, int len, int lcup, int scup)
{
    int pos = lcup - scup, size = scup - 1, unstuff = 0;
    int num = 4 - (pos & 3);
    for (int i = 0; i < num && size > 0; i++) {
        if (unstuff && pos < len && data[pos] > 0x8F) return 0;
        uint8_t b;
        if (size > 0 && pos < len) { b = data[pos]; pos++; size--; }
        else b = 0xFF;
        if (size == 1) b |= 0x0F;
        unstuff = (b == 0xFF);
    }
    return 1;
}

/* revRead ht.go:317-378 */
static void rev_read(rev_t *v)
{
    if (v->bits > 32) return;
    uint32_t val = 0;
    if (v->size > 3) {
        int p = v->pos - 3;
        if (p >= 0 && p + 3 < v->len)
            val = (uint32_t)v->data[p] | (uint32_t)v->data[p + 1] << 8 |
                  (uint32_t)v->data[p + 2] << 16 | (uint32_t)v->data[p + 3] << 24;
        v->pos -= 4; v->size -= 4;
    } else if (v->size > 0) {
        int i = 24;
        while (v->size > 0) {
            if (v->pos >= 0 && v->pos < v->len) { val |= (uint32_t)v->data[v->pos] << i; v->pos--; }
            v->size--; i -= 8;
        }
    }
    uint32_t tmp = val >> 24;
    uint32_t bits = (v->unstuff && ((val >> 24) & 0x7F) == 0x7F) ? 7 : 8;
    int unstuff = (val >> 24) > 0x8F;

    tmp |= ((val >> 16) & 0xFF) << bits;
    bits += (unstuff && ((val >> 16) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 16) & 0xFF) > 0x8F;

    tmp |= ((val >> 8) & 0xFF) << bits;
    bits += (unstuff && ((val >> 8) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 8) & 0xFF) > 0x8F;

    tmp |= (val & 0xFF) << bits;
    bits += (unstuff && (val & 0x7F) == 0x7F) ? 7 : 8;
    v->unstuff = (val & 0xFF) > 0x8F;

    v->tmp |= shl64((uint64_t)tmp, v->bits);
    v->bits += bits;
}

static uint32_t rev_fetch(rev_t *v)                       /* ht.go:381-389 */
{
    if (v->bits < 32) { rev_read(v); if (v->bits < 32) rev_read(v); }
    return (uint32_t)v->tmp;
}

static void rev_advance(rev_t *v, uint32_t n)             /* ht.go:392-396 */
{
    v->tmp = shr64(v->tmp, n);
    v->bits -= n;
}

/* initVLC ht.go:276-314 */
static void vlc_init(rev_t *v, const uint8_t *data, int len, int lcup, int scup)
{
    v->data = data; v->len = len;
    v->pos = lcup - 2; v->size = scup - 2; v->tmp = 0; v->bits = 0; v->unstuff = 0;
    if (v->pos >= 0 && v->pos < len) {
        uint8_t b = data[v->pos];
        v->pos--;
        v->tmp = (uint64_t)(b >> 4);
        v->bits = 4 - (uint32_t)((v->tmp & 7) >> 2);
        v->unstuff = (b | 0x0F) > 0x8F;
    }
    int num = 1 + (v->pos & 3);          /* Go: -1 & 3 == 3, same in two's-complement C */
    if (num > v->size) num = v->size;
    for (int i = 0; i < num; i++) {
        uint8_t b = 0;
        if (v->pos >= 0 && v->pos < len) { b = data[v->pos]; v->pos--; }
        uint32_t dbits = (v->unstuff && (b & 0x7F) == 0x7F) ? 7 : 8;
        v->tmp |= shl64((uint64_t)b, v->bits);
        v->bits += dbits;
        v->unstuff = b > 0x8F;
    }
    v->size -= num;
    rev_read(v);
}

/* frwdRead ht.go:432-501 */
static void fwd_read(fwd_t *f)
{
    if (f->bits > 32) return;
    uint32_t val = 0;
    if (f->size > 3) {
        if (f->pos + 3 < f->len)
            val = (uint32_t)f->data[f->pos] | (uint32_t)f->data[f->pos + 1] << 8 |
                  (uint32_t)f->data[f->pos + 2] << 16 | (uint32_t)f->data[f->pos + 3] << 24;
        f->pos += 4; f->size -= 4;
    } else if (f->size > 0) {
        if (f->x != 0) val = 0xFFFFFFFFu;
        int i = 0;
        while (f->size > 0) {
            if (f->pos < f->len) {
                uint32_t b = f->data[f->pos];
                uint32_t m = ~((uint32_t)0xFF << i);
                val = (val & m) | (b << i);
                f->pos++;
            }
            f->size--; i += 8;
        }
    } else if (f->x != 0) {
        val = 0xFFFFFFFFu;
    }
    uint32_t bits = f->unstuff ? 7 : 8;
    uint32_t t = val & 0xFF;
    int unstuff = (val & 0xFF) == 0xFF;

    t |= ((val >> 8) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    unstuff = ((val >> 8) & 0xFF) == 0xFF;

    t |= ((val >> 16) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    unstuff = ((val >> 16) & 0xFF) == 0xFF;

    t |= ((val >> 24) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    f->unstuff = ((val >> 24) & 0xFF) == 0xFF;

    f->tmp |= shl64((uint64_t)t, f->bits);
    f->bits += bits;
}

static uint32_t fwd_fetch(fwd_t *f)                       /* ht.go:504-512 */
{
    if (f->bits < 32) { fwd_read(f); if (f->bits < 32) fwd_read(f); }
    return (uint32_t)f->tmp;
}

static void fwd_advance(fwd_t *f, uint32_t n)             /* ht.go:515-519 */
{
    f->tmp = shr64(f->tmp, n);
    f->bits -= n;                                         /* uint32 wrap kept */
}

/* initMagSgn ht.go:399-429 */
static void magsgn_init(fwd_t *f, const uint8_t *data, int len, int size)
{
    f->data = data; f->len = len;
    f->pos = 0; f->size = size; f->tmp = 0; f->bits = 0; f->unstuff = 0; f->x = 0xFF;
    int num = 4 - (f->pos & 3);
    for (int i = 0; i < num; i++) {
        uint8_t b;
        if (f->size > 0 && f->pos < len) { b = data[f->pos]; f->pos++; f->size--; }
        else b = (uint8_t)f->x;
        uint32_t dbits = f->unstuff ? 7 : 8;
        f->tmp |= shl64((uint64_t)b, f->bits);
        f->bits += dbits;
        f->unstuff = (b == 0xFF);
    }
    fwd_read(f);
}

/* UVLC prefix table ht.go:718-727: prefix_len | suffix_len<<2 | base<<5 */
static const uint8_t k_uvlc_dec[8] = {
    3 | (5 << 2) | (5 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5),
    3 | (1 << 2) | (3 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5),
};

/* decodeInitUVLC ht.go:716-805 / decodeNonInitUVLC ht.go:808-864 (mode is 1..3 here) */
static uint32_t uvlc_decode(uint32_t vlc, uint32_t mode, int initial, uint32_t u[2])
{
    uint32_t consumed = 0;
    if (mode == 0) { u[0] = 1; u[1] = 1; return 0; }
    if (mode <= 2) {
        uint32_t t = k_uvlc_dec[vlc & 7];
        uint32_t plen = t & 3;
        vlc >>= plen; consumed += plen;
        uint32_t slen = (t >> 2) & 7;
        consumed += slen;
        uint32_t val = (t >> 5) + (vlc & ((1u << slen) - 1));
        if (mode == 1) { u[0] = val + 1; u[1] = 1; } else { u[0] = 1; u[1] = val + 1; }
        return consumed;
    }
    /* mode == 3 */
    uint32_t t1 = k_uvlc_dec[vlc & 7];
    uint32_t p1 = t1 & 3;
    vlc >>= p1; consumed += p1;
    if (initial && p1 > 2) {                              /* ht.go:756-764 */
        u[1] = (vlc & 1) + 2;
        consumed++; vlc >>= 1;
        uint32_t slen = (t1 >> 2) & 7;
        consumed += slen;
        u[0] = (t1 >> 5) + (vlc & ((1u << slen) - 1)) + 1;
        return consumed;
    }
    uint32_t t2 = k_uvlc_dec[vlc & 7];
    uint32_t p2 = t2 & 3;
    vlc >>= p2; consumed += p2;
    uint32_t s1 = (t1 >> 2) & 7;
    consumed += s1;
    u[0] = (t1 >> 5) + (vlc & ((1u << s1) - 1)) + 1;
    vlc >>= s1;
    uint32_t s2 = (t2 >> 2) & 7;
    consumed += s2;
    u[1] = (t2 >> 5) + (vlc & ((1u << s2) - 1)) + 1;
    return consumed;
}

/* one MagSgn sample, ht.go:664-684 */
static void magsgn_sample(fwd_t *ms, uint32_t emb, int32_t *out, int idx, int n_out)
{
    uint32_t mv = fwd_fetch(ms);
    uint32_t m = (mv & (shl32(1, emb) - 1)) + shl32(1, emb - 1);
    fwd_advance(ms, emb);
    uint32_t sign = fwd_fetch(ms) & 1;
    fwd_advance(ms, 1);
    if (idx < n_out) out[idx] = (int32_t)(sign ? 0u - m : m);
}

/* HTDecoder.Decode ht.go:93-150 + decodeCleanup ht.go:583-713 */
void orc_ht_decode(const uint8_t *data, int len, int w, int h, int32_t *out)
{
    int n_out = w * h;
    memset(out, 0, sizeof(int32_t) * (size_t)n_out);
    if (len < 2) return;
    int scup = (int)data[len - 1] + ((int)(data[len - 2] & 0x0F) << 8);
    if (scup < 2 || scup > len) return;
    int lcup = len;
    if (!mel_init_ok(data, len, lcup, scup)) return;
    rev_t vlc; fwd_t ms;
    vlc_init(&vlc, data, len, lcup, scup);
    magsgn_init(&ms, data, len, lcup - scup);

    int quad_cols = (w + 3) / 4;
    uint8_t *sigma1 = (uint8_t *)calloc((size_t)quad_cols + 2, 1);
    uint8_t *line_state = (uint8_t *)calloc((size_t)quad_cols + 2, 1);

    for (int y = 0; y < h; y += 4) {
        int initial = (y == 0);
        const uint16_t *tbl = initial ? k_vlc_tbl0 : k_vlc_tbl1;
        for (int qx = 0; qx < quad_cols; qx += 2) {
            uint32_t vv = rev_fetch(&vlc);
            uint8_t ctx = 0;
            if (initial) { if (qx > 0) ctx = sigma1[qx - 1] >> 4; }
            else ctx = (uint8_t)((sigma1[qx] >> 4) | (line_state[qx] >> 4));
            uint16_t q1 = tbl[((uint32_t)ctx << 7) | (vv & 0x7F)];
            uint32_t len1 = q1 & 0x0F, rho1 = (q1 >> 4) & 0x0F, uoff1 = (q1 >> 3) & 1;
            rev_advance(&vlc, len1);
            vv = rev_fetch(&vlc);
            uint8_t ctx2 = (uint8_t)((rho1 >> 2) | (sigma1[qx + 1] >> 4));
            uint16_t q2 = tbl[((uint32_t)ctx2 << 7) | (vv & 0x7F)];
            uint32_t len2 = q2 & 0x0F, rho2 = (q2 >> 4) & 0x0F, uoff2 = (q2 >> 3) & 1;
            rev_advance(&vlc, len2);
            sigma1[qx] = (uint8_t)rho1;
            sigma1[qx + 1] = (uint8_t)rho2;

            uint32_t u[2] = {0, 0};
            uint32_t mode = (uoff1 << 1) | uoff2;
            if (mode > 0) {
                vv = rev_fetch(&vlc);
                rev_advance(&vlc, uvlc_decode(vv, mode, initial, u));
            } else { u[0] = 1; u[1] = 1; }

            for (int i = 0; i < 4 && qx * 4 + i < w; i++)
                if (rho1 & (1u << i)) magsgn_sample(&ms, u[0], out, y * w + qx * 4 + i, n_out);
            for (int i = 0; i < 4 && (qx + 1) * 4 + i < w; i++)
                if (rho2 & (1u << i)) magsgn_sample(&ms, u[1], out, y * w + (qx + 1) * 4 + i, n_out);
        }
    }
    free(sigma1); free(line_state);
}

