/*
 * gen_ht.c -- restatement of the reference "HT" block ENCODER, HTEncoder.Encode
 * (internal/entropy/ht.go:942-1391).  Part of datagen/ (synthetic-input generator).  The reference
 * encoder is not ISO/IEC 15444-15 and its output does not round-trip through the reference decoder
 * (zero-filled MEL segment of maxSize/4 bytes, byte-reversed VLC segment, ht.go:978,1019,1036-1038);
 * it is restated as it is because its bytes are what the reference would hand its own decoder.
 */
#include "datagen.h"
#include <stdlib.h>
#include <string.h>

#include "ht_vlc_tables.inc"
static const uint16_t k_vlc_tbl0[1024] = HT_VLC_TBL0_INIT;
static const uint16_t k_vlc_tbl1[1024] = HT_VLC_TBL1_INIT;

static inline uint32_t shl32(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
static inline uint64_t shl64(uint64_t v, uint32_t n) { return n >= 64 ? 0u : v << n; }

/* ------------------------------ encoder ---------------------------------------- */
typedef struct { uint8_t *data; int cap; int pos; uint64_t tmp; int bits; uint8_t last; int ovf; } bw_t;

static void vlc_write(bw_t *v, uint32_t val, uint32_t nbits)       /* ht.go:1266-1286 */
{
    v->tmp |= shl64((uint64_t)val, (uint32_t)v->bits);
    v->bits += (int)nbits;
    while (v->bits >= 8) {
        uint8_t b = (uint8_t)(v->tmp & 0xFF);
        if (v->last > 0x8F && (b & 0x7F) == 0x7F) b &= 0x7F;
        if (v->pos < 0) { v->ovf = 1; return; }
        v->data[v->pos] = b; v->pos--;
        v->last = b;
        v->tmp >>= 8; v->bits -= 8;
    }
}

static void vlc_flush(bw_t *v)                                      /* ht.go:1289-1300 */
{
    while (v->bits > 0) {
        if (v->pos < 0) { v->ovf = 1; return; }
        v->data[v->pos] = (uint8_t)(v->tmp & 0xFF); v->pos--;
        v->tmp >>= 8; v->bits -= 8;
        if (v->bits < 0) v->bits = 0;
    }
}

static void ms_write(bw_t *m, uint32_t val, uint32_t nbits)        /* ht.go:1303-1327 */
{
    m->tmp |= shl64((uint64_t)val, (uint32_t)m->bits);
    m->bits += (int)nbits;
    while (m->bits >= 8) {
        uint8_t b = (uint8_t)(m->tmp & 0xFF);
        if (m->pos >= m->cap) { m->ovf = 1; return; }
        if (m->last == 0xFF) {
            b &= 0x7F;
            m->data[m->pos++] = b; m->tmp >>= 7; m->bits -= 7;
        } else {
            m->data[m->pos++] = b; m->tmp >>= 8; m->bits -= 8;
        }
        m->last = b;
    }
}

static void ms_flush(bw_t *m)                                       /* ht.go:1330-1341 */
{
    while (m->bits > 0) {
        if (m->pos >= m->cap) { m->ovf = 1; return; }
        m->data[m->pos++] = (uint8_t)(m->tmp & 0xFF);
        m->tmp >>= 8; m->bits -= 8;
        if (m->bits < 0) m->bits = 0;
    }
}

static void enc_vlc_quad(bw_t *v, uint8_t ctx, uint8_t rho, int initial)   /* ht.go:1199-1226 */
{
    const uint16_t *tbl = initial ? k_vlc_tbl0 : k_vlc_tbl1;
    for (uint32_t cwd = 0; cwd < 128; cwd++) {
        uint16_t e = tbl[((uint32_t)ctx << 7) | cwd];
        uint32_t elen = e & 0x0F, erho = (e >> 4) & 0x0F;
        if ((uint8_t)erho == rho && elen > 0) { vlc_write(v, cwd, elen); return; }
    }
    vlc_write(v, 0, 1);
}

static void enc_uvlc_one(bw_t *v, uint32_t u)                       /* ht.go:1242-1249 */
{
    if (u <= 1) vlc_write(v, 1, 1);
    else if (u <= 2) vlc_write(v, 2, 2);
    else { vlc_write(v, 0, 3); vlc_write(v, u - 3, 5); }
}

static void enc_magsgn(bw_t *m, int32_t v)                          /* ht.go:1149-1166 */
{
    uint32_t sign = 0;
    if (v < 0) { sign = 1; v = (int32_t)(0u - (uint32_t)v); }
    uint32_t mag = (uint32_t)v, emb = 1;
    while (emb < 32 && mag >= shl32(1, emb)) emb++;   /* mag < 2^31 required (Go loops forever otherwise) */
    ms_write(m, mag & (shl32(1, emb - 1) - 1), emb - 1);
    ms_write(m, sign, 1);
}

/* HTEncoder.Encode ht.go:942-1045; returns byte count, 0 for nil, -1 if the
 * reference would index out of range (panic) or cap is too small. */
int gen_ht_encode(const int32_t *d, int w, int h, int band, uint8_t *out, int cap)
{
    (void)band;
    int n = w * h;
    int32_t maxmag = 0;
    for (int i = 0; i < n; i++) {
        int32_t v = d[i] < 0 ? (int32_t)(0u - (uint32_t)d[i]) : d[i];
        if (v > maxmag) maxmag = v;
    }
    if (maxmag == 0) return 0;
    int max_size = n * 2;
    if (max_size < 64) max_size = 64;
    int mel_len = max_size / 4;                  /* zero bytes: make([]byte, maxSize/4), never written */
    bw_t vlc = {0}, ms = {0};
    vlc.cap = max_size / 2; vlc.data = (uint8_t *)calloc((size_t)vlc.cap, 1); vlc.pos = vlc.cap - 1;
    ms.cap = max_size / 2;  ms.data = (uint8_t *)calloc((size_t)ms.cap, 1);
    int quad_cols = (w + 3) / 4;
    uint8_t *sigma1 = (uint8_t *)calloc((size_t)quad_cols + 2, 1);

    for (int y = 0; y < h; y += 4) {                              /* encodeCleanup ht.go:1048-1196 */
        int initial = (y == 0);
        for (int qx = 0; qx < quad_cols; qx += 2) {
            uint8_t rho1 = 0, rho2 = 0;
            for (int i = 0; i < 4 && qx * 4 + i < w; i++) {
                int idx = y * w + qx * 4 + i;
                if (idx < n && d[idx] != 0) rho1 |= (uint8_t)(1 << i);
            }
            for (int i = 0; i < 4 && (qx + 1) * 4 + i < w; i++) {
                int idx = y * w + (qx + 1) * 4 + i;
                if (idx < n && d[idx] != 0) rho2 |= (uint8_t)(1 << i);
            }
            uint8_t ctx = 0;
            if (initial) { if (qx > 0) ctx = sigma1[qx - 1] >> 4; }
            else ctx = sigma1[qx] >> 4;
            enc_vlc_quad(&vlc, ctx, rho1, initial);
            sigma1[qx] = rho1;
            uint8_t ctx2 = (uint8_t)((rho1 >> 2) | (sigma1[qx + 1] >> 4));
            enc_vlc_quad(&vlc, ctx2, rho2, initial);
            sigma1[qx + 1] = rho2;

            int uoff1 = rho1 != 0, uoff2 = rho2 != 0;
            if (uoff1 || uoff2) {
                uint32_t u1 = 1, u2 = 1;
                for (int i = 0; i < 4 && qx * 4 + i < w; i++) {
                    int idx = y * w + qx * 4 + i;
                    if (idx < n) {
                        int32_t v = d[idx] < 0 ? (int32_t)(0u - (uint32_t)d[idx]) : d[idx];
                        if ((uint32_t)v >= shl32(1, u1)) u1++;
                    }
                }
                for (int i = 0; i < 4 && (qx + 1) * 4 + i < w; i++) {
                    int idx = y * w + (qx + 1) * 4 + i;
                    if (idx < n) {
                        int32_t v = d[idx] < 0 ? (int32_t)(0u - (uint32_t)d[idx]) : d[idx];
                        if ((uint32_t)v >= shl32(1, u2)) u2++;
                    }
                }
                uint32_t mode = (uoff1 ? 1u : 0u) | (uoff2 ? 2u : 0u);
                if (mode == 1) enc_uvlc_one(&vlc, u1);              /* encodeUVLC ht.go:1229-1263 */
                else if (mode == 2) enc_uvlc_one(&vlc, u2);
                else { enc_uvlc_one(&vlc, u1); enc_uvlc_one(&vlc, u2); }
            }
            for (int i = 0; i < 4 && qx * 4 + i < w; i++) {
                int idx = y * w + qx * 4 + i;
                if ((rho1 & (1 << i)) && idx < n) enc_magsgn(&ms, d[idx]);
            }
            for (int i = 0; i < 4 && (qx + 1) * 4 + i < w; i++) {
                int idx = y * w + (qx + 1) * 4 + i;
                if ((rho2 & (1 << i)) && idx < n) enc_magsgn(&ms, d[idx]);
            }
        }
    }
    /* melFlush is a no-op (run never counted); then vlcFlush, magSgnFlush ht.go:1008-1010 */
    vlc_flush(&vlc);
    ms_flush(&ms);

    int ret = -1;
    if (!vlc.ovf && !ms.ovf) {
        int ms_len = ms.pos;
        int vlc_len = vlc.cap - vlc.pos - 1;
        int scup = mel_len + vlc_len + 2;
        int total = ms_len + scup;
        if (total <= cap) {
            memcpy(out, ms.data, (size_t)ms_len);
            memset(out + ms_len, 0, (size_t)mel_len);
            for (int i = 0; i < vlc_len; i++) out[ms_len + mel_len + i] = vlc.data[vlc.cap - 1 - i];
            out[total - 2] = (uint8_t)(scup >> 8);
            out[total - 1] = (uint8_t)(scup & 0xFF);
            ret = total;
        }
    }
    free(vlc.data); free(ms.data); free(sigma1);
    return ret;
}
