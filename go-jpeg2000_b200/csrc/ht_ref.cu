// ht_ref.cu -- "HT" block decoder, REF semantics, one warp per code block (sm_100a).
//
// Replaces entropy.HTDecoder.Decode (reference internal/entropy/ht.go:93-150): MEL init check
// (ht.go:153-195), backward VLC reader (ht.go:276-396), forward MagSgn reader (ht.go:399-519),
// decodeCleanup (ht.go:583-713) and the two UVLC routines (ht.go:716-864).  The reference coder
// is not ISO/IEC 15444-15: only sample row y of each 4-row stripe is produced,
// the MEL stream is never consumed, the VLC length field is read with mask 0x0F.  REF mode
// reproduces exactly that; the conformant decoder is the ISO-mode kernel.
//
// The bit readers are serial, so the warp runs them lock-step on uniform registers; the lanes
// share the parallel part (zero-filling the block in the tile-component plane, coalesced).
// Go semantics kept: uint32 shifts >= 32 give 0; the uint32 "bits" counters wrap when a MagSgn
// field is longer than the buffered bits (emb up to 37, ht.go:668-669).
#include "common.h"
#include <cstdlib>

namespace {

constexpr int kWarpsPerCta = 4;

#include "ht_vlc_tables.inc"
__device__ const uint16_t d_vlc_tbl0[1024] = HT_VLC_TBL0_INIT;
__device__ const uint16_t d_vlc_tbl1[1024] = HT_VLC_TBL1_INIT;

// UVLC prefix rows, ht.go:718-727: prefix_len | suffix_len << 2 | base << 5
__constant__ uint8_t c_uvlc_dec[8] = {
    3 | (5 << 2) | (5 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5),
    3 | (1 << 2) | (3 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5)};

__device__ __forceinline__ uint32_t shl32(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
__device__ __forceinline__ uint64_t shl64(uint64_t v, uint32_t n) { return n >= 64 ? 0ull : v << n; }
__device__ __forceinline__ uint64_t shr64(uint64_t v, uint32_t n) { return n >= 64 ? 0ull : v >> n; }

struct Rev { const uint8_t *d; int len, pos, size; uint64_t tmp; uint32_t bits; bool unstuff; };
struct Fwd { const uint8_t *d; int len, pos, size; uint64_t tmp; uint32_t bits; bool unstuff; };

__device__ __forceinline__ uint32_t byte_at(const uint8_t *d, int len, int pos)
{
    return (pos >= 0 && pos < len) ? (uint32_t)__ldg(d + pos) : 0u;
}

// initMEL ht.go:153-195: only the boolean is observable
__device__ bool mel_init_ok(const uint8_t *d, int len, int lcup, int scup)
{
    int pos = lcup - scup, size = scup - 1;
    bool unstuff = false;
    int num = 4 - (pos & 3);
    for (int i = 0; i < num && size > 0; i++) {
        if (unstuff && pos < len && byte_at(d, len, pos) > 0x8F) return false;
        uint32_t b;
        if (size > 0 && pos < len) { b = byte_at(d, len, pos); pos++; size--; }
        else b = 0xFF;
        if (size == 1) b |= 0x0F;
        unstuff = (b == 0xFF);
    }
    return true;
}

// revRead ht.go:317-378
__device__ void rev_read(Rev &v)
{
    if (v.bits > 32) return;
    uint32_t val = 0;
    if (v.size > 3) {
        int p = v.pos - 3;
        if (p >= 0 && p + 3 < v.len)
            val = byte_at(v.d, v.len, p) | byte_at(v.d, v.len, p + 1) << 8 |
                  byte_at(v.d, v.len, p + 2) << 16 | byte_at(v.d, v.len, p + 3) << 24;
        v.pos -= 4; v.size -= 4;
    } else if (v.size > 0) {
        int i = 24;
        while (v.size > 0) {
            if (v.pos >= 0 && v.pos < v.len) { val |= byte_at(v.d, v.len, v.pos) << i; v.pos--; }
            v.size--; i -= 8;
        }
    }
    uint32_t tmp = val >> 24;
    uint32_t bits = (v.unstuff && ((val >> 24) & 0x7F) == 0x7F) ? 7 : 8;
    bool unstuff = (val >> 24) > 0x8F;
    tmp |= ((val >> 16) & 0xFF) << bits;
    bits += (unstuff && ((val >> 16) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 16) & 0xFF) > 0x8F;
    tmp |= ((val >> 8) & 0xFF) << bits;
    bits += (unstuff && ((val >> 8) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 8) & 0xFF) > 0x8F;
    tmp |= (val & 0xFF) << bits;
    bits += (unstuff && (val & 0x7F) == 0x7F) ? 7 : 8;
    v.unstuff = (val & 0xFF) > 0x8F;
    v.tmp |= shl64((uint64_t)tmp, v.bits);
    v.bits += bits;
}

__device__ __forceinline__ uint32_t rev_fetch(Rev &v)                 // ht.go:381-389
{
    if (v.bits < 32) { rev_read(v); if (v.bits < 32) rev_read(v); }
    return (uint32_t)v.tmp;
}
__device__ __forceinline__ void rev_advance(Rev &v, uint32_t n) { v.tmp = shr64(v.tmp, n); v.bits -= n; }

// initVLC ht.go:276-314
__device__ void vlc_init(Rev &v, const uint8_t *d, int len, int lcup, int scup)
{
    v.d = d; v.len = len; v.pos = lcup - 2; v.size = scup - 2; v.tmp = 0; v.bits = 0; v.unstuff = false;
    if (v.pos >= 0 && v.pos < len) {
        uint32_t b = byte_at(d, len, v.pos);
        v.pos--;
        v.tmp = b >> 4;
        v.bits = 4 - (uint32_t)((v.tmp & 7) >> 2);
        v.unstuff = (b | 0x0F) > 0x8F;
    }
    int num = 1 + (v.pos & 3);
    if (num > v.size) num = v.size;
    for (int i = 0; i < num; i++) {
        uint32_t b = 0;
        if (v.pos >= 0 && v.pos < len) { b = byte_at(d, len, v.pos); v.pos--; }
        uint32_t dbits = (v.unstuff && (b & 0x7F) == 0x7F) ? 7 : 8;
        v.tmp |= shl64((uint64_t)b, v.bits);
        v.bits += dbits;
        v.unstuff = b > 0x8F;
    }
    v.size -= num;
    rev_read(v);
}

// frwdRead ht.go:432-501 (MagSgn: exhausted stream feeds 0xFF)
__device__ void fwd_read(Fwd &f)
{
    if (f.bits > 32) return;
    uint32_t val = 0;
    if (f.size > 3) {
        if (f.pos + 3 < f.len)
            val = byte_at(f.d, f.len, f.pos) | byte_at(f.d, f.len, f.pos + 1) << 8 |
                  byte_at(f.d, f.len, f.pos + 2) << 16 | byte_at(f.d, f.len, f.pos + 3) << 24;
        f.pos += 4; f.size -= 4;
    } else if (f.size > 0) {
        val = 0xFFFFFFFFu;
        int i = 0;
        while (f.size > 0) {
            if (f.pos < f.len) {
                uint32_t b = byte_at(f.d, f.len, f.pos);
                val = (val & ~(0xFFu << i)) | (b << i);
                f.pos++;
            }
            f.size--; i += 8;
        }
    } else {
        val = 0xFFFFFFFFu;
    }
    uint32_t bits = f.unstuff ? 7 : 8;
    uint32_t t = val & 0xFF;
    bool unstuff = (val & 0xFF) == 0xFF;
    t |= ((val >> 8) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    unstuff = ((val >> 8) & 0xFF) == 0xFF;
    t |= ((val >> 16) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    unstuff = ((val >> 16) & 0xFF) == 0xFF;
    t |= ((val >> 24) & 0xFF) << bits;
    bits += unstuff ? 7 : 8;
    f.unstuff = ((val >> 24) & 0xFF) == 0xFF;
    f.tmp |= shl64((uint64_t)t, f.bits);
    f.bits += bits;
}

__device__ __forceinline__ uint32_t fwd_fetch(Fwd &f)                 // ht.go:504-512
{
    if (f.bits < 32) { fwd_read(f); if (f.bits < 32) fwd_read(f); }
    return (uint32_t)f.tmp;
}
__device__ __forceinline__ void fwd_advance(Fwd &f, uint32_t n) { f.tmp = shr64(f.tmp, n); f.bits -= n; }

// initMagSgn ht.go:399-429
__device__ void magsgn_init(Fwd &f, const uint8_t *d, int len, int size)
{
    f.d = d; f.len = len; f.pos = 0; f.size = size; f.tmp = 0; f.bits = 0; f.unstuff = false;
    for (int i = 0; i < 4; i++) {
        uint32_t b;
        if (f.size > 0 && f.pos < len) { b = byte_at(d, len, f.pos); f.pos++; f.size--; }
        else b = 0xFF;
        uint32_t dbits = f.unstuff ? 7 : 8;
        f.tmp |= shl64((uint64_t)b, f.bits);
        f.bits += dbits;
        f.unstuff = (b == 0xFF);
    }
    fwd_read(f);
}

// decodeInitUVLC ht.go:716-805 / decodeNonInitUVLC ht.go:808-864 (mode 1..3)
__device__ uint32_t uvlc_decode(uint32_t vlc, uint32_t mode, bool initial, uint32_t &u0, uint32_t &u1)
{
    uint32_t consumed = 0;
    if (mode <= 2) {
        uint32_t t = c_uvlc_dec[vlc & 7], plen = t & 3;
        vlc >>= plen; consumed += plen;
        uint32_t slen = (t >> 2) & 7;
        consumed += slen;
        uint32_t val = (t >> 5) + (vlc & ((1u << slen) - 1));
        if (mode == 1) { u0 = val + 1; u1 = 1; } else { u0 = 1; u1 = val + 1; }
        return consumed;
    }
    uint32_t t1 = c_uvlc_dec[vlc & 7], p1 = t1 & 3;
    vlc >>= p1; consumed += p1;
    if (initial && p1 > 2) {
        u1 = (vlc & 1) + 2;
        consumed++; vlc >>= 1;
        uint32_t slen = (t1 >> 2) & 7;
        consumed += slen;
        u0 = (t1 >> 5) + (vlc & ((1u << slen) - 1)) + 1;
        return consumed;
    }
    uint32_t t2 = c_uvlc_dec[vlc & 7], p2 = t2 & 3;
    vlc >>= p2; consumed += p2;
    uint32_t s1 = (t1 >> 2) & 7;
    consumed += s1;
    u0 = (t1 >> 5) + (vlc & ((1u << s1) - 1)) + 1;
    vlc >>= s1;
    uint32_t s2 = (t2 >> 2) & 7;
    consumed += s2;
    u1 = (t2 >> 5) + (vlc & ((1u << s2) - 1)) + 1;
    return consumed;
}

// one MagSgn sample, ht.go:664-684
__device__ __forceinline__ int32_t magsgn_sample(Fwd &ms, uint32_t emb)
{
    uint32_t mv = fwd_fetch(ms);
    uint32_t m = (mv & (shl32(1, emb) - 1)) + shl32(1, emb - 1);
    fwd_advance(ms, emb);
    uint32_t sign = fwd_fetch(ms) & 1;
    fwd_advance(ms, 1);
    return (int32_t)(sign ? 0u - m : m);
}

// Mapping (template BPW = code blocks per warp), like ht_iso.cu: the block's bit streams are one serial chain, so
//   BPW = 1   one warp per block: every lane runs the chain on uniform registers, lane 0 stores;
//   BPW = 32  one thread per block: 32 independent chains per warp (short divergent branches), 32x fewer
//             warp-instructions for the same work; the zero fill of the 32 blocks is done by the whole warp first.
// The launcher picks by J2KGPU_HTREF_MAP (default thread per block).
// ZSTEP: rows of a block that are cleared before decoding.  The reference decoder only ever writes sample row y of each
// 4-row stripe (ht.go:589-593, 677, 701), so inside a job -- whose planes were cleared once when the job was created and
// are written by nothing else -- clearing every 4th row is enough (ZSTEP = 4); the stage API clears all rows (ZSTEP = 1).
template <int BPW, typename OT, int ZSTEP>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_ht_ref(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
         OT *__restrict__ coef)
{
    __shared__ uint8_t s_sigma[20][kWarpsPerCta * 32];      // quadCols + 1 <= 17 for w <= 64; one column per decoder
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t blk;
    bool do_store;
    uint8_t *sigma1;                                        // element i at sigma1[i * kWarpsPerCta * 32]
    if (BPW == 32) {
        // fresh decoder: output pre-zeroed (ht.go:81) -- the warp clears its 32 blocks together (coalesced rows)
        const uint32_t first = (blockIdx.x * kWarpsPerCta + warp) * 32;
        for (uint32_t bb = first; bb < first + 32 && bb < n; bb++) {
            const DevCblk c0 = cblks[bb];
            OT *o = coef + c0.out_off;
            for (int y = 0; y < c0.h; y += ZSTEP)
                for (int x = lane; x < c0.w; x += 32) o[(size_t)y * c0.out_stride + x] = 0;
        }
        __syncwarp();
        blk = first + lane;
        do_store = true;
        sigma1 = &s_sigma[0][threadIdx.x];
    } else {
        blk = blockIdx.x * kWarpsPerCta + warp;
        do_store = lane == 0;
        sigma1 = &s_sigma[0][warp * 32];
    }
    constexpr int SS = kWarpsPerCta * 32;
    if (blk >= n) return;
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h, len = (int)cb.data_len;
    OT *out = coef + cb.out_off;
    const uint32_t ostride = cb.out_stride;
    const uint8_t *d = blob + cb.data_off;

    if (BPW == 1) {
        for (int y = 0; y < h; y += ZSTEP)
            for (int x = lane; x < w; x += 32) out[(size_t)y * ostride + x] = 0;
        if (lane < 20) sigma1[lane * SS] = 0;
        __syncwarp();
    } else {
        for (int i = 0; i < 20; i++) sigma1[i * SS] = 0;
    }

    if (len < 2) return;                                                    // ht.go:94-100
    int scup = (int)byte_at(d, len, len - 1) + (int)((byte_at(d, len, len - 2) & 0x0F) << 8);
    if (scup < 2 || scup > len) return;                                     // ht.go:105-111
    const int lcup = len;
    if (!mel_init_ok(d, len, lcup, scup)) return;                           // ht.go:117-122
    Rev vlc; Fwd ms;
    vlc_init(vlc, d, len, lcup, scup);
    magsgn_init(ms, d, len, lcup - scup);

    const int quad_cols = (w + 3) / 4;
    for (int y = 0; y < h; y += 4) {                                        // ht.go:589
        const bool initial = (y == 0);
        const uint16_t *tbl = initial ? d_vlc_tbl0 : d_vlc_tbl1;
        for (int qx = 0; qx < quad_cols; qx += 2) {
            uint32_t vv = rev_fetch(vlc);
            uint32_t ctx = 0;
            if (initial) { if (qx > 0) ctx = sigma1[(qx - 1) * SS] >> 4; }
            else ctx = sigma1[qx * SS] >> 4;                                // lineState is never written: 0
            uint32_t q1 = tbl[(ctx << 7) | (vv & 0x7F)];
            uint32_t len1 = q1 & 0x0F, rho1 = (q1 >> 4) & 0x0F, uoff1 = (q1 >> 3) & 1;
            rev_advance(vlc, len1);
            vv = rev_fetch(vlc);
            uint32_t ctx2 = (rho1 >> 2) | (sigma1[(qx + 1) * SS] >> 4);
            uint32_t q2 = tbl[(ctx2 << 7) | (vv & 0x7F)];
            uint32_t len2 = q2 & 0x0F, rho2 = (q2 >> 4) & 0x0F, uoff2 = (q2 >> 3) & 1;
            rev_advance(vlc, len2);
            if (BPW == 1) __syncwarp();
            if (do_store) { sigma1[qx * SS] = (uint8_t)rho1; sigma1[(qx + 1) * SS] = (uint8_t)rho2; }
            if (BPW == 1) __syncwarp();

            uint32_t u0 = 1, u1 = 1;
            uint32_t mode = (uoff1 << 1) | uoff2;
            if (mode > 0) {
                vv = rev_fetch(vlc);
                rev_advance(vlc, uvlc_decode(vv, mode, initial, u0, u1));
            }
            for (int i = 0; i < 4 && qx * 4 + i < w; i++)
                if (rho1 & (1u << i)) {
                    int32_t v = magsgn_sample(ms, u0);
                    if (do_store) out[(size_t)y * ostride + qx * 4 + i] = (OT)v;
                }
            for (int i = 0; i < 4 && (qx + 1) * 4 + i < w; i++)
                if (rho2 & (1u << i)) {
                    int32_t v = magsgn_sample(ms, u1);
                    if (do_store) out[(size_t)y * ostride + (qx + 1) * 4 + i] = (OT)v;
                }
        }
    }
}

}  // namespace

int j2k_htref_map()
{
    static int map = -1;
    if (map < 0) { const char *e = getenv("J2KGPU_HTREF_MAP"); map = (e && atoi(e) == 1) ? 1 : 32; }
    return map;
}

// planes_precleared: the destination was zeroed once and only this decoder writes it (whole-path jobs): clear every 4th row
cudaError_t launch_ht_ref(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          int planes_precleared, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const int map = j2k_htref_map();
    const uint32_t per = map == 32 ? kWarpsPerCta * 32 : kWarpsPerCta, grid = (n + per - 1) / per;
#define J2K_HTREF_GO(BPW, OT, Z) J2K_LAUNCH((k_ht_ref<BPW, OT, Z>), grid, kWarpsPerCta * 32, 0, s, d_cblks, n, d_blob, (OT *)d_coef)
    if (map == 32) {
        if (coef16) { if (planes_precleared) J2K_HTREF_GO(32, int16_t, 4); else J2K_HTREF_GO(32, int16_t, 1); }
        else { if (planes_precleared) J2K_HTREF_GO(32, int32_t, 4); else J2K_HTREF_GO(32, int32_t, 1); }
    } else {
        if (coef16) { if (planes_precleared) J2K_HTREF_GO(1, int16_t, 4); else J2K_HTREF_GO(1, int16_t, 1); }
        else { if (planes_precleared) J2K_HTREF_GO(1, int32_t, 4); else J2K_HTREF_GO(1, int32_t, 1); }
    }
#undef J2K_HTREF_GO
    return cudaGetLastError();
}
