// ht_ref.cu -- "HT" block decoder, REF semantics (sm_100a): a per-thread VLC kernel + a warp-per-block MagSgn kernel.
//
// Replaces entropy.HTDecoder.Decode (reference internal/entropy/ht.go:93-150): MEL init check
// (ht.go:153-195), backward VLC reader (ht.go:276-396), forward MagSgn reader (ht.go:399-519),
// decodeCleanup (ht.go:583-713) and the two UVLC routines (ht.go:716-864).  The reference coder
// is not ISO/IEC 15444-15: only sample row y of each 4-row stripe is produced,
// the MEL stream is never consumed, the VLC length field is read with mask 0x0F.  REF mode
// reproduces exactly that; the conformant decoder is the ISO-mode kernel.
//
// The statement-level restatement of ht.go is the CPU checker (test side); nothing of it lives here.
// Go semantics kept: uint32 shifts >= 32 give 0; the uint32 "bits" counters wrap when a MagSgn
// field is longer than the buffered bits (emb up to 37, ht.go:668-669).
#include "common.h"
#include <cstdlib>

namespace {

#include "ht_vlc_tables.inc"
__device__ const uint16_t d_vlc_tbl0[1024] = HT_VLC_TBL0_INIT;
__device__ const uint16_t d_vlc_tbl1[1024] = HT_VLC_TBL1_INIT;

// UVLC prefix rows, ht.go:718-727: prefix_len | suffix_len << 2 | base << 5
__constant__ uint8_t c_uvlc_dec[8] = {
    3 | (5 << 2) | (5 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5),
    3 | (1 << 2) | (3 << 5), 1 | (1 << 5), 2 | (2 << 5), 1 | (1 << 5)};

__device__ __forceinline__ uint32_t shl32(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }


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

// ---- two-kernel mapping (default) ------------------------------------------------------------------------------------
// The block's chain splits where its data dependencies do.  Both streams of the reference coder are pure functions of
// the block's bytes: every byte contributes 8 bits, or 7 after a 0xFF (MagSgn) / after a byte > 0x8F when its low 7
// bits are all ones (VLC); an exhausted MagSgn stream continues with 0xFF bytes, an exhausted VLC stream with zeros.
// WHEN the reference refills its 64-bit buffers is therefore unobservable, with one exception that real inputs of the
// reference's own encoder do hit: a MagSgn field wider than the 32 bits a fetch guarantees (emb = 33..37 from the
// U-VLC) makes the uint32 bit counter wrap when fewer than emb bits are buffered (ht.go:668-669); from then on the
// decoder never refills and reads zeros.  The buffered amount at a fetch is itself a pure function of the position:
// refills come in 4-byte chunks, so it is (first chunk boundary >= P + 32) - P; the MagSgn kernel keeps the chunk
// boundaries as a second bit string and finds the first wrapping sample of the block with a warp minimum.
//   A  k_htref_vlc     one thread per block: the VLC / U-VLC chain only (context of a quad = 0 or rho1 >> 2, see
//                      above), 128 quad pairs -> 16 bits per quad (rho | emb << 4) in a scratch table, and the
//                      block's status (not coded / decodable / decodable with fields wider than 32 bits);
//   B  k_htref_magsgn  one warp per block: bit length of every quad = popc(rho) * (emb + 1), warp prefix sums give
//                      every quad's position in the MagSgn stream; the warp removes the stuffing once (prefix sum
//                      of the byte widths, bytes OR-ed into a dense bit string in shared memory) and then every lane
//                      extracts the 4 samples of one quad per step: 32 quads = two full sample rows per step,
//                      written as 16-byte stores, zeros included (so no row has to be cleared first).
constexpr int kQuadWords = 128;                         // 16 stripe rows x 8 quad pairs, one uint32 per pair
constexpr int kWarpsB = 4;
constexpr int kStreamWords = (1024 * 38 + 64) / 32 + 2; // 1024 samples x (37 + 1) bits, + the look-ahead windows
enum { ST_ZERO = 0, ST_FAST = 1, ST_WIDE = 2 };

// The VLC stream as the pure function it is (see above): bytes backwards from Lcup - 3, `left` of them, then zeros; no
// inner branches, so the 32 chains of a warp diverge only on whether they refill.
struct VlcStream { const uint8_t *d; int pos, left; uint64_t tmp; uint32_t bits; bool gt8f; };

// `lim` = the blob and its size (0 when the blob is not 4-byte aligned): inside it the four bytes come from two
// aligned 32-bit loads instead of four byte loads
struct BlobLim { const uint8_t *base; uint64_t bytes; };

__device__ __forceinline__ void vlc_read4(VlcStream &v, const BlobLim &lim)
{
    uint32_t b[4], nb[4];
    const uint64_t off = (uint64_t)(v.d + v.pos - 3 - lim.base);
    if (v.left >= 4 && off + 8 <= lim.bytes) {
        const uint32_t *a = reinterpret_cast<const uint32_t *>(lim.base + (off & ~(uint64_t)3));
        const uint32_t x = __funnelshift_r(__ldg(a), __ldg(a + 1), (uint32_t)(off & 3) * 8);      // bytes pos-3 .. pos
        b[0] = x >> 24; b[1] = (x >> 16) & 0xFFu; b[2] = (x >> 8) & 0xFFu; b[3] = x & 0xFFu;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) b[i] = (i < v.left) ? (uint32_t)__ldg(v.d + v.pos - i) : 0u;
    }
    v.pos -= 4; v.left -= 4;
    bool g = v.gt8f;
#pragma unroll
    for (int i = 0; i < 4; i++) { nb[i] = (g && (b[i] & 0x7Fu) == 0x7Fu) ? 7u : 8u; g = b[i] > 0x8Fu; }
    const uint32_t t = b[0] | (b[1] << nb[0]) | (b[2] << (nb[0] + nb[1])) | (b[3] << (nb[0] + nb[1] + nb[2]));
    v.tmp |= (uint64_t)t << v.bits;
    v.bits += nb[0] + nb[1] + nb[2] + nb[3];
    v.gt8f = g;
}

// at least 32 bits afterwards (four stuffed bytes give only 28: then a second read, which is rare)
__device__ __forceinline__ void vlc_refill(VlcStream &v, const BlobLim &lim)
{
    if (v.bits < 32) {
        vlc_read4(v, lim);
        if (v.bits < 32) vlc_read4(v, lim);
    }
}

// U-VLC of a quad pair (ht.go:716-864) from a 192-entry table built at kernel start: kind (0 one code, 1 two codes,
// 2 two codes in the initial row: when the first prefix is the long one the second value is one bit + 2) x the next
// 6 bits -> prefix bits (3) | first suffix length (3) | second suffix length (3) | first base (3) | second base (3).
__device__ __forceinline__ uint16_t uvlc_entry(int kind, uint32_t bits6)
{
    const uint32_t t1 = c_uvlc_dec[bits6 & 7];
    const uint32_t p1 = t1 & 3, s1 = (t1 >> 2) & 7, b1 = t1 >> 5;
    if (kind == 0) return (uint16_t)(p1 | (s1 << 3) | (b1 << 9));
    const uint32_t rest = bits6 >> p1;
    if (kind == 2 && p1 > 2) return (uint16_t)((p1 + 1) | (s1 << 3) | (b1 << 9) | (((rest & 1) + 1) << 12));
    const uint32_t t2 = c_uvlc_dec[rest & 7];
    return (uint16_t)((p1 + (t2 & 3)) | (s1 << 3) | (((t2 >> 2) & 7) << 6) | (b1 << 9) | ((t2 >> 5) << 12));
}

__device__ __forceinline__ uint32_t uvlc_pair(const uint16_t *utab, uint32_t vlc, uint32_t mode, bool initial, uint32_t &u0, uint32_t &u1)
{
    const uint32_t kind = mode < 3 ? 0u : (initial ? 2u : 1u);
    const uint32_t t = utab[kind * 64 + (vlc & 63)];
    const uint32_t pl = t & 7, s1 = (t >> 3) & 7, s2 = (t >> 6) & 7;
    vlc >>= pl;
    const uint32_t ua = ((t >> 9) & 7) + (vlc & ((1u << s1) - 1)) + 1;
    vlc >>= s1;
    const uint32_t ub = (t >> 12) + (vlc & ((1u << s2) - 1)) + 1;
    u0 = (mode == 2) ? 1u : ua;
    u1 = (mode == 1) ? 1u : (mode == 2 ? ua : ub);
    return pl + s1 + s2;
}

__global__ void __launch_bounds__(128)
k_htref_vlc(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob, uint64_t blob_bytes,
            uint32_t *__restrict__ qinfo, uint32_t *__restrict__ status)
{
    const BlobLim lim = {blob, blob_bytes};
    __shared__ uint16_t s_tbl[2048];
    __shared__ uint16_t s_utab[192];
    for (int i = threadIdx.x; i < 1024; i += 128) { s_tbl[i] = d_vlc_tbl0[i]; s_tbl[1024 + i] = d_vlc_tbl1[i]; }
    for (int i = threadIdx.x; i < 192; i += 128) s_utab[i] = uvlc_entry(i >> 6, (uint32_t)i & 63);
    __syncthreads();
    const uint32_t blk = blockIdx.x * 128 + threadIdx.x;
    if (blk >= n) return;
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h, len = (int)cb.data_len;
    const uint8_t *d = blob + cb.data_off;
    int scup = 0;
    bool ok = len >= 2;
    if (ok) {
        scup = (int)byte_at(d, len, len - 1) + (int)((byte_at(d, len, len - 2) & 0x0F) << 8);
        ok = scup >= 2 && scup <= len && mel_init_ok(d, len, len, scup);
    }
    if (!ok) { status[blk] = ST_ZERO; return; }
    VlcStream v;                                                        // initVLC ht.go:276-314
    {
        const uint32_t b = __ldg(d + len - 2);
        v.d = d; v.pos = len - 3; v.left = scup - 2;
        v.tmp = b >> 4;
        v.bits = 4 - (uint32_t)((v.tmp & 7) >> 2);
        v.gt8f = (b | 0x0F) > 0x8F;
    }
    vlc_refill(v, lim);
    uint32_t *qi = qinfo + (size_t)blk * kQuadWords;
    const int quad_cols = (w + 3) >> 2, rows = (h + 3) >> 2;
    uint32_t umax = 0;
    for (int r = 0; r < rows; r++) {
        const bool initial = (r == 0);
        const uint16_t *tbl = s_tbl + (initial ? 0 : 1024);
        for (int qx = 0; qx < quad_cols; qx += 2) {
            vlc_refill(v, lim);                                              // >= 32 bits: both codewords (<= 15 + 15)
            const uint32_t q1 = tbl[(uint32_t)v.tmp & 0x7F];
            const uint32_t rho1 = (q1 >> 4) & 0x0F, l1 = q1 & 0x0F;
            const uint32_t q2 = tbl[((rho1 >> 2) << 7) | ((uint32_t)(v.tmp >> l1) & 0x7F)];
            const uint32_t rho2 = (q2 >> 4) & 0x0F, l12 = l1 + (q2 & 0x0F);
            v.tmp >>= l12; v.bits -= l12;
            uint32_t u0 = 1, u1 = 1;
            const uint32_t mode = (((q1 >> 3) & 1) << 1) | ((q2 >> 3) & 1);
            if (mode > 0) {
                vlc_refill(v, lim);
                const uint32_t c = uvlc_pair(s_utab, (uint32_t)v.tmp, mode, initial, u0, u1);
                v.tmp >>= c; v.bits -= c;
            }
            umax = max(umax, max(u0, u1));
            qi[r * 8 + (qx >> 1)] = rho1 | (u0 << 4) | ((rho2 | (u1 << 4)) << 16);
        }
    }
    status[blk] = (umax > 32 ? ST_WIDE : ST_FAST) | ((uint32_t)(len - scup) << 2);      // + the MagSgn segment length
}

template <typename OT, int ZSTEP>
__global__ void __launch_bounds__(kWarpsB * 32)
k_htref_magsgn(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
               const uint32_t *__restrict__ qinfo, const uint32_t *__restrict__ status, OT *__restrict__ coef)
{
    __shared__ uint32_t s_all[kWarpsB][2][kStreamWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t blk = blockIdx.x * kWarpsB + warp;
    if (blk >= n) return;
    // three independent loads: the block, its status word and (below) its quad table
    const uint32_t stw = status[blk];
    const uint16_t *q16 = reinterpret_cast<const uint16_t *>(qinfo + (size_t)blk * kQuadWords);
    uint32_t qraw[8];
#pragma unroll
    for (int j = 0; j < 8; j++) qraw[j] = q16[j * 32 + lane];           // entries the VLC kernel did not write are masked below
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h;
    OT *out = coef + cb.out_off;
    const uint32_t ostride = cb.out_stride;
    const uint8_t *d = blob + cb.data_off;
    const int st = (int)(stw & 3);
    if (st == ST_ZERO) {
        for (int y = 0; y < h; y += ZSTEP)
            for (int x = lane; x < w; x += 32) out[(size_t)y * ostride + x] = 0;
        return;
    }
    uint32_t *s = s_all[warp][0];                                       // the MagSgn bits
    uint32_t *sb = s_all[warp][1];                                      // bit C set: a 4-byte refill chunk ends at bit C
    const bool wide = (st == ST_WIDE);
    const int quad_cols = (w + 3) >> 2, rows = (h + 3) >> 2;
    const int qx = lane & 15;
    const int ncols = min(4, w - qx * 4);                               // <= 0 for quads right of the block
    const uint32_t colmask = ncols > 0 ? (1u << ncols) - 1u : 0u;

    // ---- position of every quad in the MagSgn stream ----
    uint32_t qv[8], P[8], T = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int row = 2 * j + (lane >> 4);
        const uint32_t q = (row < rows && qx < quad_cols) ? qraw[j] : 0u;
        const uint32_t rho = q & colmask, emb = q >> 4;
        const uint32_t bits = (uint32_t)__popc(rho) * (emb + 1);
        uint32_t incl = bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        P[j] = T + incl - bits;
        T += __shfl_sync(0xffffffffu, incl, 31);
        qv[j] = rho | (emb << 4);
    }

    // ---- the MagSgn bytes (then 0xFF for ever) without their stuffing, as one dense little-endian bit string ----
    const uint32_t need = T + 64;                                       // + the widest look-ahead (32-bit field, 5-bit window)
    const int nW = (int)(need >> 5) + 2;
    for (int i = lane; i < nW; i += 32) { s[i] = 0; if (wide) sb[i] = 0; }
    __syncwarp();
    const int L = (int)(stw >> 2);
    uint32_t base = 0, prev_ff = 0;
    uint32_t bn[4];                                                     // the next chunk's bytes, loaded one iteration early
#pragma unroll
    for (int i = 0; i < 4; i++) bn[i] = (4 * lane + i < L) ? (uint32_t)__ldg(d + 4 * lane + i) : 0xFFu;
    for (int k0 = 0; base < need; k0 += 128) {
        uint32_t b[4], nb[4];
        const int kn = k0 + 128 + 4 * lane;
#pragma unroll
        for (int i = 0; i < 4; i++) { b[i] = bn[i]; bn[i] = (kn + i < L) ? (uint32_t)__ldg(d + kn + i) : 0xFFu; }
        uint32_t pb = __shfl_up_sync(0xffffffffu, b[3], 1);
        if (lane == 0) pb = prev_ff ? 0xFFu : 0u;
#pragma unroll
        for (int i = 0; i < 4; i++) { nb[i] = pb == 0xFFu ? 7u : 8u; pb = b[i]; }
        const uint32_t v = b[0] | (b[1] << nb[0]) | (b[2] << (nb[0] + nb[1])) | (b[3] << (nb[0] + nb[1] + nb[2]));
        const uint32_t tot = nb[0] + nb[1] + nb[2] + nb[3];
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const uint32_t pos = base + incl - tot;
        const int wi = (int)(pos >> 5);
        const uint32_t sh = pos & 31;
        if (wi < nW) {
            atomicOr(&s[wi], v << sh);
            const uint32_t hi = sh ? v >> (32 - sh) : 0u;
            if (hi && wi + 1 < nW) atomicOr(&s[wi + 1], hi);
        }
        // the decoder starts with two chunks buffered (ht.go:399-429): the end of the first one never limits a fetch
        const uint32_t cend = pos + tot;
        if (wide && (k0 | lane) != 0 && (int)(cend >> 5) < nW) atomicOr(&sb[cend >> 5], 1u << (cend & 31));
        base += __shfl_sync(0xffffffffu, incl, 31);
        prev_ff = __shfl_sync(0xffffffffu, pb, 31) == 0xFFu;
    }
    __syncwarp();

    // ---- first sample whose field is wider than the buffered bits (decode order = quad order, then sample order) ----
    uint32_t wrap_at = 0xFFFFFFFFu;
    if (wide) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t rho = qv[j] & 15, emb = qv[j] >> 4;
            if (emb < 33 || rho == 0) continue;
            uint32_t p = P[j];
            for (int i = 0; i < 4; i++)
                if ((rho >> i) & 1) {
                    const uint32_t a = p + 32;
                    const uint32_t win = __funnelshift_r(sb[a >> 5], sb[(a >> 5) + 1], a & 31) & 0x1Fu;
                    if (win && emb > 31u + (uint32_t)__ffs((int)win)) { wrap_at = min(wrap_at, (uint32_t)((j * 32 + lane) * 4 + i)); break; }
                    p += emb + 1;
                }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) wrap_at = min(wrap_at, __shfl_xor_sync(0xffffffffu, wrap_at, o));
    }

    // ---- 32 quads (two sample rows) per step ----
    const bool vec_ok = ((cb.out_off | ostride) & 3) == 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int row = 2 * j + (lane >> 4);
        if (row >= rows || ncols <= 0) continue;
        const uint32_t rho = qv[j] & 15, emb = qv[j] >> 4;
        uint32_t p = P[j];
        int32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            v[i] = 0;
            if ((rho >> i) & 1) {
                const uint32_t idx = (uint32_t)((j * 32 + lane) * 4 + i);
                uint32_t x = __funnelshift_r(s[p >> 5], s[(p >> 5) + 1], p & 31);
                if (idx > wrap_at) x = 0;                               // after the wrap the decoder reads zeros
                const uint32_t m = (emb >= 32 ? x : (x & ((1u << emb) - 1u))) + shl32(1u, emb - 1);
                p += emb;
                uint32_t sign = (s[p >> 5] >> (p & 31)) & 1u;
                if (idx >= wrap_at) sign = 0;
                p++;
                v[i] = (int32_t)(sign ? 0u - m : m);
            }
        }
        const int y = 4 * row;
        OT *o = out + (size_t)y * ostride + qx * 4;
        if (ncols == 4 && vec_ok) {
            if (sizeof(OT) == 4) *reinterpret_cast<int4 *>(o) = make_int4(v[0], v[1], v[2], v[3]);
            else *reinterpret_cast<uint2 *>(o) = make_uint2(((uint32_t)v[0] & 0xFFFFu) | ((uint32_t)v[1] << 16), ((uint32_t)v[2] & 0xFFFFu) | ((uint32_t)v[3] << 16));
            if (ZSTEP == 1)
                for (int yy = y + 1; yy < y + 4 && yy < h; yy++) {
                    OT *z = out + (size_t)yy * ostride + qx * 4;
                    if (sizeof(OT) == 4) *reinterpret_cast<int4 *>(z) = make_int4(0, 0, 0, 0);
                    else *reinterpret_cast<uint2 *>(z) = make_uint2(0u, 0u);
                }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < ncols) {
                    o[i] = (OT)v[i];
                    if (ZSTEP == 1)
                        for (int yy = y + 1; yy < y + 4 && yy < h; yy++) out[(size_t)yy * ostride + qx * 4 + i] = 0;
                }
        }
    }
}

}  // namespace

size_t j2k_htref_scratch_bytes(uint32_t n) { return (size_t)n * (kQuadWords * 4 + 4) + 16; }
int j2k_htref_launches() { return 2; }

// planes_precleared: the destination was zeroed once and only this decoder writes it (whole-path jobs): clear every 4th row
// d_scratch: j2k_htref_scratch_bytes(n) bytes of device memory (the quad table between the two kernels)
cudaError_t launch_ht_ref(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          int planes_precleared, void *d_scratch, uint64_t blob_bytes, cudaStream_t s)
{
    if ((uintptr_t)d_blob & 3) blob_bytes = 0;           // unaligned blob: byte loads only
    if (n == 0) return cudaSuccess;
    uint32_t *qinfo = (uint32_t *)d_scratch;
    uint32_t *status = qinfo + (size_t)n * kQuadWords;
    J2K_LAUNCH((k_htref_vlc), (n + 127) / 128, 128, 0, s, d_cblks, n, d_blob, blob_bytes, qinfo, status);
    const uint32_t grid = (n + kWarpsB - 1) / kWarpsB;
#define J2K_HTREF_B(OT, Z) J2K_LAUNCH((k_htref_magsgn<OT, Z>), grid, kWarpsB * 32, 0, s, d_cblks, n, d_blob, qinfo, status, (OT *)d_coef)
    if (coef16) { if (planes_precleared) J2K_HTREF_B(int16_t, 4); else J2K_HTREF_B(int16_t, 1); }
    else { if (planes_precleared) J2K_HTREF_B(int32_t, 4); else J2K_HTREF_B(int32_t, 1); }
#undef J2K_HTREF_B
    return cudaGetLastError();
}
