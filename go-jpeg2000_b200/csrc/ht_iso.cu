// ht_iso.cu -- ISO/IEC 15444-15 (ITU-T T.814) HT cleanup-pass block decoder for J2KGPU_MODE_ISO (sm_100a).
//
// What north_star calls "HTJ2K cleanup decoding (MEL, VLC and MagSgn bitstreams)".  The reference's ht.go is
// not conformant and is covered by ht_ref.cu; this kernel follows the published algorithm
// (T.814 clause 7): MagSgn forward from byte 0, MEL forward from Lcup-Scup, VLC backward from Lcup-2, 2x2 quads
// in pairs, CxtVLC tables, U-VLC, exponent predictor from the previous quad row.  Checked in the test suite against
// a CPU statement of the same algorithm that is itself pinned by OpenJPEG decoding the same streams.
//
// Mapping: two kernels, k_htiso_vlc (one thread per block: the MEL / VLC context chain) and k_htiso_magsgn2 (half a
// warp per block: the MagSgn rows in parallel).  The single-chain statement of the same algorithm is the CPU checker
// (test side).
#include "common.h"
#include <cstdlib>

namespace {

#include "ht_vlc_tables.inc"
__device__ const uint16_t d_tbl0[1024] = HT_VLC_TBL0_INIT;
__device__ const uint16_t d_tbl1[1024] = HT_VLC_TBL1_INIT;

constexpr int kThreads = 128;

struct Mel {
    int pos, left; uint32_t tmp; int bits; bool unstuff;
    int k, zeros; bool one_after;
};

__device__ __forceinline__ int mel_exp(int k)
{
    // {0,0,0,1,1,1,2,2,2,3,3,4,5} packed 4 bits each
    return (int)((0x5433222111000ull >> (4 * k)) & 0xF);
}

__device__ __forceinline__ int mel_bit(Mel &m, const uint8_t *d)
{
    if (m.bits == 0) {
        uint32_t b = 0xFF;
        if (m.left > 0) {
            b = __ldg(d + m.pos); m.pos++; m.left--;
            if (m.left == 0) b |= 0x0F;                 // last byte is shared with the VLC stream
        }
        m.bits = m.unstuff ? 7 : 8;
        m.tmp = m.unstuff ? (b & 0x7F) : b;
        m.unstuff = (b == 0xFF);
    }
    m.bits--;
    return (int)((m.tmp >> m.bits) & 1);
}

__device__ __forceinline__ int mel_event(Mel &m, const uint8_t *d)
{
    if (m.zeros == 0 && !m.one_after) {
        const int e = mel_exp(m.k);
        if (mel_bit(m, d)) {
            m.zeros = 1 << e;
            if (m.k < 12) m.k++;
        } else {
            int r = 0;
            for (int i = 0; i < e; i++) r = (r << 1) | mel_bit(m, d);
            m.zeros = r; m.one_after = true;
            if (m.k > 0) m.k--;
        }
    }
    if (m.zeros > 0) { m.zeros--; return 0; }
    m.one_after = false;
    return 1;
}

// U-VLC prefix rows (T.814 Table 3): prefix_len | suffix_len << 2 | base << 5, indexed by the 3 LSBs
__device__ __forceinline__ uint32_t uvlc_row(uint32_t b3)
{
    // {183, 33, 66, 33, 103, 33, 66, 33}
    return (uint32_t)((0x2142216721422100ull | 0xB7ull) >> (8 * b3)) & 0xFF;
}

// value written for one decoded sample: reversible -> integer; irreversible -> dequantised float bits
__device__ __forceinline__ int32_t sample_value(uint32_t mu, uint32_t sign, int shift, float step, bool irrev)
{
    if (!irrev) {
        const uint32_t mag = mu << shift;
        return (int32_t)(sign ? 0u - mag : mag);
    }
    // mid-point reconstruction: (mu + 1/2) * 2^shift * step
    const float f = ((float)mu + 0.5f) * (float)(1u << shift) * step;
    return __float_as_int(sign ? -f : f);
}

// ---- two-kernel mapping (default) ------------------------------------------------------------------------------------
// A block's chain splits where its data dependencies do (the same cut as in ht_ref.cu):
//   A  k_htiso_vlc     one thread per block: MEL + CxtVLC + U-VLC only -- the context chain through the previous quad and
//                      the previous quad row.  Output: 16 bits per quad in a scratch table, row-major 32 quads per quad
//                      row: 2 bits per sample (0 insignificant, 1 significant, 2 + known EMB bit = 0, 3 + EMB bit = 1;
//                      the CxtVLC tables satisfy e_1 <= e_k <= rho, so this loses nothing) and u_q << 8.
//   B  k_htiso_magsgn  one warp per block, one lane per quad of a quad row: the exponent predictor needs the previous
//                      row's exponents (two shuffles), then U_q, the four field widths, a warp prefix sum for the
//                      position of every quad in the MagSgn stream, and the extraction of the row's 128 samples in
//                      parallel; rows are written as contiguous 8-byte stores.  The stuffing is removed by the warp in
//                      128-byte chunks (prefix sum of byte widths, bytes OR-ed into a dense bit string) into a 16 Kbit
//                      shared-memory ring that runs at least one quad row (32 x 4 x 31 bits) ahead of the decoder.
// The streams are pure functions of the block's bytes (see ht_ref.cu), so the result equals the single-chain decoder's
// for every input, malformed ones included (same `bad` conditions, same zero block).
constexpr int kQTabWords = 512;                         // 32 quad rows x 32 quads x 16 bits
constexpr int kRingWords = 512;
enum { ST_ZERO = 0, ST_OK = 1 };

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

// 4 bits -> the even bit positions of a byte
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return (x & 1) | ((x & 2) << 1) | ((x & 4) << 2) | ((x & 8) << 3); }

// U-VLC of a quad pair from a 192-entry table built at kernel start: kind (0 one code, 1 two codes, 2 two codes in the
// initial row: when the first prefix is the long one the second code is a single bit) x the next 6 bits ->
// prefix bits (3) | first suffix length (3) | second suffix length (3) | first base (3) | second base (3).
__device__ __forceinline__ uint16_t uvlc_entry(int kind, uint32_t bits6)
{
    const uint32_t t1 = uvlc_row(bits6 & 7);
    const uint32_t p1 = t1 & 3, s1 = (t1 >> 2) & 7, b1 = t1 >> 5;
    if (kind == 0) return (uint16_t)(p1 | (s1 << 3) | (b1 << 9));
    const uint32_t rest = bits6 >> p1;
    if (kind == 2 && p1 > 2) return (uint16_t)((p1 + 1) | (s1 << 3) | (b1 << 9) | (((rest & 1) + 1) << 12));
    const uint32_t t2 = uvlc_row(rest & 7);
    return (uint16_t)((p1 + (t2 & 3)) | (s1 << 3) | (((t2 >> 2) & 7) << 6) | (b1 << 9) | ((t2 >> 5) << 12));
}

__device__ __forceinline__ int uvlc_pair(const uint16_t *utab, uint32_t vlc, int mode, bool initial, int &u0, int &u1)
{
    const int kind = mode < 3 ? 0 : ((mode == 3 && initial) ? 2 : 1);
    const uint32_t t = utab[kind * 64 + (vlc & 63)];
    const int pl = t & 7, s1 = (t >> 3) & 7, s2 = (t >> 6) & 7;
    vlc >>= pl;
    const int ua = (int)(((t >> 9) & 7) + (vlc & ((1u << s1) - 1)));
    vlc >>= s1;
    const int ub = (int)((t >> 12) + (vlc & ((1u << s2) - 1)));
    const int add = (mode == 4) ? 2 : 0;
    u0 = (mode == 2) ? 0 : ua + add;
    u1 = (mode == 1) ? 0 : (mode == 2 ? ua : ub + add);
    return pl + s1 + s2;
}

__global__ void __launch_bounds__(kThreads)
k_htiso_vlc(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob, uint64_t blob_bytes,
            uint32_t *__restrict__ qtab, uint32_t *__restrict__ status)
{
    const BlobLim lim = {blob, blob_bytes};
    // table entries re-packed: len (3) | u_off (1) | rho (4) | sample states (8)
    __shared__ uint16_t s_tbl[2048];
    for (int i = threadIdx.x; i < 2048; i += kThreads) {
        const uint32_t e = i < 1024 ? d_tbl0[i] : d_tbl1[i - 1024];
        const uint32_t st8 = spread4((e >> 4) & 15) + spread4((e >> 12) & 15) + spread4((e >> 8) & 15);
        s_tbl[i] = (uint16_t)((e & 0xFF) | (st8 << 8));
    }
    __shared__ uint16_t s_utab[192];
    for (int i = threadIdx.x; i < 192; i += kThreads) s_utab[i] = uvlc_entry(i >> 6, (uint32_t)i & 63);
    __syncthreads();
    const uint32_t blk = blockIdx.x * kThreads + threadIdx.x;
    if (blk >= n) return;
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h;
    const uint8_t *d = blob + cb.data_off;
    const int lcup = (int)cb.data_len;
    bool ok = lcup >= 2 && cb.num_bps >= 1 && cb.num_bps <= 30;
    int scup = 0;
    if (ok) {
        scup = ((int)__ldg(d + lcup - 1) << 4) + (int)(__ldg(d + lcup - 2) & 0x0F);
        ok = scup >= 2 && scup <= lcup && scup <= 4079;
    }
    if (!ok) { status[blk] = ST_ZERO; return; }
    Mel mel; mel.pos = lcup - scup; mel.left = scup - 1; mel.tmp = 0; mel.bits = 0; mel.unstuff = false;
    mel.k = 0; mel.zeros = 0; mel.one_after = false;
    VlcStream v;
    {
        const uint32_t b = __ldg(d + lcup - 2);
        v.d = d; v.pos = lcup - 3; v.left = scup - 2;
        v.tmp = b >> 4;
        v.bits = 4 - (((v.tmp & 7) == 7) ? 1 : 0);
        v.gt8f = (b | 0x0F) > 0x8F;
    }
    uint32_t *qt = qtab + (size_t)blk * kQTabWords;
    const int nq = (w + 1) >> 1;
    uint64_t sigprev = 0;                                // bottom-row significance of the previous quad row, bit c = column c
    uint32_t badbits = 0;
    for (int y = 0; y < h; y += 2) {
        const bool initial = (y == 0);
        const uint16_t *tbl = s_tbl + (initial ? 0 : 1024);
        const uint32_t rowmask = (y + 1 < h) ? 0u : 0xA0u;              // rho bits of samples below the block
        uint64_t sp = sigprev;                           // bit 0 = column 2q of the current pair
        uint32_t spc = 0;                                // column 2q - 1
        uint64_t sn = 0;                                 // new significance, shifted in from the top, 4 columns per pair
        int cw = 0, iters = 0;
        uint32_t *qrow = qt + (y >> 1) * 16;
        for (int q = 0; q < nq; q += 2, iters++) {
            const bool pair = (q + 1 < nq);
            vlc_refill(v, lim);                               // >= 32 bits: two codewords (<= 7 each) and the U-VLC (<= 16)
            uint32_t win = (uint32_t)v.tmp, used = 0;    // the pair decodes from one 32-bit window; one 64-bit shift at the end
            const uint32_t s6 = (((uint32_t)sp & 0x1F) << 1) | spc;     // columns 2q - 1 .. 2q + 4 (zero in the initial row)
            spc = ((uint32_t)sp >> 3) & 1;
            sp >>= 4;
            uint32_t e0, e1 = 0;
            {
                const int c_q = cw | (int)((s6 & 3) != 0) | ((int)((s6 & 0xC) != 0) << 2);
                uint32_t e = tbl[(c_q << 7) | (win & 0x7F)];
                if (c_q == 0 && !mel_event(mel, d)) e = 0;
                win >>= (e & 7); used += (e & 7);
                e0 = e;
                const uint32_t rho = (e >> 4) & 0xF;
                cw = initial ? (int)(((rho & 3) != 0) | (((rho >> 2) & 3) << 1))
                             : (int)((rho >> 2) != 0) << 1;
            }
            if (pair) {
                const int c_q = cw | (int)((s6 & 0xC) != 0) | ((int)((s6 & 0x30) != 0) << 2);
                uint32_t e = tbl[(c_q << 7) | (win & 0x7F)];
                if (c_q == 0 && !mel_event(mel, d)) e = 0;
                win >>= (e & 7); used += (e & 7);
                e1 = e;
                const uint32_t rho = (e >> 4) & 0xF;
                cw = initial ? (int)(((rho & 3) != 0) | (((rho >> 2) & 3) << 1))
                             : (int)((rho >> 2) != 0) << 1;
            }
            int mode = (int)(((e0 >> 3) & 1) | (((e1 >> 3) & 1) << 1));
            if (initial && mode == 3 && mel_event(mel, d)) mode = 4;
            int u0 = 0, u1 = 0;
            if (mode) used += (uint32_t)uvlc_pair(s_utab, win, mode, initial, u0, u1);
            v.tmp >>= used; v.bits -= used;
            // bottom samples (rho bits 1 and 3) of both quads: columns 2q .. 2q + 3
            const uint32_t nb4 = ((e0 >> 5) & 1) | ((e0 >> 6) & 2) | ((e1 >> 3) & 4) | ((e1 >> 4) & 8);
            sn = (sn >> 4) | ((uint64_t)nb4 << 60);
            // significance outside the block: malformed, the block is zero (only the last row / last pair can have it)
            const int xq = 2 * q;
            if (rowmask | (uint32_t)(xq + 3 >= w)) {
                badbits |= e0 & (rowmask | (xq + 1 >= w ? 0xC0u : 0u));
                badbits |= e1 & (rowmask | (xq + 3 >= w ? 0xC0u : 0u));
            }
            // u >= 31 makes U_q > 31 whatever the predictor says: malformed either way, 5 bits are enough
            qrow[q >> 1] = (e0 >> 8) | ((uint32_t)min(u0, 31) << 8) | (((e1 >> 8) | ((uint32_t)min(u1, 31) << 8)) << 16);
        }
        sigprev = sn >> (64 - 4 * iters);
    }
    status[blk] = badbits ? (uint32_t)ST_ZERO : ((uint32_t)ST_OK | ((uint32_t)(lcup - scup) << 2));
}

// B', the default: TWO blocks per warp, one per half-warp, each lane two neighbouring quads (4 columns) of a quad row.
// Kernel B spends about as many instructions per quad row on what is per row (loop, ring upkeep, predictor shuffles,
// prefix sum) as on the samples; with two quads per lane and two blocks per instruction stream that fixed part is
// shared by four times as many samples.  Same arithmetic, same ring (one per block, 128-byte chunks), same results.
constexpr int kWarpsIsoB2 = 8;

template <typename OT, bool IRREV>
__global__ void __launch_bounds__(kWarpsIsoB2 * 32)
k_htiso_magsgn2(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
                const uint32_t *__restrict__ qtab, const uint32_t *__restrict__ status, OT *__restrict__ coef,
                const float *__restrict__ steps, int coef_bits)
{
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ uint32_t s_ring[kWarpsIsoB2 * 2][kRingWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, sl = lane & 15;
    const uint32_t blk0 = (blockIdx.x * kWarpsIsoB2 + warp) * 2;
    if (blk0 >= n) return;
    const bool have = blk0 + half < n;
    const uint32_t blk = have ? blk0 + half : blk0;
    const uint32_t stw = status[blk];
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h;
    OT *out = coef + cb.out_off;
    const size_t ostride = cb.out_stride;
    const uint8_t *d = blob + cb.data_off;
    const bool zero = (stw & 3) == ST_ZERO;
    if (have && zero)
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += 16) out[(size_t)y * ostride + x] = 0;
    const bool live = have && !zero;                     // this half-warp decodes a block
    const float step = (IRREV && steps) ? steps[blk] : 1.0f;
    const int L = (int)(stw >> 2);
    const int shift = cb.num_bps - 1;
    const int nq = (w + 1) >> 1, nrows = live ? (h + 1) >> 1 : 0;
    const bool active = live && 2 * sl < nq;             // the lane's first quad exists
    const int ncols = w - 4 * sl;                        // columns of the block right of (and including) the lane's first
    const bool vec_ok = ((cb.out_off | ostride) & 3) == 0;
    const uint32_t *qt = qtab + (size_t)blk * kQTabWords;
    uint32_t *ring = s_ring[warp * 2 + half];
    for (int i = sl; i < kRingWords; i += 16) ring[i] = 0;
    __syncwarp();
    uint32_t built = 0, prev_ff = 0, P = 0;              // uniform per half-warp
    int kbyte = 0;
    int Eb[4] = {0, 0, 0, 0};                            // bottom-sample exponents of the lane's 4 columns, previous quad row
    bool bad = false;
    const int nrows_max = max(__shfl_sync(FULL, nrows, 0), __shfl_sync(FULL, nrows, 16));
    uint32_t code_next = (active && nrows > 0) ? qt[sl] : 0u;
    for (int r = 0; r < nrows_max; r++) {
        const bool row_on = r < nrows;
        const uint32_t code2 = row_on ? code_next : 0u;
        if (r + 1 < nrows) code_next = active ? qt[(r + 1) * 16 + sl] : 0u;
        // ---- keep each ring one full quad row ahead ----
        bool need = row_on && built < P + 4096u;
        while (__any_sync(FULL, need)) {
            const int k = kbyte + 8 * sl;
            uint32_t b[8], nb[8];
#pragma unroll
            for (int i = 0; i < 8; i++) b[i] = (need && k + i < L) ? (uint32_t)__ldg(d + k + i) : 0xFFu;
            if (need) {                                  // words beyond the one `built` points into hold bits 16 Kbit old
                const uint32_t w0 = (built >> 5) + 1;
                ring[(w0 + sl) & (kRingWords - 1)] = 0;
                ring[(w0 + 16 + sl) & (kRingWords - 1)] = 0;
                if (sl < 2) ring[(w0 + 32 + sl) & (kRingWords - 1)] = 0;
            }
            uint32_t pb = __shfl_up_sync(FULL, b[7], 1, 16);
            if (sl == 0) pb = prev_ff ? 0xFFu : 0u;
#pragma unroll
            for (int i = 0; i < 8; i++) { nb[i] = pb == 0xFFu ? 7u : 8u; pb = b[i]; }
            const uint32_t va = b[0] | (b[1] << nb[0]) | (b[2] << (nb[0] + nb[1])) | (b[3] << (nb[0] + nb[1] + nb[2]));
            const uint32_t vb = b[4] | (b[5] << nb[4]) | (b[6] << (nb[4] + nb[5])) | (b[7] << (nb[4] + nb[5] + nb[6]));
            const uint32_t ta = nb[0] + nb[1] + nb[2] + nb[3], tot = ta + nb[4] + nb[5] + nb[6] + nb[7];
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o, 16); if (sl >= o) incl += t; }
            const uint32_t pos = built + incl - tot;
            __syncwarp();
            if (need) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t pg = g ? pos + ta : pos, vg = g ? vb : va;
                    const uint32_t sh = pg & 31, wi = pg >> 5;
                    atomicOr(&ring[wi & (kRingWords - 1)], vg << sh);
                    const uint32_t hi = sh ? vg >> (32 - sh) : 0u;
                    if (hi) atomicOr(&ring[(wi + 1) & (kRingWords - 1)], hi);
                }
            }
            const uint32_t total = __shfl_sync(FULL, incl, 15, 16);
            const uint32_t lastb = __shfl_sync(FULL, pb, 15, 16);
            if (need) { built += total; prev_ff = lastb == 0xFFu; kbyte += 128; }
            __syncwarp();
            need = row_on && built < P + 4096u;
        }
        // ---- U_q and the field widths of the lane's two quads ----
        const int eL = __shfl_up_sync(FULL, Eb[3], 1, 16), eR = __shfl_down_sync(FULL, Eb[0], 1, 16);
        const int eLeft = sl ? eL : 0, eRight = sl < 15 ? eR : 0;
        uint32_t st8[2];
        int U[2], m[8];
#pragma unroll
        for (int qd = 0; qd < 2; qd++) {
            const uint32_t code = (code2 >> (16 * qd)) & 0xFFFFu;
            st8[qd] = code & 0xFF;
            const int u = (int)(code >> 8);
            const uint32_t sig = (st8[qd] | (st8[qd] >> 1)) & 0x55u;
            int Uq = u + 1;
            if (r > 0 && (sig & (sig - 1))) {
                const int E = qd == 0 ? max(max(eLeft, Eb[0]), max(Eb[1], Eb[2])) : max(max(Eb[1], Eb[2]), max(Eb[3], eRight));
                Uq = u + max(1, E - 1);
            }
            if (Uq > 31) bad = true;
            if (coef_bits && Uq + shift > coef_bits + 1) bad = true;
            U[qd] = min(Uq, 31);
#pragma unroll
            for (int i = 0; i < 4; i++) { const uint32_t s2 = (st8[qd] >> (2 * i)) & 3; m[4 * qd + i] = s2 ? U[qd] - (int)(s2 >> 1) : 0; }
        }
        const uint32_t tot = (uint32_t)(m[0] + m[1] + m[2] + m[3] + m[4] + m[5] + m[6] + m[7]);
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o, 16); if (sl >= o) incl += t; }
        uint32_t p = P + incl - tot;
        P += __shfl_sync(FULL, incl, 15, 16);
        // ---- samples: per quad n = 0 (y, x), 1 (y + 1, x), 2 (y, x + 1), 3 (y + 1, x + 1) ----
        int32_t val[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t s2 = (st8[i >> 2] >> (2 * (i & 3))) & 3;
            val[i] = 0;
            if (i & 1) Eb[i >> 1] = 0;
            if (s2) {
                const uint32_t wi = p >> 5;
                const uint32_t x = __funnelshift_r(ring[wi & (kRingWords - 1)], ring[(wi + 1) & (kRingWords - 1)], p & 31);
                uint32_t vv = x & ((1u << m[i]) - 1u);
                const uint32_t sign = vv & 1;
                vv |= (uint32_t)(s2 == 3) << m[i];
                vv |= 1;
                val[i] = sample_value((vv >> 1) + 1, sign, shift, step, IRREV);
                if (i & 1) Eb[i >> 1] = 32 - __clz((int)vv);
                p += (uint32_t)m[i];
            }
        }
        if (active && row_on) {
            const int y = 2 * r;
            const bool row2 = (y + 1 < h);
            OT *p0 = out + (size_t)y * ostride + 4 * sl;
            if (ncols >= 4 && vec_ok) {
                if (sizeof(OT) == 4) {
                    *reinterpret_cast<int4 *>(p0) = make_int4(val[0], val[2], val[4], val[6]);
                    if (row2) *reinterpret_cast<int4 *>(p0 + ostride) = make_int4(val[1], val[3], val[5], val[7]);
                } else {
                    *reinterpret_cast<uint2 *>(p0) = make_uint2(((uint32_t)val[0] & 0xFFFFu) | ((uint32_t)val[2] << 16),
                                                                ((uint32_t)val[4] & 0xFFFFu) | ((uint32_t)val[6] << 16));
                    if (row2) *reinterpret_cast<uint2 *>(p0 + ostride) = make_uint2(((uint32_t)val[1] & 0xFFFFu) | ((uint32_t)val[3] << 16),
                                                                                     ((uint32_t)val[5] & 0xFFFFu) | ((uint32_t)val[7] << 16));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < ncols) {
                        p0[c] = (OT)val[2 * c];
                        if (row2) p0[ostride + c] = (OT)val[2 * c + 1];
                    }
            }
        }
    }
    // a malformed block is zero as a whole
    bad = bad && live;
    const uint32_t badm = __ballot_sync(FULL, bad);
    if (have && !zero && (badm & (half ? 0xFFFF0000u : 0x0000FFFFu))) {
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += 16) out[(size_t)y * ostride + x] = 0;
    }
}

}  // namespace

size_t j2k_htiso_scratch_bytes(uint32_t n) { return (size_t)n * (kQTabWords * 4 + 4) + 16; }
int j2k_htiso_launches() { return 2; }

template <typename OT>
static void launch_ht_iso_t(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, OT *d_coef,
                            const float *d_steps, int irrev, int coef_bits, void *d_scratch, uint64_t blob_bytes, cudaStream_t s)
{
    if ((uintptr_t)d_blob & 3) blob_bytes = 0;           // unaligned blob: byte loads only
    uint32_t *qtab = (uint32_t *)d_scratch, *status = qtab + (size_t)n * kQTabWords;
    J2K_LAUNCH((k_htiso_vlc), (n + kThreads - 1) / kThreads, kThreads, 0, s, d_cblks, n, d_blob, blob_bytes, qtab, status);
    const uint32_t grid = (n + 2 * kWarpsIsoB2 - 1) / (2 * kWarpsIsoB2);
    if (irrev) J2K_LAUNCH((k_htiso_magsgn2<OT, true>), grid, kWarpsIsoB2 * 32, 0, s, d_cblks, n, d_blob, qtab, status, d_coef, d_steps, coef_bits);
    else J2K_LAUNCH((k_htiso_magsgn2<OT, false>), grid, kWarpsIsoB2 * 32, 0, s, d_cblks, n, d_blob, qtab, status, d_coef, d_steps, coef_bits);
}

// d_scratch: j2k_htiso_scratch_bytes(n) bytes of device memory (quad table + status between the two kernels)
cudaError_t launch_ht_iso(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          const float *d_steps, int irrev, int coef_bits, void *d_scratch, uint64_t blob_bytes,
                          cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    if (coef16 && !irrev) launch_ht_iso_t<int16_t>(d_cblks, n, d_blob, (int16_t *)d_coef, d_steps, irrev, coef_bits, d_scratch, blob_bytes, s);
    else launch_ht_iso_t<int32_t>(d_cblks, n, d_blob, (int32_t *)d_coef, d_steps, irrev, coef_bits, d_scratch, blob_bytes, s);
    return cudaGetLastError();
}
