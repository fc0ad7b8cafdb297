// ht_iso.cu -- ISO/IEC 15444-15 (ITU-T T.814) HT block decoder for J2KGPU_MODE_ISO (sm_100a): cleanup, SigProp, MagRef.
//
// What north_star calls "HTJ2K cleanup/SigProp/MagRef decoding (MEL, VLC and MagSgn bitstreams)".  The reference's ht.go is
// not conformant and is covered by ht_ref.cu; this kernel follows the published algorithm
// (T.814 clause 7): MagSgn forward from byte 0, MEL forward from Lcup-Scup, VLC backward from Lcup-2, 2x2 quads
// in pairs, CxtVLC tables, U-VLC, exponent predictor from the previous quad row.  Checked in the test suite against
// a CPU statement of the same algorithm that is itself pinned by OpenJPEG decoding the same streams.
//
// Mapping: k_htiso_vlc (one thread per block: the MEL / VLC context chain), then -- only when some block of the launch
// carries SigProp / MagRef passes -- k_htiso_refine (one warp per block: significance propagation and the MagRef bits,
// left as three row bitmaps per block), then k_htiso_magsgn4 (eight lanes per block: the MagSgn rows in parallel and
// the final value of every sample, written once).  The single-chain statement of the same algorithm is the CPU checker
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

// Value written for one decoded sample.  q = magnitude in QUARTER units (integer LSB = bit 2) including the mid-point
// bit below the last decoded bit-plane, the reconstruction OpenJPEG uses (the pin of the CPU checker):
// cleanup only (2 mu + 1) << (P + 1); after MagRef mu << (P + 2) | bit << (P + 1) | 1 << P; new in SigProp 3 << P.
// Reversible -> sign * (q >> 2); irreversible -> dequantised float bits sign * q * step / 4.  uint32 arithmetic throughout.
__device__ __forceinline__ int32_t sample_value(uint32_t q, uint32_t sign, float qstep, bool irrev)
{
    if (!irrev) {
        const uint32_t mag = q >> 2;
        return (int32_t)(sign ? 0u - mag : mag);
    }
    const float f = (float)(int32_t)(sign ? 0u - q : q) * qstep;          // qstep = 0.25 * step
    return __float_as_int(f);
}

// ---- two-kernel mapping (default) ------------------------------------------------------------------------------------
// A block's chain splits where its data dependencies do (the same cut as in ht_ref.cu):
//   A  k_htiso_vlc     one thread per block: MEL + CxtVLC + U-VLC only -- the context chain through the previous quad and
//                      the previous quad row.  Output: 16 bits per quad in a scratch table, row-major 32 quads per quad
//                      row: 2 bits per sample (0 insignificant, 1 significant, 2 + known EMB bit = 0, 3 + EMB bit = 1;
//                      the CxtVLC tables satisfy e_1 <= e_k <= rho, so this loses nothing) and u_q << 8.
//   B  k_htiso_magsgn4 eight lanes per block, four blocks per warp (below): exponent predictor from the previous quad row,
//                      U_q, field widths, prefix sum for the positions in the MagSgn stream, the samples of a quad row in
//                      parallel out of a shared-memory ring that holds the stream with its stuffing removed.
// The streams are pure functions of the block's bytes (see ht_ref.cu), so the result equals the single-chain decoder's
// for every input, malformed ones included (same `bad` conditions, same zero block).
constexpr int kQTabWords = 512;                         // 32 quad rows x 32 quads x 16 bits
constexpr int kRingWords = 512;
enum { ST_ZERO = 0, ST_OK = 1 };

// The VLC stream is read backwards, four bytes at a time, from aligned words of the blob that travel through a thread-private
// ring of eight words in shared memory: a read takes the word below `hi` (the word of the next byte to read) from the ring
// and asks for the word seven below it with a 4-byte cp.async, so a word has seven reads' time to arrive and nothing
// loop-carried ever waits on global memory (a register-held read-ahead made the compiler copy the word in flight at the
// loop's merge point: one third of all stall samples).  Words below the blob's first are not read (index clamped: such
// bytes lie outside the segment and are masked by `left`); bytes beyond the segment are zeros.
struct VlcStream { const uint32_t *base; uint32_t *ring; int wi; uint32_t hi, sh; int left; uint64_t tmp; uint32_t bits; bool gt8f; };

__device__ __forceinline__ void vlc_fetch(const VlcStream &v, int wi)     // word wi -> ring slot wi % 8, one commit group
{
    const uint32_t *src = v.base + max(wi, 0);
    uint32_t *dst = v.ring + (wi & 7) * kThreads;
#ifdef J2K_EMU
    *dst = *src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.commit_group;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}

// next byte to read: d[pos]; ring: this thread's column of the CTA's [8][kThreads] words
__device__ __forceinline__ void vlc_open(VlcStream &v, const uint8_t *d, int pos, const uint8_t *blob, uint32_t *ring)
{
    v.base = reinterpret_cast<const uint32_t *>((uintptr_t)blob & ~(uintptr_t)3);
    v.ring = ring;
    const int64_t a = (int64_t)((d + pos) - reinterpret_cast<const uint8_t *>(v.base));       // -1 at the least (then nothing is left to read)
    const int w = (int)(a >> 2);
    v.sh = ((uint32_t)(a & 3) + 1) * 8;                  // d[pos - 3 .. pos] = (hi : lo) >> sh, sh = 8 .. 32
    v.hi = __ldg(v.base + max(w, 0));
    v.wi = w - 1;
#pragma unroll
    for (int i = 0; i < 7; i++) vlc_fetch(v, v.wi - i);
}

__device__ __forceinline__ void vlc_read4(VlcStream &v)
{
#ifndef J2K_EMU
    asm volatile("cp.async.wait_group 6;" ::: "memory");  // all but the six youngest requests have landed: word wi is there
#endif
    const uint32_t lo = v.ring[(v.wi & 7) * kThreads];
    vlc_fetch(v, v.wi - 7);                              // into the slot the previous read emptied
#ifdef J2K_EMU
    const uint32_t x = v.sh == 32 ? v.hi : (uint32_t)((((uint64_t)v.hi << 32) | lo) >> v.sh);
#else
    uint32_t x;
    asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(x) : "r"(lo), "r"(v.hi), "r"(v.sh));
#endif
    v.hi = lo; v.wi--;
    uint32_t y = __byte_perm(x, 0, 0x0123);              // stream order: the first byte read is bits 0 .. 7
    if (v.left < 4) y = v.left <= 0 ? 0u : (y & ((1u << (8 * v.left)) - 1u));
    v.left -= 4;
    uint32_t t = y, nbits = 32;
    if (((y & 0x7F7F7F7Fu) + 0x01010101u) & 0x80808080u) {   // some byte has its low 7 bits all ones: it carries 7 bits after a byte > 0x8F
        bool g = v.gt8f;
        t = 0; nbits = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t bi = (y >> (8 * i)) & 0xFFu;
            t |= bi << nbits;
            nbits += (g && (bi & 0x7Fu) == 0x7Fu) ? 7u : 8u;
            g = bi > 0x8Fu;
        }
    }
    v.gt8f = (y >> 24) > 0x8Fu;
    v.tmp |= (uint64_t)t << v.bits;
    v.bits += nbits;
}

// at least 32 bits afterwards (four stuffed bytes give only 28: then a second read, which is rare)
__device__ __forceinline__ void vlc_refill(VlcStream &v)
{
    if (v.bits < 32) {
        vlc_read4(v);
        if (v.bits < 32) vlc_read4(v);
    }
}

// 4 bits -> the even bit positions of a byte
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return (x & 1) | ((x & 2) << 1) | ((x & 4) << 2) | ((x & 8) << 3); }

// U-VLC of a quad pair from a 320-entry table built at kernel start: kind (0 only the first quad has a code, 1 only the
// second, 2 both, 3 both in the initial row: when the first prefix is the long one the second code is a single bit,
// 4 both in the initial row with the MEL offset: u + 2) x the next 6 bits ->
// prefix bits (3) | first suffix length (3) | second suffix length (3) | first base (3) | second base (3); an absent code has
// length 0 and base 0, so that u = base + suffix needs no selection.
__device__ __forceinline__ uint16_t uvlc_entry(int kind, uint32_t bits6)
{
    const uint32_t t1 = uvlc_row(bits6 & 7);
    const uint32_t p1 = t1 & 3, s1 = (t1 >> 2) & 7, b1 = t1 >> 5;
    if (kind == 0) return (uint16_t)(p1 | (s1 << 3) | (b1 << 9));
    if (kind == 1) return (uint16_t)(p1 | (s1 << 6) | (b1 << 12));
    const uint32_t rest = bits6 >> p1;
    if (kind == 3 && p1 > 2) return (uint16_t)((p1 + 1) | (s1 << 3) | (b1 << 9) | (((rest & 1) + 1) << 12));
    const uint32_t t2 = uvlc_row(rest & 7), add = kind == 4 ? 2u : 0u;
    return (uint16_t)((p1 + (t2 & 3)) | (s1 << 3) | (((t2 >> 2) & 7) << 6) | ((b1 + add) << 9) | (((t2 >> 5) + add) << 12));
}

// mode: 1 first quad, 2 second quad, 3 both, 4 both + MEL offset (initial row)
__device__ __forceinline__ int uvlc_pair(const uint16_t *utab, uint32_t vlc, int mode, bool initial, int &u0, int &u1)
{
    const int kind = (initial && mode == 3) ? 3 : (mode == 4 ? 4 : mode - 1);
    const uint32_t t = utab[kind * 64 + (vlc & 63)];
    const int pl = t & 7, s1 = (t >> 3) & 7, s2 = (t >> 6) & 7;
    vlc >>= pl;
    u0 = (int)(((t >> 9) & 7) + (vlc & ((1u << s1) - 1)));
    vlc >>= s1;
    u1 = (int)((t >> 12) + (vlc & ((1u << s2) - 1)));
    return pl + s1 + s2;
}

__global__ void __launch_bounds__(kThreads, 6)
k_htiso_vlc(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
            uint32_t *__restrict__ qtab, uint32_t *__restrict__ status)
{
    // table entries re-packed: len (3) | u_off (1) | rho (4) | sample states (8)
    __shared__ uint16_t s_tbl[2048];
    for (int i = threadIdx.x; i < 2048; i += kThreads) {
        const uint32_t e = i < 1024 ? d_tbl0[i] : d_tbl1[i - 1024];
        const uint32_t st8 = spread4((e >> 4) & 15) + spread4((e >> 12) & 15) + spread4((e >> 8) & 15);
        s_tbl[i] = (uint16_t)((e & 0xFF) | (st8 << 8));
    }
    __shared__ uint32_t s_vring[8 * kThreads];
    __shared__ uint16_t s_utab[320];
    for (int i = threadIdx.x; i < 320; i += kThreads) s_utab[i] = uvlc_entry(i >> 6, (uint32_t)i & 63);
    __syncthreads();
    const uint32_t blk = blockIdx.x * kThreads + threadIdx.x;
    if (blk >= n) return;
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h;
    const uint8_t *d = blob + cb.data_off;
    const int lcup = (int)cb.len_cup;                    // the cleanup segment; a refinement segment may follow it
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
        vlc_open(v, d, lcup - 3, blob, s_vring + threadIdx.x);
        v.left = scup - 2;
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
            vlc_refill(v);                                    // >= 32 bits: two codewords (<= 7 each) and the U-VLC (<= 16)
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

// ---- SigProp + MagRef (T.814 clause 7.4 / 7.5) -----------------------------------------------------------------------
// k_htiso_refine, one warp per block, between the VLC kernel and the MagSgn kernel; only blocks with num_passes > 1 and
// a non-empty refinement segment do anything.  Both passes refine bit-plane P - 1 below the cleanup pass:
//   * SigProp: 4-row stripes, column by column; an insignificant sample with a significant neighbour (cleanup
//     significance of all eight neighbours, SigProp significance of the ones visited before it) takes one bit from a
//     forward-growing stream (LSB first, 7 bits after 0xFF, zeros when exhausted); the sign bits of the samples that
//     turned significant come after the (up to 16) significance bits of each group of 4 stripe columns.  The chain of
//     decisions is serial (every bit's meaning depends on the ones before): lane 0 walks it on 64-bit row bitmaps with
//     the candidate-column mask of the EBCOT kernel; the warp first removes the stuffing in parallel so that the chain
//     reads bit i of a dense bit string.
//   * MagRef: one bit per cleanup-significant sample in stripe scan order from a backward-growing stream -- no chain:
//     a sample's bit index is the number of significant samples before it, i.e. popcounts and a warp prefix sum.
// Output: three 64 x 64 bitmaps per block (new in SigProp, its sign, the MagRef bit) that the MagSgn kernel folds into
// the values it writes.  Input significance comes from the VLC kernel's quad table.
constexpr int kRefWarps = 4;
constexpr int kRefWords = 192;                          // u64 per block in the scratch: 64 rows x (new, sign, magref)
constexpr int kSppWords = 304;                          // SigProp bits: at most 2 per sample
constexpr int kMrpWords = 152;                          // MagRef bits: at most 1 per sample
constexpr int kSppBytes = 1184, kMrpBytes = 592;        // bytes that can hold them, stuffing included

__device__ __forceinline__ uint64_t spread_even(uint32_t x)          // bit i -> bit 2 i
{
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

__device__ __forceinline__ uint32_t win3r(uint64_t row, int x) { return (uint32_t)(x ? (row >> (x - 1)) : (row << 1)) & 7u; }

// the warp unstuffs `nbytes` bytes of a stream into a dense LSB-first bit string in shared memory (zeroed before).
// FWD: bytes d[0], d[1], ...; a byte after 0xFF carries 7 bits.  !FWD: bytes d[0], d[-1], ...; a byte whose low 7 bits are
// all ones after a byte > 0x8F carries 7 bits, and the byte before the first counts as > 0x8F.  Returns the bit count.
template <bool FWD>
__device__ uint32_t unstuff_stream(const uint8_t *d, int nbytes, uint32_t *bits, int lane)
{
    uint32_t total = 0;
    uint32_t carry = FWD ? 0u : 0xFFu;                   // the byte before this chunk, as stored
    for (int k0 = 0; k0 < nbytes; k0 += 128) {
        uint32_t ob[4], nb[4], val[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const int k = k0 + 4 * lane + i; ob[i] = k < nbytes ? (uint32_t)__ldg(FWD ? d + k : d - k) : 0u; }
        uint32_t prev = __shfl_up_sync(0xffffffffu, ob[3], 1);
        if (lane == 0) prev = carry;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const bool seven = FWD ? (prev == 0xFFu) : (prev > 0x8Fu && (ob[i] & 0x7Fu) == 0x7Fu);
            nb[i] = (k0 + 4 * lane + i < nbytes) ? (seven ? 7u : 8u) : 0u;
            val[i] = seven ? (ob[i] & 0x7Fu) : ob[i];
            prev = ob[i];
        }
        const uint32_t v = val[0] | (val[1] << nb[0]) | (val[2] << (nb[0] + nb[1])) | (val[3] << (nb[0] + nb[1] + nb[2]));
        const uint32_t tot = nb[0] + nb[1] + nb[2] + nb[3];
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const uint32_t pos = total + incl - tot;
        if (tot) {
            const uint32_t sh = pos & 31, wi = pos >> 5;
            atomicOr(&bits[wi], v << sh);
            const uint32_t hi = sh ? v >> (32 - sh) : 0u;
            if (hi) atomicOr(&bits[wi + 1], hi);
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
        carry = __shfl_sync(0xffffffffu, ob[3], 31);
    }
    __syncwarp();
    return total;
}

__global__ void __launch_bounds__(kRefWarps * 32)
k_htiso_refine(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
               const uint32_t *__restrict__ qtab, const uint32_t *__restrict__ status, uint64_t *__restrict__ ref)
{
    __shared__ uint64_t s_sg[kRefWarps][66];             // cleanup significance, rows -1 .. 64
    __shared__ uint64_t s_out[kRefWarps][kRefWords];     // new, sign, magref rows
    __shared__ uint32_t s_spp[kRefWarps][kSppWords];
    __shared__ uint32_t s_mrp[kRefWarps][kMrpWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t blk = blockIdx.x * kRefWarps + warp;
    if (blk >= n) return;
    const DevCblk cb = cblks[blk];
    const int lcup = (int)cb.len_cup, lref = (int)cb.data_len - lcup;
    if (cb.num_passes < 2 || lref <= 0 || (status[blk] & 3) == ST_ZERO) return;
    const int w = cb.w, h = cb.h;
    const int np = cb.num_passes > 3 ? 3 : cb.num_passes;
    uint64_t *sg = s_sg[warp] + 1, *onew = s_out[warp], *osgn = s_out[warp] + 64, *omr = s_out[warp] + 128;
    uint32_t *spp = s_spp[warp], *mrp = s_mrp[warp];
    for (int i = lane; i < 66; i += 32) s_sg[warp][i] = 0;
    for (int i = lane; i < kRefWords; i += 32) s_out[warp][i] = 0;
    for (int i = lane; i < kSppWords; i += 32) spp[i] = 0;
    for (int i = lane; i < kMrpWords; i += 32) mrp[i] = 0;
    __syncwarp();
    // ---- cleanup significance from the quad table: lane = quad of a quad row ----
    const uint16_t *qt = reinterpret_cast<const uint16_t *>(qtab + (size_t)blk * kQTabWords);
    const int nq = (w + 1) >> 1, nrows = (h + 1) >> 1;
    for (int r = 0; r < nrows; r++) {
        const uint32_t st8 = lane < nq ? ((uint32_t)qt[r * 32 + lane] & 0xFFu) : 0u;
        const uint32_t b0 = __ballot_sync(0xffffffffu, (st8 & 0x03u) != 0), b1 = __ballot_sync(0xffffffffu, (st8 & 0x0Cu) != 0),
                       b2 = __ballot_sync(0xffffffffu, (st8 & 0x30u) != 0), b3 = __ballot_sync(0xffffffffu, (st8 & 0xC0u) != 0);
        if (lane == 0) {
            sg[2 * r] = spread_even(b0) | (spread_even(b2) << 1);
            if (2 * r + 1 < 64) sg[2 * r + 1] = spread_even(b1) | (spread_even(b3) << 1);
        }
    }
    const uint8_t *dref = blob + cb.data_off + lcup;
    const uint32_t nspp = unstuff_stream<true>(dref, lref < kSppBytes ? lref : kSppBytes, spp, lane);
    uint32_t nmrp = 0;
    if (np == 3) nmrp = unstuff_stream<false>(dref + lref - 1, lref < kMrpBytes ? lref : kMrpBytes, mrp, lane);
    __syncwarp();
    const uint64_t wmask = (w >= 64) ? ~0ull : ((1ull << w) - 1);
    // ---- SigProp: the serial chain, lane 0 ----
    if (lane == 0) {
        uint32_t pos = 0;
        uint64_t above = 0;                                  // row y0 - 1: cleanup or SigProp significance
        for (int y0 = 0; y0 < h; y0 += 4) {
            const int rows = (y0 + 4 <= h) ? 4 : (h - y0);
            uint64_t a[6], nw[4] = {0, 0, 0, 0}, sn[4] = {0, 0, 0, 0};
            a[0] = above;
#pragma unroll
            for (int k = 1; k < 6; k++) a[k] = (k <= rows || k == rows + 1) ? sg[y0 + k - 1] : 0;
            if (y0 + rows >= h) a[rows + 1] = 0;
            uint64_t colmask = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k >= rows) break;
                const uint64_t u = a[k], m = a[k + 1], c = a[k + 2];
                colmask |= ~m & (u | (u << 1) | (u >> 1) | (m << 1) | (m >> 1) | c | (c << 1) | (c >> 1));
            }
            colmask &= wmask;
            int grp = -1;
            while (true) {
                const int x = colmask ? __ffsll((long long)colmask) - 1 : 64;
                if ((x >> 2) != grp) {
                    if (grp >= 0) {                              // the signs of the group just finished
                        for (int c = 4 * grp; c < 4 * grp + 4; c++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                if ((nw[k] >> c) & 1) {
                                    const uint32_t i = pos++;
                                    if (i < nspp && ((spp[i >> 5] >> (i & 31)) & 1)) sn[k] |= 1ull << c;
                                }
                    }
                    grp = x >> 2;
                }
                if (x >= 64) break;
                colmask &= colmask - 1;
                bool grew = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= rows) break;
                    if ((a[k + 1] >> x) & 1) continue;
                    const uint32_t idx9 = win3r(a[k], x) | (win3r(a[k + 1], x) << 3) | (win3r(a[k + 2], x) << 6);
                    if ((idx9 & 0x1EFu) == 0) continue;          // no significant neighbour (bit 4 is the sample itself)
                    const uint32_t i = pos++;
                    if (i < nspp && ((spp[i >> 5] >> (i & 31)) & 1)) { a[k + 1] |= 1ull << x; nw[k] |= 1ull << x; grew = true; }
                }
                if (grew && x + 1 < w) colmask |= 1ull << (x + 1);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) if (k < rows) { onew[y0 + k] = nw[k]; osgn[y0 + k] = sn[k]; }
            above = a[rows];
        }
    }
    // ---- MagRef: lane = columns 2 lane, 2 lane + 1 of a stripe ----
    if (np == 3) {
        uint32_t base = 0;
        for (int y0 = 0; y0 < h; y0 += 4) {
            uint32_t c0 = 0, c1 = 0;                             // the column's 4 significance bits, top to bottom
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (y0 + k < h) { const uint64_t rw = sg[y0 + k]; c0 |= (uint32_t)((rw >> (2 * lane)) & 1) << k; c1 |= (uint32_t)((rw >> (2 * lane + 1)) & 1) << k; }
            const uint32_t n0 = __popc(c0), tot = n0 + __popc(c1);
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const uint32_t off = base + incl - tot;
            base += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t bit0 = 0, bit1 = 0;
                if ((c0 >> k) & 1) { const uint32_t i = off + __popc(c0 & ((1u << k) - 1)); bit0 = i < nmrp ? (mrp[i >> 5] >> (i & 31)) & 1 : 0; }
                if ((c1 >> k) & 1) { const uint32_t i = off + n0 + __popc(c1 & ((1u << k) - 1)); bit1 = i < nmrp ? (mrp[i >> 5] >> (i & 31)) & 1 : 0; }
                const uint32_t e = __ballot_sync(0xffffffffu, bit0), o = __ballot_sync(0xffffffffu, bit1);
                if (lane == 0 && y0 + k < h) omr[y0 + k] = spread_even(e) | (spread_even(o) << 1);
            }
        }
    }
    __syncwarp();
    uint64_t *dst = ref + (size_t)blk * kRefWords;
    for (int i = lane; i < kRefWords; i += 32) dst[i] = s_out[warp][i];
}

// B (k_htiso_magsgn4): FOUR blocks per warp, eight lanes per block, each lane four neighbouring quads (8 columns, 16
// samples) of a quad row.  The MagSgn work is issue-bound, so the mapping is chosen for instructions per sample:
//   * what is per quad row and not per sample (loop, ring upkeep, predictor shuffles, prefix sum over 8 lanes instead of 32)
//     is shared by 512 samples per warp step;
//   * a sample costs ~20 straight-line instructions: no branch per sample (an insignificant sample is a zero-width field),
//     one funnel shift over two ring words (the ring keeps a copy of word 0 behind its last word, so the second load needs
//     no wrap), reconstruction as one add and one shift;
//   * the stuffing is removed 16 bytes per lane (two aligned 16-byte loads funnelled to the stream position): 0xFF bytes are
//     found with three word-wide operations per 4 bytes and the rare bit removals run in a short loop; blocks of a warp
//     refill together once one of them has to (hysteresis), so the refill runs about every other quad row.
// Same arithmetic as the single-chain checker for every input, malformed ones included: a byte after 0xFF contributes 7 bits
// and its top bit is OR-ed onto the next byte's first bit, an exhausted stream continues with 0xFF.
constexpr int kWarpsB4 = 4;
constexpr uint32_t kRingMask = kRingWords - 1;

// 4 stream bytes (x, little endian) -> dense bits; *nbits = 32 - number of bytes that follow a 0xFF byte; prev_ff: the byte
// before x was 0xFF; *last_ff: the last byte of x is 0xFF
__device__ __forceinline__ uint32_t unstuff_word(uint32_t x, uint32_t prev_ff, uint32_t *nbits, uint32_t *last_ff)
{
    const uint32_t ff = ((x & 0x7F7F7F7Fu) + 0x01010101u) & x & 0x80808080u;      // bit 7 of every byte that is 0xFF
    uint32_t st = (ff << 8) | (prev_ff ? 0x80u : 0u);                            // bit 7 of every byte that carries 7 bits
    *nbits = 32u - (uint32_t)__popc(st);
    *last_ff = ff >> 31;
    while (st) {                                                                 // top down, so that lower positions stay valid
        const uint32_t pos = 31u - (uint32_t)__clz((int)st);
        const uint32_t low = (1u << pos) - 1u;
        x = (x & ((low << 1) | 1u)) | ((x >> 1) & ~low);                         // bit pos stays and the bits above close onto it
        st &= low;
    }
    return x;
}

// REFINE: some block of the launch has SigProp / MagRef passes; their bitmaps (k_htiso_refine) are folded in here.
template <typename OT, bool IRREV, bool REFINE>
__global__ void __launch_bounds__(kWarpsB4 * 32)
k_htiso_magsgn4(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob, const uint8_t *blob_end,
                const uint32_t *__restrict__ qtab, const uint32_t *__restrict__ status, OT *__restrict__ coef,
                const float *__restrict__ steps, int coef_bits, const uint64_t *__restrict__ ref)
{
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ uint32_t s_ring[kWarpsB4 * 4][kRingWords + 1];            // [kRingWords] mirrors word 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, sl = lane & 7;
    const uint32_t blk0 = (blockIdx.x * kWarpsB4 + warp) * 4;
    if (blk0 >= n) return;
    const bool have = blk0 + grp < n;
    const uint32_t blk = have ? blk0 + grp : blk0;
    const uint32_t stw = status[blk];
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h;
    OT *out = coef + cb.out_off;
    const size_t ostride = cb.out_stride;
    const uint8_t *d = blob + cb.data_off;
    const bool zero = (stw & 3) == ST_ZERO;
    if (have && zero)
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += 8) out[(size_t)y * ostride + x] = 0;
    const bool live = have && !zero;                     // these eight lanes decode a block
    const float qstep = (IRREV && steps) ? 0.25f * steps[blk] : 0.25f;
    const int L = (int)(stw >> 2);                       // bytes of the MagSgn stream
    const int shift = cb.num_bps - 1;                    // P: bit-plane of the cleanup pass
    // passes of the HT set that are decoded: 1 = cleanup, 2 = + SigProp, 3 = + MagRef (none without refinement bytes)
    const int np = (REFINE && cb.num_passes > 1 && cb.data_len > cb.len_cup) ? (cb.num_passes > 3 ? 3 : (int)cb.num_passes) : 1;
    const uint64_t *rf = REFINE ? ref + (size_t)blk * kRefWords : nullptr;
    const int nq = (w + 1) >> 1, nrows = live ? (h + 1) >> 1 : 0;
    const uint32_t pairA = (live && 4 * sl < nq) ? ~0u : 0u, pairB = (live && 4 * sl + 2 < nq) ? ~0u : 0u;   // the lane's quad pairs exist
    const int ncols = w - 8 * sl;                        // columns of the block right of (and including) the lane's first
    const bool vec_ok = ((cb.out_off | ostride) & (sizeof(OT) == 2 ? 7 : 3)) == 0;
    const int ulimit = min(28, coef_bits ? coef_bits + 1 : 28) - shift;   // a wider field would not fit 31 bits in quarter units / the plane
    // The sample path is bound by the integer ALU pipe (LOP3 / SHF / ISETP issue every other cycle per scheduler) while the
    // FMA pipe, which executes IMAD at the same rate, idles: shifts by a per-block amount are multiplies by a power of two,
    // the sign is applied as a multiply, and additions are written as a * one + b with `one` opaque to the compiler.
    const uint32_t one = min(n, 1u), minus2 = 0u - (one + one);
    uint32_t mulq = 1u << (shift + 1);
#ifndef J2K_EMU
    asm volatile("" : "+r"(mulq));                       // keep it a multiplier (the compiler would turn the product back into a shift)
#endif
    const uint2 *qt = reinterpret_cast<const uint2 *>(qtab + (size_t)blk * kQTabWords) + sl;
    uint32_t *ring = s_ring[warp * 4 + grp];
    for (int i = sl; i <= kRingWords; i += 8) ring[i] = 0;
    __syncwarp();
    const uint32_t mis = (uint32_t)((uintptr_t)d & 3u);  // byte 0 of the stream inside its aligned word
    uint32_t built = 0, prev_ff = 0, P = 0;              // uniform over the block's lanes
    int kbyte = 0;
    // the five aligned words that hold the lane's 16 stream bytes from k on; a word is read only if it starts inside the blob
    uint32_t pre[5] = {FULL, FULL, FULL, FULL, FULL};
    auto load5 = [&](int k) {
        if (k >= L) return;
        const uint8_t *a = d + k - mis;
        if (a + 20 <= blob_end) {
#pragma unroll
            for (int j = 0; j < 5; j++) pre[j] = __ldg(reinterpret_cast<const uint32_t *>(a) + j);
        } else {
#pragma unroll
            for (int j = 0; j < 5; j++) pre[j] = (a + 4 * j < blob_end) ? __ldg(reinterpret_cast<const uint32_t *>(a) + j) : FULL;
        }
    };
    if (live) load5(16 * sl);
    int Eb[8] = {-1, -1, -1, -1, -1, -1, -1, -1};        // bottom-sample exponents - 1 of the lane's 8 columns, previous quad row
    bool bad = false;
    int nrows_max = nrows;
    nrows_max = max(nrows_max, __shfl_xor_sync(FULL, nrows_max, 8));
    nrows_max = max(nrows_max, __shfl_xor_sync(FULL, nrows_max, 16));
    uint2 code_next = make_uint2(0u, 0u);
    if (nrows > 0) { code_next = __ldg(qt); code_next.x &= pairA; code_next.y &= pairB; }
    for (int r = 0; r < nrows_max; r++) {
        const bool row_on = r < nrows;
        const uint2 code = row_on ? code_next : make_uint2(0u, 0u);
        if (r + 1 < nrows) { code_next = __ldg(qt + (r + 1) * 8); code_next.x &= pairA; code_next.y &= pairB; }
        // ---- keep each ring one full quad row (32 x 4 x 31 bits) ahead; when one block has to refill, every block that is
        // less than 7 Kbit ahead refills with it ----
        if (__any_sync(FULL, row_on && built < P + 4096u)) {
            bool join = row_on && built < P + 7168u;
            while (__any_sync(FULL, join)) {
                if (join) {                              // words beyond the one `built` points into hold bits 16 Kbit old
                    const uint32_t w0 = (built >> 5) + 1 + sl;
#pragma unroll
                    for (int i = 0; i < 5; i++) ring[(w0 + 8 * i) & kRingMask] = 0;
                }
                __syncwarp();
                const int k = kbyte + 16 * sl;           // the lane's stream bytes k .. k + 15: loaded one refill ahead (pre[])
                uint32_t x[4] = {FULL, FULL, FULL, FULL};
                if (join && k < L) {
#pragma unroll
                    for (int j = 0; j < 4; j++) x[j] = __funnelshift_r(pre[j], pre[j + 1], mis * 8);
                    if (k + 16 > L) {                    // past the end of the stream: 0xFF
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int vb = L - (k + 4 * j);
                            if (vb < 4) x[j] |= vb <= 0 ? FULL : (FULL << (8 * vb));
                        }
                    }
                }
                if (join) load5(k + 128);                // the next refill's bytes are on their way while these are used
                uint32_t pf = __shfl_up_sync(FULL, x[3] >> 24, 1, 8) == 0xFFu;
                if (sl == 0) pf = prev_ff;
                uint32_t nb[4], v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = unstuff_word(x[j], pf, &nb[j], &pf);
                const uint32_t tot = nb[0] + nb[1] + nb[2] + nb[3];
                uint32_t incl = tot;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o, 8); if (sl >= o) incl += t; }
                uint32_t pos = built + incl - tot;
                if (join) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t wi = pos >> 5;
                        atomicOr(&ring[wi & kRingMask], v[j] << (pos & 31));
                        atomicOr(&ring[(wi + 1) & kRingMask], __funnelshift_l(v[j], 0u, pos));   // the bits that spill over (none at pos % 32 == 0)
                        pos += nb[j];
                    }
                }
                const uint32_t total = __shfl_sync(FULL, incl, 7, 8);
                const uint32_t lastff = __shfl_sync(FULL, pf, 7, 8);
                if (join) { built += total; prev_ff = lastff; kbyte += 128; }
                __syncwarp();
                if (sl == 0) ring[kRingWords] = ring[0];  // the word behind the last mirrors word 0: readers take two words without a wrap
                __syncwarp();
                join = row_on && built < P + 4096u;      // (a second round only if 1024 bits were not enough)
            }
        }
        // ---- U_q and the field widths of the lane's four quads ----
        const int eL = __shfl_up_sync(FULL, Eb[7], 1, 8), eR = __shfl_down_sync(FULL, Eb[0], 1, 8);
        const int c_m1 = max(sl ? eL : -1, Eb[0]), c_1 = max(Eb[1], Eb[2]), c_3 = max(Eb[3], Eb[4]), c_5 = max(Eb[5], Eb[6]),
                  c_7 = max(Eb[7], sl < 7 ? eR : -1);
        const int Eq[4] = {max(c_m1, c_1), max(c_1, c_3), max(c_3, c_5), max(c_5, c_7)};
        uint32_t sigm[4], e1m[4];
        int m[16];
        uint32_t tot = 0;
#pragma unroll
        for (int qd = 0; qd < 4; qd++) {
            const uint32_t c16 = (qd & 1) ? ((qd >> 1) ? code.y : code.x) >> 16 : ((qd >> 1) ? code.y : code.x) & 0xFFFFu;
            const uint32_t st8 = c16 & 0xFFu;
            const int u = (int)(c16 >> 8);
            const uint32_t sig = (st8 | (st8 >> 1)) & 0x55u, ek = (st8 >> 1) & 0x55u;
            sigm[qd] = sig; e1m[qd] = st8 & ek;
            int Uq = u + 1;
            if (r > 0 && (sig & (sig - 1))) Uq = u + max(1, Eq[qd]);      // Eq holds E - 1
            bad |= Uq > ulimit;
            const int U = min(Uq, 31);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                // width = significant ? U - e_k : 0, as multiply-adds (e_k is set only where the sample is significant)
                const uint32_t mi = ((sig >> (2 * i)) & 1u) * (uint32_t)U - ((ek >> (2 * i)) & 1u);
                m[4 * qd + i] = (int)mi;
                tot = mi * one + tot;
            }
        }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o, 8); if (sl >= o) incl += t; }
        uint32_t p = P + incl - tot;
        P += __shfl_sync(FULL, incl, 7, 8);
        // ---- samples: per quad n = 0 (y, x), 1 (y + 1, x), 2 (y, x + 1), 3 (y + 1, x + 1) ----
        // refinement bitmaps of the two sample rows: the lane's 8 columns are bits 8 sl .. 8 sl + 7
        uint32_t rnew[2] = {0, 0}, rsgn[2] = {0, 0}, rmr[2] = {0, 0};
        if (REFINE && np > 1 && row_on) {
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int y = 2 * r + t;
                if (y < h) {
                    rnew[t] = (uint32_t)(__ldg(rf + y) >> (8 * sl)) & 0xFFu;
                    rsgn[t] = (uint32_t)(__ldg(rf + 64 + y) >> (8 * sl)) & 0xFFu;
                    if (np == 3) rmr[t] = (uint32_t)(__ldg(rf + 128 + y) >> (8 * sl)) & 0xFFu;
                }
            }
        }
        int32_t val[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int qd = i >> 2, sn = i & 3;
            const uint32_t sigb = (sigm[qd] >> (2 * sn)) & 1u, e1 = (e1m[qd] >> (2 * sn)) & 1u;
            const uint32_t *rw = ring + ((p >> 5) & kRingMask);
            const uint32_t xb = __funnelshift_r(rw[0], rw[1], p);
            const uint32_t pw = 1u << m[i];
            const uint32_t fld = xb & (pw * one - 1u);
            const uint32_t vv = fld | (e1 * pw + sigb);                  // 2 (mu - 1) + 1, or 0 for an insignificant sample
            p = (uint32_t)m[i] * one + p;
            if (sn & 1) Eb[2 * qd + (sn >> 1)] = 31 - __clz((int)vv);     // exponent - 1 (-1 for an insignificant sample)
            const int col = 2 * qd + (sn >> 1), rowt = sn & 1;           // position inside the lane's 8 x 2 patch
            uint32_t q = (sigb * 2u + vv) * mulq;                        // (2 mu + 1) << (P + 1)
            if (REFINE && np == 3) q = sigb ? ((((vv >> 1) + 1u) << (shift + 2)) | (((rmr[rowt] >> col) & 1u) << (shift + 1)) | (1u << shift)) : 0u;
            uint32_t sign = fld & 1u;
            if (REFINE && !sigb && ((rnew[rowt] >> col) & 1u)) { q = 3u << shift; sign = (rsgn[rowt] >> col) & 1u; }
            if (!IRREV) { val[i] = (int32_t)((q >> 2) * (sign * minus2 + one)); continue; }   // magnitude * (1 - 2 sign)
            val[i] = sample_value(q, sign, qstep, IRREV);
        }
        if (row_on && ncols > 0) {
            const int y = 2 * r;
            const bool row2 = (y + 1 < h);
            OT *p0 = out + (size_t)y * ostride + 8 * sl;
            if (ncols >= 8 && vec_ok) {
                if (sizeof(OT) == 4) {
                    *reinterpret_cast<int4 *>(p0) = make_int4(val[0], val[2], val[4], val[6]);
                    *reinterpret_cast<int4 *>(p0 + 4) = make_int4(val[8], val[10], val[12], val[14]);
                    if (row2) {
                        *reinterpret_cast<int4 *>(p0 + ostride) = make_int4(val[1], val[3], val[5], val[7]);
                        *reinterpret_cast<int4 *>(p0 + ostride + 4) = make_int4(val[9], val[11], val[13], val[15]);
                    }
                } else {
#define J2K_PK16(a, b) (((uint32_t)(a) & 0xFFFFu) | ((uint32_t)(b) << 16))
                    *reinterpret_cast<uint4 *>(p0) = make_uint4(J2K_PK16(val[0], val[2]), J2K_PK16(val[4], val[6]), J2K_PK16(val[8], val[10]), J2K_PK16(val[12], val[14]));
                    if (row2) *reinterpret_cast<uint4 *>(p0 + ostride) = make_uint4(J2K_PK16(val[1], val[3]), J2K_PK16(val[5], val[7]), J2K_PK16(val[9], val[11]), J2K_PK16(val[13], val[15]));
#undef J2K_PK16
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++)
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
    if (live && (badm & (0xFFu << (8 * grp)))) {
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += 8) out[(size_t)y * ostride + x] = 0;
    }
}

}  // namespace

// quad table + status (+ the refinement bitmaps when some block has SigProp / MagRef passes)
size_t j2k_htiso_scratch_bytes(uint32_t n, int refine)
{
    return (size_t)n * (kQTabWords * 4 + 4) + 64 + (refine ? (size_t)n * kRefWords * 8 : 0);
}
int j2k_htiso_launches(int refine) { return refine ? 3 : 2; }

template <typename OT>
static void launch_ht_iso_t(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, OT *d_coef,
                            const float *d_steps, int irrev, int coef_bits, int refine, void *d_scratch, uint64_t blob_bytes, cudaStream_t s)
{
    const uint64_t blob_total = blob_bytes;
    uint32_t *qtab = (uint32_t *)d_scratch, *status = qtab + (size_t)n * kQTabWords;
    uint64_t *ref = (uint64_t *)(((uintptr_t)(status + n) + 63) & ~(uintptr_t)63);
    J2K_LAUNCH((k_htiso_vlc), (n + kThreads - 1) / kThreads, kThreads, 0, s, d_cblks, n, d_blob, qtab, status);
    if (refine) J2K_LAUNCH((k_htiso_refine), (n + kRefWarps - 1) / kRefWarps, kRefWarps * 32, 0, s, d_cblks, n, d_blob, qtab, status, ref);
    const uint32_t grid = (n + 4 * kWarpsB4 - 1) / (4 * kWarpsB4);
    const uint8_t *blob_end = d_blob + blob_total;
#define J2K_HTISO_B(IRR, REF) J2K_LAUNCH((k_htiso_magsgn4<OT, IRR, REF>), grid, kWarpsB4 * 32, 0, s, d_cblks, n, d_blob, blob_end, qtab, status, d_coef, d_steps, coef_bits, ref)
    if (irrev) { if (refine) J2K_HTISO_B(true, true); else J2K_HTISO_B(true, false); }
    else { if (refine) J2K_HTISO_B(false, true); else J2K_HTISO_B(false, false); }
#undef J2K_HTISO_B
}

// d_scratch: j2k_htiso_scratch_bytes(n, refine) bytes of device memory; refine: some block has num_passes > 1
cudaError_t launch_ht_iso(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          const float *d_steps, int irrev, int coef_bits, int refine, void *d_scratch, uint64_t blob_bytes,
                          cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    if (blob_bytes >> 33) return cudaErrorInvalidValue;  // the VLC reader indexes the blob's words with an int
    if (coef16 && !irrev) launch_ht_iso_t<int16_t>(d_cblks, n, d_blob, (int16_t *)d_coef, d_steps, irrev, coef_bits, refine, d_scratch, blob_bytes, s);
    else launch_ht_iso_t<int32_t>(d_cblks, n, d_blob, (int32_t *)d_coef, d_steps, irrev, coef_bits, refine, d_scratch, blob_bytes, s);
    return cudaGetLastError();
}
