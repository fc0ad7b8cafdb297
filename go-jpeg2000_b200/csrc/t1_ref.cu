// t1_ref.cu -- EBCOT tier-1 block decoder, REF semantics, one warp per code block (sm_100a).
//
// Replaces entropy.T1.Decode (reference internal/entropy/t1.go:1261-1410) together with
// MQDecoder (internal/entropy/mqc.go:370-497) and the context rules of t1.go:349-479 /
// t1_luts.go:35-110.  REF quirks kept: raster-order SPP/MRP, stripe-order cleanup, run-length
// only on full 4-row columns, all MQ contexts start in state 0 (UNI 92), no pass truncation.
//
// Design (not a port of the Go loops): the (w+2)(h+2) byte flag array and the int32 magnitude
// array of the reference become ROW BITMAPS -- one 64-bit word per code-block row for each of
// significance / sign / visited / refined, plus one bitmap per magnitude bit-plane -- held in
// shared memory (about 2 KB + 512 B per bit-plane per block instead of 21 KB).  Neighbourhood
// tests are 64-bit shifts of three row words; the candidate set of a whole row is one boolean
// expression; the zero-coding context is a 9-bit window (3 rows x 3 columns) indexing a 512-entry
// table in constant memory.  The MQ decision chain is inherently serial: one lane of the warp runs
// it; all lanes split the work that IS parallel: clearing state, assembling magnitudes from the
// bit-plane bitmaps, applying signs and the coalesced store into the tile-component plane.
#include "common.h"
#include <type_traits>

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kCtxZC = 0, kCtxSC = 9, kCtxMag = 14, kCtxRL = 17, kCtxUni = 18, kNumCtx = 19;
// shared memory per block in 64-bit words: significance rows -1 .. 64, sign, visited, refined (+ coded-last for ISO), one
// magnitude bit-plane, the 19 context states
constexpr int kRefWords = 66 + 64 * 4 + 4, kIsoWords = 66 + 64 * 5 + 4;

__constant__ uint32_t c_mq[94];          // qe | nmps << 16 | nlps << 24      (mqc.go:21-116)
__constant__ uint8_t  c_zc9[4 * 512];    // band, 3x3 significance window -> ZC context (t1_luts.go:35-110)
__constant__ uint8_t  c_sc[256];         // W,E,N,S (sig,neg) pairs -> (SC context - 9) << 1 | prediction (t1.go:387-460)
// the same two tables filled from ISO/IEC 15444-1 Tables D.1 - D.3 (J2KGPU_MODE_ISO)
__constant__ uint8_t  c_zc9_iso[4 * 512];
__constant__ uint8_t  c_sc_iso[256];

struct MQ {
    uint32_t A, C, CT;
    int bp, len;
    uint32_t cur;                        // data[bp] when bp < len
    const uint8_t *d;
};

// byteIn, mqc.go:402-439
__device__ __forceinline__ void mq_bytein(MQ &m)
{
    if (m.bp >= m.len) { m.C += 0xFF00; m.CT = 8; return; }
    uint32_t nxt = (m.bp + 1 < m.len) ? (uint32_t)__ldg(m.d + m.bp + 1) : 0xFFu;
    if (m.cur == 0xFF) {
        if (nxt > 0x8F) { m.C += 0xFF00; m.CT = 8; }
        else { m.bp++; m.cur = nxt; m.C += nxt << 9; m.CT = 7; }
    } else {
        m.bp++; m.cur = nxt; m.C += nxt << 8; m.CT = 8;
    }
}

// NewMQDecoder, mqc.go:370-399
__device__ __forceinline__ void mq_init(MQ &m, const uint8_t *d, int len)
{
    m.d = d; m.len = len; m.A = 0x8000; m.CT = 0; m.bp = 0;
    if (len == 0) { m.C = 0xFFu << 16; m.cur = 0; }
    else { m.cur = __ldg(d); m.C = m.cur << 16; }
    mq_bytein(m);
    m.C <<= 7;
    m.CT -= 7;
    m.A = 0x8000;
}

// Decode + renormDec, mqc.go:443-497.  ctxs = 19 state indices in shared memory (warp-uniform).
__device__ __forceinline__ uint32_t mq_decode(MQ &m, uint8_t *ctxs, int ctx)
{
    uint32_t st = ctxs[ctx];
    uint32_t row = c_mq[st];
    uint32_t qe = row & 0xFFFF;
    uint32_t mps = st & 1, d;
    m.A -= qe;
    if ((m.C >> 16) < qe) {
        if (m.A < qe) { d = mps;     st = (row >> 16) & 0xFF; }
        else          { d = mps ^ 1; st = row >> 24; }
        m.A = qe;
    } else {
        m.C -= qe << 16;
        if (m.A & 0x8000) return mps;
        if (m.A < qe) { d = mps ^ 1; st = row >> 24; }
        else          { d = mps;     st = (row >> 16) & 0xFF; }
    }
    ctxs[ctx] = (uint8_t)st;
    do {
        if (m.CT == 0) mq_bytein(m);
        m.A <<= 1; m.C <<= 1; m.CT--;
    } while ((m.A & 0x8000) == 0);
    return d;
}

// bits (x-1, x, x+1) of a row word as a 3-bit value; columns outside 0..63 read as 0
__device__ __forceinline__ uint32_t win3(uint64_t row, int x)
{
    return (uint32_t)(x ? (row >> (x - 1)) : (row << 1)) & 7u;
}

// decodeSign t1.go:1322-1328 with getSCContext t1.go:387-460; returns 1 for negative
__device__ __forceinline__ uint32_t decode_sign(MQ &m, uint8_t *ctxs, int x,
                                                uint64_t sup, uint64_t smid, uint64_t sdn,
                                                uint64_t nup, uint64_t nmid, uint64_t ndn, const uint8_t *sc = c_sc)
{
    uint32_t ms = win3(smid, x), mn = win3(nmid, x);
    uint32_t idx = (ms & 1) | ((mn & 1) << 1) | ((ms >> 2) << 2) | ((mn >> 2) << 3) |
                   ((uint32_t)((sup >> x) & 1) << 4) | ((uint32_t)((nup >> x) & 1) << 5) |
                   ((uint32_t)((sdn >> x) & 1) << 6) | ((uint32_t)((ndn >> x) & 1) << 7);
    uint32_t e = sc[idx];
    return mq_decode(m, ctxs, kCtxSC + (e >> 1)) ^ (e & 1);
}

// One magnitude bit-plane lives in shared memory; when it is complete the group's lanes OR it into the block's samples in
// the coefficient plane (which therefore hold plain magnitudes until the final pass applies signs / mid-points) and
// clear it.  Lane sl owns columns sl * 64 / G ... of every row.  Keeping all num_bps planes in shared memory (512 bytes
// each) allowed 24 blocks per SM; one plane allows about 80 chains per SM inside the register budget of 25 warps.
template <typename OT, int G>
__device__ __forceinline__ void t1_flush_plane(uint64_t *plane, OT *out, uint32_t ostride, int h, int sl, int bp, uint32_t gmask)
{
    constexpr int W = 64 / G;
    typedef typename std::conditional<sizeof(OT) == 2, uint16_t, uint32_t>::type UT;
    for (int y = 0; y < h; y++) {
        uint32_t bits = (uint32_t)(plane[y] >> (sl * W)) & (uint32_t)((1ull << W) - 1);
        UT *row = reinterpret_cast<UT *>(out + (size_t)y * ostride + sl * W);
        while (bits) {
            const int j = __ffs((int)bits) - 1;
            bits &= bits - 1;
            row[j] = (UT)(row[j] | (UT)(1u << bp));
        }
    }
    __syncwarp(gmask);                                   // every lane has read the plane
    for (int y = sl; y < 64; y += G) plane[y] = 0;
    __syncwarp(gmask);
}

// OT = element type of the coefficient arena: int32_t, or int16_t when every block of the job has num_bps <= 15.
// G = lanes per code block: a warp decodes 32 / G blocks at once.  The MQ decision chain of a block is serial and runs in
// the first lane of its group; with one block per warp every issued instruction did one lane's worth of work.  With
// several chains per warp the SIMT hardware issues an instruction once for all the chains that are at it -- the passes are
// loops over rows, candidate columns and the MQ routine, so the chains of a warp walk the same code and mostly meet --
// and the group's lanes share the parallel parts (clearing state, assembling magnitudes, stores).
template <typename OT, int G>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 6)
k_t1_ref(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob,
         OT *__restrict__ coef, int skip_empty)
{
    J2K_DYN_SMEM(uint64_t, smem);
    constexpr int BPW = 32 / G, W = 64 / G;              // blocks per warp, columns per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / G, sl = lane % G;
    const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp * G));
    const uint32_t blk_raw = (blockIdx.x * (blockDim.x >> 5) + warp) * BPW + grp;
    const bool have = blk_raw < n;
    const uint32_t blk = have ? blk_raw : n - 1;         // lanes without a block follow along and write nothing
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h, nbps = cb.num_bps, band = cb.band & 3;

    uint64_t *base = smem + (size_t)(warp * BPW + grp) * kRefWords;
    uint64_t *sig = base + 1;            // rows -1 .. 64
    uint64_t *neg = base + 66;
    uint64_t *visit = base + 130;
    uint64_t *refine = base + 194;
    uint64_t *plane = base + 258;        // the bit-plane being decoded
    uint8_t *ctxs = (uint8_t *)(base + 322);

    OT *out = coef + cb.out_off;
    const uint32_t ostride = cb.out_stride;

    const bool coded = !((cb.data_len == 0 && skip_empty) || nbps == 0);
    if (have && !coded) {                                          // tcd.go:394-396: not coded -> zeros
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += G) out[(size_t)y * ostride + x] = 0;
    }
    const bool run = have && coded;
    if (run) {
        for (int i = sl; i < 322; i += G) base[i] = 0;
        for (int i = sl; i < kNumCtx; i += G) ctxs[i] = (i == kCtxUni) ? 92 : 0;   // mqc.go:378-383
        for (int y = 0; y < h; y++)                                    // the samples collect their magnitude bits plane by plane
#pragma unroll
            for (int j = 0; j < W; j++)
                if (sl * W + j < w) out[(size_t)y * ostride + sl * W + j] = 0;
    }
    __syncwarp();

    MQ mq;
    mq_init(mq, blob + cb.data_off, (int)cb.data_len);
    const uint64_t wmask = (w >= 64) ? ~0ull : ((1ull << w) - 1);
    const uint8_t *zc = c_zc9 + band * 512;

    // the serial MQ chain: the first lane of the group runs it, the others wait for the plane
    for (int bp = nbps - 1; bp >= 0 && run; bp--) {
      if (sl == 0) {

        // ---- significance propagation, raster order (t1.go:1295-1319) ----
        for (int y = 0; y < h; y++) {
            const uint64_t up = sig[y - 1], dn = sig[y + 1];
            uint64_t mid = sig[y];
            const uint64_t nbs = up | (up << 1) | (up >> 1) | dn | (dn << 1) | (dn >> 1);
            if (((nbs | mid) & wmask) == 0) continue;               // nothing can be a candidate
            const uint64_t nup = y > 0 ? neg[y - 1] : 0, ndn = y < 63 ? neg[y + 1] : 0;
            uint64_t nmid = neg[y], todo = wmask, vis = 0, bits = 0;
            for (;;) {
                uint64_t cand = ~mid & (nbs | (mid << 1) | (mid >> 1)) & todo;
                if (!cand) break;
                int x = __ffsll((long long)cand) - 1;
                todo = (x >= 63) ? 0 : (todo & (~0ull << (x + 1)));
                uint32_t idx9 = win3(up, x) | (win3(mid, x) << 3) | (win3(dn, x) << 6);
                if (mq_decode(mq, ctxs, kCtxZC + zc[idx9])) {
                    bits |= 1ull << x;
                    if (decode_sign(mq, ctxs, x, up, mid, dn, nup, nmid, ndn)) nmid |= 1ull << x;
                    mid |= 1ull << x;
                }
                vis |= 1ull << x;
            }
            sig[y] = mid; neg[y] = nmid; visit[y] = vis;
            if (bits) plane[y] |= bits;
        }

        // ---- magnitude refinement, raster order (t1.go:1331-1347) ----
        for (int y = 0; y < h; y++) {
            const uint64_t mid = sig[y];
            uint64_t cand = mid & ~visit[y];
            if (!cand) continue;
            const uint64_t up = sig[y - 1], dn = sig[y + 1], ref = refine[y];
            // significance does not change during this pass: "has a significant neighbour" is one mask per row
            const uint64_t nbm = up | (up << 1) | (up >> 1) | dn | (dn << 1) | (dn >> 1) | (mid << 1) | (mid >> 1);
            uint64_t bits = 0;
            refine[y] = ref | cand;
            // the masks are fixed during the loop: walk the row as two 32-bit halves (64-bit find-first-set, clear-lowest and
            // variable shifts cost twice the instructions)
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                uint32_t c32 = (uint32_t)(cand >> (32 * half));
                if (!c32) continue;
                const uint32_t r32 = (uint32_t)(ref >> (32 * half)), n32 = (uint32_t)(nbm >> (32 * half));
                uint32_t b32 = 0;
                while (c32) {
                    const int x = __ffs((int)c32) - 1;
                    c32 &= c32 - 1;
                    const int ctx = kCtxMag + (((r32 >> x) & 1) ? 2 : (int)((n32 >> x) & 1));
                    if (mq_decode(mq, ctxs, ctx)) b32 |= 1u << x;
                }
                bits |= (uint64_t)b32 << (32 * half);
            }
            if (bits) plane[y] |= bits;
        }

        // ---- cleanup, 4-row stripes, column by column (t1.go:1350-1410) ----
        for (int y0 = 0; y0 < h; y0 += 4) {
            uint64_t s[6], ng[6], v[4], pb[4];
#pragma unroll
            for (int k = 0; k < 6; k++) { s[k] = sig[y0 - 1 + k]; ng[k] = neg[(y0 - 1 + k) & 63]; }
            // neg[] has rows 0..63 only: rows -1 and 64 are never negative, mask them out
            if (y0 == 0) ng[0] = 0;
            if (y0 + 4 >= 64) ng[5] = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) { v[k] = visit[(y0 + k) & 63]; pb[k] = 0; }
            const bool full = (y0 + 4 <= h);
            const int rows = full ? 4 : (h - y0);
            // only columns with a sample that is neither significant nor already coded in this bit-plane have anything
            // to decode (elsewhere the run-length test fails and every row is skipped); the set can only shrink
            uint64_t need = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if (k < rows) need |= ~(s[k + 1] | v[k]);
            need &= wmask;
            // run-length test of a column (canUseRunLength t1.go:1195-1208) = no significant sample in the 6 x 3 window and
            // nothing coded in the column: kept as a column mask, updated when a sample turns significant
            uint64_t anym = 0;
            if (full) { const uint64_t a = s[0] | s[1] | s[2] | s[3] | s[4] | s[5]; anym = a | (a << 1) | (a >> 1) | v[0] | v[1] | v[2] | v[3]; }
            while (need) {
                const int x = __ffsll((long long)need) - 1;
                need &= need - 1;
                bool rl = false;
                int pos = 0;
                if (full) {
                    if (!((anym >> x) & 1)) {
                        if (!mq_decode(mq, ctxs, kCtxRL)) continue;
                        pos = (int)(mq_decode(mq, ctxs, kCtxUni) << 1);
                        pos |= (int)mq_decode(mq, ctxs, kCtxUni);
                        rl = true;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= rows) break;
                    bool newsig;
                    if (rl) {
                        if (k < pos) continue;
                        if (k == pos) newsig = true;
                        else {
                            uint32_t idx9 = win3(s[k], x) | (win3(s[k + 1], x) << 3) | (win3(s[k + 2], x) << 6);
                            newsig = mq_decode(mq, ctxs, kCtxZC + zc[idx9]) != 0;
                        }
                    } else {
                        if ((v[k] >> x) & 1) continue;            // visited in SPP: flag cleared below
                        if ((s[k + 1] >> x) & 1) continue;
                        uint32_t idx9 = win3(s[k], x) | (win3(s[k + 1], x) << 3) | (win3(s[k + 2], x) << 6);
                        newsig = mq_decode(mq, ctxs, kCtxZC + zc[idx9]) != 0;
                    }
                    if (newsig) {
                        pb[k] |= 1ull << x;
                        if (decode_sign(mq, ctxs, x, s[k], s[k + 1], s[k + 2], ng[k], ng[k + 1], ng[k + 2]))
                            ng[k + 1] |= 1ull << x;
                        s[k + 1] |= 1ull << x;
                        anym |= x ? (7ull << (x - 1)) : 3ull;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < rows) {
                    sig[y0 + k] = s[k + 1]; neg[y0 + k] = ng[k + 1]; visit[y0 + k] = 0;
                    if (pb[k]) plane[y0 + k] |= pb[k];
                }
            }
        }
      }
        __syncwarp(gmask);
        t1_flush_plane<OT, G>(plane, out, ostride, h, sl, bp, gmask);
    }
    __syncwarp();

    // ---- apply signs (t1.go:1282-1289) ----
    for (int y = 0; run && y < h; y++) {
        uint32_t nb = (uint32_t)(neg[y] >> (sl * W)) & (uint32_t)((1ull << W) - 1);
        OT *row = out + (size_t)y * ostride + sl * W;
        while (nb) {
            const int j = __ffs((int)nb) - 1;
            nb &= nb - 1;
            row[j] = (OT)(0 - row[j]);
        }
    }
}

// ---- ISO/IEC 15444-1 Annex D block decoder (J2KGPU_MODE_ISO) -------------------------------------------------------
// Same data structures as k_t1_ref (row bitmaps in shared memory, MQ state in registers, lanes in lock-step), with what
// the standard prescribes and the reference does not: all three passes scan 4-row stripes column by
// column, Tables D.1 - D.3 contexts, the initial states of Table D.7, and decoding stops after num_passes coding passes
// (quality layers).  Output: sign * (2 * magnitude + mid-point of the last decoded bit-plane) / 2 as an integer for the
// reversible path, or that twice-scale value * step / 2 as float32 bits for the irreversible one (the convention of the
// CPU checker, which OpenJPEG pins).  Code-block styles RESET, VCAUSAL and SEGSYM (DevCblk.pad) are decoded, PREDTERM needs
// nothing from a decoder; BYPASS / TERMALL blocks never get here (job_build refuses them).
// raw (selective bypass) bit, D.6: MSB first, the byte after an 0xFF carries 7 bits, 0xFF past the end of the segment.
// The MQ struct is reused: C = current byte, CT = bits left in it, bp = next byte.
__device__ __forceinline__ uint32_t raw_decode(MQ &m)
{
    if (m.CT == 0) {
        const uint32_t nb = m.bp < m.len ? (uint32_t)__ldg(m.d + m.bp) : 0xFFu;
        if (m.C == 0xFF) {
            if (nb > 0x8F) { m.C = 0xFF; m.CT = 8; }
            else { m.C = nb; m.bp++; m.CT = 7; }
        } else { m.C = nb; m.bp++; m.CT = 8; }
    }
    m.CT--;
    return (m.C >> m.CT) & 1u;
}

// SEG: the job has blocks with a non-default code-block style (DevCblk.pad).  BYPASS / TERMALL blocks consist of several
// codeword segments whose byte counts follow the block's data_len bytes as little-endian 32-bit words.  A separate
// instantiation, so that the kernel of the default style carries none of it (measured: 3 % on 4K EBCOT frames).
template <typename OT, int G, bool SEG>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_t1_iso(const DevCblk *__restrict__ cblks, uint32_t n, const uint8_t *__restrict__ blob, OT *__restrict__ coef,
         const float *__restrict__ steps, int irrev)
{
    J2K_DYN_SMEM(uint64_t, smem);
    constexpr int BPW = 32 / G, W = 64 / G;              // blocks per warp, columns per lane (see k_t1_ref)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / G, sl = lane % G;
    const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp * G));
    const uint32_t blk_raw = (blockIdx.x * (blockDim.x >> 5) + warp) * BPW + grp;
    const bool have = blk_raw < n;
    const uint32_t blk = have ? blk_raw : n - 1;
    const DevCblk cb = cblks[blk];
    const int w = cb.w, h = cb.h, nbps = cb.num_bps, band = cb.band & 3;
    const uint32_t style = SEG ? cb.pad : 0u;            // code-block style bits (Table A.19); the plain instantiation has none
    const bool vcausal = (style & 0x08u) != 0;
    int npasses = cb.num_passes ? cb.num_passes : 3 * nbps - 2;
    if (npasses > 3 * nbps - 2) npasses = 3 * nbps - 2;

    uint64_t *base = smem + (size_t)(warp * BPW + grp) * kIsoWords;
    uint64_t *sig = base + 1;            // rows -1 .. 64
    uint64_t *neg = base + 66;
    uint64_t *visit = base + 130;
    uint64_t *refine = base + 194;
    uint64_t *lastc = base + 258;        // coded (became significant or refined) in the bit-plane decoded last
    uint64_t *plane = base + 322;        // the bit-plane being decoded
    uint8_t *ctxs = (uint8_t *)(base + 386);

    OT *out = coef + cb.out_off;
    const uint32_t ostride = cb.out_stride;
    const bool coded = !(cb.data_len == 0 || nbps == 0 || npasses <= 0);
    if (have && !coded) {                                         // not included in any layer: all zero
        for (int y = 0; y < h; y++)
            for (int x = sl; x < w; x += G) out[(size_t)y * ostride + x] = (OT)0;
    }
    const bool run = have && coded;
    if (run) {
        for (int i = sl; i < 386; i += G) base[i] = 0;
        for (int i = sl; i < kNumCtx; i += G) ctxs[i] = (i == kCtxUni) ? 92 : (i == kCtxRL ? 6 : (i == 0 ? 8 : 0));   // Table D.7
        for (int y = 0; y < h; y++)                                    // the samples collect their magnitude bits plane by plane
#pragma unroll
            for (int j = 0; j < W; j++)
                if (sl * W + j < w) out[(size_t)y * ostride + sl * W + j] = (OT)0;
    }
    __syncwarp();

    MQ mq;
    mq_init(mq, blob + cb.data_off, (int)cb.data_len);
    const uint64_t wmask = (w >= 64) ? ~0ull : ((1ull << w) - 1);
    const uint8_t *zc = c_zc9_iso + band * 512;
    int p_end = nbps - 1;

    // pass sequence: cleanup of the top bit-plane, then (significance, refinement, cleanup) per lower bit-plane
    int bp = nbps - 1, type = 2;
    int seg = 0, seg_pos = 0, seg_left = 0;
    bool raw = false;
    for (int pass = 0; pass < npasses && run; pass++) {
        p_end = bp;
      if (sl == 0) {
        if (SEG && (style & 0x05u)) {
            raw = (style & 0x01u) && pass >= 10 && type != 2;          // D.6: raw significance / refinement passes from the 5th bit-plane on
            if (seg_left == 0) {                                       // this pass opens a codeword segment
                const uint8_t *t = blob + cb.data_off + cb.data_len + 4 * seg;
                const uint32_t sl32 = (uint32_t)__ldg(t) | ((uint32_t)__ldg(t + 1) << 8) | ((uint32_t)__ldg(t + 2) << 16) | ((uint32_t)__ldg(t + 3) << 24);
                const int left = (int)cb.data_len - seg_pos;
                const int avail = sl32 > (uint32_t)left ? left : (int)sl32;
                if (raw) { mq.d = blob + cb.data_off + seg_pos; mq.len = avail; mq.bp = 0; mq.C = 0; mq.CT = 0; }
                else mq_init(mq, blob + cb.data_off + seg_pos, avail);
                seg_pos += avail;
                seg_left = (style & 0x04u) ? 1 : (seg == 0 ? 10 : ((seg & 1) ? 2 : 1));
                seg++;
            }
            seg_left--;
        }
        if (pass && (style & 0x02u))                                   // RESET: every pass starts from Table D.7
            for (int i = 0; i < kNumCtx; i++) ctxs[i] = (i == kCtxUni) ? 92 : (i == kCtxRL ? 6 : (i == 0 ? 8 : 0));
        for (int y0 = 0; y0 < h; y0 += 4) {
            const bool full = (y0 + 4 <= h);
            const int rows = full ? 4 : (h - y0);
            uint64_t s[6], ng[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { s[k] = sig[y0 - 1 + k]; ng[k] = neg[(y0 - 1 + k) & 63]; }
            if (y0 == 0) ng[0] = 0;                               // neg[] has rows 0..63 only
            if (y0 + 4 >= 64) ng[5] = 0;
            if (vcausal) { s[5] = 0; ng[5] = 0; }                 // D.7: the stripe below does not exist for context formation
            if (type == 0) {
                // ---- significance propagation (D.3.1): insignificant samples with a significant neighbour ----
                uint64_t colmask = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= rows) break;
                    const uint64_t a = s[k], m = s[k + 1], c = s[k + 2];
                    colmask |= ~m & (a | (a << 1) | (a >> 1) | (m << 1) | (m >> 1) | c | (c << 1) | (c >> 1));
                }
                colmask &= wmask;
                if (!colmask) continue;
                uint64_t v[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0};
                while (colmask) {
                    const int x = __ffsll((long long)colmask) - 1;
                    colmask &= colmask - 1;
                    bool grew = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k >= rows) break;
                        if ((s[k + 1] >> x) & 1) continue;
                        const uint32_t idx9 = win3(s[k], x) | (win3(s[k + 1], x) << 3) | (win3(s[k + 2], x) << 6);
                        if ((idx9 & 0x1EF) == 0) continue;        // no significant neighbour (bit 4 is the sample itself)
                        if ((SEG && raw) ? raw_decode(mq) : mq_decode(mq, ctxs, kCtxZC + zc[idx9])) {
                            pb[k] |= 1ull << x;
                            if ((SEG && raw) ? raw_decode(mq)
                                             : decode_sign(mq, ctxs, x, s[k], s[k + 1], s[k + 2], ng[k], ng[k + 1], ng[k + 2], c_sc_iso))
                                ng[k + 1] |= 1ull << x;
                            s[k + 1] |= 1ull << x;
                            grew = true;
                        }
                        v[k] |= 1ull << x;
                    }
                    if (grew && x + 1 < w) colmask |= 1ull << (x + 1);    // the next column may have become a candidate
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < rows) {
                        sig[y0 + k] = s[k + 1]; neg[y0 + k] = ng[k + 1]; visit[y0 + k] = v[k];
                        if (pb[k]) { plane[y0 + k] |= pb[k]; lastc[y0 + k] |= pb[k]; }
                    }
            } else if (type == 1) {
                // ---- magnitude refinement (D.3.3): significant samples not coded in this bit-plane's first pass ----
                uint64_t cand[4], pb[4] = {0, 0, 0, 0}, rf[4], cols = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    cand[k] = k < rows ? (s[k + 1] & ~visit[(y0 + k) & 63]) : 0;
                    rf[k] = refine[(y0 + k) & 63];
                    cols |= cand[k];
                }
                if (!cols) continue;
                uint64_t nbm[4];                                  // significance is fixed during this pass: one neighbour mask per row
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t a = s[k], m = s[k + 1], c = s[k + 2];
                    nbm[k] = a | (a << 1) | (a >> 1) | c | (c << 1) | (c >> 1) | (m << 1) | (m >> 1);
                }
                while (cols) {
                    const int x = __ffsll((long long)cols) - 1;
                    cols &= cols - 1;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (!((cand[k] >> x) & 1)) continue;
                        const int ctx = kCtxMag + (((rf[k] >> x) & 1) ? 2 : (int)((nbm[k] >> x) & 1));
                        if ((SEG && raw) ? raw_decode(mq) : mq_decode(mq, ctxs, ctx)) pb[k] |= 1ull << x;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < rows && cand[k]) {
                        refine[y0 + k] = rf[k] | cand[k];
                        lastc[y0 + k] |= cand[k];
                        if (pb[k]) plane[y0 + k] |= pb[k];
                    }
            } else {
                // ---- cleanup (D.3.4) ----
                uint64_t v[4], pb[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { v[k] = visit[(y0 + k) & 63]; pb[k] = 0; }
                // only columns with a sample that is neither significant nor coded in this bit-plane have work left
                uint64_t need = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) if (k < rows) need |= ~(s[k + 1] | v[k]);
                need &= wmask;
                // run-length columns: no significant sample in the 6 x 3 window and nothing coded in the column yet
                uint64_t anym = 0;
                if (full) { const uint64_t a = s[0] | s[1] | s[2] | s[3] | s[4] | s[5]; anym = a | (a << 1) | (a >> 1) | v[0] | v[1] | v[2] | v[3]; }
                while (need) {
                    const int x = __ffsll((long long)need) - 1;
                    need &= need - 1;
                    bool rl = false;
                    int pos = 0;
                    if (full) {
                        if (!((anym >> x) & 1)) {
                            if (!mq_decode(mq, ctxs, kCtxRL)) continue;
                            pos = (int)(mq_decode(mq, ctxs, kCtxUni) << 1);
                            pos |= (int)mq_decode(mq, ctxs, kCtxUni);
                            rl = true;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k >= rows) break;
                        bool newsig;
                        if (rl) {
                            if (k < pos) continue;
                            if (k == pos) newsig = true;
                            else {
                                const uint32_t idx9 = win3(s[k], x) | (win3(s[k + 1], x) << 3) | (win3(s[k + 2], x) << 6);
                                newsig = mq_decode(mq, ctxs, kCtxZC + zc[idx9]) != 0;
                            }
                        } else {
                            if ((v[k] >> x) & 1) continue;
                            if ((s[k + 1] >> x) & 1) continue;
                            const uint32_t idx9 = win3(s[k], x) | (win3(s[k + 1], x) << 3) | (win3(s[k + 2], x) << 6);
                            newsig = mq_decode(mq, ctxs, kCtxZC + zc[idx9]) != 0;
                        }
                        if (newsig) {
                            pb[k] |= 1ull << x;
                            if (decode_sign(mq, ctxs, x, s[k], s[k + 1], s[k + 2], ng[k], ng[k + 1], ng[k + 2], c_sc_iso))
                                ng[k + 1] |= 1ull << x;
                            s[k + 1] |= 1ull << x;
                            anym |= x ? (7ull << (x - 1)) : 3ull;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k < rows) {
                        sig[y0 + k] = s[k + 1]; neg[y0 + k] = ng[k + 1]; visit[y0 + k] = 0;
                        if (pb[k]) { plane[y0 + k] |= pb[k]; lastc[y0 + k] |= pb[k]; }
                    }
            }
        }
        if (type == 2 && (style & 0x20u))                           // SEGSYM: four UNIFORM decisions (1010) close a cleanup pass (D.5)
            for (int i = 0; i < 4; i++) (void)mq_decode(mq, ctxs, kCtxUni);
        if (type == 2 && pass + 1 < npasses)                        // the next bit-plane starts: nothing coded in it yet
            for (int y = 0; y < h; y++) lastc[y] = 0;
      }
        if (type == 2 || pass + 1 == npasses) {                     // the bit-plane is complete (or the block's passes end in it)
            __syncwarp(gmask);
            t1_flush_plane<OT, G>(plane, out, ostride, h, sl, bp, gmask);
        }
        if (type == 2) bp--;
        type = type == 2 ? 0 : type + 1;
    }
    __syncwarp();

    // ---- final values: twice-scale magnitude with the mid-point of the last decoded bit-plane, sign, dequantisation ----
    const float hstep = irrev ? 0.5f * steps[blk] : 0.0f;
    typedef typename std::conditional<sizeof(OT) == 2, uint16_t, uint32_t>::type UT;
    for (int y = 0; run && y < h; y++) {
        uint32_t sb = (uint32_t)(sig[y] >> (sl * W)) & (uint32_t)((1ull << W) - 1);
        const uint32_t nb = (uint32_t)(neg[y] >> (sl * W)), lb = (uint32_t)(lastc[y] >> (sl * W));
        OT *row = out + (size_t)y * ostride + sl * W;
        while (sb) {
            const int j = __ffs((int)sb) - 1;
            sb &= sb - 1;
            const uint32_t m = (uint32_t)*reinterpret_cast<UT *>(row + j);
            const uint32_t m2 = (m << 1) | (1u << (((lb >> j) & 1) ? p_end : p_end + 1));
            const bool ngt = (nb >> j) & 1;
            if (irrev) {
                const float f = (float)m2 * hstep;
                row[j] = (OT)__float_as_int(ngt ? -f : f);
            } else {
                const uint32_t v = m2 >> 1;
                row[j] = (OT)(int32_t)(ngt ? 0u - v : v);
            }
        }
    }
}

// ---- host-side table construction -------------------------------------------------------------------
// ISO/IEC 15444-1 Table C.2 rows (Qe, NMPS, NLPS, SWITCH); the reference's 94-entry table
// (mqc.go:21-116) is this machine expanded to index 2*state + mps, entry 46 being UNI.
const uint16_t kQe[47] = {
    0x5601, 0x3401, 0x1801, 0x0AC1, 0x0521, 0x0221, 0x5601, 0x5401, 0x4801, 0x3801, 0x3001, 0x2401,
    0x1C01, 0x1601, 0x5601, 0x5401, 0x5101, 0x4801, 0x3801, 0x3401, 0x3001, 0x2801, 0x2401, 0x2201,
    0x1C01, 0x1801, 0x1601, 0x1401, 0x1201, 0x1101, 0x0AC1, 0x09C1, 0x08A1, 0x0521, 0x0441, 0x02A1,
    0x0221, 0x0141, 0x0111, 0x0085, 0x0049, 0x0025, 0x0015, 0x0009, 0x0005, 0x0001, 0x5601};
const uint8_t kNmps[47] = {1, 2, 3, 4, 5, 38, 7, 8, 9, 10, 11, 12, 13, 29, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
                           25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 45, 46};
const uint8_t kNlps[47] = {1, 6, 9, 12, 29, 33, 6, 14, 14, 14, 17, 18, 20, 21, 14, 14, 15, 16, 17, 18, 19, 19, 20, 21,
                           22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 46};
const uint8_t kSwitch[47] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1};

int zc_context(int band, int hc, int vc, int dc)                 // t1_luts.go:52-107
{
    if (band == J2KGPU_BAND_HH) {
        int hv = hc + vc;
        if (hv >= 3) return 8;
        if (hv == 2) return dc >= 2 ? 7 : (dc >= 1 ? 6 : 5);
        if (hv == 1) return dc >= 2 ? 4 : 3;
        return dc >= 2 ? 2 : (dc >= 1 ? 1 : 0);
    }
    if (band == J2KGPU_BAND_HL) { int t = hc; hc = vc; vc = t; }
    if (hc == 2) return 8;
    if (hc == 1) return vc >= 1 ? 7 : (dc >= 1 ? 6 : 5);
    if (vc == 2) return 4;
    if (vc == 1) return dc >= 1 ? 3 : 2;
    return dc >= 2 ? 1 : 0;
}

// __constant__ symbols exist once per device: the tables are uploaded once per device of the process, on the launching
// stream (ordered before the first kernel that reads them; later launches on other streams of the same device happen
// after this call has returned, and the copies out of the pageable host arrays below are staged before it returns --
// the final synchronise makes the first upload visible to every stream)
cudaError_t upload_tables(cudaStream_t s)
{
    static bool done[64] = {};
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (done[dev]) return cudaSuccess;
    uint32_t mq[94];
    for (int i = 0; i < 47; i++)
        for (int m = 0; m < 2; m++) {
            uint32_t nm = 2u * kNmps[i] + m, nl = 2u * kNlps[i] + (m ^ kSwitch[i]);
            mq[2 * i + m] = kQe[i] | (nm << 16) | (nl << 24);
        }
    uint8_t zc[4 * 512];
    for (int band = 0; band < 4; band++)
        for (int i = 0; i < 512; i++) {
            // window bit layout: row (up, mid, down) * 3 + column (x-1, x, x+1)
            int nw = i & 1, n = (i >> 1) & 1, ne = (i >> 2) & 1, w = (i >> 3) & 1, e = (i >> 5) & 1,
                sw = (i >> 6) & 1, s = (i >> 7) & 1, se = (i >> 8) & 1;
            zc[band * 512 + i] = (uint8_t)zc_context(band, w + e, n + s, nw + ne + sw + se);
        }
    uint8_t sc[256];
    for (int i = 0; i < 256; i++) {                                // t1.go:392-457
        int hc = 0, vc = 0;
        if (i & 1)  hc += (i & 2) ? -1 : 1;
        if (i & 4)  hc += (i & 8) ? -1 : 1;
        if (i & 16) vc += (i & 32) ? -1 : 1;
        if (i & 64) vc += (i & 128) ? -1 : 1;
        int pred = 0, ctx = 0;
        if (hc < 0) { pred = 1; hc = -hc; }
        if (hc == 0 && vc < 0) { pred = 1; vc = -vc; }
        if (hc == 1) ctx = vc == 1 ? 4 : (vc == 0 ? 2 : 1);
        else if (hc == 0) ctx = vc == 1 ? 1 : 0;
        else if (hc == 2) ctx = 3;
        sc[i] = (uint8_t)((ctx << 1) | pred);
    }
    // ISO/IEC 15444-1 Table D.1 (zero coding) and Tables D.2 / D.3 (sign coding) in the same index formats
    uint8_t zci[4 * 512], sci[256];
    for (int band = 0; band < 4; band++)
        for (int i = 0; i < 512; i++) {
            int nw = i & 1, n = (i >> 1) & 1, ne = (i >> 2) & 1, w = (i >> 3) & 1, e = (i >> 5) & 1,
                sw = (i >> 6) & 1, s = (i >> 7) & 1, se = (i >> 8) & 1;
            int hc = w + e, vc = n + s, dc = nw + ne + sw + se, ctx;
            if (band == J2KGPU_BAND_HH) {
                int hv = hc + vc;
                if (dc >= 3) ctx = 8;
                else if (dc == 2) ctx = hv >= 1 ? 7 : 6;
                else if (dc == 1) ctx = hv >= 2 ? 5 : (hv == 1 ? 4 : 3);
                else ctx = hv >= 2 ? 2 : (hv == 1 ? 1 : 0);
            } else {
                if (band == J2KGPU_BAND_HL) { int t = hc; hc = vc; vc = t; }
                if (hc == 2) ctx = 8;
                else if (hc == 1) ctx = vc >= 1 ? 7 : (dc >= 1 ? 6 : 5);
                else if (vc == 2) ctx = 4;
                else if (vc == 1) ctx = 3;
                else ctx = dc >= 2 ? 2 : (dc == 1 ? 1 : 0);
            }
            zci[band * 512 + i] = (uint8_t)ctx;
        }
    for (int i = 0; i < 256; i++) {
        int hc = 0, vc = 0;
        if (i & 1)  hc += (i & 2) ? -1 : 1;
        if (i & 4)  hc += (i & 8) ? -1 : 1;
        if (i & 16) vc += (i & 32) ? -1 : 1;
        if (i & 64) vc += (i & 128) ? -1 : 1;
        hc = hc > 1 ? 1 : (hc < -1 ? -1 : hc);
        vc = vc > 1 ? 1 : (vc < -1 ? -1 : vc);
        int flip = 0;
        if (hc < 0 || (hc == 0 && vc < 0)) { flip = 1; hc = -hc; vc = -vc; }
        int ctx = hc == 1 ? (vc == 1 ? 4 : (vc == 0 ? 3 : 2)) : (vc == 1 ? 1 : 0);      // contexts 9..13, minus 9
        sci[i] = (uint8_t)((ctx << 1) | flip);
    }
    if ((e = cudaMemcpyToSymbolAsync(c_zc9_iso, zci, sizeof zci, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_sc_iso, sci, sizeof sci, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_mq, mq, sizeof mq, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_zc9, zc, sizeof zc, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_sc, sc, sizeof sc, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    done[dev] = true;
    return cudaSuccess;
}

}  // namespace

// warps per CTA and dynamic shared memory for 32 / G blocks per warp of `block_words` 64-bit words each: as many warps
// (at most kWarpsPerCta) as leave room for two CTAs per SM; 0 when even one warp does not fit (the caller takes a wider group)
static int t1_warps_per_cta(int G, size_t block_words, size_t *smem)
{
    const size_t per_warp = (size_t)(32 / G) * block_words * sizeof(uint64_t);
    int wpc = kWarpsPerCta;
    while (wpc > 1 && (size_t)wpc * per_warp > 48 * 1024) wpc >>= 1;
    *smem = (size_t)wpc * per_warp;
    return *smem <= 220 * 1024 ? wpc : 0;
}

template <typename OT, int G>
static cudaError_t launch_t1_ref_g(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, OT *d_coef, int skip_empty, cudaStream_t s)
{
    size_t smem;
    const int wpc = t1_warps_per_cta(G, kRefWords, &smem);
    if (!wpc) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_t1_ref<OT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t per_cta = (uint32_t)wpc * (32 / G);
    J2K_LAUNCH((k_t1_ref<OT, G>), (n + per_cta - 1) / per_cta, wpc * 32, smem, s, d_cblks, n, d_blob, d_coef, skip_empty);
    return cudaGetLastError();
}

template <typename OT, int G, bool SEG>
static cudaError_t launch_t1_iso_g(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, OT *d_coef, const float *d_steps, int irrev,
                                   cudaStream_t s)
{
    size_t smem;
    const int wpc = t1_warps_per_cta(G, kIsoWords, &smem);
    if (!wpc) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_t1_iso<OT, G, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t per_cta = (uint32_t)wpc * (32 / G);
    J2K_LAUNCH((k_t1_iso<OT, G, SEG>), (n + per_cta - 1) / per_cta, wpc * 32, smem, s, d_cblks, n, d_blob, d_coef, d_steps, irrev);
    return cudaGetLastError();
}

// Lanes per block.  A chain issues an instruction every 8 cycles or so (dependent latency), so an SM wants as many chains
// in flight as it can hold; registers allow 25 warps, shared memory about 80 blocks.  With few blocks per SM one chain
// per warp finishes first (no chain waits for a diverged neighbour): measured on B200, ms per launch for 32 / 16 / 8 / 4
// lanes per block: 1 536 blocks 6.6 / 8.7 / 11.9 / 16.8; 12 240 blocks 22.9 / 16.3 / 18.7 / 23.4; 14 010 blocks
// 44.0 / 33.1 / 32.8 / 36.5.
static int t1_group(int opt, uint32_t n)
{
    if (opt == 4 || opt == 8 || opt == 16 || opt == 32) return opt;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t per_sm = n / (uint32_t)(sms > 0 ? sms : 148);
    return per_sm < 28 ? 32 : (per_sm < 120 ? 16 : 8);
}

static cudaError_t launch_t1_ref_impl(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                                      int max_bps, int skip_empty, int group, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    cudaError_t e = upload_tables(s);
    if (e != cudaSuccess) return e;
    (void)max_bps;
#define J2K_T1_REF(G) (coef16 ? launch_t1_ref_g<int16_t, G>(d_cblks, n, d_blob, (int16_t *)d_coef, skip_empty, s) \
                              : launch_t1_ref_g<int32_t, G>(d_cblks, n, d_blob, (int32_t *)d_coef, skip_empty, s))
    switch (t1_group(group, n)) {
    case 4: return J2K_T1_REF(4);
    case 8: return J2K_T1_REF(8);
    case 16: return J2K_T1_REF(16);
    default: return J2K_T1_REF(32);
    }
#undef J2K_T1_REF
}

cudaError_t launch_t1_ref(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          int max_bps, int group, cudaStream_t s)
{
    return launch_t1_ref_impl(d_cblks, n, d_blob, d_coef, coef16, max_bps, 1, group, s);
}

// stage API form: T1.Decode is also defined for empty data (decodes the 0xFF fill, mqc.go:387-388)
cudaError_t launch_t1_ref_stage(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, int32_t *d_coef,
                                int max_bps, int group, cudaStream_t s)
{
    return launch_t1_ref_impl(d_cblks, n, d_blob, d_coef, 0, max_bps, 0, group, s);
}

// ISO/IEC 15444-1 Annex D decoder (J2KGPU_MODE_ISO): num_bps magnitude bit-planes, num_passes coding passes per block
// (0 = all); irrev: the planes receive float32 bits = value * steps[block]
cudaError_t launch_t1_iso(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          const float *d_steps, int irrev, int segmented, int group, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    cudaError_t e = upload_tables(s);
    if (e != cudaSuccess) return e;
#define J2K_T1_ISO_S(G, S) ((coef16 && !irrev) ? launch_t1_iso_g<int16_t, G, S>(d_cblks, n, d_blob, (int16_t *)d_coef, d_steps, irrev, s) \
                                               : launch_t1_iso_g<int32_t, G, S>(d_cblks, n, d_blob, (int32_t *)d_coef, d_steps, irrev, s))
#define J2K_T1_ISO(G) (segmented ? J2K_T1_ISO_S(G, true) : J2K_T1_ISO_S(G, false))
    switch (t1_group(group, n)) {
    case 4: return J2K_T1_ISO(4);
    case 8: return J2K_T1_ISO(8);
    case 16: return J2K_T1_ISO(16);
    default: return J2K_T1_ISO(32);
    }
#undef J2K_T1_ISO
#undef J2K_T1_ISO_S
}
