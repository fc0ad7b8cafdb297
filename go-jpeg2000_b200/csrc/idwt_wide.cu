// idwt_wide.cu -- the fast path of the fused "IDWT levels 1+0 + inverse RCT + DC shift + clamp + RGBA8 pack" kernel
// for the common case (3 unsigned 8-bit components, reversible 5-3 + RCT, tile width a multiple of 16).
//
// Same arithmetic and the same level-1 -> level-0 in-lane hand-over as idwt_fused.cu (see its header; reference
// dwt.go:122-147, 410-429, 534-548, mct.go:56-66, 113-118, decoder.go:417-588), but organised to cut the
// instruction count per pixel, which is what bounds that kernel (ncu: issue-bound at 2.9 warp-instructions/pixel):
//   * a lane owns SIXTEEN output columns (level 0: 8 L + 8 H columns per band row; level 1: 4 L + 4 H), so the two
//     warp shuffles of a horizontal lifting pass, every address and every copy / load instruction are amortised over
//     4x more samples, and all global / shared accesses are 16 bytes wide;
//   * a tile row of up to 512 columns is ONE warp: no halo lanes at all (tile edges are the symmetric-extension
//     edges).  Wider tiles use 30 owned lanes + 1 halo lane each side (one lane is enough at 16 columns per lane);
//   * the three components are processed one after the other, so only one component's transient rows are live;
//     the finished rows of components 0 and 1 wait in a warp-private shared-memory stage for component 2, then the
//     epilogue runs per 4 pixels (two LDS.128, 8 ALU per pixel, one 16-byte streaming store);
//   * band rows arrive through warp-private cp.async slots, one row pair ahead per component (each lane reads back
//     only what it copied itself: no barrier anywhere in the kernel);
//   * vertical lifting state (one row pair of delay) lives in registers: 144 per lane; the kernel runs 8 warps / SM.
// Eligibility (host): fast RGBA8 epilogue conditions of idwt_fused.cu + every tile width % 16 == 0, height % 4 == 0.
#include "common.h"
#include <cstdlib>
#include "tail.cuh"

namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ int even_upd(int x, int l, int r)      // x -= (l + r + 2) >> 2   (dwt.go:132-138)
{
    return (int)((uint32_t)x - (uint32_t)((int)((uint32_t)l + (uint32_t)r + 2u) >> 2));
}
__device__ __forceinline__ int odd_upd(int x, int l, int r)       // x += (l + r) >> 1       (dwt.go:141-143)
{
    return (int)((uint32_t)x + (uint32_t)((int)((uint32_t)l + (uint32_t)r) >> 1));
}
__device__ __forceinline__ int odd_last(int x, int l)             // x += l                  (dwt.go:144-146)
{
    return (int)((uint32_t)x + (uint32_t)l);
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
#ifdef J2K_EMU
    memcpy(smem_dst, gsrc, BYTES);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc), "n"(BYTES) : "memory");
#endif
}
__device__ __forceinline__ void cp_commit()
{
#ifndef J2K_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_wait()
{
#ifndef J2K_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
// d = sat_u8(b) | sat_u8(a) << 8 | c << 16
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c)
{
#ifdef J2K_EMU
    const uint32_t sa = (uint32_t)(a < 0 ? 0 : (a > 255 ? 255 : a)), sb = (uint32_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
    return sb | (sa << 8) | (c << 16);
#else
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

// ---- N consecutive elements of a plane (N = 4 or 8) <-> one lane's 32-byte shared-memory slot -------------------
// The slot is two uint4 at [0] and [32] (stride 32 uint4 = one warp row), so that a warp's LDS.128 is conflict free.
template <int N> struct Part;       // N elements per lane and part

// copy N elements of type CT from global memory into the lane's slot
template <int N>
__device__ __forceinline__ void part_copy(uint4 *slot, const int32_t *g)
{
    cp_async<16>(slot, g);
    if (N == 8) cp_async<16>(slot + 32, g + 4);
}
template <int N>
__device__ __forceinline__ void part_copy(uint4 *slot, const int16_t *g)
{
    if (N == 8) cp_async<16>(slot, g);
    else cp_async<8>(slot, g);
}
// read them back as int
template <int N>
__device__ __forceinline__ void part_read(const uint4 *slot, int32_t, int *v)
{
    const uint4 a = slot[0];
    v[0] = (int)a.x; v[1] = (int)a.y; v[2] = (int)a.z; v[3] = (int)a.w;
    if (N == 8) {
        const uint4 b = slot[32];
        v[4] = (int)b.x; v[5] = (int)b.y; v[6] = (int)b.z; v[7] = (int)b.w;
    }
}
__device__ __forceinline__ void unpack2(uint32_t r, int &a, int &b) { a = (int)(int16_t)(r & 0xFFFFu); b = (int)r >> 16; }
template <int N>
__device__ __forceinline__ void part_read(const uint4 *slot, int16_t, int *v)
{
    if (N == 8) {
        const uint4 a = slot[0];
        unpack2(a.x, v[0], v[1]); unpack2(a.y, v[2], v[3]); unpack2(a.z, v[4], v[5]); unpack2(a.w, v[6], v[7]);
    } else {
        const uint2 a = *reinterpret_cast<const uint2 *>(slot);
        unpack2(a.x, v[0], v[1]); unpack2(a.y, v[2], v[3]);
    }
}
// direct (prologue) load of N elements
template <int N>
__device__ __forceinline__ void part_ldg(const int32_t *g, int *v)
{
    const int4 a = __ldg(reinterpret_cast<const int4 *>(g));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    if (N == 8) {
        const int4 b = __ldg(reinterpret_cast<const int4 *>(g + 4));
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
template <int N>
__device__ __forceinline__ void part_ldg(const int16_t *g, int *v)
{
    if (N == 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(g));
        unpack2(a.x, v[0], v[1]); unpack2(a.y, v[2], v[3]); unpack2(a.z, v[4], v[5]); unpack2(a.w, v[6], v[7]);
    } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(g));
        unpack2(a.x, v[0], v[1]); unpack2(a.y, v[2], v[3]);
    }
}

// horizontal synthesis of one row held in band order (V[0..N) = L columns, V[N..2N) = H columns of this lane), in
// place -> 2N interleaved samples.  Every lane of the warp takes part in the two shuffles.
template <int N>
__device__ __forceinline__ void hsynth(int *V, bool first, bool last)
{
    int E[N];
    int left = __shfl_up_sync(0xffffffffu, V[2 * N - 1], 1);
    left = first ? V[N] : left;                                        // x[0] -= (x[1] + x[1] + 2) >> 2
#pragma unroll
    for (int i = 0; i < N; i++) E[i] = even_upd(V[i], i ? V[N + i - 1] : left, V[N + i]);
    const int right = __shfl_down_sync(0xffffffffu, E[0], 1);
    int O[N];
#pragma unroll
    for (int i = 0; i < N - 1; i++) O[i] = odd_upd(V[N + i], E[i], E[i + 1]);
    {
        const int t = odd_upd(V[2 * N - 1], E[N - 1], right);
        O[N - 1] = last ? odd_last(V[2 * N - 1], E[N - 1]) : t;        // last odd of an even-length line: x += x[n-2]
    }
#pragma unroll
    for (int i = 0; i < N; i++) { V[2 * i] = E[i]; V[2 * i + 1] = O[i]; }
}

// shared memory of one warp (uint4 units), NC components
__host__ __device__ constexpr int ring0_u4(int nc) { return nc * 4 * 64; }     // [comp][part loL, loH, hiL, hiH][2 halves][32 lanes]
__host__ __device__ constexpr int ring1_u4(int nc) { return nc * 4 * 32; }     // [comp][part][32 lanes]
__host__ __device__ constexpr int stage_u4(int nc) { return nc == 3 ? 2 * 2 * 4 * 32 : 0; }   // [comp 0,1][row even, odd][quad][32 lanes]
__host__ __device__ constexpr int warp_smem_u4(int nc) { return ring0_u4(nc) + ring1_u4(nc) + stage_u4(nc); }   // 3 components: 26 KB per warp
// ISO: the odd level-1 output row is wanted one step later; it waits in ring0's "lo L" part, which ISO never copies into
// when level 1 exists (that part IS the level-1 output)

// NC = 3: RGB + RCT -> RGBA8.  NC = 1: one unsigned component of 8 or 16 bits -> Gray8 / big-endian Gray16.
// RGB24 (NC = 3, host-buffer runs): the pixels are written as packed R G B, 3 bytes each, row stride 3/4 of the tile
// table's -- the alpha byte of an RGBA8 image is the constant 255, so it does not cross the PCIe link; the host side
// of the library widens the rows to RGBA8 while the next chunk is in flight (api.cu).
template <int NC, typename CT, bool ISO, bool RGB24>
__global__ void __launch_bounds__(kWarps * 32, 2)
k_idwt53_wide(const DevTileComp *__restrict__ tcs, const DevTile *__restrict__ tiles, const CT *__restrict__ coef,
              const int32_t *__restrict__ tmp, uint8_t *__restrict__ pix, int nlevels, int strip_pairs, int prec, int iso_pack)
{
    constexpr int kRing0 = ring0_u4(NC), kRing1 = ring1_u4(NC), kWarpSmem = warp_smem_u4(NC);
    J2K_DYN_SMEM(uint4, smem_all);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DevTile tile = tiles[blockIdx.y];
    const int w = (int)tile.w, h = (int)tile.h;
    const int nlx = w >> 1, nly = h >> 1, nl = w >> 4;            // nl = lanes' worth of columns in a row
    const int halo = nl > 32 ? 1 : 0, own = 32 - 2 * halo;
    const int nwx = (nl + own - 1) / own;
    const int nstrips = (nly + strip_pairs - 1) / strip_pairs;
    const int unit = blockIdx.x * kWarps + warp;
    if (unit >= nwx * nstrips) return;
    const int strip = unit / nwx, wi = unit - strip * nwx;
    const int lq = wi * own - halo + lane;
    const bool lvalid = lq >= 0 && lq < nl;
    const bool store_lane = lvalid && lane >= halo && lane < 32 - halo;
    const bool first = lq == 0, last = lq == nl - 1;
    const int lc = lvalid ? lq : 0;
    const int ka = strip * strip_pairs, kb = min(ka + strip_pairs, nly);
    const int rlast = min(kb, nly - 1);
    const bool l1on = nlevels >= 2, l2on = nlevels >= 3;
    const int nlx1 = nlx >> 1, nly1 = nly >> 1;
    const int half0 = nly >> 1;

    const CT *plane[NC];
    const int32_t *prev2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DevTileComp tc = tcs[tile.tc[c]];
        plane[c] = coef + tc.coef_off;
        prev2[c] = tmp + tc.tmp_off;                  // level 2 wrote ping-pong buffer 0
        J2K_OPAQUE_PTR(plane[c]);
        J2K_OPAQUE_PTR(prev2[c]);
    }
    uint4 *ring0 = smem_all + (size_t)warp * kWarpSmem + lane;
    uint4 *ring1 = ring0 + kRing0;
    uint4 *stage = ring1 + kRing1;
    uint4 *stash = ring0;                             // [c * 256 + half * 32]
    const uint32_t uw = (uint32_t)w;
    const uint32_t col0 = 8u * (uint32_t)lc;          // level-0 L columns 8lq..8lq+7; H columns at nlx + col0
    const uint32_t col1 = 4u * (uint32_t)lc;          // level-1 L columns 4lq..4lq+3; H columns at nlx1 + col1

    // ---- copies into the warp-private slots ------------------------------------------------------------------------
    // level-0 band row pair r of component c: lo = row r (unless it is level-1 output), hi = row nly + r
    auto issue_l0 = [&](int c, int r) {
        const bool fl = l1on && (ISO || r < half0), fh = l1on && !ISO && r < half0;
        const uint32_t o0 = (uint32_t)r * uw + col0, o2 = o0 + (uint32_t)nly * uw;
        uint4 *sl = ring0 + c * 256;
        if (!fl) part_copy<8>(sl, plane[c] + o0);
        if (!fh) part_copy<8>(sl + 64, plane[c] + o0 + nlx);
        part_copy<8>(sl + 128, plane[c] + o2);
        part_copy<8>(sl + 192, plane[c] + o2 + nlx);
    };
    // level-1 band row pair j: lo = row j (its L / H part may be the level-2 output), hi = row nly1 + j
    auto l1_src = [&](int j, uint32_t &o_lo, uint32_t &o_hi, uint32_t &o_p, bool &pL, bool &pH) {
        if (ISO) {
            o_lo = (uint32_t)j * uw + col1; o_hi = (uint32_t)(nly1 + j) * uw + col1;
            o_p = (uint32_t)j * (uint32_t)nlx1 + col1;
            pL = l2on; pH = false;
        } else {
            o_lo = (uint32_t)j * (uint32_t)nlx + col1; o_hi = (uint32_t)(nly1 + j) * (uint32_t)nlx + col1;
            o_p = o_lo;
            pL = l2on && 2 * j + 1 <= nly1; pH = l2on && 2 * j + 2 <= nly1;
        }
    };
    auto issue_l1 = [&](int c, int j) {
        uint32_t o_lo, o_hi, o_p; bool pL, pH;
        l1_src(j, o_lo, o_hi, o_p, pL, pH);
        uint4 *sl = ring1 + c * 128;
        if (pL) part_copy<4>(sl, prev2[c] + o_p); else part_copy<4>(sl, plane[c] + o_lo);
        if (pH) part_copy<4>(sl + 32, prev2[c] + o_p + nlx1); else part_copy<4>(sl + 32, plane[c] + o_lo + nlx1);
        part_copy<4>(sl + 64, plane[c] + o_hi);
        part_copy<4>(sl + 96, plane[c] + o_hi + nlx1);
    };
    auto read_l1 = [&](int c, int j, int lo[8], int hi[8]) {            // band order: [0..4) L, [4..8) H
        uint32_t o_lo, o_hi, o_p; bool pL, pH;
        l1_src(j, o_lo, o_hi, o_p, pL, pH);
        const uint4 *sl = ring1 + c * 128;
        if (pL) part_read<4>(sl, int32_t(), lo); else part_read<4>(sl, CT(), lo);
        if (pH) part_read<4>(sl + 32, int32_t(), lo + 4); else part_read<4>(sl + 32, CT(), lo + 4);
        part_read<4>(sl + 64, CT(), hi);
        part_read<4>(sl + 96, CT(), hi + 4);
    };

    int hp[NC][16], ep[NC][16];          // level 0: Hi[k] and E[k] of the lane's 16 columns (REF: band order; ISO: interleaved)
    // level 1: hi1[jn] and E1[jn] of the lane's 8 columns.  ISO keeps them in registers.  REF needs level 1 only while
    // the strip is in the top half of the tile, where ring0's two "lo" parts are never copied into (they ARE the
    // level-1 output): the state lives there (4 uint4 per lane and component) and 48 registers are saved -- without
    // this the REF variants spill their loop counters (ncu: 26 % of the stall samples were the spill reloads)
    int h1[NC][8], e1[NC][8];
    auto l1_get = [&](int c, int *hh, int *ee) {
        if (ISO) {
#pragma unroll
            for (int j = 0; j < 8; j++) { hh[j] = h1[c][j]; ee[j] = e1[c][j]; }
        } else {
            const uint4 *st = ring0 + c * 256;
            part_read<8>(st, int32_t(), hh);
            part_read<8>(st + 64, int32_t(), ee);
        }
    };
    auto l1_put = [&](int c, const int *hh, const int *ee) {
        if (ISO) {
#pragma unroll
            for (int j = 0; j < 8; j++) { h1[c][j] = hh[j]; e1[c][j] = ee[j]; }
        } else {
            uint4 *st = ring0 + c * 256;
            st[0] = make_uint4((uint32_t)hh[0], (uint32_t)hh[1], (uint32_t)hh[2], (uint32_t)hh[3]);
            st[32] = make_uint4((uint32_t)hh[4], (uint32_t)hh[5], (uint32_t)hh[6], (uint32_t)hh[7]);
            st[64] = make_uint4((uint32_t)ee[0], (uint32_t)ee[1], (uint32_t)ee[2], (uint32_t)ee[3]);
            st[96] = make_uint4((uint32_t)ee[4], (uint32_t)ee[5], (uint32_t)ee[6], (uint32_t)ee[7]);
        }
    };
    int jn = 0, j1last = -1;           // next level-1 pair to emit; last pair this strip needs

    // ---- prologue ------------------------------------------------------------------------------------------------
    const int k0 = ka - 1;
    bool l1need = false;
    if (l1on) {
        if (ISO) { l1need = true; jn = (ka >> 1) - 1; j1last = rlast >> 1; }
        else if (ka < half0) { l1need = true; jn = ka - 1; j1last = min(rlast, half0 - 1); }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {                                      // group c: first level-0 row pair (+ level-1 pair jn+1)
        if (lvalid) {
            issue_l0(c, k0 + 1);
            if (l1need) issue_l1(c, jn + 1);
        }
        cp_commit();
    }
    if (l1need) {                                                      // hi1[max(jn, 0)] -> h1; then one silent step builds E1
        const int jp = jn < 0 ? 0 : jn;
        const uint32_t o_hi = ISO ? (uint32_t)(nly1 + jp) * uw + col1 : (uint32_t)(nly1 + jp) * (uint32_t)nlx + col1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int hh[8], ee[8];
            part_ldg<4>(plane[c] + o_hi, &hh[0]);
            part_ldg<4>(plane[c] + o_hi + nlx1, &hh[4]);
            if (ISO) hsynth<4>(hh, first, last);
#pragma unroll
            for (int j = 0; j < 8; j++) ee[j] = 0;
            l1_put(c, hh, ee);
        }
    }
    {
        const int rp = k0 < 0 ? 0 : k0;                                // top edge: Hi[-1] := Hi[0]
        const uint32_t ohi = (uint32_t)(nly + rp) * uw + col0;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            part_ldg<8>(plane[c] + ohi, &hp[c][0]);
            part_ldg<8>(plane[c] + ohi + nlx, &hp[c][8]);
            if (ISO) hsynth<8>(hp[c], first, last);
#pragma unroll
            for (int j = 0; j < 16; j++) ep[c][j] = 0;
        }
    }

    // one level-1 step of component c: consumes band row pair jn+1 (in ring1), leaves output rows 2jn (E) and 2jn+1 (O),
    // interleaved, in rE / rO; refills the slot with pair jn+2 when the strip still needs it
    auto l1_step = [&](int c, int rE[8], int rO[8]) {
        int lo[8], hi[8], hh[8];
        read_l1(c, jn + 1, lo, hi);
        if (lvalid && (jn + 1 <= j1last) && (jn + 2 < nly1)) issue_l1(c, jn + 2);
        const bool inner = jn + 1 < nly1;
        if (ISO) { hsynth<4>(lo, first, last); hsynth<4>(hi, first, last); }
        l1_get(c, hh, rE);                                             // rE = E1[jn] is output row 2jn
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int e = even_upd(lo[j], hh[j], hi[j]);
            rO[j] = inner ? odd_upd(hh[j], rE[j], e) : odd_last(hh[j], rE[j]);   // bottom edge
            lo[j] = e;
        }
        l1_put(c, hi, lo);                                             // hi1[jn+1], E1[jn+1]
        if (!ISO) { hsynth<4>(rE, first, last); hsynth<4>(rO, first, last); }
    };

    if (l1need) {                                                      // the silent level-1 step (outputs discarded)
#pragma unroll
        for (int c = 0; c < NC; c++) {
            cp_wait<NC - 1>();
            int a[8], b[8];
            l1_step(c, a, b);
            cp_commit();
        }
        jn++;
    }

    // ---- output addressing ------------------------------------------------------------------------------------------
    const int bpp = NC == 3 ? (RGB24 ? 3 : 4) : (prec > 8 ? 2 : 1);
    const size_t ostride = RGB24 ? (size_t)(tile.out_stride >> 2) * 3 : (size_t)tile.out_stride;
    uint8_t *orow = pix + tile.out_off + (size_t)(tile.img_y0 + 2u * (uint32_t)ka) * ostride +
                    (size_t)bpp * (tile.img_x0 + 16u * (uint32_t)lc);

    // ---- stream the strip: step k consumes band row pair k+1 and finishes output rows 2k (even) and 2k+1 (odd) ------
    for (int k = k0; k < kb; k++) {
        const int r = k + 1;
        const bool inner = r < nly;                                    // false only below the last row pair of the tile
        const bool fl = inner && l1on && (ISO || r < half0), fh = inner && l1on && !ISO && r < half0;
        const bool do_l1 = fl && jn <= (ISO ? (r >> 1) : r);           // a new level-1 pair is due (warp-uniform)
        const bool emit = k >= ka;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            cp_wait<NC - 1>();                                         // component c's copies of the previous step have landed
            int lo[16], hi[16];
            if (fl) {
                int rE[8], rO[8];
                if (do_l1) {
                    l1_step(c, rE, rO);
                    if (ISO) {                                         // row 2jn+1 is wanted one step later
                        stash[c * 256] = make_uint4((uint32_t)rO[0], (uint32_t)rO[1], (uint32_t)rO[2], (uint32_t)rO[3]);
                        stash[c * 256 + 32] = make_uint4((uint32_t)rO[4], (uint32_t)rO[5], (uint32_t)rO[6], (uint32_t)rO[7]);
                    }
                }
                if (ISO) {
                    if (r & 1) part_read<8>(stash + c * 256, int32_t(), lo);
                    else {
#pragma unroll
                        for (int j = 0; j < 8; j++) lo[j] = rE[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) { lo[j] = rE[j]; lo[8 + j] = rO[j]; }
                }
            }
            const uint4 *sl = ring0 + c * 256;
            if (!fl) part_read<8>(sl, CT(), lo);
            if (!fh) part_read<8>(sl + 64, CT(), lo + 8);
            part_read<8>(sl + 128, CT(), hi);
            part_read<8>(sl + 192, CT(), hi + 8);
            if (lvalid && r + 1 <= rlast) issue_l0(c, r + 1);          // refill the slot: it has one full step to land
            cp_commit();
            if (ISO) { hsynth<8>(lo, first, last); hsynth<8>(hi, first, last); }
            int ev[16], od[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int e = even_upd(lo[j], hp[c][j], hi[j]);
                od[j] = inner ? odd_upd(hp[c][j], ep[c][j], e) : odd_last(hp[c][j], ep[c][j]);   // bottom edge (h even)
                ev[j] = ep[c][j];
                ep[c][j] = e; hp[c][j] = hi[j];
            }
            if (emit) {
                if (!ISO) { hsynth<8>(ev, first, last); hsynth<8>(od, first, last); }
                if (NC == 1) {
                    // DC shift (mct.go:113-118), clamp + scale exactly as createImage (decoder.go:427-466), 16 pixels per row
                    if (store_lane) {
#pragma unroll
                        for (int row = 0; row < 2; row++) {
                            const int *X = row ? od : ev;
                            uint8_t *o = orow + (size_t)row * ostride;
                            if (prec <= 8) {
                                uint32_t wd[4];
#pragma unroll
                                for (int g = 0; g < 4; g++) {
                                    wd[g] = 0;
#pragma unroll
                                    for (int p2 = 0; p2 < 4; p2++)
                                        wd[g] |= pack_value((int32_t)((uint32_t)X[4 * g + p2] + 128u), 8, 255, iso_pack != 0) << (8 * p2);
                                }
                                __stcs(reinterpret_cast<uint4 *>(o), make_uint4(wd[0], wd[1], wd[2], wd[3]));
                            } else {
#pragma unroll
                                for (int half = 0; half < 2; half++) {
                                    uint32_t wd[4];
#pragma unroll
                                    for (int g = 0; g < 4; g++) {
                                        uint32_t hv[2];
#pragma unroll
                                        for (int p2 = 0; p2 < 2; p2++) {
                                            const uint32_t t = pack_value((int32_t)((uint32_t)X[8 * half + 2 * g + p2] + 32768u), 16, 65535, iso_pack != 0);
                                            hv[p2] = (t >> 8) | ((t & 0xFFu) << 8);                   // big-endian
                                        }
                                        wd[g] = hv[0] | (hv[1] << 16);
                                    }
                                    __stcs(reinterpret_cast<uint4 *>(o) + half, make_uint4(wd[0], wd[1], wd[2], wd[3]));
                                }
                            }
                        }
                    }
                } else if (c < 2) {                                    // park the two rows until component 2 is done
                    uint4 *st = stage + c * 256;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        st[g * 32] = make_uint4((uint32_t)ev[4 * g], (uint32_t)ev[4 * g + 1], (uint32_t)ev[4 * g + 2], (uint32_t)ev[4 * g + 3]);
                        st[128 + g * 32] = make_uint4((uint32_t)od[4 * g], (uint32_t)od[4 * g + 1], (uint32_t)od[4 * g + 2], (uint32_t)od[4 * g + 3]);
                    }
                } else if (store_lane) {                               // mct.go:56-66, mct.go:113-118, decoder.go:468-487
#pragma unroll
                    for (int row = 0; row < 2; row++) {
                        const int *X2 = row ? od : ev;
                        uint32_t w3[12];                               // RGB24: the row's 16 pixels as 48 packed bytes
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            const uint4 a = stage[row * 128 + g * 32], b = stage[256 + row * 128 + g * 32];
                            const uint32_t Y[4] = {a.x, a.y, a.z, a.w}, U[4] = {b.x, b.y, b.z, b.w};
                            uint32_t px[4];
#pragma unroll
                            for (int p = 0; p < 4; p++) {
                                const uint32_t v = (uint32_t)X2[4 * g + p];
                                const uint32_t gg = Y[p] - (uint32_t)((int32_t)(U[p] + v) >> 2);
                                const int r8 = (int)(v + gg + 128u), g8 = (int)(gg + 128u), b8 = (int)(U[p] + gg + 128u);
                                px[p] = pack_sat_u8(g8, r8, pack_sat_u8(255, b8, 0u));
                            }
                            if (RGB24) {                               // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
                                w3[3 * g] = __byte_perm(px[0], px[1], 0x4210);
                                w3[3 * g + 1] = __byte_perm(px[1], px[2], 0x5421);
                                w3[3 * g + 2] = __byte_perm(px[2], px[3], 0x6542);
                            } else {
                                __stcs(reinterpret_cast<uint4 *>(orow + (size_t)row * ostride) + g, make_uint4(px[0], px[1], px[2], px[3]));
                            }
                        }
                        if (RGB24) {
#pragma unroll
                            for (int g = 0; g < 3; g++)
                                __stcs(reinterpret_cast<uint4 *>(orow + (size_t)row * ostride) + g, make_uint4(w3[4 * g], w3[4 * g + 1], w3[4 * g + 2], w3[4 * g + 3]));
                        }
                    }
                }
            }
        }
        if (do_l1) jn++;
        if (emit) orow += 2 * ostride;
    }
    cp_wait<0>();
}

template <int NC, typename CT, bool ISO, bool RGB24>
cudaError_t run_k(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    const size_t smem = (size_t)kWarps * warp_smem_u4(NC) * sizeof(uint4);
    const DevTile *tiles = p.d_tiles + p.tile_first;
    cudaError_t e = cudaFuncSetAttribute(k_idwt53_wide<NC, CT, ISO, RGB24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    J2K_LAUNCH((k_idwt53_wide<NC, CT, ISO, RGB24>), grid, kWarps * 32, smem, s, p.d_tcs, tiles, (const CT *)p.d_coef,
               (const int32_t *)p.d_tmp, p.d_pix, p.nlevels, strip_pairs, p.tail.prec[0], p.tail.iso);
    return cudaGetLastError();
}

template <int NC, typename CT>
cudaError_t run_ct(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    if (NC == 3 && p.rgb24) return p.iso ? run_k<3, CT, true, true>(p, grid, strip_pairs, s) : run_k<3, CT, false, true>(p, grid, strip_pairs, s);
    return p.iso ? run_k<NC, CT, true, false>(p, grid, strip_pairs, s) : run_k<NC, CT, false, false>(p, grid, strip_pairs, s);
}

}  // namespace

// Same contract as launch_idwt53_fused; the caller has also checked p.wide_ok (widths % 16 == 0) and p.fast_epi.
cudaError_t launch_idwt53_wide(const IdwtLaunch &p, cudaStream_t s)
{
    if (p.n_tiles == 0 || p.max_w < 16 || p.max_h < 4) return cudaSuccess;
    const uint32_t nl = p.max_w / 16, own = nl > 32 ? 30 : 32, nwx = (nl + own - 1) / own, nly = p.max_h / 2;
    // strip height: tall strips amortise the two silent prologue steps, short ones even out the last wave of 8 warps per
    // SM.  Measured on the bench batch (640 tiles of 512 x 512): REF order 64 / 32 / 16 row pairs -> 0.626 / 0.533 / 0.515 ms
    // (int32 planes), ISO order 0.419 / 0.359 / 0.367 ms (int16): at least 8 waves for REF, 4 for ISO.
    const uint64_t min_warps = 148ull * 8 * (p.iso ? 4 : 8);
    int sp = 64;
    while (sp > 8 && (uint64_t)p.n_tiles * nwx * ((nly + sp - 1) / sp) < min_warps) sp >>= 1;
    if (p.wide_sp >= 2) sp = p.wide_sp & ~1;             // J2kOpts.wide_sp: A/B runs
    const uint32_t units = nwx * ((nly + sp - 1) / sp);
    dim3 grid((units + kWarps - 1) / kWarps, p.n_tiles, 1);
    if (p.tail.ncomp == 1) return p.coef16 ? run_ct<1, int16_t>(p, grid, sp, s) : run_ct<1, int32_t>(p, grid, sp, s);
    return p.coef16 ? run_ct<3, int16_t>(p, grid, sp, s) : run_ct<3, int32_t>(p, grid, sp, s);
}
