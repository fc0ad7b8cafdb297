// idwt_fused.cu -- the last TWO inverse 5-3 DWT levels of every tile, fused with the pixel epilogue (inverse
// RCT/ICT + DC shift + clamp + pack), in one kernel: the level-1 output (a quarter of the samples) never goes to
// HBM.  This is the HBM-roofline kernel of the path.
//
// Same arithmetic as idwt.cu / idwt_stream.cu (reference dwt.go:122-147 per line; columns then rows dwt.go:410-429
// and dense-prefix addressing dwt.go:534-548 in REF mode, rows then columns on the Mallat layout in ISO mode;
// epilogue decoder.go:321-348 + 417-588).  Organisation:
//   * a WARP owns 28 quads (112 output columns) x a strip of row pairs of one tile, all components; lanes 0,1 and
//     30,31 are halo lanes that recompute what the neighbouring warps own, so there is no block-level barrier;
//   * lane q holds level-0 quad q (L columns 2q,2q+1 and H columns 2q,2q+1 -> output columns 4q..4q+3) and level-1
//     "half quad" q (L column q, H column q -> level-1 output columns 2q,2q+1).  Those two level-1 output columns
//     are exactly the lane's own level-0 low-pass inputs, in both layouts: in ISO mode level-1 output row k is the
//     LL row of level-0 row pair k; in REF mode (dense prefix) level-1 output rows 2k and 2k+1 are the L half and
//     the H half of level-0 band row k, for the top half of the tile.  So the hand-over stays in registers;
//   * vertical lifting streams down the strip with one row pair of delay (registers); horizontal lifting needs the
//     neighbour lane's last H value and first even output: two warp shuffles per row and component;
//   * the level-0 band rows are staged through a per-warp shared-memory ring with cp.async (LDGSTS), kDepth - 1
//     row pairs ahead, so the bytes in flight do not cost registers; every lane reads back only what it copied
//     itself, so the ring needs no barrier either.  Level-1 rows (1/4 of the data) use a register prefetch;
//   * the coefficient planes are int32, or int16 when the job's magnitudes provably fit (half the read traffic).
// Eligibility (host): every tile-component width a multiple of 8 and height a multiple of 4, reversible filter.
#include "common.h"
#include "tail.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kDepth = 4;      // ring slots (row pairs); kDepth - 1 are in flight behind the arithmetic
constexpr int kOwn = 28;       // quads stored per warp

// reference edge-exact 5-3 steps with Go's wrapping int32 arithmetic
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

// ---- cp.async (LDGSTS) ----------------------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
#ifdef J2K_EMU
    memcpy(smem_dst, gsrc, BYTES);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc), "n"(BYTES) : "memory");
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

// two adjacent elements / one element of a plane
__device__ __forceinline__ int2 ldpair(const int32_t *p) { return __ldg(reinterpret_cast<const int2 *>(p)); }
__device__ __forceinline__ int2 ldpair(const int16_t *p)
{
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(p));
    return make_int2((int)(int16_t)(r & 0xFFFFu), (int)r >> 16);
}
__device__ __forceinline__ int ld1(const int32_t *p) { return __ldg(p); }
__device__ __forceinline__ int ld1(const int16_t *p) { return (int)__ldg(p); }
// a ring slot (8 bytes per lane) holding a pair copied from a plane of element type CT
__device__ __forceinline__ int2 slot_pair(const uint2 *s, int32_t) { const uint2 v = *s; return make_int2((int)v.x, (int)v.y); }
__device__ __forceinline__ int2 slot_pair(const uint2 *s, int16_t)
{
    const uint32_t r = *reinterpret_cast<const uint32_t *>(s);
    return make_int2((int)(int16_t)(r & 0xFFFFu), (int)r >> 16);
}

// generic pixel store of one lane's 4 columns of one row (any format / precision / MCT): kept out of line
template <int NC>
__device__ J2K_NOINLINE void put_quad_generic(uint8_t *orow, uint32_t gx0, uint32_t img_w, const int *X, TailParams tp)
{
#pragma unroll
    for (int p = 0; p < 4; p++) {
        if (gx0 + p >= img_w) continue;
        int32_t v[4] = {X[p], NC > 1 ? X[(NC > 1 ? 4 : 0) + p] : 0, NC > 2 ? X[(NC > 2 ? 8 : 0) + p] : 0,
                        NC > 3 ? X[(NC > 3 ? 12 : 0) + p] : 0};
        tail_mct_dc(v, tp);
        if (NC >= 3) tail_colour(v, tp);                         // decoder.go:350-356 (a call only when the job has a conversion)
        store_pixel(orow, gx0 + p, v, tp);
    }
}

// FAST: 3 unsigned 8-bit components, RCT, RGBA8 output, every tile inside the image and 16-byte aligned rows
// (checked on the host): the epilogue is 8 instructions per pixel and one 16-byte streaming store per lane and row.
template <int NC, typename CT, bool ISO, bool FAST>
__global__ void __launch_bounds__(kWarps * 32, (NC <= 3 ? 4 : 3))
k_idwt53_fused(const DevTileComp *__restrict__ tcs, const DevTile *__restrict__ tiles, const CT *__restrict__ coef,
               const int32_t *__restrict__ tmp, uint8_t *__restrict__ pix, int nlevels, int strip_pairs, TailParams tp)
{
    J2K_DYN_SMEM(uint2, ring_all);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DevTile tile = tiles[blockIdx.y];
    const int w = (int)tile.w, h = (int)tile.h;
    const int nlx = w >> 1, nly = h >> 1, nq = w >> 2;
    const int nwx = (nq + kOwn - 1) / kOwn;
    const int nstrips = (nly + strip_pairs - 1) / strip_pairs;
    const int unit = blockIdx.x * kWarps + warp;
    if (unit >= nwx * nstrips) return;
    const int strip = unit / nwx, wi = unit - strip * nwx;
    const int q = wi * kOwn - 2 + lane;
    const bool qvalid = q >= 0 && q < nq;
    const bool store_lane = qvalid && lane >= 2 && lane <= 29;
    const bool q_first = q == 0, q_last = q == nq - 1;
    const int qc = qvalid ? q : 0;
    const int ka = strip * strip_pairs, kb = min(ka + strip_pairs, nly);
    const int rlast = min(kb, nly - 1);               // last level-0 band row pair this strip reads
    const bool l1on = nlevels >= 2, l2on = nlevels >= 3;
    const int nlx1 = nq, nly1 = nly >> 1;             // level-1 image: nlx x nly, low-pass part nlx1 x nly1
    const int half0 = nly >> 1;                       // REF: level-0 band rows below this ARE the level-1 output

    const CT *plane[NC];
    const int32_t *prev2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DevTileComp tc = tcs[tile.tc[c]];
        plane[c] = coef + tc.coef_off;
        prev2[c] = tmp + tc.tmp_off;                  // level 2 wrote ping-pong buffer 0
        J2K_OPAQUE_PTR(plane[c]);                     // keep the base pointers in registers: every address below is
        J2K_OPAQUE_PTR(prev2[c]);                     // base + 32-bit offset (one IMAD.WIDE), not a 64-bit carry chain
    }
    uint2 *ring = ring_all + (size_t)warp * (kDepth * NC * 4 * 32) + lane;
    const uint32_t uw = (uint32_t)w;
    const uint32_t col0 = 2u * (uint32_t)qc;          // level-0 L column pair; H pair at nlx + col0

    // ---- level-0 band row pair r -> ring slot r % kDepth (the parts that come from the coefficient planes) ----
    // rows are issued in increasing order: o_iss is the element offset of (row r, column col0), advanced per call
    const uint32_t d_hi = (uint32_t)nly * uw;
    uint32_t o_iss = 0;
    auto issue_l0 = [&](int r) {
        const int s = r & (kDepth - 1);
        const bool fl = l1on && (ISO || r < half0), fh = l1on && !ISO && r < half0;   // those parts are level-1 output
        const uint32_t o0 = o_iss, o1 = o_iss + (uint32_t)nlx, o2 = o_iss + d_hi, o3 = o2 + (uint32_t)nlx;
        uint2 *sl = ring + (size_t)(s * NC * 4) * 32;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            if (!fl) cp_async<2 * sizeof(CT)>(sl + c * 128, plane[c] + o0);
            if (!fh) cp_async<2 * sizeof(CT)>(sl + c * 128 + 32, plane[c] + o1);
            cp_async<2 * sizeof(CT)>(sl + c * 128 + 64, plane[c] + o2);
            cp_async<2 * sizeof(CT)>(sl + c * 128 + 96, plane[c] + o3);
        }
    };

    // ---- horizontal synthesis; every lane takes part in the two shuffles ----
    // level 0: a row held as (L0, L1, H0, H1) per lane -> 4 interleaved samples, in place
    auto hsynth0 = [&](int V[4]) {
        const int VL0 = V[0], VL1 = V[1], VH0 = V[2], VH1 = V[3];
        int left = __shfl_up_sync(0xffffffffu, VH1, 1);
        left = q_first ? VH0 : left;                                  // x[0] -= (x[1] + x[1] + 2) >> 2
        const int X0 = even_upd(VL0, left, VH0);
        const int X2 = even_upd(VL1, VH0, VH1);
        const int right = __shfl_down_sync(0xffffffffu, X0, 1);
        const int t = odd_upd(VH1, X2, right);
        V[0] = X0;
        V[1] = odd_upd(VH0, X0, X2);
        V[2] = X2;
        V[3] = q_last ? odd_last(VH1, X2) : t;
    };
    // level 1: a row held as (L, H) per lane -> 2 interleaved samples, in place (level-1 rows are 2 * nq samples wide)
    auto hsynth1 = [&](int V[2]) {
        const int VL = V[0], VH = V[1];
        int left = __shfl_up_sync(0xffffffffu, VH, 1);
        left = q_first ? VH : left;
        const int X0 = even_upd(VL, left, VH);
        const int right = __shfl_down_sync(0xffffffffu, X0, 1);
        const int t = odd_upd(VH, X0, right);
        V[0] = X0;
        V[1] = q_last ? odd_last(VH, X0) : t;
    };

    // ---- level 1 -----------------------------------------------------------------------------------------------
    // band row pair j of the level-1 image: lo = row j, hi = row nly1 + j; v = lo (L, H), hi (L, H) of this lane
    auto l1_load = [&](int j, int v[NC][4]) {
        uint32_t o_lo, o_hi, o_p;          // element offsets of this lane's L element; the H element is nlx1 further
        bool pL, pH;                       // lo L / lo H come from the level-2 output (int32) instead of the plane
        if (ISO) {
            o_lo = (uint32_t)j * uw + (uint32_t)qc; o_hi = (uint32_t)(nly1 + j) * uw + (uint32_t)qc;
            o_p = (uint32_t)j * (uint32_t)nlx1 + (uint32_t)qc;
            pL = l2on; pH = false;
        } else {
            o_lo = (uint32_t)j * (uint32_t)nlx + (uint32_t)qc; o_hi = (uint32_t)(nly1 + j) * (uint32_t)nlx + (uint32_t)qc;
            o_p = o_lo;
            pL = l2on && 2 * j + 1 <= nly1; pH = l2on && 2 * j + 2 <= nly1;
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            v[c][0] = pL ? __ldg(prev2[c] + o_p) : ld1(plane[c] + o_lo);
            v[c][1] = pH ? __ldg(prev2[c] + o_p + nlx1) : ld1(plane[c] + o_lo + nlx1);
            v[c][2] = ld1(plane[c] + o_hi);
            v[c][3] = ld1(plane[c] + o_hi + nlx1);
        }
    };
    int h1p[NC][2], e1p[NC][2];        // hi1[jn] and E1[jn] in the vertical-lifting domain
    int r1[2][NC][2];                  // the last emitted level-1 output row pair (rows 2(jn-1), 2(jn-1)+1), cols 2q, 2q+1
    int pf1[NC][4];                    // prefetched band row pair jn+1
    int jn = 0, j1last = -1;           // next level-1 pair to emit; last pair this strip needs

    // one level-1 step: consumes band row pair jn+1 (prefetched), emits output rows 2jn, 2jn+1 into r1, prefetches jn+2
    auto l1_step = [&]() {
        int nx[NC][4];
        const bool pfnext = (jn + 1 <= j1last) && (jn + 2 < nly1);
        if (pfnext) l1_load(jn + 2, nx);
        const bool inner = jn + 1 < nly1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int lo[2] = {pf1[c][0], pf1[c][1]}, hi[2] = {pf1[c][2], pf1[c][3]};
            if (ISO) { hsynth1(lo); hsynth1(hi); }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int e = even_upd(lo[j], h1p[c][j], hi[j]);
                const int o = inner ? odd_upd(h1p[c][j], e1p[c][j], e) : odd_last(h1p[c][j], e1p[c][j]);   // bottom edge
                r1[0][c][j] = e1p[c][j]; r1[1][c][j] = o;
                e1p[c][j] = e; h1p[c][j] = hi[j];
            }
            if (!ISO) { hsynth1(r1[0][c]); hsynth1(r1[1][c]); }
#pragma unroll
            for (int j = 0; j < 4; j++) pf1[c][j] = nx[c][j];
        }
        jn++;
    };

    // ---- prologue: fill the ring, load the row above the strip; the first loop step then only builds E[ka] -----------
    const int k0 = ka - 1;
    o_iss = (uint32_t)ka * uw + col0;
#pragma unroll
    for (int i = 1; i < kDepth; i++) {
        if (qvalid && k0 + i <= rlast) issue_l0(k0 + i);
        o_iss += uw;
        cp_commit();
    }
    bool l1need = false;
    if (l1on) {
        if (ISO) { l1need = true; jn = (ka >> 1) - 1; j1last = rlast >> 1; }
        else if (ka < half0) { l1need = true; jn = ka - 1; j1last = min(rlast, half0 - 1); }
    }
    if (l1need) {                                                      // hi1[max(jn, 0)] -> h1p, band row pair jn+1 -> pf1
        int t[NC][4];
        l1_load(jn < 0 ? 0 : jn, t);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int hi[2] = {t[c][2], t[c][3]};
            if (ISO) hsynth1(hi);
            h1p[c][0] = hi[0]; h1p[c][1] = hi[1];
            e1p[c][0] = e1p[c][1] = 0;
        }
        l1_load(jn + 1, pf1);
    }
    int hp[NC][4], ep[NC][4];          // Hi[k] and E[k] of the lane's 4 columns
    {
        const int rp = k0 < 0 ? 0 : k0;                                // top edge: Hi[-1] := Hi[0]
        const uint32_t ohi = (uint32_t)(nly + rp) * uw + col0;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const int2 a = ldpair(plane[c] + ohi), b = ldpair(plane[c] + ohi + nlx);
            hp[c][0] = a.x; hp[c][1] = a.y; hp[c][2] = b.x; hp[c][3] = b.y;
            if (ISO) hsynth0(hp[c]);
#pragma unroll
            for (int j = 0; j < 4; j++) ep[c][j] = 0;
        }
    }

    // ---- pixel epilogue of one finished row (columns interleaved) ---------------------------------------------------
    const uint32_t gx0 = tile.img_x0 + 4u * (uint32_t)qc;
    uint8_t *orow = pix + tile.out_off + (size_t)(tile.img_y0 + 2u * (uint32_t)ka) * tile.out_stride;
    uint32_t gy = tile.img_y0 + 2u * (uint32_t)ka;
    auto put_row = [&](int X[NC][4]) {
        if (FAST && NC == 1) {
            // one unsigned component of 8 or 16 bits: DC shift (mct.go:113-118), clamp + scale exactly as createImage
            // (decoder.go:427-466; constant divisors), Gray8 or big-endian Gray16, one 4 / 8 byte store per lane and row
            if (store_lane) {
                if (tp.prec[0] == 8) {
                    uint32_t px = 0;
#pragma unroll
                    for (int p = 0; p < 4; p++) px |= pack_value((int32_t)((uint32_t)X[0][p] + 128u), 8, 255, tp.iso) << (8 * p);
                    __stcs(reinterpret_cast<uint32_t *>(orow + (size_t)gx0), px);
                } else {
                    uint32_t hv[4];
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const uint32_t t = pack_value((int32_t)((uint32_t)X[0][p] + 32768u), 16, 65535, tp.iso);
                        hv[p] = (t >> 8) | ((t & 0xFFu) << 8);
                    }
                    __stcs(reinterpret_cast<uint2 *>(orow + 2 * (size_t)gx0), make_uint2(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16)));
                }
            }
        } else if (FAST) {
            if (store_lane) {
                uint32_t px[4];
#pragma unroll
                for (int p = 0; p < 4; p++) {                          // mct.go:56-66, mct.go:113-118, decoder.go:468-487
                    const uint32_t y0 = (uint32_t)X[0][p], u = (uint32_t)X[NC > 1 ? 1 : 0][p], v = (uint32_t)X[NC > 2 ? 2 : 0][p];
                    const uint32_t g = y0 - (uint32_t)((int32_t)(u + v) >> 2);
                    const int r8 = (int)(v + g + 128u), g8 = (int)(g + 128u), b8 = (int)(u + g + 128u);
                    px[p] = pack_sat_u8(g8, r8, pack_sat_u8(255, b8, 0u));
                }
                __stcs(reinterpret_cast<uint4 *>(orow + 4 * (size_t)gx0), make_uint4(px[0], px[1], px[2], px[3]));
            }
        } else {
            if (store_lane && gy < tile.img_h) put_quad_generic<NC>(orow, gx0, tile.img_w, &X[0][0], tp);   // decoder.go:398-410
            gy++;
        }
        orow += tile.out_stride;
    };

    // ---- stream the strip: step k consumes band row pair k+1 and finishes output rows 2k (even) and 2k+1 (odd) ------
    for (int k = k0; k < kb; k++) {
        if (qvalid && k + kDepth <= rlast) issue_l0(k + kDepth);
        o_iss += uw;
        cp_commit();
        const int r = k + 1;
        const bool inner = r < nly;                                    // false only below the last row pair of the tile
        const bool fl = l1on && (ISO || r < half0), fh = l1on && !ISO && r < half0;
        if (inner && fl) {
            const int need = ISO ? (r >> 1) : r;                       // the level-1 pair that holds level-0 row r
            while (jn <= need) l1_step();                              // warp-uniform; twice on the first step only
        }
        cp_wait<kDepth - 1>();
        const int s = r & (kDepth - 1);
        const bool odd = ISO && (r & 1);
        int evn[NC][4], o[NC][4];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint2 *sl = ring + (size_t)((s * NC + c) * 4) * 32;
            int lo[4], hi[4];
            int2 t;
            t = slot_pair(sl, CT()); lo[0] = t.x; lo[1] = t.y;
            t = slot_pair(sl + 32, CT()); lo[2] = t.x; lo[3] = t.y;
            t = slot_pair(sl + 64, CT()); hi[0] = t.x; hi[1] = t.y;
            t = slot_pair(sl + 96, CT()); hi[2] = t.x; hi[3] = t.y;
            if (fl) { lo[0] = odd ? r1[1][c][0] : r1[0][c][0]; lo[1] = odd ? r1[1][c][1] : r1[0][c][1]; }
            if (fh) { lo[2] = r1[1][c][0]; lo[3] = r1[1][c][1]; }
            if (ISO) { hsynth0(lo); hsynth0(hi); }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = even_upd(lo[j], hp[c][j], hi[j]);
                o[c][j] = inner ? odd_upd(hp[c][j], ep[c][j], e) : odd_last(hp[c][j], ep[c][j]);   // bottom edge (h even)
                evn[c][j] = ep[c][j];
                ep[c][j] = e; hp[c][j] = hi[j];
            }
            if (!ISO && k >= ka) { hsynth0(evn[c]); hsynth0(o[c]); }
        }
        if (k >= ka) { put_row(evn); put_row(o); }
    }
    cp_wait<0>();
}

template <int NC, typename CT, bool FAST>
cudaError_t run_ct(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    const size_t smem = (size_t)kWarps * kDepth * NC * 4 * 32 * sizeof(uint2);
    const DevTile *tiles = p.d_tiles + p.tile_first;
    cudaError_t e;
    if (p.iso) {
        if ((e = cudaFuncSetAttribute(k_idwt53_fused<NC, CT, true, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        J2K_LAUNCH((k_idwt53_fused<NC, CT, true, FAST>), grid, kWarps * 32, smem, s, p.d_tcs, tiles, (const CT *)p.d_coef,
                   (const int32_t *)p.d_tmp, p.d_pix, p.nlevels, strip_pairs, p.tail);
    } else {
        if ((e = cudaFuncSetAttribute(k_idwt53_fused<NC, CT, false, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        J2K_LAUNCH((k_idwt53_fused<NC, CT, false, FAST>), grid, kWarps * 32, smem, s, p.d_tcs, tiles, (const CT *)p.d_coef,
                   (const int32_t *)p.d_tmp, p.d_pix, p.nlevels, strip_pairs, p.tail);
    }
    return cudaGetLastError();
}

template <int NC, bool FAST>
cudaError_t run(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    return p.coef16 ? run_ct<NC, int16_t, FAST>(p, grid, strip_pairs, s) : run_ct<NC, int32_t, FAST>(p, grid, strip_pairs, s);
}

}  // namespace

// Levels 1 and 0 of the tiles [tile_first, tile_first + n_tiles) + the pixel epilogue.  The caller has run the
// coarser levels (down to level 2) and checked eligibility (j2k_fused_ok for every tile-component, reversible).
cudaError_t launch_idwt53_fused(const IdwtLaunch &p, cudaStream_t s)
{
    if (p.n_tiles == 0 || p.max_w < 8 || p.max_h < 4) return cudaSuccess;
    const uint32_t nq = p.max_w / 4, nwx = (nq + kOwn - 1) / kOwn, nly = p.max_h / 2;
    // strip height: as tall as possible (the halo costs ~3 band rows per level and strip) while keeping the machine full
    int sp = 64;
    while (sp > 8 && (uint64_t)p.n_tiles * nwx * ((nly + sp - 1) / sp) < 148ull * 16 * 2) sp >>= 1;
    const uint32_t units = nwx * ((nly + sp - 1) / sp);
    dim3 grid((units + kWarps - 1) / kWarps, p.n_tiles, 1);
    const bool fast = p.fast_epi && ((uintptr_t)p.d_pix & 15) == 0;
    if (fast && p.wide_ok) return launch_idwt53_wide(p, s);
    switch (p.tail.ncomp) {
    case 1: return fast ? run<1, true>(p, grid, sp, s) : run<1, false>(p, grid, sp, s);
    case 3: return fast ? run<3, true>(p, grid, sp, s) : run<3, false>(p, grid, sp, s);
    case 4: return run<4, false>(p, grid, sp, s);
    }
    return cudaErrorInvalidValue;
}
