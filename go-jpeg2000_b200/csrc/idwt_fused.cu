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

template <int NC, typename CT, bool ISO>
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
    const int qc = qvalid ? q : 0;
    const int ka = strip * strip_pairs, kb = min(ka + strip_pairs, nly);
    const int rlast = min(kb, nly - 1);               // last level-0 band row pair this strip reads
    const bool l1on = nlevels >= 2, l2on = nlevels >= 3;
    const int nlx1 = nq, nly1 = nly >> 1, w1 = nlx;   // level-1 image: w1 x nly, low-pass part nlx1 x nly1
    const int half0 = nly >> 1;                       // REF: level-0 band rows below this come from the level-1 output

    const CT *plane[NC];
    const int32_t *prev2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DevTileComp tc = tcs[tile.tc[c]];
        plane[c] = coef + tc.coef_off;
        prev2[c] = tmp + tc.tmp_off;                  // level 2 wrote ping-pong buffer 0
    }
    uint2 *ring = ring_all + (size_t)warp * (kDepth * NC * 4 * 32) + lane;
    const uint32_t uw = (uint32_t)w;
    const uint32_t col0 = 2u * (uint32_t)qc;          // level-0 L column pair; H pair at nlx + col0

    // which parts of level-0 band row r (r < nly) are the level-1 output
    auto l0_L_from_l1 = [&](int r) { return l1on && (ISO || r < half0); };
    auto l0_H_from_l1 = [&](int r) { return l1on && !ISO && r < half0; };

    // ---- level-0 band row pair r -> ring slot r % kDepth (the parts that come from the coefficient planes) ----
    auto issue_l0 = [&](int r) {
        if (!qvalid) return;
        const int s = r & (kDepth - 1);
        const bool fl = l0_L_from_l1(r), fh = l0_H_from_l1(r);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            uint2 *sl = ring + (size_t)((s * NC + c) * 4) * 32;
            const CT *lo = plane[c] + (size_t)((uint32_t)r * uw + col0);
            const CT *hi = plane[c] + (size_t)((uint32_t)(nly + r) * uw + col0);
            if (!fl) cp_async<2 * sizeof(CT)>(sl, lo);
            if (!fh) cp_async<2 * sizeof(CT)>(sl + 32, lo + nlx);
            cp_async<2 * sizeof(CT)>(sl + 64, hi);
            cp_async<2 * sizeof(CT)>(sl + 96, hi + nlx);
        }
    };

    // ---- horizontal synthesis; every lane takes part in the two shuffles ----
    // level 0: a row held as (L0, L1, H0, H1) per lane -> 4 interleaved samples
    auto hsynth0 = [&](const int V[4], int X[4]) {
        const int VL0 = V[0], VL1 = V[1], VH0 = V[2], VH1 = V[3];
        int left = __shfl_up_sync(0xffffffffu, VH1, 1);
        if (q == 0) left = VH0;                                       // x[0] -= (x[1] + x[1] + 2) >> 2
        const int X0 = even_upd(VL0, left, VH0);
        const int X2 = even_upd(VL1, VH0, VH1);
        const int right = __shfl_down_sync(0xffffffffu, X0, 1);
        X[0] = X0;
        X[1] = odd_upd(VH0, X0, X2);
        X[2] = X2;
        X[3] = (q == nq - 1) ? odd_last(VH1, X2) : odd_upd(VH1, X2, right);
    };
    // level 1: a row held as (L, H) per lane -> 2 interleaved samples (level-1 rows are w1 = 2 * nq samples wide)
    auto hsynth1 = [&](const int V[2], int X[2]) {
        const int VL = V[0], VH = V[1];
        int left = __shfl_up_sync(0xffffffffu, VH, 1);
        if (q == 0) left = VH;
        const int X0 = even_upd(VL, left, VH);
        const int right = __shfl_down_sync(0xffffffffu, X0, 1);
        X[0] = X0;
        X[1] = (q == nq - 1) ? odd_last(VH, X0) : odd_upd(VH, X0, right);
    };

    // ---- level 1 -----------------------------------------------------------------------------------------------
    // band row rr of the level-1 image (rr < nly1: low-pass row, else high-pass row rr - nly1): (L, H) of this lane
    auto l1_load = [&](int c, int rr, int v[2]) {
        if (!qvalid) { v[0] = v[1] = 0; return; }
        if (ISO) {
            const CT *rowp = plane[c] + (size_t)((uint32_t)rr * uw);
            v[1] = ld1(rowp + nlx1 + qc);
            if (rr < nly1 && l2on) v[0] = __ldg(prev2[c] + (size_t)((uint32_t)rr * (uint32_t)nlx1 + (uint32_t)qc));
            else v[0] = ld1(rowp + qc);
        } else {
            const uint32_t lin = (uint32_t)rr * (uint32_t)w1 + (uint32_t)qc;
            v[0] = (l2on && 2 * rr + 1 <= nly1) ? __ldg(prev2[c] + lin) : ld1(plane[c] + lin);
            v[1] = (l2on && 2 * rr + 2 <= nly1) ? __ldg(prev2[c] + lin + nlx1) : ld1(plane[c] + lin + nlx1);
        }
    };
    int h1p[NC][2], e1p[NC][2];        // hi1[jn] and E1[jn] in the vertical-lifting domain
    int r1[2][NC][2];                  // the last emitted level-1 output row pair (rows 2(jn-1), 2(jn-1)+1), cols 2q, 2q+1
    int pf1[NC][4];                    // prefetched band rows jn+1: lo (L,H), hi (L,H)
    int jn = 0, j1last = -1;           // next level-1 pair to emit; last pair this strip needs

    auto l1_emit = [&]() {
        int nx[NC][4];
        const bool pfnext = (jn + 1 <= j1last) && (jn + 2 < nly1);
        if (pfnext) {
#pragma unroll
            for (int c = 0; c < NC; c++) { l1_load(c, jn + 2, &nx[c][0]); l1_load(c, nly1 + jn + 2, &nx[c][2]); }
        }
        const bool inner = jn + 1 < nly1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int o[2];
            if (inner) {
                int lo[2] = {pf1[c][0], pf1[c][1]}, hi[2] = {pf1[c][2], pf1[c][3]};
                if (ISO) { int t[2]; hsynth1(lo, t); lo[0] = t[0]; lo[1] = t[1]; hsynth1(hi, t); hi[0] = t[0]; hi[1] = t[1]; }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int e = even_upd(lo[j], h1p[c][j], hi[j]);
                    o[j] = odd_upd(h1p[c][j], e1p[c][j], e);
                    r1[0][c][j] = e1p[c][j];
                    e1p[c][j] = e; h1p[c][j] = hi[j];
                }
            } else {                                                   // bottom edge (nly even): last odd row += x[n-2]
#pragma unroll
                for (int j = 0; j < 2; j++) { o[j] = odd_last(h1p[c][j], e1p[c][j]); r1[0][c][j] = e1p[c][j]; }
            }
            r1[1][c][0] = o[0]; r1[1][c][1] = o[1];
            if (!ISO) {
                int t[2];
                hsynth1(r1[0][c], t); r1[0][c][0] = t[0]; r1[0][c][1] = t[1];
                hsynth1(r1[1][c], t); r1[1][c][0] = t[0]; r1[1][c][1] = t[1];
            }
            if (pfnext) {
#pragma unroll
                for (int j = 0; j < 4; j++) pf1[c][j] = nx[c][j];
            }
        }
        jn++;
    };

    // ---- prologue ------------------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 1; i < kDepth; i++) {
        if (ka + i <= rlast) issue_l0(ka + i);
        cp_commit();
    }
    bool l1need = false;
    if (l1on) {
        if (ISO) { l1need = true; jn = ka >> 1; j1last = rlast >> 1; }
        else if (ka < half0) { l1need = true; jn = ka; j1last = min(rlast, half0 - 1); }
    }
    if (l1need) {
        const int ja = jn;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int a[2], b[2], p[2];
            l1_load(c, ja, a); l1_load(c, nly1 + ja, b);
            if (ja > 0) l1_load(c, nly1 + ja - 1, p);
            if (ISO) {
                int t[2];
                hsynth1(a, t); a[0] = t[0]; a[1] = t[1];
                hsynth1(b, t); b[0] = t[0]; b[1] = t[1];
                if (ja > 0) { hsynth1(p, t); p[0] = t[0]; p[1] = t[1]; }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int hprev = ja > 0 ? p[j] : b[j];                // top edge: Hi[-1] := Hi[0]
                e1p[c][j] = even_upd(a[j], hprev, b[j]);
                h1p[c][j] = b[j];
            }
            if (ja + 1 < nly1) { l1_load(c, ja + 1, &pf1[c][0]); l1_load(c, nly1 + ja + 1, &pf1[c][2]); }
        }
        l1_emit();                                                     // the pair that holds level-0 row ka
    }

    int hp[NC][4], ep[NC][4];          // Hi[k-1] and E[k-1] of the lane's 4 columns
    {
        const bool fl = l0_L_from_l1(ka), fh = l0_H_from_l1(ka);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            int lo[4], hi[4], pv[4];
            if (qvalid) {
                const CT *lop = plane[c] + (size_t)((uint32_t)ka * uw + col0);
                const CT *hip = plane[c] + (size_t)((uint32_t)(nly + ka) * uw + col0);
                int2 t;
                if (fl) {
                    const bool odd = ISO && (ka & 1);
                    lo[0] = odd ? r1[1][c][0] : r1[0][c][0]; lo[1] = odd ? r1[1][c][1] : r1[0][c][1];
                }
                else { t = ldpair(lop); lo[0] = t.x; lo[1] = t.y; }
                if (fh) { lo[2] = r1[1][c][0]; lo[3] = r1[1][c][1]; }
                else { t = ldpair(lop + nlx); lo[2] = t.x; lo[3] = t.y; }
                t = ldpair(hip); hi[0] = t.x; hi[1] = t.y;
                t = ldpair(hip + nlx); hi[2] = t.x; hi[3] = t.y;
                if (ka > 0) {
                    t = ldpair(hip - uw); pv[0] = t.x; pv[1] = t.y;
                    t = ldpair(hip - uw + nlx); pv[2] = t.x; pv[3] = t.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) lo[j] = hi[j] = pv[j] = 0;
            }
            if (ISO) {
                int t[4];
                hsynth0(lo, t);
#pragma unroll
                for (int j = 0; j < 4; j++) lo[j] = t[j];
                hsynth0(hi, t);
#pragma unroll
                for (int j = 0; j < 4; j++) hi[j] = t[j];
                if (ka > 0) {
                    hsynth0(pv, t);
#pragma unroll
                    for (int j = 0; j < 4; j++) pv[j] = t[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int hprev = ka > 0 ? pv[j] : hi[j];              // top edge: Hi[-1] := Hi[0]
                ep[c][j] = even_upd(lo[j], hprev, hi[j]);
                hp[c][j] = hi[j];
            }
        }
    }

    // ---- pixel epilogue of one finished row (columns already interleaved) ---------------------------------------
    const uint32_t gx0 = tile.img_x0 + 4u * (uint32_t)qc;
    const bool fast_rgba8 = tp.fmt == J2KGPU_FMT_RGBA8 && NC == 3 && tp.prec[0] == 8 && tp.prec[1] == 8 && tp.prec[2] == 8 &&
                            tp.mct && tp.reversible && !tp.sgnd[0] && !tp.sgnd[1] && !tp.sgnd[2] &&
                            ((tile.out_stride & 15) == 0) && ((tile.img_x0 & 3) == 0) && (gx0 + 3 < tile.img_w);
    uint8_t *orow = pix + tile.out_off + (size_t)(tile.img_y0 + 2u * (uint32_t)ka) * tile.out_stride;
    uint32_t gy = tile.img_y0 + 2u * (uint32_t)ka;
    auto put_row = [&](int X[NC][4]) {
        if (store_lane && gy < tile.img_h) {                           // decoder.go:398-410 clipping
            if (fast_rgba8) {
                uint32_t px[4];
#pragma unroll
                for (int p = 0; p < 4; p++) {                          // mct.go:56-66, mct.go:113-118, decoder.go:468-487
                    const uint32_t y0 = (uint32_t)X[0][p], u = (uint32_t)X[NC > 1 ? 1 : 0][p], v = (uint32_t)X[NC > 2 ? 2 : 0][p];
                    const uint32_t g = y0 - (uint32_t)((int32_t)(u + v) >> 2);
                    const int r8 = (int)(v + g + 128u), g8 = (int)(g + 128u), b8 = (int)(u + g + 128u);
                    px[p] = pack_sat_u8(g8, r8, pack_sat_u8(255, b8, 0u));
                }
                __stcs(reinterpret_cast<uint4 *>(orow + 4 * (size_t)gx0), make_uint4(px[0], px[1], px[2], px[3]));
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    if (gx0 + p >= tile.img_w) continue;
                    int32_t v[4] = {X[0][p], NC > 1 ? X[NC > 1 ? 1 : 0][p] : 0, NC > 2 ? X[NC > 2 ? 2 : 0][p] : 0,
                                    NC > 3 ? X[NC > 3 ? 3 : 0][p] : 0};
                    tail_mct_dc(v, tp);
                    store_pixel(orow, gx0 + p, v, tp);
                }
            }
        }
        orow += tile.out_stride;
        gy++;
    };
    // REF: the finished row is in band-column order and still needs the horizontal synthesis; ISO: it is final
    auto emit_row = [&](int V[NC][4]) {
        if (ISO) {
            put_row(V);
        } else {
            int X[NC][4];
#pragma unroll
            for (int c = 0; c < NC; c++) hsynth0(V[c], X[c]);
            put_row(X);
        }
    };

    // ---- stream the strip: step k finishes output rows 2k (even) and 2k+1 (odd) --------------------------------------
    for (int k = ka; k < kb; k++) {
        if (k + kDepth <= rlast) issue_l0(k + kDepth);
        cp_commit();
        int o[NC][4];
        if (k + 1 < nly) {
            const int r = k + 1;
            const bool fl = l0_L_from_l1(r), fh = l0_H_from_l1(r);
            if (fl && jn <= (ISO ? (r >> 1) : r)) l1_emit();           // warp-uniform
            cp_wait<kDepth - 1>();
            const int s = r & (kDepth - 1);
            int e[NC][4];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const uint2 *sl = ring + (size_t)((s * NC + c) * 4) * 32;
                int lo[4], hi[4];
                int2 t;
                if (fl) {
                    const bool odd = ISO && (r & 1);
                    lo[0] = odd ? r1[1][c][0] : r1[0][c][0]; lo[1] = odd ? r1[1][c][1] : r1[0][c][1];
                }
                else { t = slot_pair(sl, CT()); lo[0] = t.x; lo[1] = t.y; }
                if (fh) { lo[2] = r1[1][c][0]; lo[3] = r1[1][c][1]; }
                else { t = slot_pair(sl + 32, CT()); lo[2] = t.x; lo[3] = t.y; }
                t = slot_pair(sl + 64, CT()); hi[0] = t.x; hi[1] = t.y;
                t = slot_pair(sl + 96, CT()); hi[2] = t.x; hi[3] = t.y;
                if (!qvalid) {
#pragma unroll
                    for (int j = 0; j < 4; j++) lo[j] = hi[j] = 0;
                }
                if (ISO) {
                    int x[4];
                    hsynth0(lo, x);
#pragma unroll
                    for (int j = 0; j < 4; j++) lo[j] = x[j];
                    hsynth0(hi, x);
#pragma unroll
                    for (int j = 0; j < 4; j++) hi[j] = x[j];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    e[c][j] = even_upd(lo[j], hp[c][j], hi[j]);
                    o[c][j] = odd_upd(hp[c][j], ep[c][j], e[c][j]);
                    hp[c][j] = hi[j];
                }
            }
            emit_row(ep);
            emit_row(o);
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) ep[c][j] = e[c][j];
        } else {                                                       // bottom edge (h even): last odd row += x[n-2]
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) o[c][j] = odd_last(hp[c][j], ep[c][j]);
            emit_row(ep);
            emit_row(o);
        }
    }
    cp_wait<0>();
}

template <int NC, typename CT>
cudaError_t run_ct(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    const size_t smem = (size_t)kWarps * kDepth * NC * 4 * 32 * sizeof(uint2);
    const DevTile *tiles = p.d_tiles + p.tile_first;
    cudaError_t e;
    if (p.iso) {
        if ((e = cudaFuncSetAttribute(k_idwt53_fused<NC, CT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        J2K_LAUNCH((k_idwt53_fused<NC, CT, true>), grid, kWarps * 32, smem, s, p.d_tcs, tiles, (const CT *)p.d_coef,
                   (const int32_t *)p.d_tmp, p.d_pix, p.nlevels, strip_pairs, p.tail);
    } else {
        if ((e = cudaFuncSetAttribute(k_idwt53_fused<NC, CT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        J2K_LAUNCH((k_idwt53_fused<NC, CT, false>), grid, kWarps * 32, smem, s, p.d_tcs, tiles, (const CT *)p.d_coef,
                   (const int32_t *)p.d_tmp, p.d_pix, p.nlevels, strip_pairs, p.tail);
    }
    return cudaGetLastError();
}

template <int NC>
cudaError_t run(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    return p.coef16 ? run_ct<NC, int16_t>(p, grid, strip_pairs, s) : run_ct<NC, int32_t>(p, grid, strip_pairs, s);
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
    switch (p.tail.ncomp) {
    case 1: return run<1>(p, grid, sp, s);
    case 3: return run<3>(p, grid, sp, s);
    case 4: return run<4>(p, grid, sp, s);
    }
    return cudaErrorInvalidValue;
}
