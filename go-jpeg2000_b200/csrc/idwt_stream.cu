// idwt_stream.cu -- register-streaming inverse 5-3 DWT level (int32), optionally fused with the pixel
// epilogue (inverse RCT/ICT + DC shift + clamp + pack).  This is the HBM-roofline kernel of the path.
//
// Same arithmetic as idwt.cu (reference dwt.go:122-147 per line, columns then rows dwt.go:410-429,
// dense-prefix addressing dwt.go:534-548, epilogue decoder.go:321-348 + 417-588) but organised for
// bandwidth instead of generality:
//   * a WARP is the unit of work: 30 quads (120 output columns) x one strip of row pairs of one tile;
//     lanes 0 and 31 are halo lanes that overlap the neighbouring warps, so no shared memory and no
//     block-level barrier is needed at all;
//   * each lane owns one quad = L columns {2q, 2q+1} and H columns {2q, 2q+1} of the level image, i.e.
//     output columns 4q..4q+3.  It walks the strip top-down: per row pair it loads four int2 (low row L|H,
//     high row L|H) per component -- 256 contiguous bytes per warp and band row -- and keeps the vertical
//     lifting state (previous high row, previous even row) in registers (one row-pair of delay);
//   * the horizontal lifting of a finished row needs the neighbour quad's last H value and first even
//     output: two warp shuffles per row and component;
//   * the epilogue packs 4 RGBA pixels per lane and row into one 16-byte store; the last level never
//     writes int32 planes.
// Eligibility (checked on the host per level): every tile-component's level width is a multiple of 4 and
// its level height is even.  Anything else takes the general tiled kernel in idwt.cu.
#include "common.h"
#include "tail.cuh"

namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ int lvl_dim(int full, int lvl) { return (full + (1 << lvl) - 1) >> lvl; }

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

// two adjacent elements of a coefficient plane (CT = int32_t, or int16_t when the job's magnitudes fit 15 bits)
__device__ __forceinline__ int2 ldpair(const int32_t *p) { return __ldg(reinterpret_cast<const int2 *>(p)); }
__device__ __forceinline__ int2 ldpair(const int16_t *p)
{
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(p));
    return make_int2((int)(int16_t)(r & 0xFFFFu), (int)r >> 16);
}

template <typename CT>
struct Src {
    const int32_t *prev;     // dense output of the coarser level
    const CT *coef;          // coefficient plane
    uint32_t nprev;          // REF: elements [0, nprev) of the level image come from prev; ISO: non-zero if LL does
};

// REF (dense prefix): one linear index space, prev below nprev
__device__ __forceinline__ int2 ld2(const Src<int32_t> &s, uint32_t lin)
{
    const int32_t *p = lin < s.nprev ? s.prev : s.coef;          // one load, selected pointer
    return ldpair(p + lin);
}
__device__ __forceinline__ int2 ld2(const Src<int16_t> &s, uint32_t lin)
{
    return lin < s.nprev ? ldpair(s.prev + lin) : ldpair(s.coef + lin);
}

template <int NC, bool PIXELS, bool ISO, typename CT>
__global__ void __launch_bounds__(kWarps * 32, 3)
k_idwt53_stream(const DevTileComp *__restrict__ tcs, const DevTile *__restrict__ tiles,
                const CT *__restrict__ coef, int32_t *__restrict__ tmp, uint8_t *__restrict__ pix,
                int nlevels, int lvl, int strip_pairs, TailParams tp)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tci[NC];
    int W0, H0;
    DevTile tile;
    if (PIXELS) {
        tile = tiles[blockIdx.y];
#pragma unroll
        for (int c = 0; c < NC; c++) tci[c] = tile.tc[c];
        W0 = (int)tile.w; H0 = (int)tile.h;
    } else {
        tci[0] = blockIdx.y;
        W0 = (int)tcs[blockIdx.y].w; H0 = (int)tcs[blockIdx.y].h;
    }
    const int w = lvl_dim(W0, lvl), h = lvl_dim(H0, lvl);
    const int nlx = w >> 1, nly = h >> 1, nq = w >> 2;
    const int nwx = (nq + 29) / 30;
    const int nstrips = (nly + strip_pairs - 1) / strip_pairs;
    const int unit = blockIdx.x * kWarps + warp;
    if (unit >= nwx * nstrips) return;
    const int strip = unit / nwx, wi = unit - strip * nwx;
    const int q = wi * 30 - 1 + lane;
    const bool qvalid = q >= 0 && q < nq;
    const bool store_lane = qvalid && lane >= 1 && lane <= 30;
    const int ka = strip * strip_pairs;
    const int kb = min(ka + strip_pairs, nly);

    Src<CT> src[NC];
    int32_t *dst = nullptr;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DevTileComp tc = tcs[tci[c]];
        int32_t *pp0 = tmp + tc.tmp_off, *pp1 = pp0 + tc.tmp_elems;
        src[c].prev = ((lvl + 1) & 1) ? pp1 : pp0;
        src[c].coef = coef + tc.coef_off;
        src[c].nprev = (lvl + 1 < nlevels) ? (uint32_t)nlx * (uint32_t)nly : 0u;
        if (!PIXELS) dst = (lvl & 1) ? pp1 : pp0;
    }
    const uint32_t colL = 2u * (uint32_t)(qvalid ? q : 0), colH = (uint32_t)nlx + colL;
    const uint32_t uw = (uint32_t)w;

    int hp[NC][4], ep[NC][4];      // Hi[k-1] and E[k-1] of the lane's 4 columns, order L0, L1, H0, H1

    // r = row of the level image in band order (rows [0, nly) low-pass, [nly, h) high-pass); v = L0, L1, H0, H1
    const uint32_t planeW = (uint32_t)W0;            // ISO: stride of the Mallat plane
    auto load_row = [&](int c, int r, int v[4]) {
        if (qvalid) {
            int2 a, b;
            if (ISO) {
                const CT *rowp = src[c].coef + (size_t)r * planeW;
                b = ldpair(rowp + colH);
                if (r < nly && src[c].nprev) a = ldpair(src[c].prev + (size_t)r * (uint32_t)nlx + colL);
                else a = ldpair(rowp + colL);
            } else {
                const uint32_t lin = (uint32_t)r * uw;
                a = ld2(src[c], lin + colL); b = ld2(src[c], lin + colH);
            }
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
            v[0] = v[1] = v[2] = v[3] = 0;
        }
    };
    // horizontal synthesis of one row held as (L0, L1, H0, H1) per lane -> 4 interleaved samples; every lane
    // takes part in the two shuffles (lane 0 / 31 are halo lanes)
    auto hsynth = [&](const int V[4], int X[4]) {
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
    // raw band-row pair k of component c (loads only, so that they can be issued one step ahead)
    auto load_pair = [&](int c, int k, int lo[4], int hi[4]) {
        load_row(c, k, lo); load_row(c, nly + k, hi);
    };
    // bring a raw row into the domain the vertical lifting works in: REF = band columns (the horizontal
    // synthesis happens on finished rows), ISO = interleaved columns (horizontal synthesis first)
    auto to_domain = [&](int v[4]) {
        if (ISO) { int x[4]; hsynth(v, x); v[0] = x[0]; v[1] = x[1]; v[2] = x[2]; v[3] = x[3]; }
    };

    // pixel addressing of this lane's 4 output columns
    const uint32_t gx0 = PIXELS ? tile.img_x0 + 4u * (uint32_t)(qvalid ? q : 0) : 0u;
    const bool fast_rgba8 = PIXELS && tp.fmt == J2KGPU_FMT_RGBA8 && tp.prec[0] == 8 && NC >= 3 &&
                            ((tile.out_stride & 15) == 0) && ((tile.img_x0 & 3) == 0) && (gx0 + 3 < tile.img_w);

    // horizontal lifting of one finished row (all lanes take part in the shuffles) + store / epilogue
    auto emit_row = [&](int y, int V[NC][4]) {
        int X[NC][4];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            if (ISO) {
#pragma unroll
                for (int j = 0; j < 4; j++) X[c][j] = V[c][j];
            } else {
                hsynth(V[c], X[c]);
            }
        }
        if (!store_lane) return;
        if (!PIXELS) {
            *reinterpret_cast<int4 *>(dst + (size_t)y * uw + 4u * (uint32_t)q) = make_int4(X[0][0], X[0][1], X[0][2], X[0][3]);
            return;
        }
        const uint32_t gy = tile.img_y0 + (uint32_t)y;
        if (gy >= tile.img_h) return;                                  // decoder.go:398-410 clipping
        uint8_t *row = pix + tile.out_off + (size_t)gy * tile.out_stride;
        if (fast_rgba8) {
            uint32_t px[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                int32_t v[4] = {X[0][p], NC > 1 ? X[NC > 1 ? 1 : 0][p] : 0, NC > 2 ? X[NC > 2 ? 2 : 0][p] : 0,
                                NC > 3 ? X[NC > 3 ? 3 : 0][p] : 0};
                tail_mct_dc(v, tp);
                const uint32_t r = (uint32_t)clampi(v[0], 0, 255), g = (uint32_t)clampi(v[1], 0, 255),
                               b = (uint32_t)clampi(v[2], 0, 255);
                const uint32_t a = NC == 4 ? (uint32_t)clampi(v[3], 0, 255) : 255u;
                px[p] = r | (g << 8) | (b << 16) | (a << 24);
            }
            __stcs(reinterpret_cast<uint4 *>(row + 4 * (size_t)gx0), make_uint4(px[0], px[1], px[2], px[3]));
        } else {
#pragma unroll
            for (int p = 0; p < 4; p++) {
                if (gx0 + p >= tile.img_w) continue;
                int32_t v[4] = {X[0][p], NC > 1 ? X[NC > 1 ? 1 : 0][p] : 0, NC > 2 ? X[NC > 2 ? 2 : 0][p] : 0,
                                NC > 3 ? X[NC > 3 ? 3 : 0][p] : 0};
                tail_mct_dc(v, tp);
                store_pixel(row, gx0 + p, v, tp);
            }
        }
    };

    // ---- prologue: E[ka] ----
#pragma unroll
    for (int c = 0; c < NC; c++) {
        int lo[4], hi[4];
        load_pair(c, ka, lo, hi);
        if (ka > 0) load_row(c, nly + ka - 1, hp[c]);
        to_domain(lo); to_domain(hi);
        if (ka > 0) to_domain(hp[c]);
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) hp[c][j] = hi[j];              // top edge: Hi[-1] := Hi[0]
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { ep[c][j] = even_upd(lo[j], hp[c][j], hi[j]); hp[c][j] = hi[j]; }
    }

    // ---- stream the strip: step k finishes rows 2k-2 (even) and 2k-1 (odd) ----
    // The band rows of step k+1 are requested before step k is computed (register double buffer), so a warp
    // always has one full step of loads (NC x 4 x 256 B) in flight behind its arithmetic.
    int lo[NC][4], hi[NC][4];
    if (ka + 1 < nly) {
#pragma unroll
        for (int c = 0; c < NC; c++) load_pair(c, ka + 1, lo[c], hi[c]);
    }
    for (int k = ka + 1; k <= kb; k++) {
        int o[NC][4];
        if (k < nly) {
            int nlo[NC][4], nhi[NC][4];
            const bool more = (k + 1 <= kb) && (k + 1 < nly);
            if (more) {
#pragma unroll
                for (int c = 0; c < NC; c++) load_pair(c, k + 1, nlo[c], nhi[c]);
            }
            int e[NC][4];
#pragma unroll
            for (int c = 0; c < NC; c++) { to_domain(lo[c]); to_domain(hi[c]); }
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    e[c][j] = even_upd(lo[c][j], hp[c][j], hi[c][j]);
                    o[c][j] = odd_upd(hp[c][j], ep[c][j], e[c][j]);
                    hp[c][j] = hi[c][j];
                }
            emit_row(2 * k - 2, ep);
            emit_row(2 * k - 1, o);
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    ep[c][j] = e[c][j];
                    if (more) { lo[c][j] = nlo[c][j]; hi[c][j] = nhi[c][j]; }
                }
        } else {                                                       // bottom edge (h even): last odd row += x[n-2]
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) o[c][j] = odd_last(hp[c][j], ep[c][j]);
            emit_row(2 * k - 2, ep);
            emit_row(2 * k - 1, o);
        }
    }
}

template <int NC, bool PIXELS, typename CT>
cudaError_t run_ct(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    // pixel launches index tile-components through the tile table (absolute), plain levels through blockIdx.y
    const DevTileComp *tcs = PIXELS ? p.d_tcs : p.d_tcs + p.tc_first;
    const DevTile *tiles = p.d_tiles ? p.d_tiles + p.tile_first : nullptr;
    if (p.iso)
        J2K_LAUNCH((k_idwt53_stream<NC, PIXELS, true, CT>), grid, kWarps * 32, 0, s, tcs, tiles, (const CT *)p.d_coef,
                   (int32_t *)p.d_tmp, p.d_pix, p.nlevels, p.lvl, strip_pairs, p.tail);
    else
        J2K_LAUNCH((k_idwt53_stream<NC, PIXELS, false, CT>), grid, kWarps * 32, 0, s, tcs, tiles, (const CT *)p.d_coef,
                   (int32_t *)p.d_tmp, p.d_pix, p.nlevels, p.lvl, strip_pairs, p.tail);
    return cudaGetLastError();
}

template <int NC, bool PIXELS>
cudaError_t run(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    return p.coef16 ? run_ct<NC, PIXELS, int16_t>(p, grid, strip_pairs, s) : run_ct<NC, PIXELS, int32_t>(p, grid, strip_pairs, s);
}

}  // namespace

// Launch the streaming kernel for level p.lvl.  The caller has checked eligibility (see file header).
cudaError_t launch_idwt53_stream(const IdwtLaunch &p, cudaStream_t s)
{
    const int lvl = p.lvl;
    const uint32_t lw = (p.max_w + (1u << lvl) - 1) >> lvl, lh = (p.max_h + (1u << lvl) - 1) >> lvl;
    const bool pixels = (lvl == 0 && p.d_tiles != nullptr);
    const uint32_t nobj = pixels ? p.n_tiles : p.n_tc;
    if (lw < 4 || lh < 2 || nobj == 0) return cudaSuccess;
    const uint32_t nq = lw / 4, nwx = (nq + 29) / 30, nly = lh / 2;
    // strip height: as tall as possible (3 band rows of halo per strip) while keeping >= ~4 warps per scheduler
    int sp = 32;
    while (sp > 4 && (uint64_t)nobj * nwx * ((nly + sp - 1) / sp) < 148ull * 4 * 6) sp >>= 1;
    const uint32_t units = nwx * ((nly + sp - 1) / sp);
    dim3 grid((units + kWarps - 1) / kWarps, nobj, 1);
    if (pixels) {
        switch (p.tail.ncomp) {
        case 1: return run<1, true>(p, grid, sp, s);
        case 3: return run<3, true>(p, grid, sp, s);
        case 4: return run<4, true>(p, grid, sp, s);
        }
        return cudaErrorInvalidValue;
    }
    return run<1, false>(p, grid, sp, s);
}
