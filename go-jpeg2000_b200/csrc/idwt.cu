// idwt.cu -- inverse 5-3 (int32) / 9-7 (float64) lifting DWT, one decomposition level per launch,
// shared-memory tiled with halo, with the fused "last level + inverse MCT + DC shift + clamp + pack"
// epilogue (sm_100a).
//
// Replaces dwt.Inverse53/Inverse97 (reference internal/dwt/dwt.go:122-147, 213-262), Inverse2D53/97
// (dwt.go:410-429, 454-473), ReconstructMultiLevel53/97 (dwt.go:534-573), tcd.ApplyInverseDWT
// (internal/tcd/tcd.go:416-437) and, in the fused epilogue, mct.InverseRCT/InverseICT (mct.go:43-66),
// mct.DCLevelShiftInverse (mct.go:113-118), decoder.go:321-348 and createImage (decoder.go:417-588).
//
// REF addressing: level l of the reference works IN PLACE on the dense prefix
// data[0 : w_l*h_l] of the plane (stride w_l).  Here a level is out of place: element i of the level's
// input comes from the previous (coarser) level's dense output when i < w_{l+1}*h_{l+1} and from the
// coefficient plane otherwise, and the output goes to a ping-pong buffer -- every coefficient is read
// from HBM exactly once over the whole reconstruction, and the last level never writes the plane at
// all: it feeds the pixel epilogue from shared memory / registers.
//
// Per CTA: an output tile TH x TW of the level image.  The tile plus HALO interleaved samples on each
// side is gathered into a shared-memory patch in INTERLEAVED order (the reference's interleave() is
// folded into the gather; the gather walks each band row contiguously so global loads coalesce),
// then columns are lifted, then rows ("columns first, then rows", dwt.go:411-428), with the reference's
// exact edge expressions.  9-7 uses __dmul_rn/__dadd_rn so that no FMA is contracted (Go/amd64 never
// fuses) -- REF parity for float64 is bit-exact.
#include "common.h"
#include "tail.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int TW = 64;          // output tile width  (even)
constexpr int TH = 32;          // output tile height (even)

enum { EPI_STORE = 0, EPI_ROUND_I32 = 1, EPI_PIXELS = 2 };

__device__ __forceinline__ int lvl_dim(int full, int lvl) { return (full + (1 << lvl) - 1) >> lvl; }

// ---- lifting policies -------------------------------------------------------------------------------
struct Lift53 {
    typedef int32_t T;
    static constexpr int HALO = 2;
    static constexpr int STEPS = 2;
    static __device__ __forceinline__ T from_i32(int32_t v) { return v; }
    static __device__ __forceinline__ T scale(T v, int) { return v; }
    static constexpr bool kScale = false;
    // step 0: even samples  x -= (l + r + 2) >> 2        (dwt.go:132-138)
    // step 1: odd samples   x += (l + r) >> 1, last odd of an even-length line: x += l   (dwt.go:141-146)
    static __device__ __forceinline__ T apply(int step, T x, T l, T r, bool has_l, bool has_r)
    {
        uint32_t ux = (uint32_t)x, ul = (uint32_t)l, ur = (uint32_t)r;
        if (step == 0) {
            if (!has_l) ul = ur;
            if (!has_r) ur = ul;
            return (T)(ux - (uint32_t)((int32_t)(ul + ur + 2u) >> 2));
        }
        if (!has_r) return (T)(ux + ul);
        return (T)(ux + (uint32_t)((int32_t)(ul + ur) >> 1));
    }
};

struct Lift97 {
    typedef double T;
    static constexpr int HALO = 4;
    static constexpr int STEPS = 4;
    static constexpr bool kScale = true;
    static __device__ __forceinline__ T from_i32(int32_t v) { return (double)v; }        // tcd.go:429-431
    static __device__ __forceinline__ T scale(T v, int odd)                               // dwt.go:222-227
    {
        return __dmul_rn(v, odd ? 0.812893066115961 : 1.230174104914001);
    }
    // steps: delta (even), gamma (odd), beta (even), alpha (odd): x -= c * (l + r); at a line end the
    // reference uses (2c) * neighbour, which is bit-identical to c * (n + n).   (dwt.go:229-261)
    static __device__ __forceinline__ T apply(int step, T x, T l, T r, bool has_l, bool has_r)
    {
        const double c = step == 0 ? 0.443506852043971 : step == 1 ? 0.882911075530934
                       : step == 2 ? -0.052980118572961 : -1.586134342059924;
        if (!has_l) l = r;
        if (!has_r) r = l;
        return __dsub_rn(x, __dmul_rn(c, __dadd_rn(l, r)));
    }
};

// ISO mode, irreversible: float32 with OpenJPEG's constants and operation order (opj_v8dwt_decode: low-pass * K,
// high-pass * 1.625732422 with the sub-band gain left out of the step size -- here the step carries the standard
// gain, so the factor is 1.625732422 / 2; then x += (l + r) * c for -delta, -gamma, -beta, -alpha).  Powers of two
// commute with float rounding, so the results are bit-identical to OpenJPEG's decoder.
struct Lift97F {
    typedef float T;
    static constexpr int HALO = 4;
    static constexpr int STEPS = 4;
    static constexpr bool kScale = true;
    static __device__ __forceinline__ T from_i32(int32_t v) { return __int_as_float(v); }       // planes hold float bits
    static __device__ __forceinline__ T scale(T v, int odd) { return __fmul_rn(v, odd ? 0.8128662109375f : 1.230174105f); }
    static __device__ __forceinline__ T apply(int step, T x, T l, T r, bool has_l, bool has_r)
    {
        const float c = step == 0 ? 0.443506852f : step == 1 ? 0.882911075f : step == 2 ? -0.052980118f : -1.586134342f;
        if (!has_l) l = r;
        if (!has_r) r = l;
        return __fsub_rn(x, __fmul_rn(__fadd_rn(l, r), c));
    }
};

// ---- one level of one tile-component into the shared patch -------------------------------------------
struct LevelGeom {
    int w, h;          // level image size
    int nlx, nly;      // number of low-pass columns / rows
    uint32_t nprev;    // REF: elements taken from the previous level's output; ISO: non-zero if LL comes from it
    uint32_t W;        // ISO: row stride of the Mallat coefficient plane (full tile-component width)
    int coef16;        // coefficient plane holds int16 instead of int32
};

__device__ __forceinline__ int32_t ld_coef_i32(const void *coef, uint32_t lin, int coef16)
{
    return coef16 ? (int32_t)((const int16_t *)coef)[lin] : ((const int32_t *)coef)[lin];
}

// REF (dense prefix): interleaved (yy, xx) -> linear index in the level image, prev below nprev.
// ISO (Mallat): LL from the previous level's dense output (or the plane when coarsest), HL/LH/HH from the plane.
template <class L, bool IN_F64, bool ISO>
__device__ __forceinline__ typename L::T load_src(const LevelGeom &g, int yy, int xx, bool do_v, bool do_h,
                                                  const typename L::T *prev, const void *coef)
{
    const int by = do_v ? (yy >> 1) : yy, bx = do_h ? (xx >> 1) : xx;
    const bool hy = do_v && (yy & 1), hx = do_h && (xx & 1);
    if (ISO) {
        if (!hy && !hx && g.nprev) return prev[(uint32_t)by * (uint32_t)g.nlx + (uint32_t)bx];
        const uint32_t lin = (uint32_t)(hy ? g.nly + by : by) * g.W + (uint32_t)(hx ? g.nlx + bx : bx);
        if (IN_F64) return (typename L::T)((const double *)coef)[lin];
        return L::from_i32(ld_coef_i32(coef, lin, g.coef16));
    }
    const uint32_t lin = (uint32_t)(hy ? g.nly + by : by) * (uint32_t)g.w + (uint32_t)(hx ? g.nlx + bx : bx);
    if (lin < g.nprev) return prev[lin];
    if (IN_F64) return (typename L::T)((const double *)coef)[lin];
    return L::from_i32(ld_coef_i32(coef, lin, g.coef16));
}

// REF: columns first, then rows (reference dwt.go:411-428).  ISO: rows first, then columns (15444-1 F.3.2:
// HOR_SR before VER_SR) -- with integer lifting the order is observable.
template <class L, bool IN_F64, bool ISO>
__device__ void transform_tile(typename L::T *P, const LevelGeom &g, int x0, int y0,
                               const typename L::T *prev, const void *coef, bool no_xform)
{
    typedef typename L::T T;
    constexpr int HALO = L::HALO, PW = TW + 2 * HALO, PH = TH + 2 * HALO, PP = PW + 1;
    const int tid = threadIdx.x;
    const bool do_v = g.h >= 2 && !no_xform, do_h = g.w >= 2 && !no_xform;

    // gather: patch row j <-> interleaved row y0-HALO+j; walk de-interleaved column order so that
    // consecutive threads read consecutive addresses of the band row
    for (int e = tid; e < PH * PW; e += kThreads) {
        int j = e / PW, k = e - j * PW;
        int i = (k < PW / 2) ? 2 * k : 2 * (k - PW / 2) + 1;
        int yy = y0 - HALO + j, xx = x0 - HALO + i;
        if (yy < 0 || yy >= g.h || xx < 0 || xx >= g.w) continue;
        T v = load_src<L, IN_F64, ISO>(g, yy, xx, do_v, do_h, prev, coef);
        if (L::kScale) {                                   // 9-7: K on even / 1/K on odd, first direction here
            if (ISO) { if (do_h) v = L::scale(v, xx & 1); }
            else if (do_v) v = L::scale(v, yy & 1);
        }
        P[j * PP + i] = v;
    }
    __syncthreads();

    // vertical lifting over patch columns [ca, cb); horizontal lifting over patch rows [ra, rb)
    auto vpass = [&](int ca, int cb, bool prescale) {
        if (!do_v) return;
        const int nc = cb - ca;
        if (prescale && L::kScale) {
            for (int e = tid; e < PH * nc; e += kThreads) {
                int j = e / nc, i = ca + (e - j * nc);
                int yy = y0 - HALO + j, xx = x0 - HALO + i;
                if (yy < 0 || yy >= g.h || xx < 0 || xx >= g.w) continue;
                P[j * PP + i] = L::scale(P[j * PP + i], yy & 1);
            }
            __syncthreads();
        }
#pragma unroll
        for (int s = 0; s < L::STEPS; s++) {
            const int margin = L::STEPS - 2 - s;
            int ya = y0 - (margin > 0 ? margin : 0), yb = y0 + TH + margin;       // inclusive range
            if (ya < 0) ya = 0;
            if (yb > g.h - 1) yb = g.h - 1;
            if ((ya & 1) != (s & 1)) ya++;
            const int nrows = yb >= ya ? ((yb - ya) >> 1) + 1 : 0;
            for (int e = tid; e < nrows * nc; e += kThreads) {
                int r = e / nc, i = ca + (e - r * nc);
                int yy = ya + 2 * r, j = yy - (y0 - HALO);
                int xx = x0 - HALO + i;
                if (xx < 0 || xx >= g.w) continue;
                bool hl = yy - 1 >= 0, hr = yy + 1 < g.h;
                T l = hl ? P[(j - 1) * PP + i] : T(0), rr = hr ? P[(j + 1) * PP + i] : T(0);
                P[j * PP + i] = L::apply(s, P[j * PP + i], l, rr, hl, hr);
            }
            __syncthreads();
        }
    };
    auto hpass = [&](int ra, int rb, bool prescale) {
        if (!do_h) return;
        const int nr = rb - ra;
        if (prescale && L::kScale) {
            for (int e = tid; e < nr * PW; e += kThreads) {
                int r = e / PW, i = e - r * PW;
                int j = ra + r, yy = y0 - HALO + j, xx = x0 - HALO + i;
                if (yy < 0 || yy >= g.h || xx < 0 || xx >= g.w) continue;
                P[j * PP + i] = L::scale(P[j * PP + i], xx & 1);
            }
            __syncthreads();
        }
#pragma unroll
        for (int s = 0; s < L::STEPS; s++) {
            const int margin = L::STEPS - 2 - s;
            int xa = x0 - (margin > 0 ? margin : 0), xb = x0 + TW + margin;
            if (xa < 0) xa = 0;
            if (xb > g.w - 1) xb = g.w - 1;
            if ((xa & 1) != (s & 1)) xa++;
            const int ncols = xb >= xa ? ((xb - xa) >> 1) + 1 : 0;
            for (int e = tid; e < nr * ncols; e += kThreads) {
                int r = e / ncols, c = e - r * ncols;
                int j = ra + r, yy = y0 - HALO + j;
                if (yy < 0 || yy >= g.h) continue;
                int xx = xa + 2 * c, i = xx - (x0 - HALO);
                bool hl = xx - 1 >= 0, hr = xx + 1 < g.w;
                T l = hl ? P[j * PP + i - 1] : T(0), rr = hr ? P[j * PP + i + 1] : T(0);
                P[j * PP + i] = L::apply(s, P[j * PP + i], l, rr, hl, hr);
            }
            __syncthreads();
        }
    };
    if (ISO) {
        hpass(0, PH, false);                  // every patch row, horizontally scaled in the gather
        vpass(HALO, HALO + TW, true);         // then the TW output columns
    } else {
        vpass(0, PW, false);                  // every patch column, vertically scaled in the gather
        hpass(HALO, HALO + TH, true);         // then the TH output rows
    }
}

// ---- kernel: one level, every tile-component (EPI_STORE / EPI_ROUND_I32) ------------------------------
template <class L, bool IN_F64, int EPI, bool ISO>
__global__ void __launch_bounds__(kThreads)
k_idwt_level(const DevTileComp *__restrict__ tcs, const void *__restrict__ coef, typename L::T *tmp,
             void *out_planes, int nlevels, int lvl, int coef16)
{
    typedef typename L::T T;
    constexpr int HALO = L::HALO, PW = TW + 2 * HALO, PH = TH + 2 * HALO, PP = PW + 1;
    J2K_DYN_SMEM(T, P);

    const DevTileComp tc = tcs[blockIdx.z];
    LevelGeom g;
    g.w = lvl_dim((int)tc.w, lvl); g.h = lvl_dim((int)tc.h, lvl);
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (x0 >= g.w || y0 >= g.h) return;
    g.nlx = (g.w + 1) >> 1; g.nly = (g.h + 1) >> 1;
    const bool no_xform = nlevels == 0;
    g.nprev = (lvl + 1 < nlevels) ? (uint32_t)g.nlx * (uint32_t)g.nly : 0u;
    g.W = tc.w;
    g.coef16 = coef16;
    T *pp0 = tmp + tc.tmp_off, *pp1 = pp0 + tc.tmp_elems;
    const T *prev = ((lvl + 1) & 1) ? pp1 : pp0;
    const void *cbase = IN_F64 ? (const void *)((const double *)coef + tc.coef_off)
                      : coef16 ? (const void *)((const int16_t *)coef + tc.coef_off)
                               : (const void *)((const int32_t *)coef + tc.coef_off);
    transform_tile<L, IN_F64, ISO>(P, g, x0, y0, prev, cbase, no_xform);

    for (int e = threadIdx.x; e < TH * TW; e += kThreads) {
        int r = e / TW, c = e - r * TW;
        int yy = y0 + r, xx = x0 + c;
        if (yy >= g.h || xx >= g.w) continue;
        T v = P[(r + HALO) * PP + c + HALO];
        size_t o = (size_t)yy * g.w + xx;
        if (lvl > 0) {
            ((lvl & 1) ? pp1 : pp0)[o] = v;
        } else if (EPI == EPI_ROUND_I32) {
            ((int32_t *)out_planes)[tc.coef_off + o] = j2k_f64_to_i32(__dadd_rn((double)v, 0.5));   // tcd.go:433-435
        } else {
            ((T *)out_planes)[tc.coef_off + o] = v;
        }
    }
}

// ---- kernel: last level of every tile, fused with inverse MCT + DC shift + clamp + pack ---------------
template <class L, bool ISO, bool CC>
__global__ void __launch_bounds__(kThreads)
k_idwt_last_pixels(const DevTileComp *__restrict__ tcs, const DevTile *__restrict__ tiles,
                   const void *__restrict__ coef, typename L::T *tmp, uint8_t *pix, int nlevels,
                   TailParams tp, int coef16)
{
    typedef typename L::T T;
    constexpr int HALO = L::HALO, PW = TW + 2 * HALO, PP = PW + 1;
    constexpr int PER = TH * TW / kThreads;
    J2K_DYN_SMEM(T, P);

    const DevTile tile = tiles[blockIdx.z];
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (x0 >= (int)tile.w || y0 >= (int)tile.h) return;
    LevelGeom g;
    g.w = (int)tile.w; g.h = (int)tile.h;
    g.nlx = (g.w + 1) >> 1; g.nly = (g.h + 1) >> 1;
    g.nprev = nlevels > 1 ? (uint32_t)g.nlx * (uint32_t)g.nly : 0u;
    g.W = tile.w;
    g.coef16 = coef16;

    constexpr bool kIsoIrrev = ISO && sizeof(T) == 4 && L::kScale;      // float32 samples stay float until after the ICT
    int32_t acc[4][PER];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c >= tp.ncomp) break;
        const DevTileComp tc = tcs[tile.tc[c]];
        const T *prev = tmp + tc.tmp_off + tc.tmp_elems;          // level 1 wrote ping-pong buffer 1
        const void *cbase = coef16 ? (const void *)((const int16_t *)coef + tc.coef_off)
                                   : (const void *)((const int32_t *)coef + tc.coef_off);
        transform_tile<L, false, ISO>(P, g, x0, y0, prev, cbase, nlevels == 0);
#pragma unroll
        for (int k = 0; k < PER; k++) {
            int e = threadIdx.x + k * kThreads;
            int r = e / TW, cc = e - r * TW;
            T v = P[(r + HALO) * PP + cc + HALO];
            if (sizeof(T) == 8) acc[c][k] = j2k_f64_to_i32(__dadd_rn((double)v, 0.5));   // tcd.go:433-435
            else if (kIsoIrrev) acc[c][k] = __float_as_int((float)v);
            else acc[c][k] = (int32_t)v;
        }
        __syncthreads();
    }
    uint8_t *img = pix + tile.out_off;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int e = threadIdx.x + k * kThreads;
        int r = e / TW, cc = e - r * TW;
        int yy = y0 + r, xx = x0 + cc;
        if (yy >= g.h || xx >= g.w) continue;
        uint32_t gx = tile.img_x0 + xx, gy = tile.img_y0 + yy;
        if (gx >= tile.img_w || gy >= tile.img_h) continue;      // decoder.go:398-410 clipping
        int32_t v[4] = {acc[0][k], tp.ncomp > 1 ? acc[1][k] : 0, tp.ncomp > 2 ? acc[2][k] : 0,
                        tp.ncomp > 3 ? acc[3][k] : 0};
        if (kIsoIrrev) {
            const float f[4] = {__int_as_float(v[0]), __int_as_float(v[1]), __int_as_float(v[2]), __int_as_float(v[3])};
            tail_iso_irrev(f, v, tp);
        } else {
            tail_mct_dc(v, tp);
        }
        if (CC) tail_colour(v, tp);                              // decoder.go:350-356
        store_pixel(img + (size_t)gy * tile.out_stride, gx, v, tp);
    }
}

// ---- unfused tail over whole planes (stage entry points) ------------------------------------------------
__global__ void k_tail(const int32_t *c0, const int32_t *c1, const int32_t *c2, const int32_t *c3,
                       int32_t *o0, int32_t *o1, int32_t *o2, int32_t *o3, uint8_t *pix, uint64_t out_stride,
                       uint32_t width, uint32_t height, TailParams tp, int apply_tail)
{
    uint64_t n = (uint64_t)width * height;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        int32_t v[4] = {c0[i], tp.ncomp > 1 ? c1[i] : 0, tp.ncomp > 2 ? c2[i] : 0, tp.ncomp > 3 ? c3[i] : 0};
        if (apply_tail) { tail_mct_dc(v, tp); tail_colour(v, tp); }
        if (o0) { o0[i] = v[0]; if (tp.ncomp > 1) o1[i] = v[1]; if (tp.ncomp > 2) o2[i] = v[2]; if (tp.ncomp > 3) o3[i] = v[3]; }
        if (pix) {
            uint32_t y = (uint32_t)(i / width), x = (uint32_t)(i - (uint64_t)y * width);
            store_pixel(pix + (size_t)y * out_stride, x, v, tp);
        }
    }
}

__global__ void k_inverse_ict_f64(double *y, double *cb, double *cr, uint64_t n)       // mct.go:43-53
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double a = y[i], b = cb[i], c = cr[i];
        y[i]  = __dadd_rn(a, __dmul_rn(1.402, c));
        cb[i] = __dsub_rn(__dsub_rn(a, __dmul_rn(0.34413, b)), __dmul_rn(0.71414, c));
        cr[i] = __dadd_rn(a, __dmul_rn(1.772, b));
    }
}

// pixels no tile covers: the reference's planes are zero there (decoder.go:305-309), so they hold the pixel that zero
// coefficients decode to; `pattern` = that pixel (bpp bytes, produced by k_tail on a 1 x 1 zero image)
__global__ void k_fill_pixels(uint8_t *pix, uint64_t out_stride, uint32_t row_bytes, uint32_t rows, int bpp, const uint8_t *pattern)
{
    const uint8_t p0 = pattern[0], p1 = pattern[1 % bpp], p2 = pattern[2 % bpp], p3 = pattern[3 % bpp],
                  p4 = pattern[4 % bpp], p5 = pattern[5 % bpp], p6 = pattern[6 % bpp], p7 = pattern[7 % bpp];
    const uint8_t pat[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    const uint64_t n = (uint64_t)row_bytes * rows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / row_bytes), x = (uint32_t)(i - (uint64_t)y * row_bytes);
        pix[(size_t)y * out_stride + x] = pat[x % (uint32_t)bpp];
    }
}

template <class L>
constexpr size_t patch_bytes() { return sizeof(typename L::T) * (size_t)(TH + 2 * L::HALO) * (TW + 2 * L::HALO + 1); }

template <class L, bool IN_F64, int EPI>
cudaError_t run_level(const IdwtLaunch &p, dim3 grid, cudaStream_t s)
{
    if (p.iso)
        J2K_LAUNCH((k_idwt_level<L, IN_F64, EPI, true>), grid, kThreads, patch_bytes<L>(), s, p.d_tcs + p.tc_first,
                   p.d_coef, (typename L::T *)p.d_tmp, (void *)p.d_plane_out, p.nlevels, p.lvl, p.coef16);
    else
        J2K_LAUNCH((k_idwt_level<L, IN_F64, EPI, false>), grid, kThreads, patch_bytes<L>(), s, p.d_tcs + p.tc_first,
                   p.d_coef, (typename L::T *)p.d_tmp, (void *)p.d_plane_out, p.nlevels, p.lvl, p.coef16);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_idwt_level(const IdwtLaunch &p, cudaStream_t s, int *n_launches)
{
    const int lvl = p.lvl;
    uint32_t lw = (p.max_w + (1u << lvl) - 1) >> lvl, lh = (p.max_h + (1u << lvl) - 1) >> lvl;
    if (lw == 0 || lh == 0) return cudaSuccess;
    if (n_launches) (*n_launches)++;
    const bool pixels = (lvl == 0 && p.d_tiles != nullptr);
    dim3 grid((lw + TW - 1) / TW, (lh + TH - 1) / TH, pixels ? p.n_tiles : p.n_tc);
    if (grid.z == 0) return cudaSuccess;
    if (p.reversible && p.nlevels > 0 && ((p.stream_levels >> lvl) & 1) && (lvl > 0 || pixels))
        return launch_idwt53_stream(p, s);
    if (!p.reversible && !p.f64_io && p.nlevels > 0 && ((p.stream_levels >> lvl) & 1) && (lvl > 0 || pixels))
        return launch_idwt97_stream(p, s);
    if (pixels) {
#define J2K_LAST_PIXELS(L, ISO_, TT)                                                                                             \
        do {                                                                                                                     \
            if (p.tail.cconv)                                                                                                    \
                J2K_LAUNCH((k_idwt_last_pixels<L, ISO_, true>), grid, kThreads, patch_bytes<L>(), s, p.d_tcs,                     \
                           p.d_tiles + p.tile_first, p.d_coef, (TT *)p.d_tmp, p.d_pix, p.nlevels, p.tail, p.coef16);              \
            else                                                                                                                 \
                J2K_LAUNCH((k_idwt_last_pixels<L, ISO_, false>), grid, kThreads, patch_bytes<L>(), s, p.d_tcs,                    \
                           p.d_tiles + p.tile_first, p.d_coef, (TT *)p.d_tmp, p.d_pix, p.nlevels, p.tail, p.coef16);              \
        } while (0)
        if (!p.reversible && p.iso) J2K_LAST_PIXELS(Lift97F, true, float);
        else if (p.reversible && p.iso) J2K_LAST_PIXELS(Lift53, true, int32_t);
        else if (p.reversible) J2K_LAST_PIXELS(Lift53, false, int32_t);
        else J2K_LAST_PIXELS(Lift97, false, double);
#undef J2K_LAST_PIXELS
        return cudaGetLastError();
    }
    if (p.reversible) return run_level<Lift53, false, EPI_STORE>(p, grid, s);
    if (p.iso)        return run_level<Lift97F, false, EPI_STORE>(p, grid, s);
    if (p.f64_io)     return run_level<Lift97, true, EPI_STORE>(p, grid, s);
    return run_level<Lift97, false, EPI_ROUND_I32>(p, grid, s);
}

cudaError_t launch_tail(const int32_t *const d_comps[4], int32_t *const d_planes_out[4], uint8_t *d_pix,
                        uint64_t out_stride, uint32_t width, uint32_t height, const TailParams &tp,
                        int apply_tail, cudaStream_t s)
{
    uint64_t n = (uint64_t)width * height;
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    J2K_LAUNCH((k_tail), blocks, 256, 0, s, d_comps[0], d_comps[1], d_comps[2], d_comps[3],
                                  d_planes_out ? d_planes_out[0] : nullptr, d_planes_out ? d_planes_out[1] : nullptr,
                                  d_planes_out ? d_planes_out[2] : nullptr, d_planes_out ? d_planes_out[3] : nullptr,
                                  d_pix, out_stride, width, height, tp, apply_tail);
    return cudaGetLastError();
}

cudaError_t launch_fill_pixels(uint8_t *d_pix, uint64_t out_stride, uint32_t row_bytes, uint32_t rows, int bpp,
                               const uint8_t *d_pattern, cudaStream_t s)
{
    const uint64_t n = (uint64_t)row_bytes * rows;
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    J2K_LAUNCH((k_fill_pixels), blocks, 256, 0, s, d_pix, out_stride, row_bytes, rows, bpp, d_pattern);
    return cudaGetLastError();
}

cudaError_t launch_inverse_ict_f64(double *y, double *cb, double *cr, uint64_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    J2K_LAUNCH((k_inverse_ict_f64), blocks, 256, 0, s, y, cb, cr, n);
    return cudaGetLastError();
}
