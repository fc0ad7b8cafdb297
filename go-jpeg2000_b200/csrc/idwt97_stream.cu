// idwt97_stream.cu -- register-streaming inverse 9-7 DWT level in float64 (REF semantics), optionally fused with the
// pixel epilogue (int32(v + 0.5) -> inverse ICT -> int32(v + 0.5) -> DC shift -> clamp -> pack).
//
// Replaces dwt.Inverse97 / Inverse2D97 / ReconstructMultiLevel97 (reference internal/dwt/dwt.go:213-262, 454-473,
// 561-573) and tcd.ApplyInverseDWT's float path (tcd.go:428-435) on the whole-path route; arithmetic is the reference's
// operation for operation (__dmul_rn / __dadd_rn / __dsub_rn, never an FMA: Go on amd64 does not fuse), so float64
// parity is bit-exact.  The general tiled kernel in idwt.cu stays the fallback for odd shapes and for the stage API.
//
// Organisation = idwt_stream.cu (a warp owns 30 quads x a strip of row pairs, lanes 0 / 31 are halo lanes, no shared
// memory, no barrier), with the 4-step lifting written as a vertical software pipeline: reading band row pair m
// (L_m, H_m) advances, per column,
//     A_m     = K L_m     - delta (Hs_{m-1} + Hs_m)        Hs = H / K            (even rows after step 1)
//     B_{m-1} = Hs_{m-1}  - gamma (A_{m-1} + A_m)                                (odd rows after step 2)
//     C_{m-1} = A_{m-1}   - beta  (B_{m-2} + B_{m-1})                            (even rows, final)
//     D_{m-2} = B_{m-2}   - alpha (C_{m-2} + C_{m-1})                            (odd rows, final)
// so output rows 2(m-2), 2(m-2)+1 leave two row pairs behind the loads; the state is 4 doubles per column.  Line ends
// mirror the missing neighbour (the reference's 2c * neighbour is bit-identical to c * (n + n)).  A strip warms the
// pipeline up over the two row pairs above it and reads two below.  Horizontal lifting of a finished row: the same
// four steps across lanes, one double shuffle per step.
// Eligibility (host, per level): width a multiple of 4, height even (same as idwt_stream.cu).
#include "common.h"
#include "tail.cuh"
#include <type_traits>

namespace {

constexpr int kWarps = 4;

// REF: float64, the reference's constants (dwt.go:150-157).  ISO: float32 with OpenJPEG's constants and operation order
// (opj_v8dwt_decode; the high-pass factor is its 1.625732422 / 2 because the step sizes here carry the standard sub-band
// gain) -- powers of two commute with float rounding, so the results are bit-identical to OpenJPEG's decoder.
template <typename T> struct C97;
template <> struct C97<double> {
    static constexpr double K = 1.230174104914001, InvK = 0.812893066115961;
    static constexpr double Delta = 0.443506852043971, Gamma = 0.882911075530934, Beta = -0.052980118572961, Alpha = -1.586134342059924;
};
template <> struct C97<float> {
    static constexpr float K = 1.230174105f, InvK = 0.8128662109375f;
    static constexpr float Delta = 0.443506852f, Gamma = 0.882911075f, Beta = -0.052980118f, Alpha = -1.586134342f;
};
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ int lvl_dim(int full, int lvl) { return (full + (1 << lvl) - 1) >> lvl; }
// x - c * (l + r), three roundings, never an FMA (dwt.go:229-261; OpenJPEG: x += (l + r) * -c, the same value)
template <typename T>
__device__ __forceinline__ T lift(T x, T c, T l, T r) { return sub_rn(x, mul_rn(c, add_rn(l, r))); }

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

// ---- band samples travel global -> shared with cp.async into warp-private slots (16 bytes per lane and part), one row
// pair ahead of the arithmetic; every lane reads back only what it copied itself, so the kernel has no barrier ----
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
#ifdef J2K_EMU
    memcpy(smem_dst, gsrc, BYTES);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
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
// two neighbouring elements: copy, and read back as T
__device__ __forceinline__ void pair_copy(uint4 *slot, const int32_t *g) { cp_async<8>(slot, g); }
__device__ __forceinline__ void pair_copy(uint4 *slot, const int16_t *g) { cp_async<4>(slot, g); }
__device__ __forceinline__ void pair_copy(uint4 *slot, const float *g) { cp_async<8>(slot, g); }
__device__ __forceinline__ void pair_copy(uint4 *slot, const double *g) { cp_async<16>(slot, g); }
__device__ __forceinline__ void pair_read(const uint4 *slot, int32_t, double &a, double &b)
{
    const int2 v = *reinterpret_cast<const int2 *>(slot);
    a = (double)v.x; b = (double)v.y;                                                     // tcd.go:429-431
}
__device__ __forceinline__ void pair_read(const uint4 *slot, int16_t, double &a, double &b)
{
    const uint32_t r = *reinterpret_cast<const uint32_t *>(slot);
    a = (double)(int)(int16_t)(r & 0xFFFFu); b = (double)((int)r >> 16);
}
__device__ __forceinline__ void pair_read(const uint4 *slot, double, double &a, double &b)
{
    const double2 v = *reinterpret_cast<const double2 *>(slot);
    a = v.x; b = v.y;
}
__device__ __forceinline__ void pair_read(const uint4 *slot, float, float &a, float &b)  // ISO: planes and levels are float32
{
    const float2 v = *reinterpret_cast<const float2 *>(slot);
    a = v.x; b = v.y;
}

// the general pixel epilogue of four neighbouring pixels, out of line (any component count, precision, signedness, format,
// colour conversion); mct_dc: REF semantics still need the inverse MCT + DC shift (ISO values arrive with both applied)
static __device__ J2K_NOINLINE void store_quad_generic(uint8_t *row, uint32_t gx0, uint32_t img_w, int32_t (*v)[4], const TailParams &tp, bool mct_dc)
{
    for (int p = 0; p < 4; p++) {
        if (mct_dc) tail_mct_dc(v[p], tp);
        if (gx0 + p < img_w) store_pixel(row, gx0 + p, v[p], tp);
    }
}

// T = double, ISO = false: REF semantics (dense prefix, columns then rows, planes int32 / int16).
// T = float,  ISO = true : ISO semantics (Mallat layout, rows then columns, planes hold dequantised float32).
template <int NC, bool PIXELS, typename CT, typename T, bool ISO>
__global__ void __launch_bounds__(kWarps * 32, sizeof(T) == 4 ? 3 : 2)
k_idwt97_stream(const DevTileComp *__restrict__ tcs, const DevTile *__restrict__ tiles, const CT *__restrict__ coef,
                T *__restrict__ tmp, uint8_t *__restrict__ pix, int nlevels, int lvl, int strip_pairs, TailParams tp)
{
    typedef C97<T> K;
    __shared__ uint4 s_slots[kWarps][NC * 4][32];        // [component][L, H of the low row; L, H of the high row][lane]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *slots = &s_slots[warp][0][lane];              // part i of component c: slots[(4 c + i) * 32]
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
    const bool q_first = q == 0, q_last = q == nq - 1;
    const int qc = qvalid ? q : 0;
    const int ka = strip * strip_pairs, kb = min(ka + strip_pairs, nly);

    const T *prev[NC];
    const CT *plane[NC];
    T *dst = nullptr;
    uint32_t nprev = 0;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DevTileComp tc = tcs[tci[c]];
        T *pp0 = tmp + tc.tmp_off, *pp1 = pp0 + tc.tmp_elems;
        prev[c] = ((lvl + 1) & 1) ? pp1 : pp0;
        plane[c] = coef + tc.coef_off;
        if (!PIXELS) dst = (lvl & 1) ? pp1 : pp0;
    }
    nprev = (lvl + 1 < nlevels) ? (uint32_t)nlx * (uint32_t)nly : 0u;      // dense prefix: elements below come from prev
    const uint32_t uw = (uint32_t)w;
    const uint32_t colL = 2u * (uint32_t)qc, colH = (uint32_t)nlx + colL;

    // band row pair m of the level image (row m of [0, nly) low-pass, row nly + m high-pass) -> (L0, L1, H0, H1) of this
    // lane for both rows.  issue_pair starts the copies into the four parts of component c's slots, read_pair takes them
    // out.  Addresses: one 64-bit base per component and source, one 32-bit element offset per part.
    const uint32_t planeW = (uint32_t)W0;             // ISO: row stride of the Mallat plane
    const uint32_t rs = ISO ? planeW : uw;            // row stride of the coefficient plane at this level
    const uint32_t hi_row0 = (uint32_t)nly * rs;
    auto issue_pair = [&](int c, int m) {
        uint4 *sl = slots + 4 * c * 32;
        const uint32_t lo = (uint32_t)m * rs, hi = hi_row0 + lo;
        if (ISO) {
            if (nprev) pair_copy(sl, prev[c] + ((uint32_t)m * (uint32_t)nlx + colL));
            else pair_copy(sl, plane[c] + (lo + colL));
            pair_copy(sl + 32, plane[c] + (lo + colH));
            pair_copy(sl + 64, plane[c] + (hi + colL));
            pair_copy(sl + 96, plane[c] + (hi + colH));
        } else {
            if (lo + colL < nprev) pair_copy(sl, prev[c] + (lo + colL)); else pair_copy(sl, plane[c] + (lo + colL));
            if (lo + colH < nprev) pair_copy(sl + 32, prev[c] + (lo + colH)); else pair_copy(sl + 32, plane[c] + (lo + colH));
            if (hi + colL < nprev) pair_copy(sl + 64, prev[c] + (hi + colL)); else pair_copy(sl + 64, plane[c] + (hi + colL));
            if (hi + colH < nprev) pair_copy(sl + 96, prev[c] + (hi + colH)); else pair_copy(sl + 96, plane[c] + (hi + colH));
        }
    };
    auto read_pair = [&](int c, int m, T lo4[4], T hi4[4]) {
        const uint4 *sl = slots + 4 * c * 32;
        if (ISO) {                                       // previous level and plane are both float32
            pair_read(sl, T(), lo4[0], lo4[1]); pair_read(sl + 32, T(), lo4[2], lo4[3]);
            pair_read(sl + 64, T(), hi4[0], hi4[1]); pair_read(sl + 96, T(), hi4[2], hi4[3]);
        } else {
            const uint32_t lo = (uint32_t)m * rs, hi = hi_row0 + lo;
            if (lo + colL < nprev) pair_read(sl, T(), lo4[0], lo4[1]); else pair_read(sl, CT(), lo4[0], lo4[1]);
            if (lo + colH < nprev) pair_read(sl + 32, T(), lo4[2], lo4[3]); else pair_read(sl + 32, CT(), lo4[2], lo4[3]);
            if (hi + colL < nprev) pair_read(sl + 64, T(), hi4[0], hi4[1]); else pair_read(sl + 64, CT(), hi4[0], hi4[1]);
            if (hi + colH < nprev) pair_read(sl + 96, T(), hi4[2], hi4[3]); else pair_read(sl + 96, CT(), hi4[2], hi4[3]);
        }
    };

    // horizontal synthesis of a finished row (dwt.go:213-262 on the row), band order in, interleaved out
    auto hsynth = [&](const T V[4], T X[4]) {
        T e0 = mul_rn(V[0], (T)K::K), e1 = mul_rn(V[1], (T)K::K), o0 = mul_rn(V[2], (T)K::InvK), o1 = mul_rn(V[3], (T)K::InvK);
        T ol = __shfl_up_sync(0xffffffffu, o1, 1);
        ol = q_first ? o0 : ol;
        e0 = lift<T>(e0, K::Delta, ol, o0); e1 = lift<T>(e1, K::Delta, o0, o1);
        T er = __shfl_down_sync(0xffffffffu, e0, 1);
        er = q_last ? e1 : er;
        o0 = lift<T>(o0, K::Gamma, e0, e1); o1 = lift<T>(o1, K::Gamma, e1, er);
        ol = __shfl_up_sync(0xffffffffu, o1, 1);
        ol = q_first ? o0 : ol;
        e0 = lift<T>(e0, K::Beta, ol, o0); e1 = lift<T>(e1, K::Beta, o0, o1);
        er = __shfl_down_sync(0xffffffffu, e0, 1);
        er = q_last ? e1 : er;
        o0 = lift<T>(o0, K::Alpha, e0, e1); o1 = lift<T>(o1, K::Alpha, e1, er);
        X[0] = e0; X[1] = o0; X[2] = e1; X[3] = o1;
    };

    const uint32_t gx0 = PIXELS ? tile.img_x0 + 4u * (uint32_t)qc : 0u;
    const bool fast_rgba8 = PIXELS && NC == 3 && tp.fmt == J2KGPU_FMT_RGBA8 && tp.prec[0] == 8 && tp.prec[1] == 8 && tp.prec[2] == 8 &&
                            !tp.sgnd[0] && !tp.sgnd[1] && !tp.sgnd[2] && (ISO || (tp.mct && !tp.reversible)) && !tp.cconv &&
                            ((tile.out_stride & 15) == 0) && ((tile.out_off & 15) == 0) && ((tile.img_x0 & 3) == 0) &&
                            (gx0 + 3 < tile.img_w) && (((uintptr_t)pix & 15) == 0);

    // horizontal lifting of one finished row (all lanes take part in the shuffles) + store / epilogue
    // REF: the finished row is in band-column order and still needs the horizontal synthesis; ISO: it is final
    auto emit_row = [&](int y, T V[NC][4]) {
        T X[NC][4];
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
            T *o = dst + (size_t)y * uw + 4u * (uint32_t)q;
            if (sizeof(T) == 8) {
                reinterpret_cast<double2 *>(o)[0] = make_double2((double)X[0][0], (double)X[0][1]);
                reinterpret_cast<double2 *>(o)[1] = make_double2((double)X[0][2], (double)X[0][3]);
            } else {
                *reinterpret_cast<float4 *>(o) = make_float4((float)X[0][0], (float)X[0][1], (float)X[0][2], (float)X[0][3]);
            }
            return;
        }
        const uint32_t gy = tile.img_y0 + (uint32_t)y;
        if (gy >= tile.img_h) return;                                  // decoder.go:398-410 clipping
        uint8_t *row = pix + tile.out_off + (size_t)gy * tile.out_stride;
        if (fast_rgba8) {
            // three unsigned 8-bit components -> RGBA8: the epilogue with everything constant folded in (the general one
            // below is out of line: inlined eight times it made the kernel 150 KB of code, beyond the instruction cache)
            uint32_t px[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                int r8, g8, b8;
                if (ISO) {                                             // opj_mct_decode_real, round to nearest even, DC shift
                    float yv = (float)X[0][p], u = (float)X[NC > 1 ? 1 : 0][p], wv = (float)X[NC > 2 ? 2 : 0][p];
                    if (tp.mct) {
                        const float r = __fadd_rn(yv, __fmul_rn(wv, 1.402f));
                        const float g = __fsub_rn(__fsub_rn(yv, __fmul_rn(u, 0.34413f)), __fmul_rn(wv, 0.71414f));
                        const float b = __fadd_rn(yv, __fmul_rn(u, 1.772f));
                        yv = r; u = g; wv = b;
                    }
                    r8 = (int)((uint32_t)__float2int_rn(yv) + 128u); g8 = (int)((uint32_t)__float2int_rn(u) + 128u); b8 = (int)((uint32_t)__float2int_rn(wv) + 128u);
                } else {                                               // tcd.go:433-435, mct.go:43-53 via decoder.go:326-340, mct.go:113-118
                    const double yv = (double)j2k_f64_to_i32(__dadd_rn((double)X[0][p], 0.5)), cb = (double)j2k_f64_to_i32(__dadd_rn((double)X[NC > 1 ? 1 : 0][p], 0.5)),
                                 cr = (double)j2k_f64_to_i32(__dadd_rn((double)X[NC > 2 ? 2 : 0][p], 0.5));
                    const double r = __dadd_rn(yv, __dmul_rn(1.402, cr));
                    const double g = __dsub_rn(__dsub_rn(yv, __dmul_rn(0.34413, cb)), __dmul_rn(0.71414, cr));
                    const double b = __dadd_rn(yv, __dmul_rn(1.772, cb));
                    r8 = (int)((uint32_t)j2k_f64_to_i32(__dadd_rn(r, 0.5)) + 128u); g8 = (int)((uint32_t)j2k_f64_to_i32(__dadd_rn(g, 0.5)) + 128u);
                    b8 = (int)((uint32_t)j2k_f64_to_i32(__dadd_rn(b, 0.5)) + 128u);
                }
                px[p] = pack_sat_u8(g8, r8, pack_sat_u8(255, b8, 0u));
            }
            __stcs(reinterpret_cast<uint4 *>(row + 4 * (size_t)gx0), make_uint4(px[0], px[1], px[2], px[3]));
            return;
        }
        int32_t v[4][4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            if (ISO) {
                const float f[4] = {(float)X[0][p], NC > 1 ? (float)X[NC > 1 ? 1 : 0][p] : 0.f, NC > 2 ? (float)X[NC > 2 ? 2 : 0][p] : 0.f,
                                    NC > 3 ? (float)X[NC > 3 ? 3 : 0][p] : 0.f};
                tail_iso_irrev(f, v[p], tp);
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) v[p][c] = c < NC ? j2k_f64_to_i32(__dadd_rn((double)X[c < NC ? c : 0][p], 0.5)) : 0;   // tcd.go:433-435
            }
        }
        store_quad_generic(row, gx0, tile.img_w, v, tp, !ISO);
    };

    // ---- vertical pipeline -------------------------------------------------------------------------------------------
    T hs[NC][4], a[NC][4], b[NC][4], cc[NC][4];   // Hs_{m-1}, A_{m-1}, B_{m-2}, C_{m-2} on entry of step m
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) hs[c][j] = a[c][j] = b[c][j] = cc[c][j] = (T)0;
    const int ms = ka >= 2 ? ka - 2 : 0;
    const int me = min(kb + 1, nly + 1);
#pragma unroll
    for (int c = 0; c < NC; c++) {                       // group c: component c's first band row pair
        if (ms < nly) issue_pair(c, ms);
        cp_commit();
    }
    // one step of the pipeline.  STEADY: 2 <= m < nly, i.e. a band row pair comes in and no line end is in reach -- the
    // flags fold away and the steady state of a strip (all but its first and last two steps) runs without selections
    auto step = [&](int m, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        const bool have_in = STEADY || m < nly;
        const bool do_b = STEADY || (m >= 1 && m - 1 < nly), do_d = STEADY || (m >= 2 && m - 2 < nly);
        const bool top0 = !STEADY && m == 0, top1 = !STEADY && m == 1;
        T cur[NC][4], dd[NC][4];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            T lo[4], hi[4];
            if (have_in) {
                cp_wait<NC - 1>();                       // component c's copies of this row pair have landed
                read_pair(c, m, lo, hi);
                if (m + 1 < nly && m + 1 <= me) issue_pair(c, m + 1);   // refill the slots: one full step to land
                cp_commit();
                if (ISO) {                                             // rows first: both band rows become interleaved samples
                    T x[4];
                    hsynth(lo, x);
#pragma unroll
                    for (int j = 0; j < 4; j++) lo[j] = x[j];
                    hsynth(hi, x);
#pragma unroll
                    for (int j = 0; j < 4; j++) hi[j] = x[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                T A = a[c][j], Hs = hs[c][j];
                if (have_in) {
                    Hs = mul_rn(hi[j], (T)K::InvK);
                    const T hl = top0 ? Hs : hs[c][j];
                    A = lift<T>(mul_rn(lo[j], (T)K::K), K::Delta, hl, Hs);
                }
                T B = b[c][j], Cn = cc[c][j];
                if (do_b) {
                    B = lift<T>(hs[c][j], K::Gamma, a[c][j], have_in ? A : a[c][j]);
                    const T bl = top1 ? B : b[c][j];
                    Cn = lift<T>(a[c][j], K::Beta, bl, B);
                }
                // D_{m-2} = B_{m-2} - alpha (C_{m-2} + C_{m-1}); below the last row pair C_{m-1} mirrors C_{m-2}
                dd[c][j] = lift<T>(b[c][j], K::Alpha, cc[c][j], do_b ? Cn : cc[c][j]);
                cur[c][j] = cc[c][j];
                hs[c][j] = Hs; a[c][j] = A; b[c][j] = B; cc[c][j] = Cn;
            }
        }
        if (do_d && m - 2 >= ka && m - 2 < kb) {
            emit_row(2 * (m - 2), cur);
            emit_row(2 * (m - 2) + 1, dd);
        }
    };
    for (int m = ms; m <= me; m++) {
        if (m >= 2 && m < nly) step(m, std::true_type());
        else step(m, std::false_type());
    }
    cp_wait<0>();
}

template <int NC, bool PIXELS, typename CT, typename T, bool ISO>
cudaError_t run_ct(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    const DevTileComp *tcs = PIXELS ? p.d_tcs : p.d_tcs + p.tc_first;
    const DevTile *tiles = p.d_tiles ? p.d_tiles + p.tile_first : nullptr;
    J2K_LAUNCH((k_idwt97_stream<NC, PIXELS, CT, T, ISO>), grid, kWarps * 32, 0, s, tcs, tiles, (const CT *)p.d_coef,
               (T *)p.d_tmp, p.d_pix, p.nlevels, p.lvl, strip_pairs, p.tail);
    return cudaGetLastError();
}

template <int NC, bool PIXELS>
cudaError_t run(const IdwtLaunch &p, dim3 grid, int strip_pairs, cudaStream_t s)
{
    if (p.iso) return run_ct<NC, PIXELS, float, float, true>(p, grid, strip_pairs, s);
    return p.coef16 ? run_ct<NC, PIXELS, int16_t, double, false>(p, grid, strip_pairs, s)
                    : run_ct<NC, PIXELS, int32_t, double, false>(p, grid, strip_pairs, s);
}

}  // namespace

// Level p.lvl of every tile-component (or, with tiles, the last level fused with the pixel epilogue) of a 9-7 job whose
// coefficient planes are int32 / int16 and whose intermediate levels are float64.  The caller has checked eligibility.
cudaError_t launch_idwt97_stream(const IdwtLaunch &p, cudaStream_t s)
{
    const int lvl = p.lvl;
    const uint32_t lw = (p.max_w + (1u << lvl) - 1) >> lvl, lh = (p.max_h + (1u << lvl) - 1) >> lvl;
    const bool pixels = (lvl == 0 && p.d_tiles != nullptr);
    const uint32_t nobj = pixels ? p.n_tiles : p.n_tc;
    if (lw < 4 || lh < 2 || nobj == 0) return cudaSuccess;
    const uint32_t nq = lw / 4, nwx = (nq + 29) / 30, nly = lh / 2;
    // a strip costs 4 extra row pairs (pipeline warm-up above, look-ahead below): keep strips tall
    int sp = 64;
    while (sp > 4 && (uint64_t)nobj * nwx * ((nly + sp - 1) / sp) < 148ull * 8 * 3) sp >>= 1;
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
