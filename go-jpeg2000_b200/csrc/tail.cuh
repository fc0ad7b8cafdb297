// tail.cuh -- device-side pixel epilogue shared by the inverse-DWT kernels: inverse MCT + DC shift
// (reference decoder.go:321-348, mct.go:43-66, 113-118) and createImage packing (decoder.go:417-588).
#pragma once
#include "common.h"

// ---- pixel epilogue: decoder.go:321-348 + createImage decoder.go:417-588 ----------------------------
__device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// clampToInt32 colorspace.go:483-491: below 0 -> 0, above max -> max, else int32(v + 0.5) (truncating)
__device__ __forceinline__ int32_t cs_clamp_round(double v, double maxv)
{
    if (v < 0.0) return 0;
    if (v > maxv) return (int32_t)maxv;
    return j2k_f64_to_i32(__dadd_rn(v, 0.5));
}

// colour conversion to sRGB (decoder.go:350-356 -> getColorConversion colorspace.go:54-88), float64 with Go's evaluation
// order and no FMA.  cs 1 / 2: the YCbCr family -- convertSYCCToRGB / convertYPbPr709ToRGB / convertEYCCToRGB
// (colorspace.go:90-114, 429-482) and convertYCbCr601ToRGB (:116-140); cs 3: convertPhotoYCCToRGB (:142-168); cs 4:
// convertCMYToRGB (:170-189); cs 5: convertCMYKToRGB (:191-217); cs 6: convertYCCKToRGB (:219-250).
// cs 7 / 8: convertCIELabToRGB / convertCIEJabToRGB (:250-292, :319-359: the same arithmetic); cs 9: convertESRGBToRGB (:363-389);
// cs 10: convertROMMRGBToRGB (:393-427) -- CUDA's pow() where Go has math.Pow, everything else in Go's order (1 LSB tolerance).
// Out of line and by value, and called only from epilogues that are out of line or rarely used themselves (put_quad_generic of
// the fused kernel, k_idwt_last_pixels<.., CC = true>, k_tail): measured, even an untaken call inlined into the epilogue of the
// register-heavy streaming kernels cost them 7 to 30 %.
// labInverseF colorspace.go:294-301 (6/29, 4/29 and 3 * (6/29)^2 are Go's exact constants rounded to float64)
__device__ __forceinline__ double cs_lab_inverse_f(double t)
{
    if (t > 6.0 / 29.0) return __dmul_rn(__dmul_rn(t, t), t);
    return __dmul_rn(3 * (6.0 / 29.0) * (6.0 / 29.0), __dsub_rn(t, 4.0 / 29.0));
}

// srgbGamma colorspace.go:303-309
__device__ __forceinline__ double cs_srgb_gamma(double lin)
{
    if (lin <= 0.0031308) return __dmul_rn(12.92, lin);
    return __dsub_rn(__dmul_rn(1.055, pow(lin, 1.0 / 2.4)), 0.055);
}

__device__ __forceinline__ double cs_clampf(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   // colorspace.go:494-502

// XYZ -> linear sRGB (colorspace.go:277-279) -> gamma -> integer; clamp01: convertROMMRGBToRGB clamps before the gamma
static __device__ J2K_NOINLINE int3 cs_xyz_to_srgb(double x, double y, double z, bool clamp01, double maxv)
{
    double rl = __dsub_rn(__dsub_rn(__dmul_rn(3.2404542, x), __dmul_rn(1.5371385, y)), __dmul_rn(0.4985314, z));
    double gl = __dadd_rn(__dadd_rn(__dmul_rn(-0.9692660, x), __dmul_rn(1.8760108, y)), __dmul_rn(0.0415560, z));
    double bl = __dadd_rn(__dsub_rn(__dmul_rn(0.0556434, x), __dmul_rn(0.2040259, y)), __dmul_rn(1.0572252, z));
    if (clamp01) { rl = cs_clampf(rl, 0.0, 1.0); gl = cs_clampf(gl, 0.0, 1.0); bl = cs_clampf(bl, 0.0, 1.0); }
    return make_int3(cs_clamp_round(__dmul_rn(cs_srgb_gamma(rl), maxv), maxv), cs_clamp_round(__dmul_rn(cs_srgb_gamma(gl), maxv), maxv),
                     cs_clamp_round(__dmul_rn(cs_srgb_gamma(bl), maxv), maxv));
}

static __device__ J2K_NOINLINE int3 tail_colour_rgb(int32_t v0, int32_t v1, int32_t v2, int32_t v3, int cconv, int ncomp, int prec)
{
    const int32_t maxi = (int32_t)((1u << prec) - 1u);
    const double maxv = (double)maxi;
    if (cconv == J2KGPU_CS_YCC709 || cconv == J2KGPU_CS_YCC601) {
        const bool bt709 = cconv == J2KGPU_CS_YCC709;
        const double half = (double)(int32_t)(1u << (prec - 1));
        const double y = (double)v0, cb = __dsub_rn((double)v1, half), cr = __dsub_rn((double)v2, half);
        const double r = __dadd_rn(y, __dmul_rn(bt709 ? 1.5748 : 1.402, cr));
        const double g = __dsub_rn(__dsub_rn(y, __dmul_rn(bt709 ? 0.1873 : 0.344136, cb)), __dmul_rn(bt709 ? 0.4681 : 0.714136, cr));
        const double b = __dadd_rn(y, __dmul_rn(bt709 ? 1.8556 : 1.772, cb));
        return make_int3(cs_clamp_round(r, maxv), cs_clamp_round(g, maxv), cs_clamp_round(b, maxv));
    }
    if (cconv == J2KGPU_CS_CIELAB || cconv == J2KGPU_CS_CIEJAB) {
        const double L = __dmul_rn(__ddiv_rn((double)v0, maxv), 100.0);
        const double a = __dsub_rn(__dmul_rn(__ddiv_rn((double)v1, maxv), 255.0), 128.0);
        const double b = __dsub_rn(__dmul_rn(__ddiv_rn((double)v2, maxv), 255.0), 128.0);
        const double fy = __ddiv_rn(__dadd_rn(L, 16.0), 116.0);
        const double fx = __dadd_rn(__ddiv_rn(a, 500.0), fy), fz = __dsub_rn(fy, __ddiv_rn(b, 200.0));
        return cs_xyz_to_srgb(__dmul_rn(0.96422, cs_lab_inverse_f(fx)), cs_lab_inverse_f(fy), __dmul_rn(0.82521, cs_lab_inverse_f(fz)), false, maxv);
    }
    if (cconv == J2KGPU_CS_ESRGB) {
        int32_t o[3];
        const int32_t in[3] = {v0, v1, v2};
#pragma unroll 1
        for (int c = 0; c < 3; c++) {
            const double v = __dsub_rn(__dmul_rn(__ddiv_rn((double)in[c], maxv), 1.25), 0.25);
            o[c] = cs_clamp_round(__dmul_rn(cs_srgb_gamma(cs_clampf(v, 0.0, 1.0)), maxv), maxv);
        }
        return make_int3(o[0], o[1], o[2]);
    }
    if (cconv == J2KGPU_CS_ROMM) {
        const double rr = pow(__ddiv_rn((double)v0, maxv), 1.8), gr = pow(__ddiv_rn((double)v1, maxv), 1.8), br = pow(__ddiv_rn((double)v2, maxv), 1.8);
        const double x = __dadd_rn(__dadd_rn(__dmul_rn(0.7977, rr), __dmul_rn(0.1352, gr)), __dmul_rn(0.0313, br));
        const double y = __dadd_rn(__dadd_rn(__dmul_rn(0.2880, rr), __dmul_rn(0.7119, gr)), __dmul_rn(0.0001, br));
        const double z = __dadd_rn(__dadd_rn(__dmul_rn(0.0, rr), __dmul_rn(0.0, gr)), __dmul_rn(0.8249, br));
        return cs_xyz_to_srgb(x, y, z, true, maxv);
    }
    if (cconv == J2KGPU_CS_CMY)
        return make_int3((int32_t)((uint32_t)maxi - (uint32_t)v0), (int32_t)((uint32_t)maxi - (uint32_t)v1), (int32_t)((uint32_t)maxi - (uint32_t)v2));
    if (cconv == J2KGPU_CS_CMYK) {
        if (ncomp < 4) return make_int3(v0, v1, v2);
        const double c = __ddiv_rn((double)v0, maxv), m = __ddiv_rn((double)v1, maxv), y = __ddiv_rn((double)v2, maxv),
                     k1 = __dsub_rn(1.0, __ddiv_rn((double)v3, maxv));
        return make_int3(cs_clamp_round(__dmul_rn(__dmul_rn(__dsub_rn(1.0, c), k1), maxv), maxv),
                         cs_clamp_round(__dmul_rn(__dmul_rn(__dsub_rn(1.0, m), k1), maxv), maxv),
                         cs_clamp_round(__dmul_rn(__dmul_rn(__dsub_rn(1.0, y), k1), maxv), maxv));
    }
    const bool ycck = cconv == J2KGPU_CS_YCCK;                          // else PhotoYCC
    if (ycck && ncomp < 4) return make_int3(v0, v1, v2);
    const double scale = __ddiv_rn(maxv, 255.0);
    const double y = __ddiv_rn((double)v0, scale), c1 = __dsub_rn(__ddiv_rn((double)v1, scale), 156.0),
                 c2 = __dsub_rn(__ddiv_rn((double)v2, scale), 156.0);
    double r = __dadd_rn(y, __dmul_rn(1.3584, c2));
    double g = __dsub_rn(__dsub_rn(y, __dmul_rn(0.4302, c1)), __dmul_rn(0.7915, c2));
    double b = __dadd_rn(y, __dmul_rn(2.2179, c1));
    r = __dmul_rn(r, scale); g = __dmul_rn(g, scale); b = __dmul_rn(b, scale);
    if (ycck) {
        const double k1 = __dsub_rn(1.0, __ddiv_rn((double)v3, maxv));
        r = __dmul_rn(r, k1); g = __dmul_rn(g, k1); b = __dmul_rn(b, k1);
    }
    return make_int3(cs_clamp_round(r, maxv), cs_clamp_round(g, maxv), cs_clamp_round(b, maxv));
}

__device__ __forceinline__ void tail_colour(int32_t v[4], const TailParams &tp)
{
    if (tp.cconv == 0) return;                                          // make_tail: 0 unless ncomp >= 3
    const int3 c = tail_colour_rgb(v[0], v[1], v[2], v[3], tp.cconv, tp.ncomp, tp.prec[0]);
    v[0] = c.x; v[1] = c.y; v[2] = c.z;
}

__device__ __forceinline__ void tail_mct_dc(int32_t v[4], const TailParams &tp)
{
    if (tp.mct) {
        if (tp.reversible) {                                   // mct.go:56-66
            uint32_t y = (uint32_t)v[0], u = (uint32_t)v[1], w = (uint32_t)v[2];
            uint32_t g = y - (uint32_t)((int32_t)(u + w) >> 2);
            v[0] = (int32_t)(w + g); v[1] = (int32_t)g; v[2] = (int32_t)(u + g);
        } else {                                               // mct.go:43-53 via decoder.go:326-340
            double y = (double)v[0], cb = (double)v[1], cr = (double)v[2];
            double r = __dadd_rn(y, __dmul_rn(1.402, cr));
            double g = __dsub_rn(__dsub_rn(y, __dmul_rn(0.34413, cb)), __dmul_rn(0.71414, cr));
            double b = __dadd_rn(y, __dmul_rn(1.772, cb));
            v[0] = j2k_f64_to_i32(__dadd_rn(r, 0.5));          // int32(v + 0.5) truncates toward zero
            v[1] = j2k_f64_to_i32(__dadd_rn(g, 0.5));
            v[2] = j2k_f64_to_i32(__dadd_rn(b, 0.5));
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (c < tp.ncomp && !tp.sgnd[c]) v[c] = (int32_t)((uint32_t)v[c] + (1u << (tp.prec[c] - 1)));   // mct.go:113-118
}

// ISO mode, irreversible path: float32 samples -> inverse ICT in float32 -> round to nearest even -> DC shift.
// Operation order and constants are OpenJPEG's (opj_mct_decode_real, opj_tcd_dc_level_shift_decode), which the test
// suite uses as the independent decoder: bit-identical output.  Clamping happens in store_pixel.
__device__ __forceinline__ void tail_iso_irrev(const float f[4], int32_t v[4], const TailParams &tp)
{
    float y = f[0], u = f[1], w = f[2];
    if (tp.mct) {
        const float r = __fadd_rn(y, __fmul_rn(w, 1.402f));
        const float g = __fsub_rn(__fsub_rn(y, __fmul_rn(u, 0.34413f)), __fmul_rn(w, 0.71414f));
        const float b = __fadd_rn(y, __fmul_rn(u, 1.772f));
        y = r; u = g; w = b;
    }
    const float o[4] = {y, u, w, f[3]};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int32_t q = __float2int_rn(o[c]);
        if (c < tp.ncomp && !tp.sgnd[c]) q = (int32_t)((uint32_t)q + (1u << (tp.prec[c] - 1)));
        v[c] = q;
    }
}

// scaled sample value exactly as createImage computes it (int32 product wraps in REF mode)
__device__ __forceinline__ uint32_t pack_value(int32_t v, int prec, int32_t maxv, bool iso)
{
    v = clampi(v, 0, maxv);
    if (prec <= 8) {
        if (prec != 8) v = (int32_t)((uint32_t)v * 255u) / maxv;
        return (uint32_t)v & 0xFF;
    }
    if (iso) return (uint32_t)(((uint64_t)(uint32_t)v * 65535u) / (uint32_t)maxv) & 0xFFFF;
    v = (int32_t)((uint32_t)v * 65535u) / maxv;
    return (uint32_t)v & 0xFFFF;
}

__device__ __forceinline__ void store_pixel(uint8_t *row, uint32_t x, const int32_t v[4], const TailParams &tp)
{
    const int prec = tp.prec[0];
    const int32_t maxv = (int32_t)((1u << prec) - 1u);
    switch (tp.fmt) {
    case J2KGPU_FMT_GRAY8:
        row[x] = (uint8_t)pack_value(v[0], prec, maxv, tp.iso);
        break;
    case J2KGPU_FMT_GRAY16: {
        uint32_t p = pack_value(v[0], prec, maxv, tp.iso);
        *(uint16_t *)(row + 2 * (size_t)x) = (uint16_t)((p >> 8) | ((p & 0xFF) << 8));      // big-endian
        break;
    }
    case J2KGPU_FMT_RGBA8: {
        uint32_t r = pack_value(v[0], prec, maxv, tp.iso), g = pack_value(v[1], prec, maxv, tp.iso),
                 b = pack_value(v[2], prec, maxv, tp.iso);
        uint32_t a = tp.ncomp == 4 ? pack_value(v[3], prec, maxv, tp.iso) : 255u;
        *(uint32_t *)(row + 4 * (size_t)x) = r | (g << 8) | (b << 16) | (a << 24);
        break;
    }
    default: {   // RGBA64, big-endian 16-bit channels
        uint32_t r = pack_value(v[0], prec, maxv, tp.iso), g = pack_value(v[1], prec, maxv, tp.iso),
                 b = pack_value(v[2], prec, maxv, tp.iso);
        uint32_t a = tp.ncomp == 4 ? pack_value(v[3], prec, maxv, tp.iso) : 65535u;
        uint32_t lo = ((r >> 8) | ((r & 0xFF) << 8)) | (((g >> 8) | ((g & 0xFF) << 8)) << 16);
        uint32_t hi = ((b >> 8) | ((b & 0xFF) << 8)) | (((a >> 8) | ((a & 0xFF) << 8)) << 16);
        *(uint2 *)(row + 8 * (size_t)x) = make_uint2(lo, hi);
        break;
    }
    }
}

