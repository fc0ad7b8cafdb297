// encode.cu -- the forward path: what encoder.encode does between extractImageData and
// createTileHeader, on the GPU, byte for byte:
//   extractImageData   encoder.go:79-214    Go image bytes -> int32 component planes (+ Options.Precision rescale)
//   preprocess         encoder.go:216-281   DC shift (mct.go:96-101), ForwardRCT / ForwardICT (mct.go:14-38, rounding
//                                           half away from zero), DecomposeMultiLevel53 / 97 (dwt.go:524-558: rows, then
//                                           columns, dense-prefix level layout), v / stepSize +- 0.5 truncated
//   encodeTile         encoder.go:597-743   the code-block list (component, resolution, band, block row, block column),
//                                           extractCodeBlockData (encoder.go:763-796: blocks are cut from the top-left
//                                           corner of the component plane whatever the band -- kept as written),
//                                           T1.SetData + T1.Encode per block (t1.go:292-304, t1_fast5.go:10-899,
//                                           MQ encoder mqc.go:185-341), results appended in list order
// Kernels: k_enc_prep (pixels -> planes), k_fwd_rows / k_fwd_cols (one lifting level: a row per CTA in shared memory;
// a column strip per thread, streaming, edges by whole-sample mirroring, which reproduces the reference's edge forms
// bit for bit), k_enc_quant, k_t1_enc (one warp per code block: the lanes cut the block out of the plane and build one
// bit-plane bitmap at a time, lane 0 runs the three coding passes and the MQ encoder), k_enc_scan + k_enc_gather
// (block bytes -> one contiguous tile).  float64 with explicit round-to-nearest multiplies and adds (no FMA), int32 wraps.
#include "common.h"
#include <algorithm>

namespace {

__constant__ uint32_t c_emq[94];          // qe (15 bits) | MPS << 15 | nmps << 16 | nlps << 24 (mqc.go:21-116)
__constant__ uint8_t  c_ezc[4 * 256];     // band, 8 neighbour bits -> zero-coding context (t1_luts.go:35-110)
__constant__ uint8_t  c_ezc9[4 * 512];    // band, 3 x 3 significance window (row above | row << 3 | row below << 6) -> the same
__constant__ uint8_t  c_esc[256];         // W, E, N, S (significant, negative) pairs -> (context - 9) << 1 | prediction (t1.go:387-460)

struct EncBlk {                           // one entry of encodeTile's job list
    uint64_t plane_off;                   // element offset of the component plane
    uint32_t sx, sy;                      // extractCodeBlockData's startX / startY
    uint16_t w, h;
    uint8_t  band, pad[3];
};

struct EncGeom {                          // options after the reference's defaulting rules
    int w, h, nc, pix_bits, prec, lossless, levels, num_res, cbw, cbh;
    double step;
};

// ---- pixels -> component planes ---------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t round_half_away(double v)       // encoder.go:236-242
{
    return v >= 0 ? j2k_f64_to_i32(__dadd_rn(v, 0.5)) : j2k_f64_to_i32(__dadd_rn(v, -0.5));
}

template <typename T>
__global__ void k_enc_prep(const uint8_t *__restrict__ pix, uint64_t stride, EncGeom g, T *__restrict__ planes)
{
    const uint64_t n = (uint64_t)g.w * g.h;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t y = (uint32_t)(i / (uint32_t)g.w), x = (uint32_t)(i - (uint64_t)y * g.w);
    const uint8_t *row = pix + (uint64_t)y * stride;
    int32_t v[4] = {0, 0, 0, 0};
    if (g.nc == 1) {
        v[0] = g.pix_bits == 8 ? row[x] : (int32_t)((row[2 * x] << 8) | row[2 * x + 1]);            // Gray / Gray16 (big-endian)
    } else {
        for (int c = 0; c < g.nc; c++)                                                              // RGBA / RGBA64 / NRGBA / NRGBA64
            v[c] = g.pix_bits == 8 ? row[4 * x + c] : (int32_t)((row[8 * x + 2 * c] << 8) | row[8 * x + 2 * c + 1]);
    }
    if (g.prec != g.pix_bits) {                                                                     // encoder.go:197-211
        const int32_t src_max = (int32_t)((1u << g.pix_bits) - 1u), dst_max = (int32_t)((1u << g.prec) - 1u);
        for (int c = 0; c < g.nc; c++) v[c] = (int32_t)((uint32_t)v[c] * (uint32_t)dst_max) / src_max;
    }
    const int32_t dc = (int32_t)(1u << (g.prec - 1));
    for (int c = 0; c < g.nc; c++) v[c] = (int32_t)((uint32_t)v[c] - (uint32_t)dc);                 // mct.go:96-101
    if (g.nc >= 3) {
        if (g.lossless) {                                                                           // mct.go:28-38
            const int32_t r = v[0], gg = v[1], b = v[2];
            v[0] = (int32_t)((uint32_t)r + 2u * (uint32_t)gg + (uint32_t)b) >> 2;
            v[1] = (int32_t)((uint32_t)b - (uint32_t)gg);
            v[2] = (int32_t)((uint32_t)r - (uint32_t)gg);
        } else {                                                                                    // mct.go:14-24
            const double r = (double)v[0], gg = (double)v[1], b = (double)v[2];
            const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.299, r), __dmul_rn(0.587, gg)), __dmul_rn(0.114, b));
            const double cb = __dadd_rn(__dadd_rn(__dmul_rn(-0.16875, r), -__dmul_rn(0.33126, gg)), __dmul_rn(0.5, b));
            const double cr = __dadd_rn(__dadd_rn(__dmul_rn(0.5, r), -__dmul_rn(0.41869, gg)), -__dmul_rn(0.08131, b));
            v[0] = round_half_away(yy); v[1] = round_half_away(cb); v[2] = round_half_away(cr);
        }
    }
    for (int c = 0; c < g.nc; c++) planes[(uint64_t)c * n + i] = (T)v[c];
}

// ---- one lifting level: rows --------------------------------------------------------------------------------------------
constexpr double kAlpha = -1.586134342059924, kBeta = -0.052980118572961, kGamma = 0.882911075530934,
                 kDelta = 0.443506852043971, kK = 1.230174104914001, kKInv = 0.812893066115961;    // dwt.go:150-157

__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// odd samples of a line of n (shared memory, in place): x[i] += c * (x[i-1] + x[i+1]); the last odd of an even-length line
// takes (2c) * x[n-2] (dwt.go:169-206)
__device__ __forceinline__ void lift_f(double *s, int n, int parity, double c)
{
    for (int i = parity + 2 * (int)threadIdx.x; i < n; i += 2 * (int)blockDim.x) {
        double nb;
        if (i == 0) nb = __dmul_rn(2 * c, s[1]);
        else if (i == n - 1) nb = __dmul_rn(2 * c, s[n - 2]);
        else nb = __dmul_rn(c, __dadd_rn(s[i - 1], s[i + 1]));
        s[i] = __dadd_rn(s[i], nb);
    }
    __syncthreads();
}

// row r of a w x h level (dense, row stride w) of plane blockIdx.y: src -> dst, L | H halves (Forward53 / Forward97 +
// deinterleave, dwt.go:73-118, 161-210, 265-284)
template <typename T>
__global__ void k_fwd_rows(const T *__restrict__ src, T *__restrict__ dst, int w, int h, uint64_t plane_elems)
{
    J2K_DYN_SMEM(T, s);
    const T *in = src + (uint64_t)blockIdx.y * plane_elems + (uint64_t)blockIdx.x * w;
    T *out = dst + (uint64_t)blockIdx.y * plane_elems + (uint64_t)blockIdx.x * w;
    const int n = w, t = (int)threadIdx.x, nt = (int)blockDim.x;
    if (n < 2) { if (t < n) out[t] = in[t]; return; }
    for (int i = t; i < n; i += nt) s[i] = in[i];
    __syncthreads();
    if constexpr (sizeof(T) == 4) {
        int32_t *d = reinterpret_cast<int32_t *>(s);
        for (int i = 1 + 2 * t; i < n; i += 2 * nt)                                   // predict, dwt.go:84-92
            d[i] = i < n - 1 ? wsub(d[i], wadd(d[i - 1], d[i + 1]) >> 1) : wsub(d[i], d[i - 1]);
        __syncthreads();
        for (int i = 2 * t; i < n; i += 2 * nt) {                                     // update, dwt.go:95-107
            const int32_t l = i ? d[i - 1] : d[1], r = i < n - 1 ? d[i + 1] : d[i - 1];
            d[i] = wadd(d[i], wadd(wadd(l, r), 2) >> 2);
        }
        __syncthreads();
    } else {
        double *d = reinterpret_cast<double *>(s);
        lift_f(d, n, 1, kAlpha);
        lift_f(d, n, 0, kBeta);
        lift_f(d, n, 1, kGamma);
        lift_f(d, n, 0, kDelta);
        for (int i = t; i < n; i += nt) d[i] = __dmul_rn(d[i], (i & 1) ? kK : kKInv);
        __syncthreads();
    }
    const int half = (n + 1) >> 1;
    for (int j = t; j < n; j += nt) out[j] = j < half ? s[2 * j] : s[2 * (j - half) + 1];
}

// ---- one lifting level: columns -----------------------------------------------------------------------------------------
__device__ __forceinline__ int mirror(int i, int n)               // whole-sample symmetric extension, n >= 2
{
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}

constexpr int kColPairs = 32;             // row pairs per thread of the column pass

// thread = one column x one strip of kColPairs row pairs of a w x h level of plane blockIdx.z: src -> dst (low rows, then
// high rows).  int32: the two steps in one sweep, carrying the previous high sample.
__global__ void k_fwd_cols53(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int w, int h, uint64_t plane_elems)
{
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= w) return;
    const int32_t *in = src + (uint64_t)blockIdx.z * plane_elems + x;
    int32_t *out = dst + (uint64_t)blockIdx.z * plane_elems + x;
    if (h < 2) { if (blockIdx.y == 0 && h == 1) out[0] = in[0]; return; }
    const int nl = (h + 1) >> 1, nh = h >> 1;
    const int a = (int)blockIdx.y * kColPairs, b = min(a + kColPairs, nl);
    auto high = [&](int k) -> int32_t {                           // H[k], 0 <= k < nh (dwt.go:84-92)
        const int32_t x0 = in[(uint64_t)(2 * k) * w], x1 = in[(uint64_t)(2 * k + 1) * w];
        return 2 * k + 2 < h ? wsub(x1, wadd(x0, in[(uint64_t)(2 * k + 2) * w]) >> 1) : wsub(x1, x0);
    };
    int32_t hprev = a > 0 ? high(a - 1) : 0;
    for (int k = a; k < b; k++) {
        const int32_t x0 = in[(uint64_t)(2 * k) * w];
        int32_t hk = hprev;                                        // odd-length line: the last even sample has one neighbour
        if (k < nh) { hk = high(k); out[(uint64_t)(nl + k) * w] = hk; }
        const int32_t l = k ? hprev : hk;
        out[(uint64_t)k * w] = wadd(x0, wadd(wadd(l, hk), 2) >> 2);
        hprev = hk;
    }
}

// float64: the four steps as a pipeline over the row pairs of the strip, two pairs of run-in; virtual samples outside the
// line are its mirror images, under which every intermediate sequence is symmetric as well, so the edge samples come out
// as the reference's (2c) * neighbour forms (c * (a + a) and (2c) * a round identically)
__global__ void k_fwd_cols97(const double *__restrict__ src, double *__restrict__ dst, int w, int h, uint64_t plane_elems)
{
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= w) return;
    const double *in = src + (uint64_t)blockIdx.z * plane_elems + x;
    double *out = dst + (uint64_t)blockIdx.z * plane_elems + x;
    if (h < 2) { if (blockIdx.y == 0 && h == 1) out[0] = in[0]; return; }
    const int nl = (h + 1) >> 1, nh = h >> 1;
    const int a = (int)blockIdx.y * kColPairs, b = min(a + kColPairs, nl);
    auto X = [&](int i) -> double { return in[(uint64_t)mirror(i, h) * w]; };
    double o1p = 0, e2p = 0, o3pp = 0;                             // o1[j-1], e2[j-1], o3[j-2]
    double xe = X(2 * (a - 2));
    for (int j = a - 2; j <= b; j++) {
        const double xo = X(2 * j + 1), xn = X(2 * j + 2);
        const double o1 = __dadd_rn(xo, __dmul_rn(kAlpha, __dadd_rn(xe, xn)));
        const double e2 = __dadd_rn(xe, __dmul_rn(kBeta, __dadd_rn(o1p, o1)));
        const double o3 = __dadd_rn(o1p, __dmul_rn(kGamma, __dadd_rn(e2p, e2)));          // o3[j-1]
        const double e4 = __dadd_rn(e2p, __dmul_rn(kDelta, __dadd_rn(o3pp, o3)));         // e4[j-1]
        const int k = j - 1;
        if (k >= a && k < b) {
            out[(uint64_t)k * w] = __dmul_rn(e4, kKInv);
            if (k < nh) out[(uint64_t)(nl + k) * w] = __dmul_rn(o3, kK);
        }
        o3pp = o3; e2p = e2; o1p = o1; xe = xn;
    }
}

// encoder.go:263-277: int32(v / stepSize + 0.5), int32(v / stepSize - 0.5)
__global__ void k_enc_quant(const double *__restrict__ in, int32_t *__restrict__ out, uint64_t n, double step)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = in[i], q = __ddiv_rn(v, step);
    out[i] = v >= 0 ? j2k_f64_to_i32(__dadd_rn(q, 0.5)) : j2k_f64_to_i32(__dadd_rn(q, -0.5));
}

// ---- tier-1 encoder ------------------------------------------------------------------------------------------------------
enum { F_SIG = 1, F_VISIT = 2, F_REFINE = 4, F_NEG = 8 };         // t1.go:72-91
enum { CX_SC0 = 9, CX_MAG0 = 14, CX_RL = 17, CX_UNI = 18, CX_N = 19 };

struct MqEnc {                            // mqc.go:169-201; the byte at bp lives in `cur` until the next one arrives
    uint32_t A, C, cur, cap;
    int CT, bp, ovf;
    uint8_t *out;                         // out[i] = the reference's buf[i + 1] (buf[0] is its dummy byte)
#ifndef J2K_ENC_CXROW
#define J2K_ENC_CXROW 1                   // a context holds its table row, not its state index: one load per decision on the chain
#endif
#if J2K_ENC_CXROW
    uint32_t *cx;                         // 19 contexts, shared memory: the table row of the context's state
#else
    uint8_t *cx;                          // 19 context states, shared memory
#endif
};
__device__ __forceinline__ void mq_put(MqEnc &e, uint32_t b)
{
    if (e.bp >= 1) {
        if ((uint32_t)(e.bp - 1) < e.cap) e.out[e.bp - 1] = (uint8_t)e.cur; else e.ovf = 1;
    }
    e.bp++;
    e.cur = b & 0xFFu;
}
__device__ __forceinline__ void mq_byte_out(MqEnc &e)                 // mqc.go:270-310
{
    if (e.cur == 0xFFu) { mq_put(e, e.C >> 20); e.C &= 0xFFFFFu; e.CT = 7; }
    else if ((e.C & 0x8000000u) == 0) { mq_put(e, e.C >> 19); e.C &= 0x7FFFFu; e.CT = 8; }
    else {
        e.cur++;
        if (e.cur == 0xFFu) { e.C &= 0x7FFFFFFu; mq_put(e, e.C >> 20); e.C &= 0xFFFFFu; e.CT = 7; }
        else { mq_put(e, e.C >> 19); e.C &= 0x7FFFFu; e.CT = 8; }
    }
}
__device__ __forceinline__ void mq_renorm(MqEnc &e)                // mqc.go:258-267
{
    do {
        e.A <<= 1; e.C <<= 1;
        if (--e.CT == 0) mq_byte_out(e);
    } while ((e.A & 0x8000u) == 0);
}
__device__ __forceinline__ void mq_encode(MqEnc &e, int cx, int d) // mqc.go:224-255
{
#if J2K_ENC_CXROW
    const uint32_t row = e.cx[cx], qe = row & 0x7FFFu, mps = (row >> 15) & 1u;
#else
    const uint32_t s = e.cx[cx], row = c_emq[s], qe = row & 0x7FFFu, mps = s & 1u;
#endif
    e.A -= qe;
    if ((uint32_t)(d & 1) == mps) {
        if ((e.A & 0x8000u) == 0) {
            if (e.A < qe) e.A = qe; else e.C += qe;
#if J2K_ENC_CXROW
            e.cx[cx] = c_emq[(row >> 16) & 0xFFu];
#else
            e.cx[cx] = (uint8_t)((row >> 16) & 0xFFu);
#endif
            mq_renorm(e);
        } else e.C += qe;
    } else {
        if (e.A < qe) e.C += qe; else e.A = qe;
#if J2K_ENC_CXROW
        e.cx[cx] = c_emq[row >> 24];
#else
        e.cx[cx] = (uint8_t)(row >> 24);
#endif
        mq_renorm(e);
    }
}
__device__ __forceinline__ int mq_flush(MqEnc &e)                                  // mqc.go:313-341 -> number of bytes
{
    const uint32_t tc = e.C + e.A;
    e.C |= 0xFFFFu;
    if (e.C >= tc) e.C -= 0x8000u;
    e.C <<= e.CT; mq_byte_out(e);
    e.C <<= e.CT; mq_byte_out(e);
    if (e.bp >= 1) {
        if ((uint32_t)(e.bp - 1) < e.cap) e.out[e.bp - 1] = (uint8_t)e.cur; else e.ovf = 1;
    }
    int end = e.bp + 1;
    if (e.cur == 0xFFu) end--;                                     // (end >= 1 always)
    return end > 1 ? end - 1 : 0;
}

struct T1Enc {
    uint8_t *f;                           // (w + 2) x (h + 2) flag bytes
    const uint64_t *bits;                 // current bit-plane, `words` 64-bit words per row
    int w, h, stride, words, band;
};
__device__ __forceinline__ int t1_bit(const T1Enc &t, int x, int y) { return (int)((t.bits[y * t.words + (x >> 6)] >> (x & 63)) & 1u); }
__device__ __forceinline__ int t1_has_nb(const T1Enc &t, int i)    // t1.go:1087-1092
{
    const uint8_t *f = t.f; const int s = t.stride;
    return ((f[i - 1] | f[i + 1] | f[i - s] | f[i + s] | f[i - s - 1] | f[i - s + 1] | f[i + s - 1] | f[i + s + 1]) & F_SIG) != 0;
}
__device__ __forceinline__ int t1_zc(const T1Enc &t, int i)        // t1.go:349-384
{
    const uint8_t *f = t.f; const int s = t.stride;
    const int p = (f[i - 1] & 1) | (f[i + 1] & 1) << 1 | (f[i - s] & 1) << 2 | (f[i + s] & 1) << 3 | (f[i - s - 1] & 1) << 4 |
                  (f[i - s + 1] & 1) << 5 | (f[i + s - 1] & 1) << 6 | (f[i + s + 1] & 1) << 7;
    return c_ezc[t.band * 256 + p];
}
__device__ __forceinline__ void t1_sign(const T1Enc &t, MqEnc &mq, int i)   // t1.go:387-460, 482-555
{
    const uint8_t *f = t.f; const int s = t.stride;
    auto contrib = [&](int j) { return (f[j] & F_SIG) ? ((f[j] & F_NEG) ? -1 : 1) : 0; };
    int hc = contrib(i - 1) + contrib(i + 1), vc = contrib(i - s) + contrib(i + s), pred = 0;
    if (hc < 0) { pred = 1; hc = -hc; }
    if (hc == 0 && vc < 0) { pred = 1; vc = -vc; }
    int cx = CX_SC0;
    if (hc == 1) cx = CX_SC0 + (vc == 1 ? 4 : (vc == 0 ? 2 : 1));
    else if (hc == 0) cx = CX_SC0 + (vc == 1 ? 1 : 0);
    else if (hc == 2) cx = CX_SC0 + 3;
    mq_encode(mq, cx, ((f[i] & F_NEG) ? 1 : 0) ^ pred);
}
__device__ __forceinline__ void t1_zc_and_sign(T1Enc &t, MqEnc &mq, int i, int sig)
{
    mq_encode(mq, t1_zc(t, i), sig);
    if (sig) { t1_sign(t, mq, i); t.f[i] |= F_SIG; }
}
// the three passes of one bit-plane (t1.go:558-770, 816-914), run by one lane
__device__ __forceinline__ void t1_plane(T1Enc &t, MqEnc &mq)
{
    const int w = t.w, h = t.h, s = t.stride;
    for (int y = 0; y < h; y++)                                    // significance propagation: raster order
        for (int x = 0, i = (y + 1) * s + 1; x < w; x++, i++) {
            if (t.f[i] & F_SIG) continue;
            if (!t1_has_nb(t, i)) continue;
            t1_zc_and_sign(t, mq, i, t1_bit(t, x, y));
            t.f[i] |= F_VISIT;
        }
    for (int y = 0; y < h; y++)                                    // magnitude refinement: raster order
        for (int x = 0, i = (y + 1) * s + 1; x < w; x++, i++) {
            const int fl = t.f[i];
            if (!(fl & F_SIG) || (fl & F_VISIT)) continue;
            const int cx = (fl & F_REFINE) ? CX_MAG0 + 2 : (t1_has_nb(t, i) ? CX_MAG0 + 1 : CX_MAG0);   // t1.go:463-479
            mq_encode(mq, cx, t1_bit(t, x, y));
            t.f[i] = (uint8_t)(fl | F_REFINE);
        }
    for (int y = 0; y < h; y += 4)                                 // cleanup: stripes of four rows, column by column
        for (int x = 0; x < w; x++) {
            bool rl = y + 4 <= h;                                  // t1.go:1195-1208
            for (int k = 0; rl && k < 4; k++) {
                const int i = (y + k + 1) * s + x + 1;
                if ((t.f[i] & (F_SIG | F_VISIT)) || t1_has_nb(t, i)) rl = false;
            }
            if (rl) {
                int first = -1;
                for (int k = 0; k < 4; k++)
                    if (t1_bit(t, x, y + k)) { first = k; break; }
                if (first < 0) { mq_encode(mq, CX_RL, 0); continue; }
                mq_encode(mq, CX_RL, 1);
                mq_encode(mq, CX_UNI, (first >> 1) & 1);
                mq_encode(mq, CX_UNI, first & 1);
                int i = (y + first + 1) * s + x + 1;
                t1_sign(t, mq, i);
                t.f[i] |= F_SIG;
                for (int k = first + 1; k < 4 && y + k < h; k++) {
                    i = (y + k + 1) * s + x + 1;
                    t1_zc_and_sign(t, mq, i, t1_bit(t, x, y + k));
                }
                continue;
            }
            for (int yy = y; yy < y + 4 && yy < h; yy++) {
                const int i = (yy + 1) * s + x + 1;
                if (t.f[i] & F_VISIT) { t.f[i] &= (uint8_t)~F_VISIT; continue; }
                if (t.f[i] & F_SIG) continue;
                t1_zc_and_sign(t, mq, i, t1_bit(t, x, yy));
            }
        }
}

// ---- the coder's state as 64-bit masks per row -------------------------------------------------------------------------------
// sig / neg / vis / ref hold rows -1 .. h (index y + 1; the two border rows stay zero), `ws` words per row (M = false: blocks
// at most 64 samples wide, one word, every word loop and carry below folds away; M = true: up to four words).  Instead of
// visiting every sample, each pass computes the set of samples that take part (candidates of a row word, or of a stripe's
// columns) with a few logic operations and walks its set bits in coding order; a sample that becomes significant adds its
// right-hand neighbour to the candidates (SPP) or takes the next column out of run-length mode (cleanup), which is all
// that the reference's visit-time tests (t1.go:349-384, 1087-1092, 1195-1208) can see of it.  Words are taken left to right
// and a word's sets are computed when the walk reaches it, so what happened in the word before is already in the masks.
struct T1Mask {
    uint64_t *sig, *neg, *vis, *ref;
    const uint64_t *bits;                 // current bit-plane, nw words per row
    uint64_t lastmask;                    // valid columns of the last word
    int w, h, band, ws, nw;               // ws = words per row of the state arrays (launch-wide), nw = words this block uses
};
template <bool M> __device__ __forceinline__ int mk_ws(const T1Mask &t) { return M ? t.ws : 1; }
template <bool M> __device__ __forceinline__ int mk_nw(const T1Mask &t) { return M ? t.nw : 1; }
template <bool M> __device__ __forceinline__ uint64_t mk_mask(const T1Mask &t, int k) { return (M && k + 1 < t.nw) ? ~(uint64_t)0 : t.lastmask; }
template <bool M> __device__ __forceinline__ uint64_t mk_shl(const T1Mask &t, const uint64_t *a, int r, int k)   // column x <- x - 1
{
    const uint64_t *p = a + r * mk_ws<M>(t) + k;
    return (p[0] << 1) | ((M && k > 0) ? p[-1] >> 63 : 0);
}
template <bool M> __device__ __forceinline__ uint64_t mk_shr(const T1Mask &t, const uint64_t *a, int r, int k)   // column x <- x + 1
{
    const uint64_t *p = a + r * mk_ws<M>(t) + k;
    return (p[0] >> 1) | ((M && k + 1 < t.nw) ? p[1] << 63 : 0);
}
template <bool M> __device__ __forceinline__ uint64_t mk_spread(const T1Mask &t, const uint64_t *a, int r, int k)
{
    return a[r * mk_ws<M>(t) + k] | mk_shl<M>(t, a, r, k) | mk_shr<M>(t, a, r, k);
}
template <bool M> __device__ __forceinline__ uint64_t mk_nb(const T1Mask &t, int y, int k)      // samples of row y, word k, with a significant neighbour
{
    return (mk_spread<M>(t, t.sig, y, k) | mk_spread<M>(t, t.sig, y + 2, k) | mk_shl<M>(t, t.sig, y + 1, k) | mk_shr<M>(t, t.sig, y + 1, k)) &
           mk_mask<M>(t, k);
}
// columns x - 1, x, x + 1 of row r (x = 64 k + b) as bits 0..2
template <bool M> __device__ __forceinline__ uint32_t mk_win(const T1Mask &t, const uint64_t *a, int r, int k, int b)
{
    const uint64_t *p = a + r * mk_ws<M>(t) + k;
    uint32_t u = (uint32_t)(b ? p[0] >> (b - 1) : p[0] << 1) & 7u;
    if (M) {
        if (b == 0 && k > 0) u |= (uint32_t)(p[-1] >> 63);
        if (b == 63 && k + 1 < t.nw) u |= (uint32_t)(p[1] & 1u) << 2;
    }
    return u;
}
template <bool M> __device__ __forceinline__ uint32_t mk_bit(const T1Mask &t, const uint64_t *a, int r, int k, int b)
{
    return (uint32_t)(a[r * mk_ws<M>(t) + k] >> b) & 1u;
}
template <bool M> __device__ __forceinline__ void mk_sign(const T1Mask &t, MqEnc &mq, int y, int k, int b)
{
    const uint32_t sm = mk_win<M>(t, t.sig, y + 1, k, b), nm = mk_win<M>(t, t.neg, y + 1, k, b);
    const uint32_t idx = (sm & 1u) | (nm & 1u) << 1 | (sm & 4u) | (nm & 4u) << 1 |
                         mk_bit<M>(t, t.sig, y, k, b) << 4 | mk_bit<M>(t, t.neg, y, k, b) << 5 |
                         mk_bit<M>(t, t.sig, y + 2, k, b) << 6 | mk_bit<M>(t, t.neg, y + 2, k, b) << 7;
    const uint32_t e = c_esc[idx];
    mq_encode(mq, CX_SC0 + (int)(e >> 1), (int)(mk_bit<M>(t, t.neg, y + 1, k, b) ^ (e & 1u)));
}
// zero coding of sample (64 k + b, y) with bit `sig`; returns sig after coding the sign and marking the sample significant
template <bool M> __device__ __forceinline__ int mk_zc_and_sign(const T1Mask &t, MqEnc &mq, int y, int k, int b, int sig)
{
    const uint32_t idx = mk_win<M>(t, t.sig, y, k, b) | mk_win<M>(t, t.sig, y + 1, k, b) << 3 | mk_win<M>(t, t.sig, y + 2, k, b) << 6;
    mq_encode(mq, c_ezc9[t.band * 512 + idx], sig);
    if (sig) { mk_sign<M>(t, mq, y, k, b); t.sig[(y + 1) * mk_ws<M>(t) + k] |= (uint64_t)1 << b; }
    return sig;
}
template <bool M>
__device__ __forceinline__ void t1_plane_masks(T1Mask &t, MqEnc &mq)
{
    const int h = t.h, ws = mk_ws<M>(t), nw = mk_nw<M>(t);
    for (int y = 0; y < h; y++)                                    // significance propagation, t1.go:558-639
        for (int k = 0; k < nw; k++) {
            uint64_t *srow = t.sig + (y + 1) * ws + k;
            const uint64_t mask = mk_mask<M>(t, k), plane = t.bits[y * nw + k];
            uint64_t cand = ~*srow & mk_nb<M>(t, y, k), seen = 0;
            while (cand) {
                const int b = __ffsll((long long)cand) - 1;
                const uint64_t bit = (uint64_t)1 << b;
                cand &= ~bit; seen |= bit;
                if (mk_zc_and_sign<M>(t, mq, y, k, b, (int)((plane >> b) & 1u)))
                    cand |= (bit << 1) & ~*srow & mask;            // its right-hand neighbour has a significant neighbour now
            }
            t.vis[(y + 1) * ws + k] |= seen;
        }
    for (int y = 0; y < h; y++)                                    // magnitude refinement, t1.go:642-683
        for (int k = 0; k < nw; k++) {
            const int i = (y + 1) * ws + k;
            const uint64_t cand = t.sig[i] & ~t.vis[i];
            if (!cand) continue;
            const uint64_t plane = t.bits[y * nw + k], nb = mk_nb<M>(t, y, k), ref = t.ref[i];
            t.ref[i] = ref | cand;
            // a word as two 32-bit halves: find-first-set, clear-lowest and variable shifts of 64-bit words cost twice the instructions
            auto half = [&](uint32_t c, uint32_t refh, uint32_t nbh, uint32_t pl) {
                while (c) {
                    const int x = __ffs((int)c) - 1;
                    c &= c - 1;
                    const int cx = ((refh >> x) & 1u) ? CX_MAG0 + 2 : CX_MAG0 + (int)((nbh >> x) & 1u);   // t1.go:463-479
                    mq_encode(mq, cx, (int)((pl >> x) & 1u));
                }
            };
            half((uint32_t)cand, (uint32_t)ref, (uint32_t)nb, (uint32_t)plane);
            half((uint32_t)(cand >> 32), (uint32_t)(ref >> 32), (uint32_t)(nb >> 32), (uint32_t)(plane >> 32));
        }
    for (int y = 0; y < h; y += 4) {                               // cleanup, t1.go:686-770, 816-914
        const int rows = min(4, h - y);
        for (int k = 0; k < nw; k++) {
            const uint64_t mask = mk_mask<M>(t, k);
            uint64_t todo = 0, busy = 0;
            for (int j = 0; j < rows; j++) {
                const uint64_t sv = t.sig[(y + j + 1) * ws + k] | t.vis[(y + j + 1) * ws + k];
                todo |= ~sv & mask;
                busy |= sv | mk_nb<M>(t, y + j, k);
            }
            uint64_t rl = rows == 4 ? ~busy & mask : 0;             // columns in run-length mode (t1.go:1195-1208)
            while (todo) {
                const int b = __ffsll((long long)todo) - 1;
                const uint64_t bit = (uint64_t)1 << b;
                todo &= ~bit;
                int j = 0, grew = 0;
                if (rl & bit) {
                    const uint32_t col = (uint32_t)((t.bits[y * nw + k] >> b) & 1u) | (uint32_t)((t.bits[(y + 1) * nw + k] >> b) & 1u) << 1 |
                                         (uint32_t)((t.bits[(y + 2) * nw + k] >> b) & 1u) << 2 | (uint32_t)((t.bits[(y + 3) * nw + k] >> b) & 1u) << 3;
                    if (!col) { mq_encode(mq, CX_RL, 0); continue; }
                    const int first = __ffs((int)col) - 1;
                    mq_encode(mq, CX_RL, 1);
                    mq_encode(mq, CX_UNI, (first >> 1) & 1);
                    mq_encode(mq, CX_UNI, first & 1);
                    mk_sign<M>(t, mq, y + first, k, b);
                    t.sig[(y + first + 1) * ws + k] |= bit;
                    grew = 1;
                    j = first + 1;
                }
                for (; j < rows; j++) {
                    if ((t.sig[(y + j + 1) * ws + k] | t.vis[(y + j + 1) * ws + k]) & bit) continue;
                    grew |= mk_zc_and_sign<M>(t, mq, y + j, k, b, (int)((t.bits[(y + j) * nw + k] >> b) & 1u));
                }
                if (grew) rl &= ~(bit << 1);
            }
        }
        for (int j = 0; j < rows; j++)
            for (int k = 0; k < nw; k++) t.vis[(y + j + 1) * ws + k] = 0;
    }
}

// shared memory of one warp: flag bytes (or row masks), the bit-plane bitmap, the context states
__host__ __device__ constexpr size_t t1enc_flag_bytes(int cbw, int cbh, bool masks)
{
    return masks ? (size_t)4 * (cbh + 2) * ((cbw + 63) / 64) * 8 : (((size_t)(cbw + 2) * (cbh + 2) + 15) & ~(size_t)15);
}
__host__ __device__ constexpr size_t t1enc_warp_bytes(int cbw, int cbh, bool masks)
{
    return t1enc_flag_bytes(cbw, cbh, masks) + (size_t)cbh * ((cbw + 63) / 64) * 8 + 96;
}

// longest chains first: a warp per block finds the block's bit-plane count (the length of its chain, to first order), one CTA
// then lists the blocks by falling count; k_t1_enc takes them in that order, so that the tail of the launch is made of the
// short chains (the reference's own pool hands its jobs out in list order; the bytes of a block do not depend on when it runs)
__global__ void k_enc_bps(const EncBlk *__restrict__ blks, uint32_t n, const int32_t *__restrict__ planes, int W, int H,
                          uint8_t *__restrict__ bps)
{
    const uint32_t bi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = (int)threadIdx.x & 31;
    if (bi >= n) return;
    const EncBlk blk = blks[bi];
    const int32_t *plane = planes + blk.plane_off;
    int32_t maxv = 0;
    for (int y = 0; y < blk.h; y++) {
        const uint32_t gy = blk.sy + (uint32_t)y;
        if (gy >= (uint32_t)H) break;
        for (int x = lane; x < blk.w; x += 32) {
            const uint32_t gx = blk.sx + (uint32_t)x;
            int32_t v = gx < (uint32_t)W ? plane[(uint64_t)gy * W + gx] : 0;
            v = v < 0 ? (int32_t)(0u - (uint32_t)v) : v;
            maxv = v > maxv ? v : maxv;
        }
    }
    for (int o = 16; o; o >>= 1) { const int32_t m = __shfl_xor_sync(0xffffffffu, maxv, o); maxv = m > maxv ? m : maxv; }
    int nbps = 0;
    for (int32_t m = maxv; m > 0; m >>= 1) nbps++;
    if (lane == 0) bps[bi] = (uint8_t)nbps;
}

__global__ void k_enc_order(const uint8_t *__restrict__ bps, uint32_t n, uint32_t *__restrict__ order)
{
    __shared__ uint32_t hist[33], cursor[33];
    const uint32_t t = threadIdx.x;
    if (t < 33) hist[t] = 0;
    __syncthreads();
    for (uint32_t i = t; i < n; i += blockDim.x) atomicAdd(&hist[min((uint32_t)bps[i], 32u)], 1u);
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int b = 32; b >= 0; b--) { cursor[b] = run; run += hist[b]; }
    }
    __syncthreads();
    for (uint32_t i = t; i < n; i += blockDim.x) order[atomicAdd(&cursor[min((uint32_t)bps[i], 32u)], 1u)] = i;
}

// MODE 0: flag bytes (A/B and tests: option enc_bytes); 1: row masks, blocks at most 64 wide; 2: row masks, several words per row
template <int MODE>
__global__ void k_t1_enc(const EncBlk *__restrict__ blks, const uint32_t *__restrict__ order, uint32_t n,
                         const int32_t *__restrict__ planes, int W, int H,
                         int cbw, int cbh, uint8_t *__restrict__ slab, uint32_t cap, uint32_t *__restrict__ lens,
                         uint8_t *__restrict__ bps, int *__restrict__ err)
{
    constexpr bool MASKS = MODE != 0;
    J2K_DYN_SMEM(uint8_t, smem);
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (uint32_t)warp;
    if (slot >= n) return;
    const uint32_t bi = order[slot];
    const EncBlk blk = blks[bi];
    uint8_t *base = smem + (size_t)warp * t1enc_warp_bytes(cbw, cbh, MASKS);
    uint8_t *flags = base;                                         // MODE 0: flag bytes; else four arrays of row masks
    uint64_t *rows = reinterpret_cast<uint64_t *>(base);
    uint64_t *bits = reinterpret_cast<uint64_t *>(base + t1enc_flag_bytes(cbw, cbh, MASKS));
    const int w = blk.w, h = blk.h, stride = w + 2, words = (w + 63) >> 6;
    const int ws = MODE == 1 ? 1 : (cbw + 63) >> 6, arr = (cbh + 2) * ws;   // words per row / per array of the mask state
#if J2K_ENC_CXROW
    uint32_t *cx = reinterpret_cast<uint32_t *>(bits + (size_t)cbh * ((cbw + 63) / 64));
#else
    uint8_t *cx = reinterpret_cast<uint8_t *>(bits + (size_t)cbh * ((cbw + 63) / 64));
#endif
    const int32_t *plane = planes + blk.plane_off;
    // extractCodeBlockData + SetData: sign flags, largest magnitude
    if (MASKS) { for (int i = lane; i < 4 * arr; i += 32) rows[i] = 0; }
    else { for (int i = lane; i < stride * (h + 2); i += 32) flags[i] = 0; }
#if J2K_ENC_CXROW
    if (lane < CX_N) cx[lane] = c_emq[lane == CX_UNI ? 92 : 0];    // mqc.go:194-199
#else
    if (lane < CX_N) cx[lane] = lane == CX_UNI ? 92 : 0;           // mqc.go:194-199
#endif
    __syncwarp();
    auto sample = [&](int x, int y) -> int32_t {
        const uint32_t gx = blk.sx + (uint32_t)x, gy = blk.sy + (uint32_t)y;
        return (gx < (uint32_t)W && gy < (uint32_t)H) ? plane[(uint64_t)gy * W + gx] : 0;
    };
    int32_t maxv = 0;
    if (MASKS) {
        uint64_t *neg = rows + arr;
        for (int y = 0; y < h; y++)
            for (int wd = 0; wd < words; wd++) {
                const int x0 = wd * 64 + lane, x1 = x0 + 32;
                int32_t v0 = x0 < w ? sample(x0, y) : 0, v1 = x1 < w ? sample(x1, y) : 0;
                const uint32_t lo = __ballot_sync(0xffffffffu, v0 < 0), hi = __ballot_sync(0xffffffffu, v1 < 0);
                if (lane == 0) neg[(y + 1) * ws + wd] = (uint64_t)lo | ((uint64_t)hi << 32);
                v0 = v0 < 0 ? (int32_t)(0u - (uint32_t)v0) : v0;
                v1 = v1 < 0 ? (int32_t)(0u - (uint32_t)v1) : v1;
                maxv = v0 > maxv ? v0 : maxv;
                maxv = v1 > maxv ? v1 : maxv;
            }
    } else {
        for (int y = 0; y < h; y++)
            for (int x = lane; x < w; x += 32) {
                int32_t v = sample(x, y);
                if (v < 0) { v = (int32_t)(0u - (uint32_t)v); flags[(y + 1) * stride + x + 1] = F_NEG; }
                maxv = v > maxv ? v : maxv;
            }
    }
    for (int o = 16; o; o >>= 1) { const int32_t m = __shfl_xor_sync(0xffffffffu, maxv, o); maxv = m > maxv ? m : maxv; }
    int nbps = 0;
    for (int32_t m = maxv; m > 0; m >>= 1) nbps++;                 // t1_fast5.go:23-27
    if (maxv == 0) {                                               // t1_fast5.go:20-22: nil
        if (lane == 0) { lens[bi] = 0; bps[bi] = 0; }
        return;
    }
    MqEnc mq;
    mq.A = 0x8000u; mq.C = 0; mq.CT = 12; mq.bp = 0; mq.cur = 0; mq.ovf = 0;
    mq.out = slab + (uint64_t)bi * cap; mq.cap = cap; mq.cx = cx;
    T1Enc t{flags, bits, w, h, stride, words, (int)blk.band};
    const int wl = w - 64 * (words - 1);                           // columns of the last word
    T1Mask tm{rows, rows + arr, rows + 2 * arr, rows + 3 * arr, bits, wl >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << wl) - 1),
              w, h, (int)blk.band, ws, words};
    for (int bp = nbps - 1; bp >= 0; bp--) {
        __syncwarp();
        for (int y = 0; y < h; y++)
            for (int wd = 0; wd < words; wd++) {
                const int x0 = wd * 64 + lane, x1 = x0 + 32;
                int32_t v0 = x0 < w ? sample(x0, y) : 0, v1 = x1 < w ? sample(x1, y) : 0;
                v0 = v0 < 0 ? (int32_t)(0u - (uint32_t)v0) : v0;
                v1 = v1 < 0 ? (int32_t)(0u - (uint32_t)v1) : v1;
                const uint32_t lo = __ballot_sync(0xffffffffu, ((uint32_t)v0 >> bp) & 1u);
                const uint32_t hi = __ballot_sync(0xffffffffu, ((uint32_t)v1 >> bp) & 1u);
                if (lane == 0) bits[y * words + wd] = (uint64_t)lo | ((uint64_t)hi << 32);
            }
        __syncwarp();
        if (lane == 0) {
            if (MODE == 1) t1_plane_masks<false>(tm, mq);
            else if (MODE == 2) t1_plane_masks<true>(tm, mq);
            else t1_plane(t, mq);
        }
    }
    if (lane == 0) {
        const int len = mq_flush(mq);                              // t1_fast5.go:878-898
        lens[bi] = (uint32_t)len;
        bps[bi] = (uint8_t)nbps;
        if (mq.ovf) atomicOr(err, 1);
    }
}

// ---- block bytes -> one contiguous tile ---------------------------------------------------------------------------------
__global__ void k_enc_scan(const uint32_t *__restrict__ lens, uint32_t n, uint64_t *__restrict__ offs)
{
    __shared__ uint64_t part[1024];
    const uint32_t t = threadIdx.x, seg = (n + 1023u) / 1024u;
    const uint32_t a = min(t * seg, n), b = min(a + seg, n);
    uint64_t sum = 0;
    for (uint32_t i = a; i < b; i++) sum += lens[i];
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        const uint64_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint64_t run = part[t] - sum;
    for (uint32_t i = a; i < b; i++) { offs[i] = run; run += lens[i]; }
    if (t == 1023) offs[n] = part[1023];
}

__global__ void k_enc_gather(const uint8_t *__restrict__ slab, uint32_t cap, const uint32_t *__restrict__ lens,
                             const uint64_t *__restrict__ offs, uint32_t n, uint8_t *__restrict__ out)
{
    const uint32_t bi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (bi >= n) return;
    const uint8_t *src = slab + (uint64_t)bi * cap;
    uint8_t *dst = out + offs[bi];
    for (uint32_t i = lane; i < lens[bi]; i += 32) dst[i] = src[i];
}

// ---- host side -------------------------------------------------------------------------------------------------------------
cudaError_t upload_enc_tables(cudaStream_t s)
{
    static bool done[64] = {};
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (done[dev]) return cudaSuccess;
    // ISO/IEC 15444-1 Table C.2 (Qe, NMPS, NLPS, SWITCH) expanded to the reference's 94 (state, MPS) rows (mqc.go:21-116)
    static const uint16_t qe[47] = {0x5601, 0x3401, 0x1801, 0x0AC1, 0x0521, 0x0221, 0x5601, 0x5401, 0x4801, 0x3801, 0x3001, 0x2401,
                                    0x1C01, 0x1601, 0x5601, 0x5401, 0x5101, 0x4801, 0x3801, 0x3401, 0x3001, 0x2801, 0x2401, 0x2201,
                                    0x1C01, 0x1801, 0x1601, 0x1401, 0x1201, 0x1101, 0x0AC1, 0x09C1, 0x08A1, 0x0521, 0x0441, 0x02A1,
                                    0x0221, 0x0141, 0x0111, 0x0085, 0x0049, 0x0025, 0x0015, 0x0009, 0x0005, 0x0001, 0x5601};
    static const uint8_t nmps[47] = {1, 2, 3, 4, 5, 38, 7, 8, 9, 10, 11, 12, 13, 29, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
                                     25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 45, 46};
    static const uint8_t nlps[47] = {1, 6, 9, 12, 29, 33, 6, 14, 14, 14, 17, 18, 20, 21, 14, 14, 15, 16, 17, 18, 19, 19, 20, 21,
                                     22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 46};
    static const uint8_t sw[47] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1};
    uint32_t mq[94];
    for (int i = 0; i < 47; i++)
        for (int m = 0; m < 2; m++)
            mq[2 * i + m] = qe[i] | (uint32_t)m << 15 | (uint32_t)(2 * nmps[i] + m) << 16 | (uint32_t)(2 * nlps[i] + (m ^ sw[i])) << 24;
    uint8_t zc[4 * 256];
    for (int band = 0; band < 4; band++)
        for (int p = 0; p < 256; p++) {                            // t1_luts.go:35-110
            int hc = (p & 1) + ((p >> 1) & 1), vc = ((p >> 2) & 1) + ((p >> 3) & 1);
            const int dc = ((p >> 4) & 1) + ((p >> 5) & 1) + ((p >> 6) & 1) + ((p >> 7) & 1);
            int cx;
            if (band == J2KGPU_BAND_HH) {
                const int hv = hc + vc;
                if (hv >= 3) cx = 8;
                else if (hv == 2) cx = dc >= 2 ? 7 : (dc >= 1 ? 6 : 5);
                else if (hv == 1) cx = dc >= 2 ? 4 : 3;
                else cx = dc >= 2 ? 2 : (dc >= 1 ? 1 : 0);
            } else {
                if (band == J2KGPU_BAND_HL) std::swap(hc, vc);
                if (hc == 2) cx = 8;
                else if (hc == 1) cx = vc >= 1 ? 7 : (dc >= 1 ? 6 : 5);
                else if (vc == 2) cx = 4;
                else if (vc == 1) cx = dc >= 1 ? 3 : 2;
                else cx = dc >= 2 ? 1 : 0;
            }
            zc[band * 256 + p] = (uint8_t)cx;
        }
    uint8_t zc9[4 * 512], sc[256];
    for (int band = 0; band < 4; band++)
        for (int i = 0; i < 512; i++) {                            // window bits: (row above, row, row below) x (x - 1, x, x + 1)
            const int p = ((i >> 3) & 1) | ((i >> 5) & 1) << 1 | ((i >> 1) & 1) << 2 | ((i >> 7) & 1) << 3 | (i & 1) << 4 |
                          ((i >> 2) & 1) << 5 | ((i >> 6) & 1) << 6 | ((i >> 8) & 1) << 7;
            zc9[band * 512 + i] = zc[band * 256 + p];
        }
    for (int i = 0; i < 256; i++) {                                // t1.go:387-460
        int hc = 0, vc = 0, pred = 0, cx = 0;
        if (i & 1) hc += (i & 2) ? -1 : 1;
        if (i & 4) hc += (i & 8) ? -1 : 1;
        if (i & 16) vc += (i & 32) ? -1 : 1;
        if (i & 64) vc += (i & 128) ? -1 : 1;
        if (hc < 0) { pred = 1; hc = -hc; }
        if (hc == 0 && vc < 0) { pred = 1; vc = -vc; }
        if (hc == 1) cx = vc == 1 ? 4 : (vc == 0 ? 2 : 1);
        else if (hc == 0) cx = vc == 1 ? 1 : 0;
        else if (hc == 2) cx = 3;
        sc[i] = (uint8_t)(cx << 1 | pred);
    }
    if ((e = cudaMemcpyToSymbolAsync(c_ezc9, zc9, sizeof zc9, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_esc, sc, sizeof sc, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_emq, mq, sizeof mq, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbolAsync(c_ezc, zc, sizeof zc, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    done[dev] = true;
    return cudaSuccess;
}

int enc_geometry(j2kgpu_ctx *ctx, const j2k_encode_t *p, EncGeom &g)
{
    if (p->width == 0 || p->height == 0) return j2k_set_err(ctx, J2KGPU_E_ARG, "empty image");
    if (p->width > 16384 || p->height > 16384) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "image side above 16384");
    if (p->ncomp != 1 && p->ncomp != 3 && p->ncomp != 4) return j2k_set_err(ctx, J2KGPU_E_ARG, "ncomp must be 1, 3 or 4");
    if (p->pix_bits != 8 && p->pix_bits != 16) return j2k_set_err(ctx, J2KGPU_E_ARG, "pix_bits must be 8 or 16");
    if (p->cb_x > 6 || p->cb_y > 6) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "code blocks above 256 x 256");
    g.w = (int)p->width; g.h = (int)p->height; g.nc = p->ncomp; g.pix_bits = p->pix_bits;
    g.prec = (p->precision > 0 && p->precision <= 16 && p->precision != p->pix_bits) ? p->precision : p->pix_bits;   // encoder.go:197
    g.lossless = p->lossless ? 1 : 0;
    g.levels = (int)p->num_resolutions - 1;                        // encoder.go:249-252
    if (g.levels <= 0) g.levels = 5;
    g.num_res = p->num_resolutions ? p->num_resolutions : 6;       // encoder.go:601-604
    g.cbw = 1 << (p->cb_x + 2); g.cbh = 1 << (p->cb_y + 2);        // encoder.go:606-607
    const int q = p->quality > 0 ? p->quality : 100;               // encoder.go:265-268
    g.step = 1.0 / (double)q;
    return J2KGPU_OK;
}

// encodeTile's job list, encoder.go:615-673
void enc_blocks(const EncGeom &g, std::vector<EncBlk> *out, uint32_t *count)
{
    uint32_t n = 0;
    for (int c = 0; c < g.nc; c++)
        for (int r = 0; r < g.num_res; r++) {
            const int nb = r == 0 ? 1 : 3;
            for (int b = 0; b < nb; b++) {
                const int band = r == 0 ? J2KGPU_BAND_LL : (b == 0 ? J2KGPU_BAND_HL : (b == 1 ? J2KGPU_BAND_LH : J2KGPU_BAND_HH));
                const int64_t scale = (int64_t)1 << (g.num_res - 1 - r);
                int bw = (int)((g.w + scale - 1) / scale), bh = (int)((g.h + scale - 1) / scale);
                if (r > 0) { bw = (bw + 1) / 2; bh = (bh + 1) / 2; }
                for (int cby = 0; cby * g.cbh < bh; cby++)
                    for (int cbx = 0; cbx * g.cbw < bw; cbx++) {
                        if (out) {
                            EncBlk e{};
                            e.plane_off = (uint64_t)c * g.w * g.h;
                            e.sx = (uint32_t)(cbx * g.cbw); e.sy = (uint32_t)(cby * g.cbh);
                            e.w = (uint16_t)std::min(g.cbw, bw - cbx * g.cbw);
                            e.h = (uint16_t)std::min(g.cbh, bh - cby * g.cbh);
                            e.band = (uint8_t)band;
                            out->push_back(e);
                        }
                        n++;
                    }
            }
        }
    if (count) *count = n;
}

struct PoolPtrs {                         // device blocks taken from the ctx pool, returned on every exit path
    j2kgpu_ctx *ctx;
    std::vector<void *> p;
    ~PoolPtrs() { for (void *q : p) j2k_pool_free(ctx, q); }
    void *get(size_t bytes, cudaError_t *e)
    {
        void *q = j2k_pool_alloc(ctx, bytes, e);
        if (q) p.push_back(q);
        return q;
    }
};

// pixels -> int32 component planes after preprocess, on the ctx stream; *d_planes is owned by `mem`
template <typename T>
int enc_transform(j2kgpu_ctx *ctx, const EncGeom &g, const uint8_t *d_pix, uint64_t stride, PoolPtrs &mem, int32_t **d_planes)
{
    cudaStream_t s = ctx->stream;
    const uint64_t n = (uint64_t)g.w * g.h, total = n * g.nc;
    cudaError_t e = cudaSuccess;
    T *a = (T *)mem.get(total * sizeof(T), &e);
    T *b = e == cudaSuccess ? (T *)mem.get(total * sizeof(T), &e) : nullptr;
    if (e != cudaSuccess) return j2k_set_err(ctx, J2KGPU_E_NOMEM, "forward transform planes: %s", cudaGetErrorString(e));
    J2K_LAUNCH((k_enc_prep<T>), (unsigned)((n + 255) / 256), 256, 0, s, d_pix, stride, g, a);
    ctx->launches++;
    if ((size_t)g.w * sizeof(T) > 48 * 1024)                      // (a per-device attribute: set on every call that needs it)
        J2K_CUDA(ctx, cudaFuncSetAttribute((const void *)k_fwd_rows<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(T)));
    int w = g.w, h = g.h;
    for (int l = 0; l < g.levels; l++) {                           // dwt.go:524-531, 551-558: every level works on the dense prefix
        J2K_LAUNCH((k_fwd_rows<T>), dim3((unsigned)h, (unsigned)g.nc), 256, (size_t)w * sizeof(T), s, (const T *)a, b, w, h, n);
        const dim3 grid((unsigned)((w + 127) / 128), (unsigned)((((h + 1) >> 1) + kColPairs - 1) / kColPairs), (unsigned)g.nc);
        if constexpr (sizeof(T) == 4) J2K_LAUNCH((k_fwd_cols53), grid, 128, 0, s, (const int32_t *)b, (int32_t *)a, w, h, n);
        else J2K_LAUNCH((k_fwd_cols97), grid, 128, 0, s, (const double *)b, (double *)a, w, h, n);
        ctx->launches += 2;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    if constexpr (sizeof(T) == 8) {
        J2K_LAUNCH((k_enc_quant), (unsigned)((total + 255) / 256), 256, 0, s, (const double *)a, (int32_t *)b, total, g.step);
        ctx->launches++;
        *d_planes = (int32_t *)b;
    } else *d_planes = (int32_t *)a;
    J2K_CUDA(ctx, cudaGetLastError());
    return J2KGPU_OK;
}

int enc_pixels_in(j2kgpu_ctx *ctx, const j2k_encode_t *p, const EncGeom &g, const uint8_t *pix, uint64_t stride, PoolPtrs &mem,
                  const uint8_t **d_pix)
{
    const uint64_t bpp = (uint64_t)(g.nc == 1 ? 1 : 4) * (g.pix_bits / 8);
    if (!pix) return j2k_set_err(ctx, J2KGPU_E_ARG, "null pixels");
    if (stride < (uint64_t)g.w * bpp) return j2k_set_err(ctx, J2KGPU_E_ARG, "pixel stride below the row size");
    if (p->flags & J2KGPU_ENC_DEVICE_PTRS) { *d_pix = pix; return J2KGPU_OK; }
    cudaError_t e = cudaSuccess;
    const uint64_t bytes = stride * (uint64_t)(g.h - 1) + (uint64_t)g.w * bpp;
    uint8_t *d = (uint8_t *)mem.get(bytes, &e);
    if (e != cudaSuccess) return j2k_set_err(ctx, J2KGPU_E_NOMEM, "pixel buffer: %s", cudaGetErrorString(e));
    J2K_CUDA(ctx, cudaMemcpyAsync(d, pix, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d_pix = d;
    return J2KGPU_OK;
}

}  // namespace

extern "C" uint32_t j2kgpu_encode_block_count(const j2k_encode_t *p)
{
    if (!p || p->width == 0 || p->height == 0 || p->cb_x > 6 || p->cb_y > 6 || (p->ncomp != 1 && p->ncomp != 3 && p->ncomp != 4)) return 0;
    EncGeom g{};
    g.w = (int)p->width; g.h = (int)p->height; g.nc = p->ncomp;
    g.num_res = p->num_resolutions ? p->num_resolutions : 6;
    g.cbw = 1 << (p->cb_x + 2); g.cbh = 1 << (p->cb_y + 2);
    uint32_t n = 0;
    enc_blocks(g, nullptr, &n);
    return n;
}

extern "C" int j2kgpu_encode_preprocess(j2kgpu_ctx *ctx, const j2k_encode_t *p, const uint8_t *pix, uint64_t pix_stride, int32_t *planes)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!p || !planes) return j2k_set_err(ctx, J2KGPU_E_ARG, "null argument");
    EncGeom g{};
    int rc = enc_geometry(ctx, p, g);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    PoolPtrs mem{ctx, {}};
    const uint8_t *d_pix = nullptr;
    if ((rc = enc_pixels_in(ctx, p, g, pix, pix_stride, mem, &d_pix))) return rc;
    int32_t *d_planes = nullptr;
    rc = g.lossless ? enc_transform<int32_t>(ctx, g, d_pix, pix_stride, mem, &d_planes) : enc_transform<double>(ctx, g, d_pix, pix_stride, mem, &d_planes);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    const uint64_t bytes = (uint64_t)g.w * g.h * g.nc * 4;
    cudaError_t e = cudaMemcpyAsync(planes, d_planes, bytes, (p->flags & J2KGPU_ENC_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "forward transform");
    return J2KGPU_OK;
}

extern "C" int j2kgpu_encode_tile(j2kgpu_ctx *ctx, const j2k_encode_t *p, const uint8_t *pix, uint64_t pix_stride, uint8_t *out,
                                  uint64_t out_cap, uint64_t *out_len, uint32_t *blk_len, uint8_t *blk_bps, uint32_t n_blk)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!p || !out_len || (!out && out_cap)) return j2k_set_err(ctx, J2KGPU_E_ARG, "null argument");
    *out_len = 0;
    EncGeom g{};
    int rc = enc_geometry(ctx, p, g);
    if (rc) return rc;
    std::vector<EncBlk> blks;
    enc_blocks(g, &blks, nullptr);
    const uint32_t n = (uint32_t)blks.size();
    if ((blk_len || blk_bps) && n_blk < n) return j2k_set_err(ctx, J2KGPU_E_ARG, "per-block arrays hold %u entries, the tile has %u blocks", n_blk, n);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    cudaError_t e = upload_enc_tables(s);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "encoder tables");
    PoolPtrs mem{ctx, {}};
    const uint8_t *d_pix = nullptr;
    if ((rc = enc_pixels_in(ctx, p, g, pix, pix_stride, mem, &d_pix))) return rc;
    int32_t *d_planes = nullptr;
    rc = g.lossless ? enc_transform<int32_t>(ctx, g, d_pix, pix_stride, mem, &d_planes) : enc_transform<double>(ctx, g, d_pix, pix_stride, mem, &d_planes);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    // tier-1: one warp per block, every block into its own slab slot
    const uint32_t cap = (uint32_t)(64 + 5 * g.cbw * g.cbh);
    EncBlk *d_blks = (EncBlk *)mem.get((size_t)n * sizeof(EncBlk), &e);
    uint8_t *d_slab = e == cudaSuccess ? (uint8_t *)mem.get((size_t)n * cap, &e) : nullptr;
    uint32_t *d_lens = e == cudaSuccess ? (uint32_t *)mem.get((size_t)n * 4, &e) : nullptr;
    uint8_t *d_bps = e == cudaSuccess ? (uint8_t *)mem.get((size_t)n, &e) : nullptr;
    uint64_t *d_offs = e == cudaSuccess ? (uint64_t *)mem.get((size_t)(n + 1) * 8 + 16, &e) : nullptr;
    uint32_t *d_order = e == cudaSuccess ? (uint32_t *)mem.get((size_t)n * 4, &e) : nullptr;
    if (e != cudaSuccess) { cudaStreamSynchronize(s); return j2k_set_err(ctx, J2KGPU_E_NOMEM, "tier-1 encoder buffers: %s", cudaGetErrorString(e)); }
    int *d_err = (int *)(d_offs + n + 1);
    uint64_t h_tail[2] = {0, 0};                                   // total bytes, error flag
    e = cudaMemsetAsync(d_err, 0, 8, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_blks, blks.data(), (size_t)n * sizeof(EncBlk), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { cudaStreamSynchronize(s); return j2k_cuda_err(ctx, e, "block table upload"); }
    const int mode = ctx->opt.enc_bytes ? 0 : (g.cbw <= 64 ? 1 : 2);   // option enc_bytes forces the flag-byte coder (A/B, tests)
    const size_t per_warp = t1enc_warp_bytes(g.cbw, g.cbh, mode != 0);
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * per_warp > 96 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * per_warp;
    const void *kfn = mode == 0 ? (const void *)k_t1_enc<0> : (mode == 1 ? (const void *)k_t1_enc<1> : (const void *)k_t1_enc<2>);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cudaStreamSynchronize(s); return j2k_cuda_err(ctx, e, "k_t1_enc shared memory"); }
    }
    J2K_LAUNCH((k_enc_bps), (n + 7) / 8, 256, 0, s, (const EncBlk *)d_blks, n, (const int32_t *)d_planes, g.w, g.h, d_bps);
    J2K_LAUNCH((k_enc_order), 1, 1024, 0, s, (const uint8_t *)d_bps, n, d_order);
#define J2K_T1ENC(MODE_) J2K_LAUNCH((k_t1_enc<MODE_>), (n + wpc - 1) / wpc, wpc * 32, smem, s, (const EncBlk *)d_blks, (const uint32_t *)d_order, n, \
                                    (const int32_t *)d_planes, g.w, g.h, g.cbw, g.cbh, d_slab, cap, d_lens, d_bps, d_err)
    if (mode == 0) J2K_T1ENC(0);
    else if (mode == 1) J2K_T1ENC(1);
    else J2K_T1ENC(2);
#undef J2K_T1ENC
    J2K_LAUNCH((k_enc_scan), 1, 1024, 0, s, (const uint32_t *)d_lens, n, d_offs);
    ctx->launches += 4;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_tail, d_offs + n, 16, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && blk_len) e = cudaMemcpyAsync(blk_len, d_lens, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && blk_bps) e = cudaMemcpyAsync(blk_bps, d_bps, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { cudaStreamSynchronize(s); return j2k_cuda_err(ctx, e, "tier-1 encoder"); }
    if ((int)h_tail[1] != 0) return j2k_set_err(ctx, J2KGPU_E_INTERNAL, "a code block outgrew its %u-byte slot", cap);
    *out_len = h_tail[0];
    if (h_tail[0] > out_cap) return j2k_set_err(ctx, J2KGPU_E_ARG, "output buffer holds %llu bytes, the tile needs %llu", (unsigned long long)out_cap, (unsigned long long)h_tail[0]);
    if (h_tail[0] == 0) return J2KGPU_OK;
    const bool dev = (p->flags & J2KGPU_ENC_DEVICE_PTRS) != 0;
    uint8_t *d_out = dev ? out : (uint8_t *)mem.get(h_tail[0], &e);
    if (e != cudaSuccess) return j2k_set_err(ctx, J2KGPU_E_NOMEM, "tile buffer: %s", cudaGetErrorString(e));
    J2K_LAUNCH((k_enc_gather), (n + 7) / 8, 256, 0, s, (const uint8_t *)d_slab, cap, (const uint32_t *)d_lens, (const uint64_t *)d_offs, n, d_out);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess && !dev) e = cudaMemcpyAsync(out, d_out, h_tail[0], cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "tile gather");
    return J2KGPU_OK;
}
