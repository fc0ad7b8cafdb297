// cuda_runtime.h (EMULATION SHIM) -- development tool, NOT part of the product.
//
// tools/emu builds the very same .cu sources of go-jpeg2000_b200/csrc with g++ against this header, so that
// kernel logic (indexing, halos, warp shuffles, barriers) can be debugged in a container without a GPU.  Each
// CUDA thread is a fiber; a CTA's fibers are scheduled round-robin on one OS thread and switch at
// __syncthreads / __syncwarp / shuffle points; CTAs run one after another.  Nothing in the package, the tests
// marked gpu, bench.py or smoke() loads the resulting library: it is only used by tools/emu/run_emu.py.
#pragma once
#ifndef J2K_EMU
#error "this header is the CPU emulation shim; build with -DJ2K_EMU"
#endif

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <functional>

// ---- qualifiers ------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static          /* CTAs run one at a time on one OS thread */
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ----------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct int3 { int x, y, z; };
static inline int3 make_int3(int x, int y, int z) { int3 r; r.x = x; r.y = y; r.z = z; return r; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(4))) short2 { short x, y; };
struct __attribute__((aligned(8))) short4 { short x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline short2 make_short2(short x, short y) { short2 r; r.x = x; r.y = y; return r; }
static inline short4 make_short4(short x, short y, short z, short w) { short4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- fiber scheduler (emu_runtime.cpp) ------------------------------------------------------------------
namespace emu {
struct ThreadCtx {
    uint3 tid;
    int lin, warp, lane;
};
extern ThreadCtx *cur;
extern uint3 block_idx;
extern dim3 block_dim, grid_dim;
extern unsigned char *dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
void sync_block();
void sync_warp();
uint64_t shfl_exchange(uint64_t v, int src_lane);   // every live lane of the warp calls it
unsigned ballot(int pred);
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)
#define warpSize 32

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T emu_shfl(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    u = emu::shfl_exchange(u, src);
    T r;
    memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const int lane = emu::cur->lane;
    return emu_shfl(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    const int lane = emu::cur->lane, base = lane & ~(width - 1);
    const int src = lane - (int)d;
    return emu_shfl(v, src < base ? lane : src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    const int lane = emu::cur->lane, base = lane & ~(width - 1);
    const int src = lane + (int)d;
    return emu_shfl(v, src >= base + width ? lane : src);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    (void)width;
    return emu_shfl(v, emu::cur->lane ^ m);
}
static inline unsigned __ballot_sync(unsigned, int pred) { return emu::ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return emu::ballot(pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return emu::ballot(!pred) == 0; (void)m; }
static inline unsigned __activemask() { return emu::ballot(1); }

// ---- intrinsics -------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline int __double2int_rz(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int)0x80000000;
    return (int)v;
}
static inline int __float2int_rz(float v) { return __double2int_rz((double)v); }
static inline int __float2int_rn(float v)
{
    if (v != v) return 0;
    double r = nearbyint((double)v);
    return __double2int_rz(r);
}
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline unsigned __brev(unsigned v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t t = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((t >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32);
}
using std::max;
using std::min;
static inline int max(int a, unsigned b) { return (int)std::max<long long>(a, b); }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }

// ---- runtime API (synchronous, host memory) ---------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidDevice = 101 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
#define CUDART_CB
typedef void (*cudaHostFn_t)(void *);
static inline cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void *ud) { fn(ud); return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 148; return cudaSuccess; }
// exact size (posix_memalign): under AddressSanitizer (make asan) any access past the requested bytes is reported
static inline cudaError_t cudaMalloc(void **p, size_t n) { return posix_memalign(p, 256, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
enum { cudaHostAllocPortable = 1, cudaHostRegisterPortable = 1 };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *src, size_t n) { memcpy(&sym, src, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T &sym, const void *src, size_t n, size_t, cudaMemcpyKind, cudaStream_t) { memcpy(&sym, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t y = 0; y < h; y++) memmove((char *)d + y * dp, (const char *)s + y * sp, w);
    return cudaSuccess;
}
enum { cudaHostAllocDefault = 0 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
