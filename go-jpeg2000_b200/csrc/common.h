// common.h -- internal declarations shared by the CUDA translation units of libj2kgpu.so.
// Public boundary: include/j2kgpu.h.  Nothing here is exported.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <mutex>

#include "../../include/j2kgpu.h"
#include "../host/rgb_expand.h"

#define J2K_MAX_LEVELS 10

// Kernel launch / dynamic shared memory spelled as macros so that tools/emu can build these same sources for the
// CPU fiber emulator (a debugging aid; the product is the nvcc build below).
#define J2K_UNPAREN(...) __VA_ARGS__
#ifdef J2K_EMU
#define J2K_LAUNCH(K, G, B, SM, ST, ...) emu::launch(dim3(G), dim3(B), (SM), [&]() { J2K_UNPAREN K(__VA_ARGS__); })
#define J2K_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem)
#define J2K_NOINLINE __attribute__((noinline))
#define J2K_OPAQUE_PTR(p) ((void)0)
#else
#define J2K_OPAQUE_PTR(p) asm volatile("" : "+l"(p))
#define J2K_NOINLINE __noinline__
#define J2K_LAUNCH(K, G, B, SM, ST, ...) J2K_UNPAREN K<<<(G), (B), (SM), (ST)>>>(__VA_ARGS__)
#define J2K_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char j2k_dyn_smem_[]; T *name = reinterpret_cast<T *>(j2k_dyn_smem_)
#endif

// ---- device-side tables (uploaded once per job) ---------------------------------------------
struct DevCblk {                 // one code block, 40 bytes
    uint64_t data_off;           // into the job blob
    uint64_t out_off;            // int32 element offset of sample (0,0) in the coefficient arena
    uint32_t data_len;
    uint32_t out_stride;         // row stride of the destination plane, in elements
    uint16_t w, h;
    uint8_t  band, num_bps, level, num_passes;
    uint32_t len_cup;            // ISO HT: bytes of the cleanup segment (the refinement segment follows); 0 = data_len
    uint32_t pad;                // k_t1_iso: code-block style bits (J2KGPU_CBLK_*); 0 elsewhere
};

struct DevTileComp {             // one tile-component plane
    uint64_t coef_off;           // int32 element offset of the plane in the coefficient arena
    uint64_t tmp_off;            // element offset of this plane's two ping-pong level buffers in the scratch arena
    uint32_t w, h;               // plane size
    uint32_t tmp_elems;          // elements in ONE ping-pong buffer (= w_1 * h_1)
    uint32_t pad;
};

struct DevTile {                 // ncomp tile-components that share a footprint: unit of the fused last level
    uint32_t tc[4];              // tile-component indices per component
    uint32_t img_x0, img_y0;     // position of the tile inside the image
    uint32_t w, h;               // tile-component size (equal for all components)
    uint64_t out_off;            // byte offset of the owning image's pixel buffer in d_out
    uint64_t out_stride;         // bytes per output row of the owning image
    uint32_t img_w, img_h;       // clipping bounds (decoder.go:398-410)
};

struct TailParams {              // image-wide constants of the MCT / DC / pack epilogue
    int ncomp;
    int prec[4];
    int sgnd[4];
    int mct;                     // inverse MCT requested and ncomp >= 3
    int reversible;
    int fmt;                     // J2KGPU_FMT_* (resolved, never AUTO)
    int iso;                     // 1: ISO packing (no int32-overflow quirk)
    int cconv;                   // J2KGPU_CS_*: colour conversion to sRGB after the DC shift (0 = none)
};

// ---- context -----------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
};

// A/B switches and test hooks: seeded from the environment (J2KGPU_<NAME>) ONCE, in j2kgpu_create, and changed only through
// j2kgpu_set_option; no launch path reads the environment
struct J2kOpts {
    int no_fuse = 0;        // per-level IDWT launches instead of the fused levels-1+0 kernels
    int no_wide = 0;        // 4-columns-per-lane fused kernel instead of the 16-columns-per-lane one
    int no_fast_epi = 0;    // generic pixel epilogue
    int coef32 = 0;         // int32 coefficient planes everywhere
    int no_preclear = 0;    // reference HT coder: clear every row on every run
    int wide_sp = 0;        // strip height of the wide IDWT kernel in row pairs (0 = planner)
    int t1_group = 0;       // EBCOT kernels: lanes per code block (4, 8, 16, 32; 0 = default 8)
    int split_min_mpixel = 0; // a single image of at least this many Mpixel is pipelined by groups of tiles (0 = 24)
    int host_alpha = -1;    // host-buffer runs of RGBA8 images: packed R G B over PCIe, alpha filled in by host threads (-1 = auto)
    int enc_bytes = 0;      // forward path: the flag-byte tier-1 coder also for blocks at most 64 wide (A/B, tests)
    int debug_plan = 0;     // print the chunk plan of host-buffer runs
    std::string chunks;     // explicit chunk sizes of host-buffer runs, e.g. "1,1,2,4"
};

struct J2kRgbCb { J2kExpandPool *pool; J2kExpandTask t; };   // what a copy-out stream callback hands to the pool

struct j2kgpu_ctx {
    int device = 0;
    J2kOpts opt;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // stream in use (own or external)
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy-in / copy-out streams of the pipelined host-buffer run
    cudaEvent_t ev_start = nullptr;
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    // grow-only scratch used by the host-buffer entry points
    DevBuf d_in, d_out, d_aux, d_tab;
    DevBuf h_in, h_out;                  // pinned staging
    // device buffers released by finished jobs, reused by the next job (repeated decode calls do not cudaMalloc)
    std::vector<DevBuf> pool;
    std::vector<DevBuf> hpool;           // page-locked host blocks (table staging of pipelined batch calls), same policy
    std::vector<cudaEvent_t> events;     // timing-disabled events, reused across calls
    J2kExpandPool *expand = nullptr;     // host threads that widen packed RGB rows to RGBA8 (created on first use)
};
void *j2k_pool_alloc(j2kgpu_ctx *ctx, size_t bytes, cudaError_t *err);
void j2k_pool_free(j2kgpu_ctx *ctx, void *p);
void *j2k_hpool_alloc(j2kgpu_ctx *ctx, size_t bytes, size_t *cap, cudaError_t *err);
void j2k_hpool_free(j2kgpu_ctx *ctx, void *p, size_t cap);

struct j2kgpu_job {
    j2kgpu_ctx *ctx = nullptr;
    uint32_t n_img = 0, n_tc = 0, n_tiles = 0, n_cb = 0;
    j2k_image_t hdr{};                   // shared header fields
    TailParams tail{};
    int nlevels = 0;
    uint32_t max_w = 0, max_h = 0;       // largest tile-component
    int max_bps = 0;
    bool need_clear = false;
    uint32_t stream_levels = 0;
    int iso = 0;                         // J2KGPU_MODE_ISO
    int coef16 = 0;                      // coefficient arena holds int16 (every magnitude provably < 2^15) instead of int32
    int fused_ok = 0;                    // levels 1 + 0 + pixel epilogue run as one kernel (idwt_fused.cu)
    int fast_epi = 0;                    // every tile qualifies for the fused kernel's fixed RGBA8 epilogue
    int wide_ok = 0;                     // ... and for the 16-columns-per-lane variant (idwt_wide.cu)
    int ht_refine = 0;                   // ISO HT: some block has SigProp / MagRef passes
    int t1_segmented = 0;                // ISO EBCOT: some item has a non-default code-block style (styled instantiation of k_t1_iso)
    int precleared = 0;                  // reference HT coder: planes zeroed at job creation, decoder clears every 4th row only
    int pix_fill = 0;                    // some pixel of some image is covered by no tile: pre-fill with the pixel of zero coefficients
    std::vector<uint8_t> item_fill;      // per item: needs the pre-fill
    std::vector<uint32_t> row_bytes;     // per item: width * bytes per pixel (rows are copied back without their padding)
    std::vector<uint64_t> out_stride;    // per item
    std::vector<uint32_t> img_h;         // per item
    const DevTile *h_tiles_host = nullptr;               // the tile table inside h_tables (valid while h_tables is)
    void *h_tables = nullptr; size_t h_tables_cap = 0;   // page-locked staging of the tables while their upload is in flight
    void *d_fillpix = nullptr;           // 16 bytes: the packed pixel an all-zero coefficient decodes to
    std::vector<uint32_t> item_cb, item_tc, item_tile;   // first block / tile-component / tile of each item (+ end)
    std::vector<uint64_t> tc_coef_off;                   // coefficient-arena offset of each tile-component (host copy)
    std::vector<cudaEvent_t> ev_in, ev_done;             // per chunk of the pipelined host-buffer run
    float *d_steps = nullptr;            // ISO irreversible: dequantisation step per block
    void *d_htscratch = nullptr;         // reference HT coder: quad table between its two kernels
    DevCblk *d_cblks = nullptr;
    DevTileComp *d_tcs = nullptr;
    DevTile *d_tiles = nullptr;
    void *d_coef = nullptr;     uint64_t coef_elems = 0;   // int32 or int16 elements (coef16)
    void *d_tmp = nullptr;      uint64_t tmp_bytes = 0;
    uint64_t blob_bytes = 0, out_bytes = 0;
    std::vector<uint64_t> blob_off, out_off, out_size;
    std::vector<std::pair<void *, size_t>> owned;   // device allocations taken from the ctx pool
    // staging for run_host
    void *d_blob = nullptr; void *d_pix = nullptr;
    void *h_blob = nullptr; void *h_pix = nullptr;
    // packed-RGB transfer of host-buffer runs (rgb_expand.h): possible for this job / in use by the current run
    int rgb24_ok = 0, rgb24 = 0;
    void *h_rgb = nullptr; size_t h_rgb_cap = 0;         // page-locked staging of the packed rows
    std::vector<uint64_t> rgb_off;                       // per item: offset in h_rgb
    std::deque<J2kRgbCb> rgb_tasks;                      // one per item of the current run (stable addresses for the stream callbacks)
};

// ---- error helpers -----------------------------------------------------------------------------
int j2k_set_err(j2kgpu_ctx *ctx, int code, const char *fmt, ...);
int j2k_cuda_err(j2kgpu_ctx *ctx, cudaError_t e, const char *what);
#define J2K_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return j2k_cuda_err((ctx), e__, #call); } while (0)
int j2k_reserve(j2kgpu_ctx *ctx, DevBuf &b, size_t bytes, bool pinned_host);
int j2k_ctx_copy_streams(j2kgpu_ctx *ctx);

// ---- kernel launchers (each returns a cudaError_t; all asynchronous on `s`) ---------------------
// entropy stage: one warp per code block
// d_coef: coefficient arena, int16 elements when coef16 else int32
// group: lanes per code block (4, 8, 16, 32; 0 = default), i.e. 32 / group blocks share a warp's instruction stream
cudaError_t launch_t1_ref(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          int max_bps, int group, cudaStream_t s);
cudaError_t launch_ht_ref(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          int planes_precleared, void *d_scratch, uint64_t blob_bytes, cudaStream_t s);
size_t j2k_htref_scratch_bytes(uint32_t n_blocks);   // device scratch launch_ht_ref needs for n blocks
int j2k_htref_launches();        // kernels per launch_ht_ref call
// ISO/IEC 15444-1 Annex D decoder (stripe-order passes, standard tables, pass truncation)
cudaError_t launch_t1_iso(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          const float *d_steps, int irrev, int segmented, int group, cudaStream_t s);
// ISO/IEC 15444-15 block decoder (VLC kernel + MagSgn kernel)
// refine: some block carries SigProp / MagRef passes (num_passes > 1): the refinement kernel runs between the two
cudaError_t launch_ht_iso(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, void *d_coef, int coef16,
                          const float *d_steps, int irrev, int coef_bits, int refine, void *d_scratch, uint64_t blob_bytes,
                          cudaStream_t s);
size_t j2k_htiso_scratch_bytes(uint32_t n_blocks, int refine);   // device scratch between the kernels
int j2k_htiso_launches(int refine);                              // kernels per launch_ht_iso call

// inverse DWT, REF (dense-prefix) addressing.  One call = one decomposition level of every
// tile-component in the table.  `lvl` counts from 0 (full resolution).  For lvl > 0 the output goes to the
// tile-component's ping-pong buffer; for lvl == 0 it goes to `d_plane_out` (int32 planes at coef_off, in place
// is NOT allowed) or, when tiles != nullptr, through the fused MCT + DC + pack epilogue into d_pix.
struct IdwtLaunch {
    const DevTileComp *d_tcs; uint32_t n_tc;
    const DevTile *d_tiles; uint32_t n_tiles;      // only for the fused last level
    const void *d_coef;                             // coefficient arena (int32, or int16 when coef16; double when f64_io)
    int coef16;
    uint32_t tc_first, tile_first;                  // sub-range of the tables this launch covers (batch pipelining)
    int fast_epi;                                   // fused kernel: 3 x 8-bit unsigned, RCT, RGBA8, tiles inside the image, aligned rows
    int wide_ok;                                    // + every tile width a multiple of 16: the 16-columns-per-lane kernel (idwt_wide.cu)
    void *d_tmp;                                    // ping-pong arena (int32 for 5-3, double for 9-7)
    int nlevels, lvl;
    uint32_t max_w, max_h;                          // largest tile-component (grid sizing)
    int reversible;
    int f64_io;                                     // 9-7 stage API: coefficient arena and output planes are double
    int iso;                                        // 1: Mallat addressing + ISO order (rows, then columns)
    int wide_sp;                                    // J2kOpts.wide_sp
    int rgb24;                                      // wide kernel, 3 components: packed R G B out (3 bytes per pixel, stride 3/4 of the table's)
    uint32_t stream_levels;                         // bit l set: level l of every tile-component fits the streaming kernel
    int32_t *d_plane_out;                           // lvl == 0 without tiles: output planes (same offsets as coef)
    uint8_t *d_pix;                                 // lvl == 0 with tiles: packed pixels
    TailParams tail;
};
cudaError_t launch_idwt_level(const IdwtLaunch &p, cudaStream_t s, int *n_launches);
cudaError_t launch_idwt53_stream(const IdwtLaunch &p, cudaStream_t s);
cudaError_t launch_idwt97_stream(const IdwtLaunch &p, cudaStream_t s);     // 9-7 float64, REF semantics (idwt97_stream.cu)
// levels 1 and 0 of every tile + inverse MCT + DC shift + clamp + pack in one kernel (5-3; see idwt_fused.cu)
cudaError_t launch_idwt53_fused(const IdwtLaunch &p, cudaStream_t s);
cudaError_t launch_idwt53_wide(const IdwtLaunch &p, cudaStream_t s);
// a tile-component fits the fused kernel when its width is a multiple of 8 and its height a multiple of 4
static inline bool j2k_fused_ok(uint32_t w, uint32_t h) { return w >= 8 && (w & 7) == 0 && h >= 4 && (h & 3) == 0; }
// level l fits the streaming kernel when its width is a multiple of 4 and its height is even (>= 2)
static inline bool j2k_stream_ok(uint32_t w, uint32_t h, int lvl)
{
    uint32_t wl = (w + (1u << lvl) - 1) >> lvl, hl = (h + (1u << lvl) - 1) >> lvl;
    return wl >= 4 && (wl & 3) == 0 && hl >= 2 && (hl & 1) == 0;
}

// unfused tail: planar int32 components -> (inverse MCT, DC shift) -> planes and/or packed pixels
cudaError_t launch_tail(const int32_t *const d_comps[4], int32_t *const d_planes_out[4], uint8_t *d_pix,
                        uint64_t out_stride, uint32_t width, uint32_t height, const TailParams &tp,
                        int apply_tail, cudaStream_t s);
cudaError_t launch_inverse_ict_f64(double *y, double *cb, double *cr, uint64_t n, cudaStream_t s);
// rows x row_bytes of d_pix (row pitch out_stride) <- the bpp-byte pixel at d_pattern, repeated
cudaError_t launch_fill_pixels(uint8_t *d_pix, uint64_t out_stride, uint32_t row_bytes, uint32_t rows, int bpp,
                               const uint8_t *d_pattern, cudaStream_t s);

// Go's int32(float64) as the reference runs it on amd64 (CVTTSD2SL): truncation, and 0x80000000 for NaN / out of range
// (CUDA's cvt.rzi saturates instead, which differs for positive overflow)
#if defined(__CUDACC__) || defined(J2K_EMU)
__device__ __forceinline__ int32_t j2k_f64_to_i32(double v)
{
    const int32_t r = __double2int_rz(v);
    return (v > -2147483649.0 && v < 2147483648.0) ? r : (int32_t)0x80000000u;
}
#endif

int j2k_resolve_fmt(int ncomp, int prec, int fmt);
int j2k_fmt_bpp(int fmt);
