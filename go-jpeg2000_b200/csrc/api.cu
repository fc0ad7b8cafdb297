// api.cu -- the C ABI of libj2kgpu.so (include/j2kgpu.h): context, job flattening / validation,
// whole-path launch sequence, host-buffer staging and the per-stage entry points.
// There is no CPU fallback anywhere in this file: every entry point either runs CUDA kernels or
// returns an error.
#include "common.h"
#include "../host/tier2.h"

#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <thread>
#include <map>
#include <new>
#include <tuple>

cudaError_t launch_t1_ref_stage(const DevCblk *d_cblks, uint32_t n, const uint8_t *d_blob, int32_t *d_coef,
                                int max_bps, int group, cudaStream_t s);

// ---- errors ------------------------------------------------------------------------------------------
int j2k_set_err(j2kgpu_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

int j2k_cuda_err(j2kgpu_ctx *ctx, cudaError_t e, const char *what)
{
    return j2k_set_err(ctx, J2KGPU_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int j2k_reserve(j2kgpu_ctx *ctx, DevBuf &b, size_t bytes, bool pinned_host)
{
    if (bytes <= b.cap) return J2KGPU_OK;
    size_t cap = bytes + bytes / 4 + 4096;
    if (b.p) {
        if (pinned_host) cudaFreeHost(b.p); else cudaFree(b.p);
        b.p = nullptr; b.cap = 0;
    }
    cudaError_t e = pinned_host ? cudaMallocHost(&b.p, cap) : cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) { b.p = nullptr; return j2k_set_err(ctx, J2KGPU_E_NOMEM, "allocating %zu bytes: %s", cap, cudaGetErrorString(e)); }
    b.cap = cap;
    return J2KGPU_OK;
}

int j2k_ctx_copy_streams(j2kgpu_ctx *ctx)
{
    if (ctx->s_in) return J2KGPU_OK;
    J2K_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    J2K_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    J2K_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    return J2KGPU_OK;
}

// ---- ctx-level buffer pools (grow-only caches: repeated decode calls neither cudaMalloc nor cudaHostAlloc) --------
static std::map<void *, size_t> &pool_sizes()
{
    static std::map<void *, size_t> m;
    return m;
}
static std::mutex g_pool_mu;
static constexpr size_t kPoolEntries = 1024;

void *j2k_pool_alloc(j2kgpu_ctx *ctx, size_t bytes, cudaError_t *err)
{
    *err = cudaSuccess;
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); i++)
        if (ctx->pool[i].cap >= bytes && ctx->pool[i].cap <= 2 * bytes + (1u << 20) &&
            (best < 0 || ctx->pool[i].cap < ctx->pool[best].cap)) best = (int)i;
    if (best >= 0) {
        void *p = ctx->pool[best].p;
        ctx->pool.erase(ctx->pool.begin() + best);
        return p;
    }
    void *p = nullptr;
    size_t cap = (bytes + 255) / 256 * 256;
    *err = cudaMalloc(&p, cap);
    if (*err != cudaSuccess) {                      // drop the cache and retry once
        for (auto &b : ctx->pool) { std::lock_guard<std::mutex> g(g_pool_mu); pool_sizes().erase(b.p); cudaFree(b.p); }
        ctx->pool.clear();
        *err = cudaMalloc(&p, cap);
        if (*err != cudaSuccess) return nullptr;
    }
    std::lock_guard<std::mutex> g(g_pool_mu);
    pool_sizes()[p] = cap;
    return p;
}

void j2k_pool_free(j2kgpu_ctx *ctx, void *p)
{
    if (!p) return;
    size_t cap = 0;
    { std::lock_guard<std::mutex> g(g_pool_mu); auto it = pool_sizes().find(p); if (it != pool_sizes().end()) cap = it->second; }
    if (cap == 0 || ctx->pool.size() >= kPoolEntries) {
        std::lock_guard<std::mutex> g(g_pool_mu);
        pool_sizes().erase(p);
        cudaFree(p);
        return;
    }
    DevBuf b; b.p = p; b.cap = cap;
    ctx->pool.push_back(b);
}

// page-locked host blocks (the table staging of pipelined batch calls)
void *j2k_hpool_alloc(j2kgpu_ctx *ctx, size_t bytes, size_t *cap, cudaError_t *err)
{
    *err = cudaSuccess;
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (size_t i = 0; i < ctx->hpool.size(); i++)
        if (ctx->hpool[i].cap >= bytes && (best < 0 || ctx->hpool[i].cap < ctx->hpool[best].cap)) best = (int)i;
    if (best >= 0) {
        void *p = ctx->hpool[best].p;
        *cap = ctx->hpool[best].cap;
        ctx->hpool.erase(ctx->hpool.begin() + best);
        return p;
    }
    void *p = nullptr;
    *cap = (bytes + bytes / 8 + 4095) / 4096 * 4096;
    *err = cudaHostAlloc(&p, *cap, cudaHostAllocDefault);
    return *err == cudaSuccess ? p : nullptr;
}

void j2k_hpool_free(j2kgpu_ctx *ctx, void *p, size_t cap)
{
    if (!p) return;
    if (ctx->hpool.size() >= 64) { cudaFreeHost(p); return; }
    DevBuf b; b.p = p; b.cap = cap;
    ctx->hpool.push_back(b);
}

static cudaEvent_t ctx_event(j2kgpu_ctx *ctx, cudaError_t *err)
{
    *err = cudaSuccess;
    if (!ctx->events.empty()) { cudaEvent_t e = ctx->events.back(); ctx->events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    *err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}

int j2k_resolve_fmt(int ncomp, int prec, int fmt)
{
    int want = ncomp == 1 ? (prec <= 8 ? J2KGPU_FMT_GRAY8 : J2KGPU_FMT_GRAY16)
                          : (prec <= 8 ? J2KGPU_FMT_RGBA8 : J2KGPU_FMT_RGBA64);      // decoder.go:427-523
    if (fmt == J2KGPU_FMT_AUTO || fmt == want) return want;
    return -1;
}

int j2k_fmt_bpp(int fmt)
{
    switch (fmt) {
    case J2KGPU_FMT_GRAY8: return 1;
    case J2KGPU_FMT_GRAY16: return 2;
    case J2KGPU_FMT_RGBA8: return 4;
    case J2KGPU_FMT_RGBA64: return 8;
    }
    return 0;
}

static int make_tail(j2kgpu_ctx *ctx, const j2k_image_t &im, TailParams &tp)
{
    if (im.ncomp != 1 && im.ncomp != 3 && im.ncomp != 4)
        return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "unsupported number of components: %d", (int)im.ncomp);   // decoder.go:585-586
    for (int c = 0; c < im.ncomp; c++)
        if (im.prec[c] < 1 || im.prec[c] > 16)
            return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "component %d precision %d outside 1..16", c, (int)im.prec[c]);
    memset(&tp, 0, sizeof tp);
    tp.ncomp = im.ncomp;
    for (int c = 0; c < 4; c++) { tp.prec[c] = c < im.ncomp ? im.prec[c] : 8; tp.sgnd[c] = c < im.ncomp ? (im.sgnd[c] != 0) : 1; }
    tp.mct = (im.mct != 0 && im.ncomp >= 3);                                            // decoder.go:322
    tp.reversible = im.reversible != 0;
    tp.iso = im.mode == J2KGPU_MODE_ISO;
    if (im.colorspace > J2KGPU_CS_ROMM)
        return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "colour conversion %d is not built", (int)im.colorspace);
    tp.cconv = im.ncomp >= 3 ? im.colorspace : 0;                                       // colorspace.go:93-95: fewer than 3 components -> no-op
    tp.fmt = j2k_resolve_fmt(im.ncomp, im.prec[0], im.out_fmt);
    if (tp.fmt < 0) return j2k_set_err(ctx, J2KGPU_E_ARG, "out_fmt %d does not match ncomp/precision", (int)im.out_fmt);
    return J2KGPU_OK;
}

// ---- context -------------------------------------------------------------------------------------------
extern "C" int j2kgpu_abi_version(void) { return J2KGPU_ABI_VERSION; }

extern "C" const char *j2kgpu_strerror(int code)
{
    switch (code) {
    case J2KGPU_OK: return "ok";
    case J2KGPU_E_ARG: return "invalid argument";
    case J2KGPU_E_RANGE: return "offset or block out of range";
    case J2KGPU_E_UNSUPPORTED: return "unsupported configuration";
    case J2KGPU_E_CUDA: return "CUDA error";
    case J2KGPU_E_NOMEM: return "out of memory";
    case J2KGPU_E_NODEVICE: return "no CUDA device";
    case J2KGPU_E_INTERNAL: return "internal bound exceeded";
    }
    return "unknown error";
}

// option names = the J2KGPU_<NAME> environment switches, lower case, without the prefix
static int set_option(J2kOpts &o, const char *name, const char *value)
{
    const std::string k = name ? name : "", v = value ? value : "";
    const int iv = atoi(v.c_str());
    const int on = !v.empty() && v != "0";
    if (k == "no_fuse") o.no_fuse = on;
    else if (k == "no_wide") o.no_wide = on;
    else if (k == "no_fast_epi") o.no_fast_epi = on;
    else if (k == "coef32") o.coef32 = on;
    else if (k == "no_preclear") o.no_preclear = on;
    else if (k == "wide_sp") o.wide_sp = iv;
    else if (k == "t1_group") o.t1_group = iv;
    else if (k == "host_alpha") o.host_alpha = iv;
    else if (k == "split_min_mpixel") o.split_min_mpixel = iv;
    else if (k == "enc_bytes") o.enc_bytes = on;
    else if (k == "debug_plan") o.debug_plan = on;
    else if (k == "chunks") o.chunks = v;
    else return J2KGPU_E_ARG;
    return J2KGPU_OK;
}

extern "C" int j2kgpu_create(int device, j2kgpu_ctx **out)
{
    if (!out) return J2KGPU_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return J2KGPU_E_NODEVICE;
    if (device < 0 || device >= ndev) return J2KGPU_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return J2KGPU_E_CUDA;
    j2kgpu_ctx *ctx = new (std::nothrow) j2kgpu_ctx();
    if (!ctx) return J2KGPU_E_NOMEM;
    ctx->device = device;
    // the environment is read here, once per context, and nowhere else
    static const char *const names[] = {"no_fuse", "no_wide", "no_fast_epi", "coef32", "no_preclear", "wide_sp", "t1_group", "host_alpha", "split_min_mpixel", "enc_bytes", "debug_plan", "chunks"};
    for (const char *nm : names) {
        std::string env = "J2KGPU_";
        for (const char *c = nm; *c; c++) env += (char)toupper((unsigned char)*c);
        if (const char *e = getenv(env.c_str())) set_option(ctx->opt, nm, e);
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return J2KGPU_E_CUDA; }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return J2KGPU_OK;
}

extern "C" int j2kgpu_set_option(j2kgpu_ctx *ctx, const char *name, const char *value)
{
    if (!ctx || !name) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    const int rc = set_option(ctx->opt, name, value);
    if (rc) return j2k_set_err(ctx, rc, "unknown option '%s'", name);
    return J2KGPU_OK;
}

static void free_buf(DevBuf &b, bool pinned) { if (b.p) { if (pinned) cudaFreeHost(b.p); else cudaFree(b.p); } b.p = nullptr; b.cap = 0; }

extern "C" void j2kgpu_destroy(j2kgpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_buf(ctx->d_in, false); free_buf(ctx->d_out, false); free_buf(ctx->d_aux, false); free_buf(ctx->d_tab, false);
    free_buf(ctx->h_in, true); free_buf(ctx->h_out, true);
    for (auto &b : ctx->pool) { { std::lock_guard<std::mutex> g(g_pool_mu); pool_sizes().erase(b.p); } cudaFree(b.p); }
    ctx->pool.clear();
    for (auto &b : ctx->hpool) cudaFreeHost(b.p);
    ctx->hpool.clear();
    delete ctx->expand;
    ctx->expand = nullptr;
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    ctx->events.clear();
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *j2kgpu_last_error(const j2kgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int j2kgpu_set_stream(j2kgpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return J2KGPU_OK;
}

extern "C" uint64_t j2kgpu_launch_count(const j2kgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int j2kgpu_sync(j2kgpu_ctx *ctx)
{
    if (!ctx) return J2KGPU_E_ARG;
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

// ---- job ---------------------------------------------------------------------------------------------
static uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

static void job_free(j2kgpu_job *job)
{
    if (!job) return;
    if (job->ctx) {
        cudaSetDevice(job->ctx->device);
        void *ps[] = {job->d_cblks, job->d_tcs, job->d_tiles, job->d_coef, job->d_tmp, job->d_blob, job->d_pix, job->d_steps,
                      job->d_htscratch, job->d_fillpix};
        for (void *p : ps) j2k_pool_free(job->ctx, p);
        for (cudaEvent_t e : job->ev_in) if (e) job->ctx->events.push_back(e);
        for (cudaEvent_t e : job->ev_done) if (e) job->ctx->events.push_back(e);
        if (job->h_tables) j2k_hpool_free(job->ctx, job->h_tables, job->h_tables_cap);
        if (job->h_rgb) j2k_hpool_free(job->ctx, job->h_rgb, job->h_rgb_cap);
    }
    if (job->h_blob) cudaFreeHost(job->h_blob);
    if (job->h_pix) cudaFreeHost(job->h_pix);
    delete job;
}

extern "C" void j2kgpu_job_destroy(j2kgpu_job *job)
{
    if (!job) return;
    if (job->ctx) { std::lock_guard<std::mutex> g(job->ctx->mu); cudaStreamSynchronize(job->ctx->stream); job_free(job); return; }
    job_free(job);
}

// codeword segments that np coding passes of a classic block touch (B.10.7.2, D.4, D.6); 1 without BYPASS / TERMALL
static uint32_t t1_num_segments(uint32_t style, uint32_t np)
{
    if (np == 0 || !(style & (J2KGPU_CBLK_BYPASS | J2KGPU_CBLK_TERMALL))) return 1;
    const uint32_t i = np - 1;
    if (style & J2KGPU_CBLK_TERMALL) return i + 1;
    if (i < 10) return 1;
    return 1 + 2 * ((i - 10) / 3) + ((i - 10) % 3 == 2 ? 1 : 0) + 1;
}

static bool same_header(const j2k_image_t &a, const j2k_image_t &b)
{
    return a.ncomp == b.ncomp && !memcmp(a.prec, b.prec, 4) && !memcmp(a.sgnd, b.sgnd, 4) && a.mct == b.mct &&
           a.reversible == b.reversible && a.nlevels == b.nlevels && a.ht == b.ht && a.mode == b.mode &&
           a.out_fmt == b.out_fmt && a.colorspace == b.colorspace;
}

// a table built in place in page-locked memory
template <class T>
struct Fixed {
    T *p = nullptr; size_t n = 0;
    void push_back(const T &v) { p[n++] = v; }
    T &operator[](size_t i) { return p[i]; }
    size_t size() const { return n; }
};

struct Rect { uint32_t x0, y0, x1, y1; };

// how do the rectangles (all inside a w x h plane) cover it?  0: exactly once; 1: without overlap but with holes;
// 2: two of them overlap.  The overlap test is a sweep over the rectangles sorted by their top edge.
static int rect_cover(std::vector<Rect> &r, uint32_t w, uint32_t h)
{
    uint64_t area = 0;
    for (const Rect &q : r) area += (uint64_t)(q.x1 - q.x0) * (q.y1 - q.y0);
    std::sort(r.begin(), r.end(), [](const Rect &a, const Rect &b) { return a.y0 != b.y0 ? a.y0 < b.y0 : a.x0 < b.x0; });
    for (size_t i = 0; i < r.size(); i++)
        for (size_t j = i + 1; j < r.size() && r[j].y0 < r[i].y1; j++)
            if (r[j].x0 < r[i].x1 && r[i].x0 < r[j].x1) return 2;
    return area == (uint64_t)w * h ? 0 : 1;
}
static bool tiles_exactly(std::vector<Rect> &r, uint32_t w, uint32_t h) { return rect_cover(r, w, h) == 0; }

// async_stream != nullptr: the tables are staged in page-locked memory owned by the job and uploaded on that stream
// without any synchronisation (pipelined batch calls: the host builds the next chunk's tables while the device decodes
// this one); nullptr: uploaded on the ctx stream and synchronised before returning.
// Blocks of one image in order of falling height, then width (counting sort, stable): the entropy kernels give a block
// to a thread or to a few lanes of a warp and run as long as the tallest block of the warp, so neighbours in the table
// should be alike; the tallest first also keeps the last wave short.  The table order carries no meaning otherwise.
static void sort_blocks_by_shape(DevCblk *cb, float *step, size_t n, std::vector<DevCblk> &tmp, std::vector<float> &stmp)
{
    if (n < 2) return;
    uint32_t count[65 * 65 + 1] = {0};
    auto key = [](const DevCblk &c) { return (uint32_t)(64 - c.h) * 65u + (uint32_t)(64 - c.w); };
    bool sorted = true;
    for (size_t i = 0; i < n; i++) { count[key(cb[i]) + 1]++; if (i && key(cb[i]) < key(cb[i - 1])) sorted = false; }
    if (sorted) return;
    for (int k = 0; k < 65 * 65; k++) count[k + 1] += count[k];
    tmp.assign(cb, cb + n); stmp.assign(step, step + n);
    for (size_t i = 0; i < n; i++) { const uint32_t d = count[key(tmp[i])]++; cb[d] = tmp[i]; step[d] = stmp[i]; }
}

static int job_build(j2kgpu_ctx *ctx, uint32_t n_img, const j2k_batch_item_t *items, j2kgpu_job **out, cudaStream_t async_stream)
{
    if (!ctx || !out || !items || n_img == 0) return j2k_set_err(ctx, J2KGPU_E_ARG, "null or empty batch");
    *out = nullptr;
    const j2k_image_t &hdr = items[0].image;
    const J2kOpts &opt = ctx->opt;
    TailParams tp;
    int rc = make_tail(ctx, hdr, tp);
    if (rc) return rc;
    if (hdr.mode != J2KGPU_MODE_REF && hdr.mode != J2KGPU_MODE_ISO)
        return j2k_set_err(ctx, J2KGPU_E_ARG, "unknown mode %d", (int)hdr.mode);
    const bool iso = hdr.mode == J2KGPU_MODE_ISO;
    if (hdr.nlevels > J2K_MAX_LEVELS) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "nlevels %d > %d", (int)hdr.nlevels, J2K_MAX_LEVELS);
    const int bpp = j2k_fmt_bpp(tp.fmt);

    uint64_t tot_cb = 0, tot_tc = 0;
    for (uint32_t ii = 0; ii < n_img; ii++) { tot_cb += items[ii].n_cblks; tot_tc += items[ii].n_tilecomps; }
    if (tot_cb > 0xFFFFFFF0ull || tot_tc > 0xFFFFFFF0ull) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "batch too large");

    j2kgpu_job *job = new (std::nothrow) j2kgpu_job();
    if (!job) return j2k_set_err(ctx, J2KGPU_E_NOMEM, "job");
    job->ctx = ctx; job->n_img = n_img; job->hdr = hdr; job->tail = tp; job->nlevels = hdr.nlevels; job->iso = iso;
    cudaSetDevice(ctx->device);

    // one page-locked block for all tables: [DevCblk x tot_cb][float x tot_cb][DevTileComp x tot_tc][DevTile x tot_tc]
    const size_t off_steps = align_up(tot_cb * sizeof(DevCblk), 256), off_tcs = align_up(off_steps + tot_cb * sizeof(float), 256),
                 off_tiles = align_up(off_tcs + tot_tc * sizeof(DevTileComp), 256), tab_bytes = align_up(off_tiles + tot_tc * sizeof(DevTile), 256);
    cudaError_t e = cudaSuccess;
    job->h_tables = j2k_hpool_alloc(ctx, tab_bytes, &job->h_tables_cap, &e);
    if (e != cudaSuccess) { job_free(job); return j2k_set_err(ctx, J2KGPU_E_NOMEM, "table staging: %s", cudaGetErrorString(e)); }
    uint8_t *hb = (uint8_t *)job->h_tables;
    Fixed<DevCblk> cbs;       cbs.p = (DevCblk *)hb;
    Fixed<float> steps;       steps.p = (float *)(hb + off_steps);          // ISO irreversible: dequantisation step per block
    Fixed<DevTileComp> tcs;   tcs.p = (DevTileComp *)(hb + off_tcs);
    Fixed<DevTile> tiles;     tiles.p = (DevTile *)(hb + off_tiles);
    job->h_tiles_host = tiles.p;

    uint64_t coef_elems = 0, tmp_elems = 0, blob_bytes = 0, out_bytes = 0;
    int max_bps = 0;
    bool need_clear = false;
    uint32_t stream_levels = hdr.nlevels ? ((1u << hdr.nlevels) - 1) : 0;
    bool fused_ok = hdr.reversible && hdr.nlevels >= 1 && !opt.no_fuse;
    // a colour conversion (j2k_image_t.colorspace) lives in two out-of-line epilogues only: the fused kernel's generic one
    // (put_quad_generic) and the general tiled kernel's; the streaming kernels stay free of it (see tail.cuh)
    if (tp.cconv) stream_levels &= ~1u;
    const bool will_coef16 = !opt.coef32 && iso && hdr.reversible && hdr.coef_bits >= 1 && hdr.coef_bits <= 14;
    std::vector<Rect> rects;
    std::vector<std::vector<Rect>> blk_rects;
    std::vector<DevCblk> shape_tmp;
    std::vector<float> step_tmp;

#define J2K_FAIL(...) do { const int rc__ = j2k_set_err(__VA_ARGS__); job_free(job); return rc__; } while (0)
    for (uint32_t ii = 0; ii < n_img; ii++) {
        const j2k_batch_item_t &it = items[ii];
        const j2k_image_t &im = it.image;
        if (!same_header(im, hdr)) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: header differs from item 0", ii);
        if (im.width == 0 || im.height == 0) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: empty image", ii);
        // classic blocks: all six style bits of Table A.19.  BYPASS / TERMALL blocks carry several codeword segments: their
        // byte counts follow the block's bytes in the blob (j2k_image_t.cblk_style)
        const uint32_t cstyle = (iso && !im.ht) ? im.cblk_style : 0u;
        if (cstyle & ~0x3Fu) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "item %u: code-block style %02X", ii, cstyle);
        if (cstyle) job->t1_segmented = 1;                   // the styled instantiation of k_t1_iso
        if ((!it.tilecomps && it.n_tilecomps) || (!it.cblks && it.n_cblks) || (!it.blob && it.blob_len))
            J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: null table", ii);
        if (it.out_stride < (uint64_t)im.width * bpp || it.out_stride % bpp)
            J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: out_stride %llu too small or not a multiple of %d", ii, (unsigned long long)it.out_stride, bpp);
        const uint32_t tc_base = (uint32_t)tcs.size();
        job->item_cb.push_back((uint32_t)cbs.size());
        job->item_tc.push_back(tc_base);
        job->item_tile.push_back((uint32_t)tiles.size());
        job->blob_off.push_back(blob_bytes);
        job->out_off.push_back(out_bytes);
        job->out_size.push_back(it.out_stride * im.height);
        job->row_bytes.push_back(im.width * (uint32_t)bpp);
        job->out_stride.push_back(it.out_stride);
        job->img_h.push_back(im.height);

        std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>, uint32_t> tile_of;
        for (uint32_t t = 0; t < it.n_tilecomps; t++) {
            const j2k_tilecomp_t &tc = it.tilecomps[t];
            if (tc.comp >= im.ncomp || tc.x1 <= tc.x0 || tc.y1 <= tc.y0) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u tile-component %u: bad bounds/component", ii, t);
            DevTileComp d{};
            d.w = tc.x1 - tc.x0; d.h = tc.y1 - tc.y0;
            if ((uint64_t)d.w * d.h > (1ull << 31)) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "tile-component too large");
            d.coef_off = coef_elems;
            job->tc_coef_off.push_back(coef_elems);
            coef_elems = align_up(coef_elems + (uint64_t)d.w * d.h, 32);
            d.tmp_elems = (uint32_t)align_up((uint64_t)((d.w + 1) / 2) * ((d.h + 1) / 2), 32);
            d.tmp_off = tmp_elems;
            tmp_elems += 2ull * d.tmp_elems;
            tcs.push_back(d);
            for (int l = 0; l < hdr.nlevels; l++)
                if (!j2k_stream_ok(d.w, d.h, l)) stream_levels &= ~(1u << l);
            if (!j2k_fused_ok(d.w, d.h)) fused_ok = false;
            if (iso && ((tc.x0 | tc.y0) & ((1u << hdr.nlevels) - 1)))
                J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "ISO mode: tile origin (%u,%u) is not a multiple of 2^nlevels", tc.x0, tc.y0);
            if (iso && (d.w & 3)) stream_levels = 0;          // Mallat rows must stay 8-byte aligned for the streaming kernel
            if (d.w > job->max_w) job->max_w = d.w;
            if (d.h > job->max_h) job->max_h = d.h;
            auto key = std::make_tuple(tc.x0, tc.y0, tc.x1, tc.y1);
            auto f = tile_of.find(key);
            uint32_t ti;
            if (f == tile_of.end()) {
                DevTile tl{};
                for (int c = 0; c < 4; c++) tl.tc[c] = 0xFFFFFFFFu;
                tl.img_x0 = tc.x0; tl.img_y0 = tc.y0; tl.w = d.w; tl.h = d.h;
                tl.out_off = out_bytes; tl.out_stride = it.out_stride; tl.img_w = im.width; tl.img_h = im.height;
                ti = (uint32_t)tiles.size();
                tiles.push_back(tl);
                tile_of[key] = ti;
            } else ti = f->second;
            if (tiles[ti].tc[tc.comp] != 0xFFFFFFFFu) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: duplicate tile-component", ii);
            tiles[ti].tc[tc.comp] = tc_base + t;
        }
        for (auto &kv : tile_of)
            for (int c = 0; c < im.ncomp; c++)
                if (tiles[kv.second].tc[c] == 0xFFFFFFFFu) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: a tile lacks component %d (subsampled components are not supported)", ii, c);
        // pixels no tile covers keep what the reference's zero-initialised planes decode to (decoder.go:305-309): if the
        // tiles, clipped to the image, do not cover it exactly once, the pixel buffer is pre-filled with that value
        rects.clear();
        for (auto &kv : tile_of) {
            const DevTile &t = tiles[kv.second];
            Rect q = {t.img_x0, t.img_y0, t.img_x0 + t.w < im.width ? t.img_x0 + t.w : im.width, t.img_y0 + t.h < im.height ? t.img_y0 + t.h : im.height};
            if (q.x0 < q.x1 && q.y0 < q.y1) rects.push_back(q);
        }
        // (an item whose copy-out is restricted to its tiles' rectangles never shows the uncovered pixels: no pre-fill)
        const bool fill = !(it.flags & J2KGPU_ITEM_TILES_ONLY) && !tiles_exactly(rects, im.width, im.height);
        job->item_fill.push_back(fill ? 1 : 0);
        if (fill) job->pix_fill = 1;

        blk_rects.assign(it.n_tilecomps, std::vector<Rect>());
        const size_t cb_first = cbs.size();
        for (uint32_t b = 0; b < it.n_cblks; b++) {
            const j2k_cblk_t &cb = it.cblks[b];
            if (cb.tilecomp >= it.n_tilecomps) J2K_FAIL(ctx, J2KGPU_E_RANGE, "item %u block %u: tile-component %u out of range", ii, b, cb.tilecomp);
            const DevTileComp &d = tcs[tc_base + cb.tilecomp];
            if (cb.w == 0 || cb.h == 0) continue;
            if (cb.w > 64 || cb.h > 64) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "item %u block %u: %ux%u exceeds 64x64", ii, b, cb.w, cb.h);
            if ((uint32_t)cb.x0 + cb.w > d.w || (uint32_t)cb.y0 + cb.h > d.h) J2K_FAIL(ctx, J2KGPU_E_RANGE, "item %u block %u: outside its tile-component", ii, b);
            if (cb.data_off > it.blob_len || cb.data_len > it.blob_len - cb.data_off) J2K_FAIL(ctx, J2KGPU_E_RANGE, "item %u block %u: data outside blob", ii, b);
            if (iso && (cb.num_bps < 1 || cb.num_bps > 30) && cb.data_len) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "item %u block %u: num_bps %d outside 1..30", ii, b, (int)cb.num_bps);
            if (iso && hdr.ht && cb.num_passes > 3) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "item %u block %u: %d HT coding passes (one HT set = cleanup, SigProp, MagRef)", ii, b, (int)cb.num_passes);
            if (cb.num_bps > 31 || cb.band > 3) J2K_FAIL(ctx, J2KGPU_E_UNSUPPORTED, "item %u block %u: num_bps %d / band %d", ii, b, (int)cb.num_bps, (int)cb.band);
            // int16 planes hold what coef_bits promises; an EBCOT block that claims more bit-planes than that would be truncated
            if (will_coef16 && !hdr.ht && cb.data_len && cb.num_bps > hdr.coef_bits)
                J2K_FAIL(ctx, J2KGPU_E_RANGE, "item %u block %u: num_bps %d exceeds the declared coef_bits %d", ii, b, (int)cb.num_bps, (int)hdr.coef_bits);
            if ((cstyle & (J2KGPU_CBLK_BYPASS | J2KGPU_CBLK_TERMALL)) && cb.data_len && cb.num_bps) {
                const uint32_t np = cb.num_passes ? cb.num_passes : 3u * cb.num_bps - 2u;
                if (4ull * t1_num_segments(cstyle, np) > it.blob_len - cb.data_off - cb.data_len)
                    J2K_FAIL(ctx, J2KGPU_E_RANGE, "item %u block %u: segment-length table outside blob", ii, b);
            }
            DevCblk o{};
            o.data_off = blob_bytes + cb.data_off; o.data_len = cb.data_len;
            o.out_off = d.coef_off + (uint64_t)cb.y0 * d.w + cb.x0; o.out_stride = d.w;
            o.w = cb.w; o.h = cb.h; o.band = cb.band; o.num_bps = cb.num_bps; o.level = cb.level; o.num_passes = cb.num_passes;
            o.len_cup = (cb.len_cleanup && cb.len_cleanup <= cb.data_len) ? cb.len_cleanup : cb.data_len;
            o.pad = cstyle;                                          // k_t1_iso reads the style per block
            cbs.push_back(o);
            steps.push_back(cb.step);
            blk_rects[cb.tilecomp].push_back(Rect{cb.x0, cb.y0, (uint32_t)cb.x0 + cb.w, (uint32_t)cb.y0 + cb.h});
            if (cb.data_len && cb.num_bps > max_bps) max_bps = cb.num_bps;
            if (iso && hdr.ht && cb.num_passes > 1) job->ht_refine = 1;
        }
        sort_blocks_by_shape(cbs.p + cb_first, steps.p + cb_first, cbs.size() - cb_first, shape_tmp, step_tmp);
        // a plane its blocks leave holes in is cleared before the entropy stage; blocks that overlap are refused (no
        // codestream has them, and the EBCOT kernels accumulate a block's bit-planes in place: its samples are its own)
        for (uint32_t t = 0; t < it.n_tilecomps; t++) {
            const int cover = rect_cover(blk_rects[t], tcs[tc_base + t].w, tcs[tc_base + t].h);
            if (cover == 2) J2K_FAIL(ctx, J2KGPU_E_ARG, "item %u: code blocks of tile-component %u overlap", ii, t);
            if (cover == 1) need_clear = true;
        }
        blob_bytes += it.blob_len;
        out_bytes = align_up(out_bytes + it.out_stride * im.height, 256);
    }
#undef J2K_FAIL
    job->item_cb.push_back((uint32_t)cbs.size());
    job->item_tc.push_back((uint32_t)tcs.size());
    job->item_tile.push_back((uint32_t)tiles.size());
    job->n_tc = (uint32_t)tcs.size(); job->n_tiles = (uint32_t)tiles.size(); job->n_cb = (uint32_t)cbs.size();
    // int16 coefficient planes when every magnitude provably fits: EBCOT magnitudes are < 2^num_bps; a conformant
    // codestream bounds them by coef_bits (Mb); the reference's HT coder has no bound (ht.go:664-684) and stays int32
    if (!opt.coef32) {
        if (!iso && !hdr.ht) job->coef16 = max_bps <= 15;
        if (iso && hdr.reversible) job->coef16 = hdr.coef_bits >= 1 && hdr.coef_bits <= 14;
    }
    job->fused_ok = fused_ok;
    // fixed epilogues of the fused kernels: 3 x 8-bit unsigned + RCT -> RGBA8, or 1 x 8 / 16-bit unsigned -> Gray8 / Gray16
    if (tp.ncomp == 3) {
        job->fast_epi = tp.fmt == J2KGPU_FMT_RGBA8 && tp.mct && tp.reversible && !tp.cconv;
        for (int c = 0; c < 3; c++) if (tp.prec[c] != 8 || tp.sgnd[c]) job->fast_epi = 0;
    } else if (tp.ncomp == 1) {
        job->fast_epi = (tp.prec[0] == 8 || tp.prec[0] == 16) && !tp.sgnd[0];
    }
    if (opt.no_fast_epi) job->fast_epi = 0;
    for (size_t i = 0; i < tiles.size(); i++) {
        const DevTile &t = tiles[i];
        if ((t.out_stride & 15) || (t.out_off & 15) || (t.img_x0 & 3) || t.img_x0 + t.w > t.img_w || t.img_y0 + t.h > t.img_h) job->fast_epi = 0;
    }
    job->wide_ok = job->fast_epi && fused_ok && (tp.ncomp == 3 || tp.ncomp == 1) && !opt.no_wide;
    for (size_t i = 0; i < tiles.size(); i++) if ((tiles[i].w & 15) || (tp.ncomp == 1 && (tiles[i].img_x0 & 15))) job->wide_ok = 0;   // 16-byte stores per lane
    // packed-RGB transfer (rgb_expand.h): the 16-columns-per-lane kernel writes the pixels of every tile, rows of 16 pixels are
    // 48 aligned bytes in the packed layout, and nothing else (pre-fill of uncovered pixels) writes the pixel buffer
    job->rgb24_ok = job->wide_ok && job->fused_ok && tp.ncomp == 3 && tp.fmt == J2KGPU_FMT_RGBA8 && !job->pix_fill;
    for (size_t i = 0; i < tiles.size(); i++) if ((tiles[i].img_x0 & 15) || (tiles[i].out_stride & 63)) job->rgb24_ok = 0;
    job->coef_elems = coef_elems; job->blob_bytes = blob_bytes; job->out_bytes = out_bytes; job->max_bps = max_bps;
    job->tmp_bytes = tmp_elems * ((hdr.reversible || iso) ? 4 : 8);     // int32 (5-3), float32 (ISO 9-7), float64 (REF 9-7)
    job->need_clear = need_clear;                        // some plane is not tiled exactly by its blocks
    job->stream_levels = stream_levels;

    const cudaStream_t up_stream = async_stream ? async_stream : ctx->stream;
    auto up = [&](void **dst, const void *src, size_t bytes) {
        if (e != cudaSuccess) return;
        *dst = j2k_pool_alloc(ctx, bytes, &e);
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, up_stream);
    };
    up((void **)&job->d_cblks, cbs.p, cbs.size() * sizeof(DevCblk));
    if (iso && !hdr.reversible) up((void **)&job->d_steps, steps.p, steps.size() * sizeof(float));
    up((void **)&job->d_tcs, tcs.p, tcs.size() * sizeof(DevTileComp));
    up((void **)&job->d_tiles, tiles.p, tiles.size() * sizeof(DevTile));
    if (e == cudaSuccess) job->d_coef = j2k_pool_alloc(ctx, coef_elems * (job->coef16 ? 2 : 4) + 64, &e);
    if (e == cudaSuccess) job->d_tmp = j2k_pool_alloc(ctx, job->tmp_bytes, &e);
    if (e == cudaSuccess && hdr.ht)
        job->d_htscratch = j2k_pool_alloc(ctx, iso ? j2k_htiso_scratch_bytes((uint32_t)cbs.size(), job->ht_refine) : j2k_htref_scratch_bytes((uint32_t)cbs.size()), &e);
    // reference HT coder: its decoder writes one row in four (ht.go:677, 701); zero the planes once, here, so that every
    // run only has to clear the rows it may write (3/4 of the entropy stage's zero-fill traffic saved)
    if (e == cudaSuccess && !iso && hdr.ht && !opt.no_preclear) {
        e = cudaMemsetAsync(job->d_coef, 0, coef_elems * (job->coef16 ? 2 : 4), ctx->stream);
        job->precleared = 1;
    }
    // the pixel that all-zero coefficients decode to (inverse MCT, DC shift, colour conversion, clamp, pack of zeros)
    if (e == cudaSuccess && job->pix_fill) {
        job->d_fillpix = j2k_pool_alloc(ctx, 64, &e);
        if (e == cudaSuccess) e = cudaMemsetAsync(job->d_fillpix, 0, 64, ctx->stream);
        if (e == cudaSuccess) {
            const int32_t *z = (const int32_t *)((uint8_t *)job->d_fillpix + 16);
            const int32_t *zc[4] = {z, z + 1, z + 2, z + 3};
            e = launch_tail(zc, nullptr, (uint8_t *)job->d_fillpix, 16, 1, 1, tp, 1, ctx->stream);
            ctx->launches++;
        }
    }
    if (e == cudaSuccess && !async_stream) {
        e = cudaStreamSynchronize(ctx->stream);              // the tables have left the staging block
        j2k_hpool_free(ctx, job->h_tables, job->h_tables_cap);
        job->h_tables = nullptr; job->h_tiles_host = nullptr;
    }
    if (e != cudaSuccess) { if (async_stream) cudaStreamSynchronize(async_stream); job_free(job); return j2k_cuda_err(ctx, e, "job upload"); }
    *out = job;
    return J2KGPU_OK;
}

extern "C" int j2kgpu_job_create(j2kgpu_ctx *ctx, uint32_t n_img, const j2k_batch_item_t *items, j2kgpu_job **out)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    return job_build(ctx, n_img, items, out, nullptr);
}

extern "C" uint64_t j2kgpu_job_blob_bytes(const j2kgpu_job *job) { return job ? job->blob_bytes : 0; }
extern "C" uint64_t j2kgpu_job_out_bytes(const j2kgpu_job *job) { return job ? job->out_bytes : 0; }
extern "C" uint64_t j2kgpu_job_out_offset(const j2kgpu_job *job, uint32_t item)
{
    return (job && item < job->n_img) ? job->out_off[item] : 0;
}

// ---- launch sequences over the items [ia, ib) of a job, on stream `st` ------------------------------------------------
static int run_entropy(j2kgpu_job *job, const void *d_blob, uint32_t ia, uint32_t ib, cudaStream_t st)
{
    j2kgpu_ctx *ctx = job->ctx;
    const size_t esz = job->coef16 ? 2 : 4;
    if (job->need_clear) {
        const uint32_t ta = job->item_tc[ia], tb = job->item_tc[ib];
        if (tb > ta) {
            const uint64_t ea = job->tc_coef_off[ta], eb = tb < job->n_tc ? job->tc_coef_off[tb] : job->coef_elems;
            J2K_CUDA(ctx, cudaMemsetAsync((uint8_t *)job->d_coef + ea * esz, 0, (eb - ea) * esz, st));
        }
    }
    const uint32_t ca = job->item_cb[ia], n = job->item_cb[ib] - ca;
    if (n == 0) return J2KGPU_OK;
    const DevCblk *cbs = job->d_cblks + ca;
    cudaError_t e;
    const int irrev = job->iso && !job->hdr.reversible;                  // ISO 9-7: the planes receive dequantised float32
    const float *steps = job->d_steps ? job->d_steps + ca : nullptr;
    if (job->iso && !job->hdr.ht) e = launch_t1_iso(cbs, n, (const uint8_t *)d_blob, job->d_coef, job->coef16, steps, irrev, job->t1_segmented, ctx->opt.t1_group, st);
    else if (job->iso) {
        // chunks of a pipelined run share the scratch: their kernels are ordered on one stream
        e = launch_ht_iso(cbs, n, (const uint8_t *)d_blob, job->d_coef, job->coef16, steps, irrev, job->hdr.coef_bits,
                          job->ht_refine, job->d_htscratch, job->blob_bytes, st);
        ctx->launches += j2k_htiso_launches(job->ht_refine) - 1;
    }
    else if (job->hdr.ht) {
        e = launch_ht_ref(cbs, n, (const uint8_t *)d_blob, job->d_coef, job->coef16, job->precleared,
                          job->d_htscratch, job->blob_bytes, st);
        ctx->launches += j2k_htref_launches() - 1;
    }
    else e = launch_t1_ref(cbs, n, (const uint8_t *)d_blob, job->d_coef, job->coef16, job->max_bps, ctx->opt.t1_group, st);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "entropy kernel launch");
    ctx->launches++;
    return J2KGPU_OK;
}

static void fill_launch(const j2kgpu_job *job, IdwtLaunch &p, void *d_out, uint32_t ia, uint32_t ib)
{
    p = IdwtLaunch{};
    p.d_tcs = job->d_tcs; p.d_tiles = job->d_tiles;
    p.tc_first = job->item_tc[ia]; p.n_tc = job->item_tc[ib] - p.tc_first;
    p.tile_first = job->item_tile[ia]; p.n_tiles = job->item_tile[ib] - p.tile_first;
    p.d_coef = job->d_coef; p.coef16 = job->coef16; p.d_tmp = job->d_tmp; p.nlevels = job->nlevels; p.fast_epi = job->fast_epi; p.wide_ok = job->wide_ok;
    p.max_w = job->max_w; p.max_h = job->max_h; p.reversible = job->hdr.reversible != 0; p.f64_io = 0;
    p.d_plane_out = nullptr; p.d_pix = (uint8_t *)d_out; p.tail = job->tail; p.stream_levels = job->stream_levels; p.iso = job->iso;
    p.wide_sp = job->ctx->opt.wide_sp;
    p.rgb24 = job->rgb24;
}

// one level of the reconstruction; with the fused kernel, level 1 is part of the level-0 launch
static int run_level(j2kgpu_job *job, const IdwtLaunch &base, int lvl, cudaStream_t st)
{
    j2kgpu_ctx *ctx = job->ctx;
    IdwtLaunch q = base;
    q.lvl = lvl;
    if (job->fused_ok && lvl <= 1) {
        if (lvl == 1) return J2KGPU_OK;
        cudaError_t e = launch_idwt53_fused(q, st);
        if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "fused idwt launch");
        ctx->launches++;
        return J2KGPU_OK;
    }
    if (lvl > 0) q.d_tiles = nullptr;
    int nl = 0;
    cudaError_t e = launch_idwt_level(q, st, &nl);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "idwt level launch");
    ctx->launches += nl;
    return J2KGPU_OK;
}

static int run_dwt_mct(j2kgpu_job *job, void *d_out, uint32_t ia, uint32_t ib, cudaStream_t st)
{
    j2kgpu_ctx *ctx = job->ctx;
    if (job->pix_fill)
        for (uint32_t i = ia; i < ib; i++)
            if (job->item_fill[i]) {
                cudaError_t e = launch_fill_pixels((uint8_t *)d_out + job->out_off[i], job->out_stride[i], job->row_bytes[i], job->img_h[i],
                                                   j2k_fmt_bpp(job->tail.fmt), (const uint8_t *)job->d_fillpix, st);
                if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "pixel fill launch");
                ctx->launches++;
            }
    IdwtLaunch p;
    fill_launch(job, p, d_out, ia, ib);
    if (p.n_tiles == 0) return J2KGPU_OK;
    for (int lvl = job->nlevels ? job->nlevels - 1 : 0; lvl >= 0; lvl--) {
        int rc = run_level(job, p, lvl, st);
        if (rc) return rc;
    }
    return J2KGPU_OK;
}

extern "C" int j2kgpu_job_run_entropy(j2kgpu_job *job, const void *d_blob)
{
    if (!job || (!d_blob && job->blob_bytes)) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(job->ctx->mu);
    cudaSetDevice(job->ctx->device);
    return run_entropy(job, d_blob, 0, job->n_img, job->ctx->stream);
}

extern "C" int j2kgpu_job_run_dwt_mct(j2kgpu_job *job, void *d_out)
{
    if (!job || !d_out) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(job->ctx->mu);
    cudaSetDevice(job->ctx->device);
    return run_dwt_mct(job, d_out, 0, job->n_img, job->ctx->stream);
}

extern "C" int j2kgpu_job_run_level(j2kgpu_job *job, int lvl, void *d_out)
{
    if (!job || lvl < 0 || lvl >= (job->nlevels ? job->nlevels : 1) || (lvl == 0 && !d_out)) return J2KGPU_E_ARG;
    j2kgpu_ctx *ctx = job->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    IdwtLaunch p;
    fill_launch(job, p, d_out, 0, job->n_img);
    return run_level(job, p, lvl, ctx->stream);
}

extern "C" int j2kgpu_job_fused_levels(const j2kgpu_job *job) { return job ? (job->fused_ok ? 2 : 1) : 0; }
extern "C" int j2kgpu_job_coef_bytes(const j2kgpu_job *job) { return job ? (job->coef16 ? 2 : 4) : 0; }
extern "C" int j2kgpu_job_plan(const j2kgpu_job *job)
{
    if (!job) return 0;
    return (job->fused_ok ? J2KGPU_PLAN_FUSED : 0) | (job->fused_ok && job->fast_epi ? J2KGPU_PLAN_FAST_EPILOGUE : 0) |
           (job->fused_ok && job->wide_ok ? J2KGPU_PLAN_WIDE : 0) | (job->coef16 ? J2KGPU_PLAN_COEF16 : 0);
}

extern "C" int j2kgpu_job_run(j2kgpu_job *job, const void *d_blob, void *d_out)
{
    if (!job || !d_out || (!d_blob && job->blob_bytes)) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(job->ctx->mu);
    cudaSetDevice(job->ctx->device);
    int rc = run_entropy(job, d_blob, 0, job->n_img, job->ctx->stream);
    if (rc) return rc;
    return run_dwt_mct(job, d_out, 0, job->n_img, job->ctx->stream);
}

// Chunk plan of the host-buffer run.  The device->host copy of the pixels is the bottleneck of the path (4 bytes per
// pixel out against about 2 in), and on a Gen5 x16 link under bidirectional load the host->device direction is not
// faster per frame than the device->host one (measured: 53 GB/s out, about 35 GB/s in, tools/pcie_probe.py), so chunks
// cannot grow along the batch without starving the copy-out engine.  The plan is therefore: the smallest possible first
// chunk (its copy-in + decode is the only exposed latency), then equal chunks, as many as the decoder's per-launch
// latency floor (one code block's serial chain per launch sequence) allows inside the copy-out time; compute-bound
// coders (EBCOT) get four chunks.  Measured on the bench batch (16 x 4K, reference HT coder): 16 chunks of one frame 11.16 ms,
// (1,3,3,3,3,3) 11.72 ms, (4,4,4,4) 12.26 ms, one chunk 16.6 ms; conformant HTJ2K with the round-2 kernels: 16 x 1 11.17 ms,
// (1,1,2,2,...) 11.28, (1,2,2,...,3) 11.44, 8 x 2 11.69, (1,3,3,3,3,3) 12.02, (1,3,4,4,4) 12.28.
static std::vector<uint32_t> plan_chunks(const j2kgpu_ctx *ctx, uint32_t n, const j2k_image_t &hdr, uint64_t out_bytes)
{
    if (!ctx->opt.chunks.empty()) {                      // experiments: explicit chunk sizes, e.g. "1,1,2,4"; the rest in one chunk
        const char *e = ctx->opt.chunks.c_str();
        std::vector<uint32_t> cuts(1, 0u);
        while (*e && cuts.back() < n) {
            const uint32_t k = (uint32_t)strtoul(e, (char **)&e, 10);
            if (k == 0) break;
            cuts.push_back(cuts.back() + k < n ? cuts.back() + k : n);
            if (*e == ',') e++;
        }
        if (cuts.back() < n) cuts.push_back(n);
        return cuts;
    }
    double lat, r_cmp;                                   // seconds per launch sequence, output bytes per second
    if (!hdr.ht) { lat = 5e-3; r_cmp = 2.4e9; }                         // EBCOT / MQ
    else if (hdr.mode == J2KGPU_MODE_ISO) { lat = 0.4e-3; r_cmp = 400e9; }     // round-2 kernels: one frame per chunk pays
    else { lat = 0.4e-3; r_cmp = 800e9; }
    const double t_out = (double)out_bytes / 53e9, t_thr = (double)out_bytes / r_cmp;
    uint32_t nchunks = t_thr > 0.5 * t_out ? 4u : (uint32_t)((0.8 * t_out - t_thr) / lat);
    if (nchunks > 16) nchunks = 16;
    if (nchunks > n) nchunks = n;
    if (nchunks < 1) nchunks = 1;
    std::vector<uint32_t> cuts(1, 0u);
    if (nchunks >= 3 && t_thr <= 0.5 * t_out) {          // small first chunk, the rest in nchunks - 1 equal parts
        const uint32_t k0 = (n + 15) / 16;
        cuts.push_back(k0);
        const uint32_t rest = n - k0, parts = nchunks - 1;
        for (uint32_t c = 1; c <= parts; c++) {
            const uint32_t end = k0 + (uint32_t)(((uint64_t)rest * c) / parts);
            if (end > cuts.back()) cuts.push_back(end);
        }
    } else {
        for (uint32_t c = 1; c <= nchunks; c++) {
            const uint32_t end = (uint32_t)(((uint64_t)n * c) / nchunks);
            if (end > cuts.back()) cuts.push_back(end);
        }
    }
    if (ctx->opt.debug_plan) { fprintf(stderr, "j2kgpu chunk plan:"); for (uint32_t c : cuts) fprintf(stderr, " %u", c); fprintf(stderr, "\n"); }
    return cuts;
}

// ---- packed-RGB transfer of host-buffer runs (host/rgb_expand.h) ---------------------------------------------------------
static bool rgb24_wanted(const j2kgpu_ctx *ctx)
{
    if (ctx->opt.host_alpha >= 0) return ctx->opt.host_alpha != 0;
    int ndev = 1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
    // the widening has to store 4 bytes per pixel faster than the link delivers 3: measured on a 16-core B200 host, eight
    // worker threads widen the bench batch (531 MB of RGBA per step) in 17 ms against 11 ms for the plain RGBA transfer,
    // so the automatic choice asks for 32 hardware threads per GPU
    return std::thread::hardware_concurrency() / (unsigned)ndev >= 32;
}

// decides whether this host-buffer run of `job` moves packed RGB, and prepares the staging block and the worker threads
static int rgb24_begin(j2kgpu_job *job, const j2k_batch_item_t *items, uint32_t n)
{
    j2kgpu_ctx *ctx = job->ctx;
    job->rgb24 = 0;
    job->rgb_tasks.clear();
    if (!job->rgb24_ok || !rgb24_wanted(ctx)) return J2KGPU_OK;
    for (uint32_t i = 0; i < n; i++) if (items[i].flags & J2KGPU_ITEM_TILES_ONLY) return J2KGPU_OK;
    uint64_t tot = 0;
    job->rgb_off.resize(n);
    for (uint32_t i = 0; i < n; i++) { job->rgb_off[i] = tot; tot = align_up(tot + (uint64_t)job->img_h[i] * (job->out_stride[i] / 4 * 3), 256); }
    if (!job->h_rgb || job->h_rgb_cap < tot) {
        if (job->h_rgb) j2k_hpool_free(ctx, job->h_rgb, job->h_rgb_cap);
        cudaError_t e = cudaSuccess;
        job->h_rgb = j2k_hpool_alloc(ctx, tot, &job->h_rgb_cap, &e);
        if (e != cudaSuccess) { job->h_rgb = nullptr; return j2k_cuda_err(ctx, e, "packed-RGB staging"); }
    }
    if (!ctx->expand) ctx->expand = new (std::nothrow) J2kExpandPool();
    if (!ctx->expand) return J2KGPU_OK;                  // (no pool: plain RGBA transfer)
    int ndev = 1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency()) / (unsigned)ndev;
    ctx->expand->start(std::min(8u, std::max(2u, hw / 2)));
    job->rgb24 = 1;
    if (ctx->opt.debug_plan) fprintf(stderr, "j2kgpu packed-RGB transfer: %u image(s), %u host thread(s)\n", n, ctx->expand->threads());
    return J2KGPU_OK;
}

static void CUDART_CB rgb24_callback(void *ud)
{
    J2kRgbCb *cb = static_cast<J2kRgbCb *>(ud);
    cb->pool->submit(cb->t);
}

// after the copy-out stream has drained: every row has been widened; the job returns to plain RGBA
static void rgb24_end(j2kgpu_job *job)
{
    if (!job->rgb24) return;
    job->ctx->expand->wait_idle();
    job->rgb24 = 0;
    job->rgb_tasks.clear();
}

// pixels of item i (job numbering) back to the caller.  Rows go without their padding (the bytes between width * bpp and
// out_stride belong to the caller and are left alone).  J2KGPU_ITEM_TILES_ONLY: only the rectangles this item's tiles
// cover are copied, so that several contexts (GPUs) given disjoint tile subsets of ONE image fill one host buffer.
static int copy_out_item(j2kgpu_job *job, uint32_t i, const j2k_batch_item_t &it, const DevTile *h_tiles, cudaStream_t st)
{
    j2kgpu_ctx *ctx = job->ctx;
    const uint8_t *src = (const uint8_t *)job->d_pix + job->out_off[i];
    const uint64_t stride = job->out_stride[i];
    const uint32_t rb = job->row_bytes[i];
    if (job->rgb24) {                                    // packed rows into the staging block, widened by the pool once they are there
        const uint64_t s3 = stride / 4 * 3;
        uint8_t *stage = (uint8_t *)job->h_rgb + job->rgb_off[i];
        J2K_CUDA(ctx, cudaMemcpyAsync(stage, src, (size_t)job->img_h[i] * s3, cudaMemcpyDeviceToHost, st));
        job->rgb_tasks.push_back(J2kRgbCb{ctx->expand, J2kExpandTask{stage, it.out_pix, s3, it.out_stride, rb / 4, job->img_h[i]}});
        J2K_CUDA(ctx, cudaLaunchHostFunc(st, rgb24_callback, &job->rgb_tasks.back()));
        return J2KGPU_OK;
    }
    if (!(it.flags & J2KGPU_ITEM_TILES_ONLY)) {
        if (stride == rb) J2K_CUDA(ctx, cudaMemcpyAsync(it.out_pix, src, job->out_size[i], cudaMemcpyDeviceToHost, st));
        else J2K_CUDA(ctx, cudaMemcpy2DAsync(it.out_pix, stride, src, stride, rb, job->img_h[i], cudaMemcpyDeviceToHost, st));
        return J2KGPU_OK;
    }
    const int bpp = j2k_fmt_bpp(job->tail.fmt);
    const uint32_t width = rb / (uint32_t)bpp;
    // horizontally adjacent tiles with the same rows are merged; a run of full-width rows is one contiguous copy
    std::vector<Rect> rs;
    for (uint32_t t = job->item_tile[i]; t < job->item_tile[i + 1]; t++) {
        const DevTile &tl = h_tiles[t];
        Rect q = {tl.img_x0, tl.img_y0, tl.img_x0 + tl.w < width ? tl.img_x0 + tl.w : width,
                  tl.img_y0 + tl.h < job->img_h[i] ? tl.img_y0 + tl.h : job->img_h[i]};
        if (q.x0 < q.x1 && q.y0 < q.y1) rs.push_back(q);
    }
    std::sort(rs.begin(), rs.end(), [](const Rect &a, const Rect &b) { return a.y0 != b.y0 ? a.y0 < b.y0 : a.x0 < b.x0; });
    for (size_t k = 0; k < rs.size();) {
        Rect q = rs[k++];
        while (k < rs.size() && rs[k].y0 == q.y0 && rs[k].y1 == q.y1 && rs[k].x0 == q.x1) q.x1 = rs[k++].x1;
        const size_t off = (size_t)q.y0 * stride + (size_t)q.x0 * bpp;
        if (q.x0 == 0 && q.x1 == width && stride == rb)
            J2K_CUDA(ctx, cudaMemcpyAsync(it.out_pix + off, src + off, (size_t)(q.y1 - q.y0) * stride, cudaMemcpyDeviceToHost, st));
        else
            J2K_CUDA(ctx, cudaMemcpy2DAsync(it.out_pix + off, stride, src + off, stride, (size_t)(q.x1 - q.x0) * bpp, q.y1 - q.y0,
                                            cudaMemcpyDeviceToHost, st));
    }
    return J2KGPU_OK;
}

// Host-buffer run of a prepared job.  The batch is cut into chunks of whole items; chunk c's host->device copy runs on
// the copy-in stream, its kernels on the ctx stream and its device->host copy on the copy-out stream, chained by events,
// so that the PCIe transfers of neighbouring chunks overlap the kernels (the two copy engines work in both directions
// at once).  A single-item batch degenerates to copy, compute, copy.
static int run_host_body(j2kgpu_job *job, const j2k_batch_item_t *items)
{
    j2kgpu_ctx *ctx = job->ctx;
    cudaError_t pe = cudaSuccess;
    if (!job->d_blob) { job->d_blob = j2k_pool_alloc(ctx, job->blob_bytes + 64, &pe); if (pe != cudaSuccess) return j2k_cuda_err(ctx, pe, "blob staging"); }
    if (!job->d_pix) { job->d_pix = j2k_pool_alloc(ctx, job->out_bytes, &pe); if (pe != cudaSuccess) return j2k_cuda_err(ctx, pe, "pixel staging"); }
    std::vector<DevTile> h_tiles;
    for (uint32_t i = 0; i < job->n_img; i++) {
        if (!items[i].out_pix) return j2k_set_err(ctx, J2KGPU_E_ARG, "item %u: null out_pix", i);
        if ((items[i].flags & J2KGPU_ITEM_TILES_ONLY) && h_tiles.empty()) {
            h_tiles.resize(job->n_tiles);
            J2K_CUDA(ctx, cudaMemcpy(h_tiles.data(), job->d_tiles, job->n_tiles * sizeof(DevTile), cudaMemcpyDeviceToHost));
        }
    }
    int rc = j2k_ctx_copy_streams(ctx);
    if (rc) return rc;
    const std::vector<uint32_t> cuts = plan_chunks(ctx, job->n_img, job->hdr, job->out_bytes);   // chunk c = items [cuts[c], cuts[c + 1])
    const uint32_t nchunk = (uint32_t)cuts.size() - 1;
    while (job->ev_in.size() < nchunk) {
        cudaError_t e1, e2;
        job->ev_in.push_back(ctx_event(ctx, &e1));
        job->ev_done.push_back(ctx_event(ctx, &e2));
        if (e1 != cudaSuccess || e2 != cudaSuccess) return j2k_cuda_err(ctx, e1 != cudaSuccess ? e1 : e2, "event");
    }
    // the copy streams start after whatever the caller queued on the ctx stream
    J2K_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    J2K_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
    J2K_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
    for (uint32_t c = 0; c < nchunk; c++) {
        const uint32_t ia = cuts[c], ib = cuts[c + 1];
        for (uint32_t i = ia; i < ib; i++)
            if (items[i].blob_len)
                J2K_CUDA(ctx, cudaMemcpyAsync((uint8_t *)job->d_blob + job->blob_off[i], items[i].blob, items[i].blob_len,
                                              cudaMemcpyHostToDevice, ctx->s_in));
        J2K_CUDA(ctx, cudaEventRecord(job->ev_in[c], ctx->s_in));
        J2K_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, job->ev_in[c], 0));
        rc = run_entropy(job, job->d_blob, ia, ib, ctx->stream);
        if (rc) return rc;
        rc = run_dwt_mct(job, job->d_pix, ia, ib, ctx->stream);
        if (rc) return rc;
        J2K_CUDA(ctx, cudaEventRecord(job->ev_done[c], ctx->stream));
        J2K_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, job->ev_done[c], 0));
        for (uint32_t i = ia; i < ib; i++)
            if ((rc = copy_out_item(job, i, items[i], h_tiles.data(), ctx->s_out))) return rc;
    }
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

static int run_host_locked(j2kgpu_job *job, const j2k_batch_item_t *items)
{
    j2kgpu_ctx *ctx = job->ctx;
    cudaSetDevice(ctx->device);
    int rc = rgb24_begin(job, items, job->n_img);
    if (rc == J2KGPU_OK) rc = run_host_body(job, items);
    if (job->rgb24) {                                    // also after an error: no callback may outlive the run
        if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
        rgb24_end(job);
    }
    return rc;
}

extern "C" int j2kgpu_job_run_host(j2kgpu_job *job, const j2k_batch_item_t *items)
{
    if (!job || !items) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(job->ctx->mu);
    return run_host_locked(job, items);
}

// ---- page-locked host memory -------------------------------------------------------------------------------
extern "C" int j2kgpu_host_alloc(j2kgpu_ctx *ctx, uint64_t bytes, void **out)
{
    if (!ctx || !out) return J2KGPU_E_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    J2K_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_host_free(j2kgpu_ctx *ctx, void *p)
{
    if (!ctx) return J2KGPU_E_ARG;
    if (!p) return J2KGPU_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    J2K_CUDA(ctx, cudaFreeHost(p));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_host_register(j2kgpu_ctx *ctx, void *p, uint64_t bytes)
{
    if (!ctx || !p || !bytes) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    J2K_CUDA(ctx, cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_host_unregister(j2kgpu_ctx *ctx, void *p)
{
    if (!ctx || !p) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    J2K_CUDA(ctx, cudaHostUnregister(p));
    return J2KGPU_OK;
}

// The plugin call.  The batch is cut into chunks of whole images and every chunk is a job of its own: its tables are
// validated and flattened on the host into page-locked staging and uploaded on the copy-in stream together with its
// compressed bytes, its kernels run on the ctx stream, its pixels leave on the copy-out stream.  Nothing synchronises
// until the last chunk is queued, so the host builds chunk c + 1 while the device decodes chunk c and the link carries
// chunk c - 1 out: table building, both PCIe directions and the SMs all overlap.
struct BatchPipe {
    j2kgpu_ctx *ctx = nullptr;
    std::vector<j2kgpu_job *> jobs;
    std::vector<DevTile> h_tiles;
    int rc = J2KGPU_OK;
};

static int pipe_begin(BatchPipe &bp, j2kgpu_ctx *ctx)
{
    bp.ctx = ctx;
    cudaSetDevice(ctx->device);
    int rc = j2k_ctx_copy_streams(ctx);
    if (rc) return bp.rc = rc;
    cudaError_t ce = cudaEventRecord(ctx->ev_start, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0);
    if (ce != cudaSuccess) bp.rc = j2k_cuda_err(ctx, ce, "stream setup");
    return bp.rc;
}

// queue one chunk: tables + compressed bytes in, kernels, pixels out; returns without waiting for any of it
static int pipe_submit(BatchPipe &bp, const j2k_batch_item_t *its, uint32_t n)
{
    if (bp.rc || n == 0) return bp.rc;
    j2kgpu_ctx *ctx = bp.ctx;
    j2kgpu_job *job = nullptr;
    if ((bp.rc = job_build(ctx, n, its, &job, ctx->s_in))) return bp.rc;
    bp.jobs.push_back(job);
    if ((bp.rc = rgb24_begin(job, its, n))) return bp.rc;
    cudaError_t ce = cudaSuccess;
    job->d_blob = j2k_pool_alloc(ctx, job->blob_bytes + 64, &ce);
    if (ce == cudaSuccess) job->d_pix = j2k_pool_alloc(ctx, job->out_bytes, &ce);
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    if (ce == cudaSuccess) { ev_in = ctx_event(ctx, &ce); job->ev_in.push_back(ev_in); }
    if (ce == cudaSuccess) { ev_done = ctx_event(ctx, &ce); job->ev_done.push_back(ev_done); }
    for (uint32_t i = 0; i < n && ce == cudaSuccess; i++)
        if (its[i].blob_len)
            ce = cudaMemcpyAsync((uint8_t *)job->d_blob + job->blob_off[i], its[i].blob, its[i].blob_len, cudaMemcpyHostToDevice, ctx->s_in);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev_in, ctx->s_in);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, ev_in, 0);
    if (ce != cudaSuccess) return bp.rc = j2k_cuda_err(ctx, ce, "chunk upload");
    if ((bp.rc = run_entropy(job, job->d_blob, 0, n, ctx->stream))) return bp.rc;
    if ((bp.rc = run_dwt_mct(job, job->d_pix, 0, n, ctx->stream))) return bp.rc;
    ce = cudaEventRecord(ev_done, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->s_out, ev_done, 0);
    if (ce != cudaSuccess) return bp.rc = j2k_cuda_err(ctx, ce, "chunk events");
    bool subset = false;
    for (uint32_t i = 0; i < n; i++) subset |= (its[i].flags & J2KGPU_ITEM_TILES_ONLY) != 0;
    if (subset) bp.h_tiles.assign(job->h_tiles_host, job->h_tiles_host + job->n_tiles);   // still in the page-locked staging block
    for (uint32_t i = 0; i < n && bp.rc == J2KGPU_OK; i++)
        bp.rc = copy_out_item(job, i, its[i], bp.h_tiles.data(), ctx->s_out);
    return bp.rc;
}

// everything queued (or failed): drain, then give the chunks' buffers back to the pools
static int pipe_finish(BatchPipe &bp)
{
    j2kgpu_ctx *ctx = bp.ctx;
    if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
    cudaError_t e1 = cudaStreamSynchronize(ctx->stream), e2 = ctx->s_out ? cudaStreamSynchronize(ctx->s_out) : cudaSuccess;
    for (j2kgpu_job *j : bp.jobs) { rgb24_end(j); job_free(j); }
    bp.jobs.clear();
    if (bp.rc == J2KGPU_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) bp.rc = j2k_cuda_err(ctx, e1 != cudaSuccess ? e1 : e2, "decode_batch");
    return bp.rc;
}

// One image of many tiles as several items of one batch: groups of tiles (in the order of the table) whose copy-in, decode
// and copy-out then overlap like the images of a batch do (every group's item names the same out_pix and carries
// J2KGPU_ITEM_TILES_ONLY; its blob is the byte range its blocks lie in).  Only when the tiles cover the image exactly
// (otherwise the uncovered pixels need the whole-image pre-fill) and the image is large enough for the split to pay.
struct SplitItems {
    std::vector<j2k_batch_item_t> items;
    std::vector<std::vector<j2k_tilecomp_t>> tcs;
    std::vector<std::vector<j2k_cblk_t>> cbs;
};

static bool split_by_tiles(const j2k_batch_item_t &it, SplitItems &sp, uint64_t j2k_split_min_pixels)
{
    const j2k_image_t &im = it.image;
    // every group is a launch sequence of its own, and an entropy launch costs its longest block chain (0.25 ms) however few
    // blocks it has: measured, one 4K frame 1.46 ms in one piece against 2.89 ms in eight groups; one 8192 x 8192 image 6.2 ms
    // against 3.6 ms.  The split starts at 24 Mpixel.
    if ((it.flags & J2KGPU_ITEM_TILES_ONLY) || (uint64_t)im.width * im.height < j2k_split_min_pixels || it.n_tilecomps < 4u * im.ncomp || !im.ncomp) return false;
    // tiles = distinct bounds, numbered in order of first appearance
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>, uint32_t> tile_of;
    std::vector<uint32_t> tc_tile(it.n_tilecomps);
    std::vector<Rect> rects;
    for (uint32_t t = 0; t < it.n_tilecomps; t++) {
        const j2k_tilecomp_t &tc = it.tilecomps[t];
        if (tc.x1 <= tc.x0 || tc.y1 <= tc.y0 || tc.x0 >= im.width || tc.y0 >= im.height) return false;
        auto key = std::make_tuple(tc.x0, tc.y0, tc.x1, tc.y1);
        auto f = tile_of.find(key);
        if (f == tile_of.end()) {
            f = tile_of.emplace(key, (uint32_t)tile_of.size()).first;
            rects.push_back(Rect{tc.x0, tc.y0, tc.x1 < im.width ? tc.x1 : im.width, tc.y1 < im.height ? tc.y1 : im.height});
        }
        tc_tile[t] = f->second;
    }
    const uint32_t ntiles = (uint32_t)tile_of.size();
    if (ntiles < 4 || !tiles_exactly(rects, im.width, im.height)) return false;
    const uint32_t K = ntiles < 8 ? ntiles : 8;
    sp.items.assign(K, it);
    sp.tcs.assign(K, {});
    sp.cbs.assign(K, {});
    std::vector<uint32_t> tc_new(it.n_tilecomps);
    for (uint32_t t = 0; t < it.n_tilecomps; t++) {
        const uint32_t g = (uint32_t)((uint64_t)tc_tile[t] * K / ntiles);
        tc_new[t] = (uint32_t)sp.tcs[g].size();
        sp.tcs[g].push_back(it.tilecomps[t]);
    }
    std::vector<uint64_t> lo(K, ~0ull), hi(K, 0);
    for (uint32_t b = 0; b < it.n_cblks; b++) {
        const j2k_cblk_t &cb = it.cblks[b];
        if (cb.tilecomp >= it.n_tilecomps) return false;                 // (the plain path reports it)
        const uint32_t g = (uint32_t)((uint64_t)tc_tile[cb.tilecomp] * K / ntiles);
        j2k_cblk_t c2 = cb;
        c2.tilecomp = tc_new[cb.tilecomp];
        sp.cbs[g].push_back(c2);
        if (cb.data_len) {
            if (cb.data_off > it.blob_len || cb.data_len > it.blob_len - cb.data_off) return false;
            lo[g] = std::min(lo[g], cb.data_off); hi[g] = std::max(hi[g], cb.data_off + cb.data_len);
        }
    }
    for (uint32_t g = 0; g < K; g++) {
        j2k_batch_item_t &o = sp.items[g];
        if (lo[g] > hi[g]) { lo[g] = hi[g] = 0; }
        for (j2k_cblk_t &c : sp.cbs[g]) c.data_off = c.data_len ? c.data_off - lo[g] : 0;
        o.tilecomps = sp.tcs[g].data(); o.n_tilecomps = (uint32_t)sp.tcs[g].size();
        o.cblks = sp.cbs[g].data(); o.n_cblks = (uint32_t)sp.cbs[g].size();
        o.blob = it.blob ? it.blob + lo[g] : nullptr; o.blob_len = hi[g] - lo[g];
        o.flags |= J2KGPU_ITEM_TILES_ONLY;
    }
    return true;
}

extern "C" int j2kgpu_decode_batch(j2kgpu_ctx *ctx, uint32_t n_img, const j2k_batch_item_t *items)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!items || n_img == 0) return j2k_set_err(ctx, J2KGPU_E_ARG, "null or empty batch");
    TailParams tp;
    int rc = make_tail(ctx, items[0].image, tp);
    if (rc) return rc;
    SplitItems sp;                                       // one large image of many tiles: its groups of tiles are the batch
    if (n_img == 1 && items[0].out_pix && (items[0].tilecomps || !items[0].n_tilecomps) && (items[0].cblks || !items[0].n_cblks) &&
        ctx->opt.chunks.empty() && split_by_tiles(items[0], sp, ctx->opt.split_min_mpixel > 0 ? (uint64_t)ctx->opt.split_min_mpixel << 20 : 24ull << 20)) {
        items = sp.items.data();
        n_img = (uint32_t)sp.items.size();
    }
    uint64_t out_bytes = 0;
    for (uint32_t i = 0; i < n_img; i++) {
        if (!items[i].out_pix) return j2k_set_err(ctx, J2KGPU_E_ARG, "item %u: null out_pix", i);
        if (!same_header(items[i].image, items[0].image)) return j2k_set_err(ctx, J2KGPU_E_ARG, "item %u: header differs from item 0", i);
        out_bytes += items[i].out_stride * items[i].image.height;
    }
    const std::vector<uint32_t> cuts = plan_chunks(ctx, n_img, items[0].image, out_bytes);
    BatchPipe bp;
    pipe_begin(bp, ctx);
    for (size_t c = 0; c + 1 < cuts.size() && bp.rc == J2KGPU_OK; c++) pipe_submit(bp, items + cuts[c], cuts[c + 1] - cuts[c]);
    return pipe_finish(bp);
}

// ---- codestream front door: tier-2 on the host (host/tier2.cpp), then the same pipeline --------------------------------
extern "C" int j2kgpu_parse_codestream(const uint8_t *cs, uint64_t len, uint32_t reduce, uint32_t threads, j2kgpu_parsed **out)
{
    if (!out) return J2KGPU_E_ARG;
    j2kgpu_parsed *p = new (std::nothrow) j2kgpu_parsed();
    *out = p;
    if (!p) return J2KGPU_E_NOMEM;
    return j2k_tier2_parse(cs, len, reduce, threads, *p);
}

extern "C" void j2kgpu_parsed_free(j2kgpu_parsed *p) { delete p; }
extern "C" const char *j2kgpu_parsed_error(const j2kgpu_parsed *p) { return p ? p->err.c_str() : "null"; }

extern "C" int j2kgpu_parsed_item(const j2kgpu_parsed *p, j2k_batch_item_t *item)
{
    if (!p || !item) return J2KGPU_E_ARG;
    memset(item, 0, sizeof *item);
    item->image = p->image;
    item->tilecomps = p->tilecomps.data(); item->n_tilecomps = (uint32_t)p->tilecomps.size();
    item->cblks = p->cblks.data(); item->n_cblks = (uint32_t)p->cblks.size();
    item->blob = p->blob; item->blob_len = p->blob_len;
    return J2KGPU_OK;
}

extern "C" int j2kgpu_parsed_info(const j2kgpu_parsed *p, uint32_t info[8])
{
    if (!p || !info) return J2KGPU_E_ARG;
    info[0] = p->layers; info[1] = p->tiles; info[2] = p->tile_parts; info[3] = p->packets; info[4] = p->progression;
    info[5] = p->plt_packets; info[6] = p->tlm_tile_parts; info[7] = p->owned.empty() ? 1u : 0u;   // 1: blocks point into the caller's bytes
    return J2KGPU_OK;
}

// n codestreams -> n images.  Worker threads run the tier-2 of the frames (one frame per thread at a time) while this
// thread submits finished frames, in order and in chunks, to the copy / kernel / copy pipeline: host parsing of frame
// i + k overlaps the device's work on frame i.  All frames must share the header fields j2kgpu_decode_batch requires.
extern "C" int j2kgpu_decode_codestreams(j2kgpu_ctx *ctx, uint32_t n, const uint8_t *const *cs, const uint64_t *lens, uint32_t reduce,
                                         uint8_t *const *out_pix, const uint64_t *out_stride)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!n || !cs || !lens || !out_pix || !out_stride) return j2k_set_err(ctx, J2KGPU_E_ARG, "null or empty batch");
    // Host side: the main header and tile-part index of every frame are read here (microseconds each); the tiles of all
    // frames then form one task list in frame order, worked off by one set of threads, so the first frame is ready after
    // 1 / threads of its tier-2 time and the frames finish in the order the device pipeline wants them.  The thread that
    // parses a frame's last tile merges its tables.
    std::vector<j2kgpu_parsed> parsed(n);
    std::vector<int> prc(n, 0);
    std::vector<std::atomic<int>> done(n);
    for (auto &d : done) d.store(0);
    std::vector<j2k_t2_frame *> frames(n, nullptr);
    std::vector<uint32_t> first_task(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        prc[i] = j2k_tier2_begin(cs[i], lens[i], reduce, &frames[i], parsed[i].err);
        first_task[i + 1] = first_task[i] + (prc[i] ? 0 : j2k_tier2_tiles(frames[i]));
        if (prc[i] || j2k_tier2_tiles(frames[i]) == 0) {  // failed, or nothing to parse: finished as it is
            if (!prc[i]) prc[i] = j2k_tier2_finish(frames[i], parsed[i]);
            frames[i] = nullptr;
            done[i].store(1);
        }
    }
    const uint32_t ntasks = first_task[n];
    std::vector<std::atomic<uint32_t>> left(n);
    for (uint32_t i = 0; i < n; i++) left[i].store(first_task[i + 1] - first_task[i]);
    std::atomic<uint32_t> next{0};
    std::mutex mu;
    std::condition_variable cv;
    const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t nthreads = std::max(1u, std::min<uint32_t>({hw, ntasks, 32u}));
    auto worker = [&]() {
        uint32_t f = 0;
        for (;;) {
            const uint32_t k = next.fetch_add(1);
            if (k >= ntasks) break;
            while (k >= first_task[f + 1]) f++;
            j2k_tier2_tile(frames[f], k - first_task[f]);
            if (left[f].fetch_sub(1) == 1) {             // the frame's last tile: merge, publish
                prc[f] = j2k_tier2_finish(frames[f], parsed[f]);
                { std::lock_guard<std::mutex> lk(mu); done[f].store(1); }
                cv.notify_all();
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nthreads && ntasks; t++) th.emplace_back(worker);
    auto wait_for = [&](uint32_t i) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done[i].load() != 0; }); };
    BatchPipe bp;
    int rc = J2KGPU_OK;
    std::vector<j2k_batch_item_t> items(n);
    std::vector<uint32_t> cuts;
    uint32_t submitted = 0, ci = 1;
    for (uint32_t i = 0; i < n && rc == J2KGPU_OK; i++) {
        wait_for(i);
        if (prc[i]) { rc = j2k_set_err(ctx, prc[i], "codestream %u: %s", i, parsed[i].err.c_str()); break; }
        j2kgpu_parsed_item(&parsed[i], &items[i]);
        items[i].out_pix = out_pix[i]; items[i].out_stride = out_stride[i];
        if (!out_pix[i]) { rc = j2k_set_err(ctx, J2KGPU_E_ARG, "codestream %u: null out_pix", i); break; }
        if (i == 0) {                                    // chunk plan from the first frame (frames of a batch are alike)
            cuts = plan_chunks(ctx, n, items[0].image, (uint64_t)n * out_stride[0] * items[0].image.height);
            rc = pipe_begin(bp, ctx);
        }
        if (rc == J2KGPU_OK && i + 1 == cuts[ci]) {
            rc = pipe_submit(bp, items.data() + submitted, i + 1 - submitted);
            submitted = i + 1; ci++;
        }
    }
    for (auto &t : th) t.join();                         // (after an error the remaining tiles are still parsed: their frames are freed by finish)
    if (bp.ctx) { bp.rc = bp.rc ? bp.rc : rc; return pipe_finish(bp); }
    return rc;
}

extern "C" int j2kgpu_decode_codestream(j2kgpu_ctx *ctx, const uint8_t *cs, uint64_t len, uint32_t reduce, uint8_t *out_pix, uint64_t out_stride)
{
    return j2kgpu_decode_codestreams(ctx, 1, &cs, &len, reduce, &out_pix, &out_stride);
}

extern "C" int j2kgpu_decode(j2kgpu_ctx *ctx, const j2k_image_t *img,
                             const j2k_tilecomp_t *tilecomps, uint32_t n_tilecomps,
                             const j2k_cblk_t *cblks, uint32_t n_cblks,
                             const uint8_t *blob, uint64_t blob_len,
                             uint8_t *out_pix, uint64_t out_stride)
{
    if (!ctx) return J2KGPU_E_ARG;
    if (!img || !out_pix) return j2k_set_err(ctx, J2KGPU_E_ARG, "null image or output");
    j2k_batch_item_t it;
    memset(&it, 0, sizeof it);
    it.image = *img; it.tilecomps = tilecomps; it.n_tilecomps = n_tilecomps; it.cblks = cblks; it.n_cblks = n_cblks;
    it.blob = blob; it.blob_len = blob_len; it.out_pix = out_pix; it.out_stride = out_stride;
    return j2kgpu_decode_batch(ctx, 1, &it);
}

// ---- per-stage entry points -------------------------------------------------------------------------------
static int stage_blocks(j2kgpu_ctx *ctx, int mode, int ht, const j2k_blkjob_t *jobs, uint32_t n,
                        const uint8_t *blob, uint64_t blob_len, int32_t *out, uint64_t out_len)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (mode != J2KGPU_MODE_REF && mode != J2KGPU_MODE_ISO) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "mode %d not built for this coder", mode);
    if ((!jobs && n) || (!blob && blob_len) || (!out && out_len)) return j2k_set_err(ctx, J2KGPU_E_ARG, "null argument");
    if (n == 0) return J2KGPU_OK;
    cudaSetDevice(ctx->device);
    std::vector<DevCblk> cbs(n);
    int max_bps = 1, segmented = 0;
    for (uint32_t i = 0; i < n; i++) {
        const j2k_blkjob_t &j = jobs[i];
        if (j.w == 0 || j.h == 0 || j.w > 64 || j.h > 64) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "block %u: size %ux%u", i, j.w, j.h);
        if (j.data_off > blob_len || j.data_len > blob_len - j.data_off) return j2k_set_err(ctx, J2KGPU_E_RANGE, "block %u: data outside blob", i);
        if ((uint64_t)j.out_off + (uint64_t)j.w * j.h > out_len) return j2k_set_err(ctx, J2KGPU_E_RANGE, "block %u: output outside buffer", i);
        if (j.num_bps > 31 || j.band > 3) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "block %u: num_bps %d / band %d", i, (int)j.num_bps, (int)j.band);
        DevCblk &o = cbs[i];
        memset(&o, 0, sizeof o);
        o.data_off = j.data_off; o.data_len = j.data_len; o.out_off = j.out_off; o.out_stride = j.w;
        o.w = j.w; o.h = j.h; o.band = j.band; o.num_bps = j.num_bps; o.num_passes = j.rsv0;    // ISO: coding passes (0 = all)
        o.len_cup = (j.len_cleanup && j.len_cleanup <= j.data_len) ? j.len_cleanup : j.data_len;
        if (mode == J2KGPU_MODE_ISO && !ht) {
            if (j.rsv1 & ~0x3Fu) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "block %u: code-block style %02X", i, (unsigned)j.rsv1);
            if ((j.rsv1 & (J2KGPU_CBLK_BYPASS | J2KGPU_CBLK_TERMALL)) && j.data_len && j.num_bps) {
                const uint32_t np = j.rsv0 ? j.rsv0 : 3u * j.num_bps - 2u;
                if (4ull * t1_num_segments(j.rsv1, np) > blob_len - j.data_off - j.data_len)
                    return j2k_set_err(ctx, J2KGPU_E_RANGE, "block %u: segment-length table outside blob", i);
            }
            if (j.rsv1) segmented = 1;
            o.pad = j.rsv1;
        }
        if (j.num_bps > max_bps) max_bps = j.num_bps;
    }
    int rc;
    if ((rc = j2k_reserve(ctx, ctx->d_tab, n * sizeof(DevCblk), false))) return rc;
    if ((rc = j2k_reserve(ctx, ctx->d_in, blob_len + 16, false))) return rc;
    if ((rc = j2k_reserve(ctx, ctx->d_out, out_len * sizeof(int32_t) + 16, false))) return rc;
    J2K_CUDA(ctx, cudaMemcpyAsync(ctx->d_tab.p, cbs.data(), n * sizeof(DevCblk), cudaMemcpyHostToDevice, ctx->stream));
    if (blob_len) J2K_CUDA(ctx, cudaMemcpyAsync(ctx->d_in.p, blob, blob_len, cudaMemcpyHostToDevice, ctx->stream));
    J2K_CUDA(ctx, cudaMemsetAsync(ctx->d_out.p, 0, out_len * sizeof(int32_t), ctx->stream));
    if (ht && mode != J2KGPU_MODE_ISO) { if ((rc = j2k_reserve(ctx, ctx->d_aux, j2k_htref_scratch_bytes(n), false))) return rc; ctx->launches += j2k_htref_launches() - 1; }
    int refine = 0;
    if (ht && mode == J2KGPU_MODE_ISO) {
        for (uint32_t i = 0; i < n; i++) {
            if (cbs[i].num_passes > 3) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "block %u: %d HT coding passes", i, (int)cbs[i].num_passes);
            if (cbs[i].num_passes > 1) refine = 1;
        }
        if ((rc = j2k_reserve(ctx, ctx->d_aux, j2k_htiso_scratch_bytes(n, refine), false))) return rc;
        ctx->launches += j2k_htiso_launches(refine) - 1;
    }
    cudaError_t e = (ht && mode == J2KGPU_MODE_ISO)
                        ? launch_ht_iso((const DevCblk *)ctx->d_tab.p, n, (const uint8_t *)ctx->d_in.p, ctx->d_out.p, 0, nullptr, 0, 0, refine, ctx->d_aux.p, blob_len, ctx->stream)
                    : mode == J2KGPU_MODE_ISO ? launch_t1_iso((const DevCblk *)ctx->d_tab.p, n, (const uint8_t *)ctx->d_in.p, ctx->d_out.p, 0, nullptr, 0, segmented, ctx->opt.t1_group, ctx->stream)
                    : ht ? launch_ht_ref((const DevCblk *)ctx->d_tab.p, n, (const uint8_t *)ctx->d_in.p, ctx->d_out.p, 0, 0, ctx->d_aux.p, blob_len, ctx->stream)
                       : launch_t1_ref_stage((const DevCblk *)ctx->d_tab.p, n, (const uint8_t *)ctx->d_in.p, (int32_t *)ctx->d_out.p, max_bps, ctx->opt.t1_group, ctx->stream);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "entropy kernel launch");
    ctx->launches++;
    J2K_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out.p, out_len * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_t1_decode_blocks(j2kgpu_ctx *ctx, int mode, const j2k_blkjob_t *jobs, uint32_t n,
                                       const uint8_t *blob, uint64_t blob_len, int32_t *out, uint64_t out_len)
{
    return stage_blocks(ctx, mode, 0, jobs, n, blob, blob_len, out, out_len);
}

extern "C" int j2kgpu_ht_decode_blocks(j2kgpu_ctx *ctx, int mode, const j2k_blkjob_t *jobs, uint32_t n,
                                       const uint8_t *blob, uint64_t blob_len, int32_t *out, uint64_t out_len)
{
    return stage_blocks(ctx, mode, 1, jobs, n, blob, blob_len, out, out_len);
}

// one plane through every level; kind 0: int32 5-3, 1: float64 9-7, 2: int32 -> 9-7 -> int32(v+0.5)
static int stage_idwt(j2kgpu_ctx *ctx, int mode, void *data, uint32_t w, uint32_t h, uint32_t levels, int kind)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (mode != J2KGPU_MODE_REF && !(mode == J2KGPU_MODE_ISO && kind == 0))
        return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "mode %d not built for this transform", mode);
    if (!data && w && h) return j2k_set_err(ctx, J2KGPU_E_ARG, "null data");
    if (w == 0 || h == 0) return J2KGPU_OK;
    if (levels == 0 && kind != 2) return J2KGPU_OK;                           // dwt.go:545: no levels, nothing to do
    if (levels > J2K_MAX_LEVELS) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "levels %u > %d", levels, J2K_MAX_LEVELS);
    if ((uint64_t)w * h > (1ull << 31)) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "plane too large");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)w * h;
    const size_t in_el = kind == 1 ? 8 : 4, out_el = kind == 1 ? 8 : 4, tmp_el = kind == 0 ? 4 : 8;
    DevTileComp tc{};
    tc.w = w; tc.h = h; tc.coef_off = 0; tc.tmp_off = 0;
    tc.tmp_elems = (uint32_t)align_up((uint64_t)((w + 1) / 2) * ((h + 1) / 2), 32);
    int rc;
    if ((rc = j2k_reserve(ctx, ctx->d_tab, sizeof tc, false))) return rc;
    if ((rc = j2k_reserve(ctx, ctx->d_in, n * in_el, false))) return rc;
    if ((rc = j2k_reserve(ctx, ctx->d_out, n * out_el, false))) return rc;
    if ((rc = j2k_reserve(ctx, ctx->d_aux, 2ull * tc.tmp_elems * tmp_el, false))) return rc;
    J2K_CUDA(ctx, cudaMemcpyAsync(ctx->d_tab.p, &tc, sizeof tc, cudaMemcpyHostToDevice, ctx->stream));
    J2K_CUDA(ctx, cudaMemcpyAsync(ctx->d_in.p, data, n * in_el, cudaMemcpyHostToDevice, ctx->stream));
    IdwtLaunch p{};
    p.d_tcs = (const DevTileComp *)ctx->d_tab.p; p.n_tc = 1; p.d_tiles = nullptr; p.n_tiles = 0;
    p.d_coef = (const int32_t *)ctx->d_in.p; p.d_tmp = ctx->d_aux.p; p.nlevels = (int)levels;
    p.max_w = w; p.max_h = h; p.reversible = kind == 0; p.f64_io = kind == 1;
    p.d_plane_out = (int32_t *)ctx->d_out.p; p.d_pix = nullptr;
    p.stream_levels = 0;
    p.iso = mode == J2KGPU_MODE_ISO;
    if (kind == 0 && !(p.iso && (w & 3)))
        for (int l = 1; l < (int)levels; l++)                 // level 0 of the stage API stores planes: tiled kernel
            if (j2k_stream_ok(w, h, l)) p.stream_levels |= 1u << l;
    // levels == 0 (kind 2 only): tcd.go:428-435 still runs int32 -> float64 -> int32(v + 0.5), which maps a
    // negative integer n to n + 1 (truncation); one pass-through launch reproduces it
    for (int lvl = levels ? (int)levels - 1 : 0; lvl >= 0; lvl--) {
        p.lvl = lvl;
        int nl = 0;
        cudaError_t e = launch_idwt_level(p, ctx->stream, &nl);
        if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "idwt level launch");
        ctx->launches += nl;
    }
    J2K_CUDA(ctx, cudaMemcpyAsync(data, ctx->d_out.p, n * out_el, cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_idwt53(j2kgpu_ctx *ctx, int mode, int32_t *data, uint32_t width, uint32_t height, uint32_t levels)
{
    return stage_idwt(ctx, mode, data, width, height, levels, 0);
}

extern "C" int j2kgpu_idwt97(j2kgpu_ctx *ctx, int mode, double *data, uint32_t width, uint32_t height, uint32_t levels)
{
    return stage_idwt(ctx, mode, data, width, height, levels, 1);
}

extern "C" int j2kgpu_apply_inverse_dwt(j2kgpu_ctx *ctx, int mode, int32_t *data, uint32_t width, uint32_t height,
                                        uint32_t levels, int reversible)
{
    return stage_idwt(ctx, mode, data, width, height, levels, reversible ? 0 : 2);
}

static int stage_tail(j2kgpu_ctx *ctx, const TailParams &tp, const int32_t *const *comps, int32_t *const *planes_out,
                      uint32_t width, uint32_t height, int apply_tail, uint8_t *out_pix, uint64_t out_stride)
{
    const uint64_t n = (uint64_t)width * height;
    if (n == 0) return J2KGPU_OK;
    cudaSetDevice(ctx->device);
    const int nc = tp.ncomp;
    int rc;
    if ((rc = j2k_reserve(ctx, ctx->d_in, n * 4 * nc, false))) return rc;
    const int32_t *d_c[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t *d_o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < nc; c++) {
        if (!comps[c]) return j2k_set_err(ctx, J2KGPU_E_ARG, "null component %d", c);
        int32_t *p = (int32_t *)ctx->d_in.p + (size_t)c * n;
        J2K_CUDA(ctx, cudaMemcpyAsync(p, comps[c], n * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_c[c] = p;
    }
    for (int c = nc; c < 4; c++) d_c[c] = d_c[0];
    if (planes_out) {
        if ((rc = j2k_reserve(ctx, ctx->d_aux, n * 4 * nc, false))) return rc;
        for (int c = 0; c < 4; c++) d_o[c] = (int32_t *)ctx->d_aux.p + (size_t)(c < nc ? c : 0) * n;
    }
    uint8_t *d_pix = nullptr;
    if (out_pix) {
        if ((rc = j2k_reserve(ctx, ctx->d_out, out_stride * height, false))) return rc;
        d_pix = (uint8_t *)ctx->d_out.p;
    }
    cudaError_t e = launch_tail(d_c, planes_out ? d_o : nullptr, d_pix, out_stride, width, height, tp, apply_tail, ctx->stream);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "tail kernel launch");
    ctx->launches++;
    if (planes_out)
        for (int c = 0; c < nc; c++)
            J2K_CUDA(ctx, cudaMemcpyAsync(planes_out[c], d_o[c], n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_pix) J2K_CUDA(ctx, cudaMemcpyAsync(out_pix, d_pix, out_stride * height, cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_inverse_rct(j2kgpu_ctx *ctx, int32_t *y, int32_t *u, int32_t *v, uint64_t n)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return J2KGPU_OK;
    if (!y || !u || !v) return j2k_set_err(ctx, J2KGPU_E_ARG, "null plane");
    if (n > 0xFFFFFFFFull) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "plane too large");
    TailParams tp{};
    tp.ncomp = 3; tp.mct = 1; tp.reversible = 1; tp.fmt = J2KGPU_FMT_RGBA8;
    for (int c = 0; c < 4; c++) { tp.prec[c] = 8; tp.sgnd[c] = 1; }            // signed: no DC shift
    const int32_t *in[4] = {y, u, v, nullptr};
    int32_t *out[4] = {y, u, v, nullptr};
    return stage_tail(ctx, tp, in, out, (uint32_t)n, 1, 1, nullptr, 0);
}

extern "C" int j2kgpu_inverse_ict(j2kgpu_ctx *ctx, double *y, double *cb, double *cr, uint64_t n)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return J2KGPU_OK;
    if (!y || !cb || !cr) return j2k_set_err(ctx, J2KGPU_E_ARG, "null plane");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = j2k_reserve(ctx, ctx->d_in, n * 8 * 3, false))) return rc;
    double *d = (double *)ctx->d_in.p;
    J2K_CUDA(ctx, cudaMemcpyAsync(d, y, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    J2K_CUDA(ctx, cudaMemcpyAsync(d + n, cb, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    J2K_CUDA(ctx, cudaMemcpyAsync(d + 2 * n, cr, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = launch_inverse_ict_f64(d, d + n, d + 2 * n, n, ctx->stream);
    if (e != cudaSuccess) return j2k_cuda_err(ctx, e, "ict kernel launch");
    ctx->launches++;
    J2K_CUDA(ctx, cudaMemcpyAsync(y, d, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaMemcpyAsync(cb, d + n, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaMemcpyAsync(cr, d + 2 * n, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    J2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return J2KGPU_OK;
}

extern "C" int j2kgpu_dc_level_shift_inverse(j2kgpu_ctx *ctx, int32_t *data, uint64_t n, int precision)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return J2KGPU_OK;
    if (!data) return j2k_set_err(ctx, J2KGPU_E_ARG, "null data");
    if (precision < 1 || precision > 31) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "precision %d", precision);
    if (n > 0xFFFFFFFFull) return j2k_set_err(ctx, J2KGPU_E_UNSUPPORTED, "plane too large");
    TailParams tp{};
    tp.ncomp = 1; tp.mct = 0; tp.reversible = 1; tp.fmt = J2KGPU_FMT_GRAY8;
    for (int c = 0; c < 4; c++) { tp.prec[c] = precision; tp.sgnd[c] = c != 0; }
    const int32_t *in[4] = {data, nullptr, nullptr, nullptr};
    int32_t *out[4] = {data, nullptr, nullptr, nullptr};
    return stage_tail(ctx, tp, in, out, (uint32_t)n, 1, 1, nullptr, 0);
}

extern "C" int j2kgpu_mct_dc_pack(j2kgpu_ctx *ctx, const j2k_image_t *img, const int32_t *const *comps,
                                  int apply_tail, uint8_t *out_pix, uint64_t out_stride)
{
    if (!ctx) return J2KGPU_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!img || !comps || !out_pix) return j2k_set_err(ctx, J2KGPU_E_ARG, "null argument");
    TailParams tp;
    int rc = make_tail(ctx, *img, tp);
    if (rc) return rc;
    const int bpp = j2k_fmt_bpp(tp.fmt);
    if (out_stride < (uint64_t)img->width * bpp || out_stride % bpp) return j2k_set_err(ctx, J2KGPU_E_ARG, "bad out_stride");
    return stage_tail(ctx, tp, comps, nullptr, img->width, img->height, apply_tail, out_pix, out_stride);
}
