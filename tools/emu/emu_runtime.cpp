// emu_runtime.cpp -- fiber scheduler of the CPU emulation shim (development tool, see include/cuda_runtime.h).
// One CTA at a time; each CUDA thread is a ucontext fiber; fibers yield at block / warp synchronisation points.
#include "cuda_runtime.h"

#include <ucontext.h>
#include <stdio.h>
#include <vector>

namespace emu {

ThreadCtx *cur = nullptr;
uint3 block_idx;
dim3 block_dim, grid_dim;
unsigned char *dyn_smem = nullptr;

namespace {

constexpr size_t kStack = 256 * 1024;

struct Barrier {
    int expected = 0, count = 0;
    unsigned gen = 0;
};

struct Fiber {
    ucontext_t uc;
    ThreadCtx tc;
    bool done = false;
    unsigned char *stack = nullptr;
};

struct WarpState {
    Barrier bar;
    uint64_t xchg[2][32];
    unsigned mask[2];
    bool live[32];
};

ucontext_t sched_uc;
std::vector<Fiber> fibers;
std::vector<WarpState> warps;
std::vector<unsigned> lane_parity;
Barrier block_bar;
const std::function<void()> *body_fn = nullptr;
Fiber *cur_fiber = nullptr;

void yield() { swapcontext(&cur_fiber->uc, &sched_uc); }

long events = 0;                  // barrier arrivals + thread exits: a scheduler round without any is a deadlock

void arrive(Barrier &b)
{
    events++;
    b.count++;
    if (b.count >= b.expected) { b.count = 0; b.gen++; return; }
    const unsigned g = b.gen;
    while (b.gen == g) yield();
}

void retire(Barrier &b)           // a thread that exits no longer takes part
{
    b.expected--;
    if (b.expected > 0 && b.count >= b.expected) { b.count = 0; b.gen++; }
}

void fiber_main()
{
    (*body_fn)();
    Fiber *f = cur_fiber;
    f->done = true;
    events++;
    warps[f->tc.warp].live[f->tc.lane] = false;
    retire(warps[f->tc.warp].bar);
    retire(block_bar);
    swapcontext(&f->uc, &sched_uc);
}

}  // namespace

void sync_block() { arrive(block_bar); }
void sync_warp() { arrive(warps[cur->warp].bar); }

uint64_t shfl_exchange(uint64_t v, int src_lane)
{
    WarpState &w = warps[cur->warp];
    unsigned &par = lane_parity[cur->lin];
    const unsigned p = par & 1;
    par++;
    w.xchg[p][cur->lane] = v;
    arrive(w.bar);
    if (src_lane < 0 || src_lane > 31) return v;
    return w.xchg[p][src_lane];       // still valid if the source lane has exited since
}

unsigned ballot(int pred)
{
    WarpState &w = warps[cur->warp];
    unsigned &par = lane_parity[cur->lin];
    const unsigned p = par & 1;
    par++;
    if (w.bar.count == 0) w.mask[p] = 0;          // first lane to arrive in this round
    if (pred) w.mask[p] |= 1u << cur->lane;
    arrive(w.bar);
    return w.mask[p];
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    grid_dim = grid;
    block_dim = block;
    body_fn = &body;
    std::vector<unsigned char> smem(smem_bytes + 256);
    dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    fibers.assign(nthreads, Fiber());
    for (auto &f : fibers) f.stack = (unsigned char *)malloc(kStack);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                block_idx.x = bx; block_idx.y = by; block_idx.z = bz;
                memset(dyn_smem, 0xCD, smem_bytes);
                warps.assign(nwarps, WarpState());
                lane_parity.assign(nthreads, 0);
                block_bar = Barrier();
                block_bar.expected = nthreads;
                for (int t = 0; t < nthreads; t++) {
                    Fiber &f = fibers[t];
                    f.done = false;
                    f.tc.lin = t;
                    f.tc.tid.x = t % block.x;
                    f.tc.tid.y = (t / block.x) % block.y;
                    f.tc.tid.z = t / (block.x * block.y);
                    f.tc.warp = t / 32;
                    f.tc.lane = t % 32;
                    warps[f.tc.warp].live[f.tc.lane] = true;
                    warps[f.tc.warp].bar.expected++;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = f.stack;
                    f.uc.uc_stack.ss_size = kStack;
                    f.uc.uc_link = &sched_uc;
                    makecontext(&f.uc, (void (*)())fiber_main, 0);
                }
                int live = nthreads;
                while (live > 0) {
                    const long ev0 = events;
                    for (int t = 0; t < nthreads; t++) {
                        Fiber &f = fibers[t];
                        if (f.done) continue;
                        cur_fiber = &f;
                        cur = &f.tc;
                        swapcontext(&sched_uc, &f.uc);
                        if (f.done) live--;
                    }
                    if (live > 0 && events == ev0) {
                        fprintf(stderr, "emu: CTA (%u,%u,%u) appears deadlocked at a barrier\n", bx, by, bz);
                        abort();
                    }
                }
            }
    for (auto &f : fibers) free(f.stack);
    fibers.clear();
    cur = nullptr;
    dyn_smem = nullptr;
}

}  // namespace emu
