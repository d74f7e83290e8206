// emu_sched.cpp -- TEST INFRASTRUCTURE: the fiber scheduler behind tests/emu/cudaemu/cuda_runtime.h.
// One CUDA block at a time; each of its threads is a fiber with its own stack; a fiber runs until it reaches a warp
// collective, a block barrier, a spin-wait or the end of the kernel.  A warp collective completes when every lane of
// the warp that has not returned is waiting at it (full-mask, convergent collectives -- the only kind this repo uses);
// a block barrier when every thread that has not returned is waiting at it.  Anything else is reported as a deadlock.
#include "cuda_runtime.h"

#include <sys/mman.h>
#include <vector>

#if !defined(__x86_64__)
#error "cudaemu's context switch is written for x86-64"
#endif

extern "C" void emu_ctx_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".globl emu_ctx_switch\n"
    ".type emu_ctx_switch,@function\n"
    "emu_ctx_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size emu_ctx_switch,.-emu_ctx_switch\n");

namespace emu {

enum { READY = 0, WAIT_WARP, WAIT_BAR, YIELDED, DONE };
static const size_t kStack = 256 * 1024;
static const int kMaxThreads = 1024;

struct WarpX : Warp { uint32_t part[2]; };

struct Sched {
    char*  stacks = nullptr;            // kMaxThreads stacks, mapped once per host thread (pages are touched lazily)
    void*  sp = nullptr;                // the scheduler's own saved stack pointer
    std::vector<Fiber> fibers;
    std::vector<WarpX> warps;
    void (*entry)(void*) = nullptr;
    void*  arg = nullptr;
};

thread_local Fiber* cur = nullptr;
thread_local uint3  block_idx;
thread_local dim3   block_dim, grid_dim;
thread_local void*  dyn_smem = nullptr;
static thread_local Sched* S = nullptr;
static thread_local uint64_t st_coll = 0, st_bar = 0, st_warp_max = 0;

static void to_scheduler() { Fiber* f = cur; emu_ctx_switch(&f->sp, S->sp); }

static void fiber_main()
{
    S->entry(S->arg);
    cur->state = DONE;
    to_scheduler();
    abort();                            // a finished fiber is never resumed
}

const uint64_t* warp_exchange(uint64_t v, uint32_t* mask)
{
    Fiber* f = cur;
    WarpX* w = static_cast<WarpX*>(f->warp);
    const uint32_t g = w->gen;
    w->val[g & 1][f->lane] = v;
    w->arrived |= 1u << f->lane;
    f->state = WAIT_WARP;
    to_scheduler();
    *mask = w->part[g & 1];
    return w->val[g & 1];
}

void block_barrier() { ++st_bar; cur->state = WAIT_BAR; to_scheduler(); }
void yield() { cur->state = YIELDED; to_scheduler(); }

static void run_block(unsigned nthreads)
{
    Sched* s = S;
    const unsigned nwarps = (nthreads + 31) / 32;
    s->fibers.assign(nthreads, Fiber());
    s->warps.assign(nwarps, WarpX());
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = s->fibers[t];
        f.tid.x = t % block_dim.x; f.tid.y = (t / block_dim.x) % block_dim.y; f.tid.z = t / (block_dim.x * block_dim.y);
        f.lane = (int)(t & 31);
        f.state = READY;
        f.warp = &s->warps[t >> 5];
        s->warps[t >> 5].alive |= 1u << (t & 31);
        // initial frame: six callee-saved registers (zero), the entry point, and a slot standing for its return address
        char* top = s->stacks + (size_t)(t + 1) * kStack;
        void** p = reinterpret_cast<void**>(top - 16);
        p[0] = reinterpret_cast<void*>(&fiber_main);
        p[1] = nullptr;
        for (int r = 1; r <= 6; ++r) p[-r] = nullptr;
        f.sp = p - 6;
    }
    unsigned remaining = nthreads;
    static const int order = [] { const char* e = getenv("NB_EMU_ORDER"); return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strcmp(e, "random") ? 2 : 0; }();
    static thread_local uint64_t rnd = 0x9E3779B97F4A7C15ull;
    while (remaining) {
        bool progress = false;
        for (Fiber& f : s->fibers) if (f.state == YIELDED) { f.state = READY; progress = true; }
        const unsigned wrot = order == 2 ? (unsigned)((rnd >> 40) % nwarps) : 0;
        for (unsigned wj = 0; wj < nwarps; ++wj) {
            const unsigned wi = order == 1 ? nwarps - 1 - wj : (wj + wrot) % nwarps;    // the warps' turns are permuted too
            WarpX& w = s->warps[wi];
            for (;;) {
                bool ran = false;
                const unsigned t0 = wi * 32, t1 = t0 + 32 < nthreads ? t0 + 32 : nthreads;
                for (unsigned u = t0; u < t1; ++u) {
                    // NB_EMU_ORDER: the order in which the runnable lanes of a warp get their turn.  Results must not
                    // depend on it; if they do, two lanes communicate without a __syncwarp() / collective between them.
                    unsigned t = u;
                    if (order == 1) t = t0 + (t1 - 1 - u);
                    else if (order == 2) {          // a fresh rotation and direction for every pass: still a permutation
                        if (u == t0) rnd = rnd * 6364136223846793005ull + 1442695040888963407ull;
                        const unsigned n = t1 - t0, r = (unsigned)((rnd >> 33) % n), k = u - t0;
                        t = t0 + (((rnd >> 20) & 1) ? (r + k) % n : (r + n - k) % n);
                    }
                    Fiber& f = s->fibers[t];
                    if (f.state != READY) continue;
                    cur = &f;
                    emu_ctx_switch(&s->sp, f.sp);
                    ran = true; progress = true;
                    if (f.state == DONE) { --remaining; w.alive &= ~(1u << f.lane); }
                }
                if (w.alive && w.arrived == w.alive) {          // every live lane is at the collective: it completes
                    w.part[w.gen & 1] = w.arrived;
                    w.arrived = 0;
                    w.gen++;
                    for (unsigned t = t0; t < t1; ++t) if (s->fibers[t].state == WAIT_WARP) s->fibers[t].state = READY;
                    continue;
                }
                if (!ran) break;
                bool any = false;
                for (unsigned t = t0; t < t1; ++t) any |= s->fibers[t].state == READY;
                if (!any) break;
            }
        }
        if (remaining) {
            bool all_bar = true, any_bar = false, any_yield = false;
            for (Fiber& f : s->fibers) {
                if (f.state == DONE) continue;
                if (f.state == WAIT_BAR) any_bar = true; else all_bar = false;
                if (f.state == YIELDED) any_yield = true;
            }
            if (all_bar && any_bar) { for (Fiber& f : s->fibers) if (f.state == WAIT_BAR) f.state = READY; progress = true; }
            else if (!progress && !any_yield) {
                fprintf(stderr, "cudaemu: deadlock in block (%u,%u,%u): threads wait at different synchronisation points\n", block_idx.x, block_idx.y, block_idx.z);
                for (unsigned t = 0; t < nthreads; ++t) if (s->fibers[t].state != DONE) fprintf(stderr, "  thread %u: %s\n", t, s->fibers[t].state == WAIT_WARP ? "warp collective" : s->fibers[t].state == WAIT_BAR ? "__syncthreads" : "runnable");
                abort();
            }
        }
    }
    for (unsigned wi = 0; wi < nwarps; ++wi) { st_coll += s->warps[wi].gen; if (s->warps[wi].gen > st_warp_max) st_warp_max = s->warps[wi].gen; }
    cur = nullptr;
}

// NB_EMU_STATS=1: per launch, the number of warp collectives and block barriers, in total and the most any single warp
// executed -- on the GPU the latter is a proxy for the length of the dependent chain that bounds a latency-bound kernel

void run(const char* name, dim3 grid, dim3 block, size_t smem, void (*entry)(void*), void* arg)
{
    static const bool stats = getenv("NB_EMU_STATS") != nullptr;
    st_coll = st_bar = st_warp_max = 0;
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > (unsigned)kMaxThreads) { fprintf(stderr, "cudaemu: %u threads per block\n", nthreads); abort(); }
    if (cur) { fprintf(stderr, "cudaemu: kernel launched from device code\n"); abort(); }
    if (!S) {
        S = new Sched();
        S->stacks = static_cast<char*>(mmap(nullptr, (size_t)kMaxThreads * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (S->stacks == MAP_FAILED) { perror("cudaemu: mmap"); abort(); }
    }
    S->entry = entry; S->arg = arg;
    block_dim = block; grid_dim = grid;
    void* sm = smem ? aligned_alloc(128, (smem + 127) & ~(size_t)127) : nullptr;
    if (sm) memset(sm, 0xA5, smem);
    dyn_smem = sm;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                block_idx.x = x; block_idx.y = y; block_idx.z = z;
                run_block(nthreads);
            }
    dyn_smem = nullptr;
    free(sm);
    if (stats)
        fprintf(stderr, "[cudaemu] %-34s grid %5u x %4u threads: %10llu warp collectives (longest warp %9llu), %8llu barrier waits\n", name,
                grid.x * grid.y * grid.z, nthreads, (unsigned long long)st_coll, (unsigned long long)st_warp_max, (unsigned long long)st_bar);
}

}  // namespace emu
