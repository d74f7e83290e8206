// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA runtime and the device intrinsics this repo's
// kernels use, so that the UNMODIFIED sources under necat_b200/csrc can be compiled by g++ and run on a CPU
// (tests/emu/cudaemu/prep.py rewrites only what is not C++: the <<< >>> launches, `extern __shared__` and the handful
// of inline-PTX helpers).  Every CUDA thread of a block is a fiber (own stack, cooperative switching, emu_sched.cpp);
// a warp collective or a barrier is a rendezvous of fibers, so code that forgets a __syncwarp() fails here too.
// Blocks run one after the other, kernels run synchronously at their launch.  Nothing under necat_b200/ knows about this.
#pragma once
#define NB_EMU 1
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <math.h>
#include <algorithm>
#include <type_traits>

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static

// ---------------------------------------------------------------- vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(8)))  uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v; v.x = x; v.y = y; return v; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// ---------------------------------------------------------------- fibers (emu_sched.cpp)
namespace emu {
struct Warp;
struct Fiber {
    void*    sp;            // saved stack pointer
    uint3    tid;
    int      lane, state;
    Warp*    warp;
};
struct Warp {
    uint64_t val[2][32];    // values published by the lanes, double-buffered by the generation's parity
    uint32_t gen;           // collectives completed
    uint32_t arrived;       // lanes waiting at the current collective
    uint32_t alive;         // lanes that have not returned
};
extern thread_local Fiber* cur;
extern thread_local uint3  block_idx;
extern thread_local dim3   block_dim, grid_dim;
extern thread_local void*  dyn_smem;
// publish v, wait for the other live lanes of the warp, return the published values; *mask = the lanes that took part
const uint64_t* warp_exchange(uint64_t v, uint32_t* mask);
void block_barrier();
void yield();               // spin-wait helper: lets the other fibers run
void run(const char* name, dim3 grid, dim3 block, size_t smem, void (*entry)(void*), void* arg);
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)
#define warpSize 32

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { uint32_t m; emu::warp_exchange(0, &m); }
[[noreturn]] static inline void __trap() { fprintf(stderr, "cudaemu: __trap()\n"); abort(); }

namespace emu {
template <class T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits"); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    const int lane = emu::cur->lane, base = lane & ~(width - 1), s = base + (src & (width - 1));
    return ((m >> s) & 1) ? emu::from_bits<T>(a[s]) : v;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    const int lane = emu::cur->lane, base = lane & ~(width - 1), s = lane - (int)d;
    return (s >= base && ((m >> s) & 1)) ? emu::from_bits<T>(a[s]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    const int lane = emu::cur->lane, base = lane & ~(width - 1), s = lane + (int)d;
    return (s < base + width && ((m >> s) & 1)) ? emu::from_bits<T>(a[s]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int x, int width = 32)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    const int lane = emu::cur->lane, s = lane ^ x;
    return ((s & ~(width - 1)) == (lane & ~(width - 1)) && ((m >> s) & 1)) ? emu::from_bits<T>(a[s]) : v;
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(pred ? 1 : 0, &m);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((m >> l) & 1) && a[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned k, int pred) { return __ballot_sync(k, pred) != 0; }
static inline int __all_sync(unsigned k, int pred)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(pred ? 1 : 0, &m);
    for (int l = 0; l < 32; ++l) if (((m >> l) & 1) && !a[l]) return 0;
    return 1;
}
#define EMU_REDUCE(name, T, init, expr)                                                   \
    static inline T name(unsigned, T v)                                                   \
    {                                                                                     \
        uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);          \
        T r = init;                                                                       \
        for (int l = 0; l < 32; ++l) if ((m >> l) & 1) { const T o = emu::from_bits<T>(a[l]); r = expr; } \
        return r;                                                                         \
    }
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, r | o)
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, r & o)
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, r + o)
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, (o > r ? o : r))
EMU_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (o < r ? o : r))
static inline int __reduce_max_sync(unsigned k, int v)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    int r = INT32_MIN; (void)k;
    for (int l = 0; l < 32; ++l) if ((m >> l) & 1) { const int o = emu::from_bits<int>(a[l]); r = o > r ? o : r; }
    return r;
}
static inline int __reduce_min_sync(unsigned k, int v)
{
    uint32_t m; const uint64_t* a = emu::warp_exchange(emu::to_bits(v), &m);
    int r = INT32_MAX; (void)k;
    for (int l = 0; l < 32; ++l) if ((m >> l) & 1) { const int o = emu::from_bits<int>(a[l]); r = o < r ? o : r; }
    return r;
}

// ---------------------------------------------------------------- scalar intrinsics
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned n) { const uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (n & 31)) >> 32); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned n) { const uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (n & 31)); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __hiloint2double(int hi, int lo) { const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d; }
static inline int __double2hiint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(b >> 32); }
static inline int __double2loint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(uint32_t)b; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::yield(); }

// CUDA's overloaded min / max accept mixed integer types
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)a : (T)b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a > (T)b ? (T)a : (T)b; }

// atomics: other host threads may be emulating other kernels on the same memory
template <class T, class U> static inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* p, double v)
{
    uint64_t old, neu; double d;
    do { old = __atomic_load_n((uint64_t*)p, __ATOMIC_SEQ_CST); memcpy(&d, &old, 8); d += v; memcpy(&neu, &d, 8); }
    while (!__atomic_compare_exchange_n((uint64_t*)p, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    memcpy(&d, &old, 8);
    return d;
}
template <class T, class U> static inline T atomicSub(T* p, U v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicAnd(T* p, U v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < (T)v && !__atomic_compare_exchange_n(p, &o, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > (T)v && !__atomic_compare_exchange_n(p, &o, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

// ---------------------------------------------------------------- the inline-PTX helpers of nb_common.cuh / nb_cns_kernel.cuh
// (prep.py removes the originals).  A bulk copy lands at once; the mbarrier is a count of completed phases.
static inline uint32_t nb_smem_addr(const void* p) { return (uint32_t)(uintptr_t)p; }
static inline void nb_mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
static inline void nb_mbar_expect_tx(uint64_t*, uint32_t) { }
static inline void nb_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { memcpy(dst, src, bytes); *bar += 1; }
static inline void nb_mbar_wait(uint64_t* bar, uint32_t parity) { while ((*(volatile uint64_t*)bar & 1) == parity) emu::yield(); }
static inline void cns_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
static inline void cns_cp_async_commit() { }
static inline void cns_cp_async_wait(int) { }

// ---------------------------------------------------------------- runtime API (everything is synchronous host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct emu_stream_* cudaStream_t;
typedef struct emu_event_ { double t; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem, sharedMemPerBlockOptin; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "cudaemu (host fibers standing in for sm_100a)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)180 << 30; p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *fr = (size_t)8 << 30; *tot = (size_t)180 << 30; return cudaSuccess; }
// fresh device memory is NOT zero: fill it with a pattern so that a kernel that reads before it writes shows up
static inline cudaError_t cudaMalloc(void** p, size_t n)
{
    // exactly n bytes (256-byte aligned like the device allocator): under AddressSanitizer a kernel that reads or
    // writes one byte past what the host code asked for is reported
    *p = NULL;
    if (posix_memalign(p, 256, n ? n : 1)) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, n ? n : 1);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = aligned_alloc(256, ((n + 255) & ~(size_t)255) ? ((n + 255) & ~(size_t)255) : 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned f) { return cudaHostAlloc((void**)p, n, f); }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaHostAlloc(p, n, 0); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline double emu_now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, sizeof(**e)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { if (e) e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---------------------------------------------------------------- kernel launch:  EMU_LAUNCH((kernel), grid, block, smem, stream)(args...)
namespace emu {
template <class... P> struct Launcher {
    const char* name;
    void (*k)(P...);
    dim3 g, b;
    size_t smem;
    template <class... A> void operator()(A&&... a) const
    {
        auto body = [&]() { k((P)(a)...); };          // C-style: NULL (a long) becomes a null pointer, as nvcc accepts
        typedef decltype(body) B;
        run(name, g, b, smem, [](void* p) { (*static_cast<B*>(p))(); }, &body);
    }
};
template <class... P> inline Launcher<P...> launcher(const char* name, void (*k)(P...), dim3 g, dim3 b, size_t smem = 0, cudaStream_t = 0) { return Launcher<P...>{name, k, g, b, smem}; }
}  // namespace emu
#define EMU_LAUNCH(k, ...) emu::launcher(#k, k, __VA_ARGS__)
