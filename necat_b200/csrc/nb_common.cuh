// nb_common.cuh -- shared device/host helpers of libnecat_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/necat_b200.h"

#define NB_SMS 148          // B200: 2 dies x 74 SMs
#define NB_MAX_WORKERS 4    // pipeline workers of nb_consensus_job (the context itself is worker 0)
#define NB_AUX 6            // side streams of a context

struct nb_ctx {
    int          device;
    cudaStream_t stream;
    char         err[512];
    uint64_t     launches;
    nb_timings   tm;
    nb_map_counters mc;
    nb_cns_counters cc;
    cudaEvent_t  ev0, ev1, ev2, ev3;
    cudaStream_t aux[NB_AUX];    // side streams: slices of the consensus pipeline overlap (links of one, best-path chains of others)
    cudaEvent_t  aux_ev[NB_AUX + 1];    // [i]: slice i has finished; [NB_AUX]: the fork
    cudaEvent_t  link_ev[NB_AUX];       // [i]: the link kernel of slice i has finished (the next slice's waits for it)
    void*        pinned;         // pinned host staging buffer for large device-to-host copies
    size_t       pinned_bytes;
    void*        ext_mat;        // walk-back scratch of the extension kernel (kept across calls)
    size_t       ext_mat_slots;
    void*        dev_cache;      // NbDevCache*: device blocks kept for reuse (cudaMalloc/cudaFree are slow and synchronising)
    void*        pin_cache[4];   // pinned host blocks of freed volumes, reused by the next upload (pinning 50 MB costs ~10 ms)
    size_t       pin_cache_bytes[4];
    nb_ctx*      workers[NB_MAX_WORKERS];   // [1..]: shadow contexts of the pipeline workers (own stream, events, staging,
                                            // scratch; they share dev_cache with their parent), created on first use
    int          is_worker;
    nb_ctx*      parent;         // of a worker: the context that owns the shared scratch (extension kernel launches are serialised)
    int          ext_ctas;       // k_extend CTAs per SM for launches of this context (0 = default)
    int          ext_private;    // 1: this worker's extension launches use its own walk-back scratch (they may run beside the parent's)
};

struct nb_volume {
    uint8_t*  d_pac;        // 2-bit bases, padded with >= 64 zero bytes
    uint64_t  db_size;
    uint64_t  nseq;
    uint64_t* d_off;        // [nseq + 1]; d_off[nseq] = db_size
    uint32_t* d_len;        // [nseq]
    uint64_t* h_off;        // host copies (the sequential host logic needs them)
    uint32_t* h_len;
    uint32_t  max_len;
    uint8_t*  h_pac;        // host copy of the packed bases (raw leftovers are cut on the host); pinned: it is also the
                            // staging buffer of the upload
    size_t    h_pac_bytes;
    nb_ctx*   owner;        // device blocks and the pinned copy go back to this context's caches
};

struct nb_index {
    int       k;
    int       q;
    uint64_t  n_hash;       // 4^k
    uint32_t* d_tbl;        // [n_hash + 1] CSR starts: run of hash h = [tbl[h], tbl[h+1])
    uint32_t* d_off;        // [n_off] positions, ascending inside a run
    uint64_t  n_off;
    uint64_t  n_distinct;
    int       external;     // 1: d_tbl / d_off belong to the caller (nb_index_adopt)
};

extern char g_nb_err[512];

// caching device allocator (per context, thread-safe; a block is released only after the work using it was synchronised)
int  nb_dev_alloc(nb_ctx* ctx, void** p, size_t bytes);
void nb_dev_release(nb_ctx* ctx, void* p);
void nb_dev_cache_destroy(nb_ctx* ctx);
int  nb_pinned_reserve(nb_ctx* ctx, size_t bytes);      // ctx->pinned holds at least `bytes` afterwards
int  nb_worker_ctx(nb_ctx* parent, int w, nb_ctx** out);    // the w-th pipeline worker of `parent` (1 <= w < NB_MAX_WORKERS)

#ifdef __cplusplus
#include <vector>
struct NbDevBuf {          // RAII list of cached device blocks
    nb_ctx* ctx;
    std::vector<void*> ptrs;
    explicit NbDevBuf(nb_ctx* c) : ctx(c) {}
    ~NbDevBuf() { for (void* q : ptrs) nb_dev_release(ctx, q); }
    template <typename T> int get(T** p, uint64_t n)
    {
        void* q = NULL;
        int rc = nb_dev_alloc(ctx, &q, sizeof(T) * (size_t)(n ? n : 1));
        if (rc) return rc;
        *p = (T*)q;
        ptrs.push_back(q);
        return NB_OK;
    }
};
#endif

int nb_fail(nb_ctx* ctx, int code, const char* fmt, ...);

#define NB_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return nb_fail(ctx, NB_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,  \
                           cudaGetErrorString(e__));                                    \
    } while (0)

#define NB_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return nb_fail(ctx, NB_ECUDA, "%s:%d launch -> %s", __FILE__, __LINE__,     \
                           cudaGetErrorString(e__));                                    \
    } while (0)

// ---------------------------------------------------------------- packed bases
// PackedDB layout (common/ontcns_aux.h:118-119): base i lives in byte i>>2 at bit
// shift ((~i)&3)<<1, i.e. bytes are big-endian strings of 2-bit codes.

__host__ __device__ __forceinline__ int nb_base_at(const uint8_t* pac, uint64_t i)
{
    return (pac[i >> 2] >> (((~i) & 3) << 1)) & 3;
}

#ifdef __CUDACC__
// 32 consecutive bases starting at base index 16*w as one big-endian 64-bit value
// (first base in the top bits).  `words` is the pac buffer viewed as uint32_t.
__device__ __forceinline__ uint64_t nb_load_be64(const uint32_t* __restrict__ words, uint64_t w)
{
    uint32_t a = __ldg(words + w), b = __ldg(words + w + 1);
    a = __byte_perm(a, 0, 0x0123);
    b = __byte_perm(b, 0, 0x0123);
    return ((uint64_t)a << 32) | b;
}

// k-mer (k <= 16) starting at base p: the reference's hash value
// (lookup_table.c:8-12): first base in the most significant position.
__device__ __forceinline__ uint32_t nb_kmer_at(const uint32_t* __restrict__ words, uint64_t p, int k)
{
    uint64_t v = nb_load_be64(words, p >> 4);
    int sh = 64 - 2 * (int)(p & 15) - 2 * k;
    return (uint32_t)(v >> sh) & (uint32_t)((1ull << (2 * k)) - 1);
}

// reverse complement of a 2k-bit k-mer (codes: 3 - c is the complement)
__device__ __forceinline__ uint32_t nb_revcomp_kmer(uint32_t x, int k)
{
    uint32_t r = __brev(~x);                                 // reverses bit order, complements
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1); // restore bit order inside each base
    return r >> (32 - 2 * k);
}

// ------------------------------------------------------------- TMA bulk copy
// 1-D bulk async copy global -> shared (SASS UBLKCP) completing on an mbarrier.
// src and dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ uint32_t nb_smem_addr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void nb_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nb_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nb_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nb_smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void nb_tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     nb_smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(nb_smem_addr(bar))
                 : "memory");
}
// Bounded wait: traps instead of hanging the GPU if the copy never lands.
__device__ __forceinline__ void nb_mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(nb_smem_addr(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
#endif  // __CUDACC__
