// nb_index.cu -- k-mer index build on the GPU.
//
// Replaces build_lookup_table (lookup_table/lookup_table.c:149-164): same content
// (every k-mer occurring <= q times, its positions ascending), different
// construction and layout:
//   reference: dense u64 table (count<<34|start) + u64 offset list, built by two
//              serial scans and a 4-pass LSD radix sort of (hash<<34|pos) words;
//   here:      dense u32 CSR row-pointer table tbl[4^k + 1] + u32 offsets.
//              count (RED.ADD per k-mer) -> mask > q -> exclusive scan -> fill
//              with an atomic cursor per hash -> sort the (short) runs.
// Algorithmic HBM bytes per indexed base: 2 x 0.25 B packed read +
// 2 x 4 B random RMW on the count/cursor tables + 4 B offset write + table
// sweeps 4^k x (4+4+4+4+4) B.
#include "nb_common.cuh"

#define IDX_THREADS 256

// ---- mark k-mer start positions that would run over the end of their read
__global__ void k_mark_invalid(const uint64_t* __restrict__ off, const uint32_t* __restrict__ len, uint64_t nseq,
                               int k, uint32_t* __restrict__ invalid)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nseq) return;
    uint64_t end = off[r] + len[r];
    uint64_t beg = len[r] >= (uint32_t)k ? end - (k - 1) : off[r];
    for (uint64_t p = beg; p < end; ++p) atomicOr(invalid + (p >> 5), 1u << (p & 31));
}

// ---- count pass / fill pass share the k-mer enumeration: one thread per 16 bases
template <bool FILL>
__global__ void __launch_bounds__(IDX_THREADS)
k_scan_kmers(const uint32_t* __restrict__ words, uint64_t db_size, const uint32_t* __restrict__ invalid, int k,
             uint32_t* __restrict__ cnt, uint32_t* __restrict__ tbl, uint32_t* __restrict__ offs, uint32_t h_lo, uint32_t h_n)
{
    const uint64_t nchunk = (db_size + 15) >> 4;
    const uint32_t mask = (uint32_t)((1ull << (2 * k)) - 1);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = nb_load_be64(words, c);
        uint32_t bad = __ldg(invalid + (c >> 1)) >> ((c & 1) * 16);
        uint64_t p0 = c << 4;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((bad >> i) & 1) continue;
            if (p0 + i >= db_size) continue;
            uint32_t h = ((uint32_t)(v >> (64 - 2 * i - 2 * k)) & mask) - h_lo;      // inside this build's slice of the hash space?
            if (h >= h_n) continue;
            if (!FILL) {
                atomicAdd(cnt + h, 1u);                 // RED.ADD, result unused
            } else if (__ldg(cnt + h)) {
                uint32_t slot = atomicAdd(tbl + h + 1, 1u);
                offs[slot] = (uint32_t)(p0 + i);
            }
        }
    }
}

// ---- mask (count > q -> 0) + 3-phase exclusive scan into tbl[h + 1]
#define SCAN_ITEMS 8
#define SCAN_THREADS 1024
#define SCAN_TILE (SCAN_ITEMS * SCAN_THREADS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        warp_sums[lane] = winc - w;     // exclusive
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_mask_reduce(uint32_t* __restrict__ cnt, uint64_t n, uint32_t q, uint32_t* __restrict__ tile_sum,
              uint32_t* __restrict__ tile_distinct)
{
    __shared__ uint32_t tot, totd;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0, d = 0;
    if (base + SCAN_ITEMS <= n) {
        uint4 a = *reinterpret_cast<uint4*>(cnt + base), b = *reinterpret_cast<uint4*>(cnt + base + 4);
        uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        bool dirty = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (x[i] > q) { x[i] = 0; dirty = true; }
            s += x[i];
            d += x[i] != 0;
        }
        if (dirty) {
            *reinterpret_cast<uint4*>(cnt + base) = make_uint4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<uint4*>(cnt + base + 4) = make_uint4(x[4], x[5], x[6], x[7]);
        }
    } else {
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n) {
                uint32_t x = cnt[base + i];
                if (x > q) { x = 0; cnt[base + i] = 0; }
                s += x;
                d += x != 0;
            }
    }
    block_exclusive_scan(s, &tot);
    block_exclusive_scan(d, &totd);
    if (threadIdx.x == 0) { tile_sum[blockIdx.x] = tot; tile_distinct[blockIdx.x] = totd; }
}

// single block: exclusive scan of the tile sums (in place), totals to out[0..1]
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(uint32_t* __restrict__ tile_sum, const uint32_t* __restrict__ tile_distinct, uint32_t ntiles,
             uint64_t* __restrict__ out)
{
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    uint64_t dsum = 0;
    for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sum[i] : 0;
        uint32_t ex = block_exclusive_scan(v, &tot);
        if (i < ntiles) tile_sum[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    // distinct count: plain strided sum by thread 0..1023 then reduce
    uint64_t local = 0;
    for (uint32_t i = threadIdx.x; i < ntiles; i += SCAN_THREADS) local += tile_distinct[i];
    __shared__ unsigned long long dtot;
    if (threadIdx.x == 0) dtot = 0;
    __syncthreads();
    atomicAdd(&dtot, (unsigned long long)local);
    __syncthreads();
    dsum = dtot;
    if (threadIdx.x == 0) { out[0] = carry; out[1] = dsum; }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const uint32_t* __restrict__ cnt, uint64_t n, const uint32_t* __restrict__ tile_base,
             uint32_t* __restrict__ tbl)
{
    __shared__ uint32_t tot;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        x[i] = base + i < n ? cnt[base + i] : 0;
        s += x[i];
    }
    uint32_t run = tile_base[blockIdx.x] + block_exclusive_scan(s, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) tbl[base + i + 1] = run;      // cursor of hash (base+i) = start of its run
        run += x[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) tbl[0] = 0;
}

// ---- order every run by position (the atomic cursor fills runs in arbitrary order)
#define LONG_RUN 24
__global__ void __launch_bounds__(IDX_THREADS)
k_sort_short_runs(const uint32_t* __restrict__ tbl, uint64_t n_hash, uint32_t* __restrict__ offs,
                  uint32_t* __restrict__ long_list, uint32_t* __restrict__ n_long)
{
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n_hash; h += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = __ldg(tbl + h), hi = __ldg(tbl + h + 1);
        uint32_t n = hi - lo;
        if (n < 2) continue;
        if (n > LONG_RUN) { long_list[atomicAdd(n_long, 1u)] = (uint32_t)h; continue; }
        uint32_t* a = offs + lo;
        for (uint32_t i = 1; i < n; ++i) {
            uint32_t v = a[i];
            uint32_t j = i;
            while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; --j; }
            a[j] = v;
        }
    }
}

#define LR_CAP 2048
__global__ void __launch_bounds__(128)
k_sort_long_runs(const uint32_t* __restrict__ tbl, uint32_t* __restrict__ offs, const uint32_t* __restrict__ long_list,
                 uint32_t n_long)
{
    __shared__ uint32_t s[LR_CAP];
    for (uint32_t w = blockIdx.x; w < n_long; w += gridDim.x) {
        uint32_t h = long_list[w];
        uint32_t lo = tbl[h], n = tbl[h + 1] - lo;
        uint32_t* a = offs + lo;
        if (n > LR_CAP) {       // far beyond any sensible -q: in-place heap sort by one thread
            if (threadIdx.x == 0) {
                for (uint32_t st = n / 2; st-- > 0;) {
                    uint32_t r = st, v = a[r];
                    for (;;) { uint32_t c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; }
                    a[r] = v;
                }
                for (uint32_t e = n - 1; e > 0; --e) {
                    uint32_t v = a[e]; a[e] = a[0];
                    uint32_t r = 0;
                    for (;;) { uint32_t c = 2 * r + 1; if (c >= e) break; if (c + 1 < e && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; }
                    a[r] = v;
                }
            }
            __syncthreads();
            continue;
        }
        uint32_t P = 32;
        while (P < n) P <<= 1;
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s[i] = i < n ? a[i] : 0xffffffffu;
        __syncthreads();
        for (uint32_t kk = 2; kk <= P; kk <<= 1)
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                    uint32_t l = i ^ j;
                    if (l > i) {
                        bool up = (i & kk) == 0;
                        uint32_t x = s[i], y = s[l];
                        if ((x > y) == up) { s[i] = y; s[l] = x; }
                    }
                }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) a[i] = s[i];
        __syncthreads();
    }
}

// CSR of the hash values [h_lo, h_lo + h_n) of the volume: d_tbl[h_n + 1] (starts from 0), d_off.
static int index_build_range(nb_ctx* ctx, const nb_volume* ref, int k, int q, uint64_t h_lo, uint64_t h_n,
                             uint32_t** out_tbl, uint32_t** out_off, uint64_t* out_n_off, uint64_t* out_n_distinct)
{
    cudaStream_t st = ctx->stream;
    const uint64_t nh = h_n;
    const uint32_t ntiles = (uint32_t)((nh + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t *d_cnt = NULL, *d_invalid = NULL, *d_tile = NULL, *d_tiled = NULL, *d_long = NULL, *d_nlong = NULL, *d_tbl = NULL, *d_off = NULL;
    uint64_t* d_tot = NULL;
    const uint64_t inv_words = ((ref->db_size + 31) >> 5) + 2;
    NbDevBuf B(ctx);
    {
        int rc;
        if ((rc = B.get(&d_cnt, nh + 8))) return rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&d_tbl, sizeof(uint32_t) * (nh + 8)))) return rc;
        if ((rc = B.get(&d_invalid, inv_words))) return rc;
        if ((rc = B.get(&d_tile, ntiles))) return rc;
        if ((rc = B.get(&d_tiled, ntiles))) return rc;
        if ((rc = B.get(&d_tot, 2))) return rc;
        if ((rc = B.get(&d_nlong, 1))) return rc;
    }
    NB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * (nh + 8), st));
    NB_CUDA(ctx, cudaMemsetAsync(d_invalid, 0, sizeof(uint32_t) * inv_words, st));
    NB_CUDA(ctx, cudaMemsetAsync(d_nlong, 0, sizeof(uint32_t), st));

    const uint32_t* words = reinterpret_cast<const uint32_t*>(ref->d_pac);
    const uint64_t nchunk = (ref->db_size + 15) >> 4;
    unsigned grid_scan = (unsigned)((nchunk + IDX_THREADS - 1) / IDX_THREADS);
    if (grid_scan > NB_SMS * 32) grid_scan = NB_SMS * 32;
    if (grid_scan == 0) grid_scan = 1;
    if (ref->nseq) {
        k_mark_invalid<<<(unsigned)((ref->nseq + 255) / 256), 256, 0, st>>>(ref->d_off, ref->d_len, ref->nseq, k, d_invalid);
        NB_LAUNCH_CHECK(ctx);
    }
    k_scan_kmers<false><<<grid_scan, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, k, d_cnt, NULL, NULL, (uint32_t)h_lo, (uint32_t)h_n);
    NB_LAUNCH_CHECK(ctx);
    k_mask_reduce<<<ntiles, SCAN_THREADS, 0, st>>>(d_cnt, nh, (uint32_t)q, d_tile, d_tiled);
    NB_LAUNCH_CHECK(ctx);
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(d_tile, d_tiled, ntiles, d_tot);
    NB_LAUNCH_CHECK(ctx);
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(d_cnt, nh, d_tile, d_tbl);
    NB_LAUNCH_CHECK(ctx);
    uint64_t tot[2];
    NB_CUDA(ctx, cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    {
        int rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&d_off, sizeof(uint32_t) * (tot[0] + 16)))) return rc;
        if ((rc = B.get(&d_long, tot[0] / LONG_RUN + 16))) return rc;
    }
    k_scan_kmers<true><<<grid_scan, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, k, d_cnt, d_tbl, d_off, (uint32_t)h_lo, (uint32_t)h_n);
    NB_LAUNCH_CHECK(ctx);
    unsigned grid_h = (unsigned)((nh + IDX_THREADS - 1) / IDX_THREADS);
    if (grid_h > NB_SMS * 64) grid_h = NB_SMS * 64;
    k_sort_short_runs<<<grid_h, IDX_THREADS, 0, st>>>(d_tbl, nh, d_off, d_long, d_nlong);
    NB_LAUNCH_CHECK(ctx);
    uint32_t n_long = 0;
    NB_CUDA(ctx, cudaMemcpyAsync(&n_long, d_nlong, sizeof n_long, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_long) {
        unsigned g = n_long < NB_SMS * 16 ? n_long : NB_SMS * 16;
        k_sort_long_runs<<<g, 128, 0, st>>>(d_tbl, d_off, d_long, n_long);
        NB_LAUNCH_CHECK(ctx);
    }
    *out_tbl = d_tbl; *out_off = d_off; *out_n_off = tot[0]; *out_n_distinct = tot[1];
    return NB_OK;
}

static int index_check_args(nb_ctx* ctx, const nb_volume* ref, int k, int q)
{
    if (k < 4 || k > 15) return nb_fail(ctx, NB_ELIMIT, "kmer size %d unsupported (4..15; reference HashBits = 30)", k);
    if (q < 1) return nb_fail(ctx, NB_EINVAL, "kmer count cutoff must be >= 1");
    if (ref->db_size >= (1ull << 32)) return nb_fail(ctx, NB_ELIMIT, "a volume of %llu bases cannot be indexed: index positions are 32-bit (oc2mkdb cuts volumes at 2 Gbp)", (unsigned long long)ref->db_size);
    return NB_OK;
}

extern "C" int nb_index_build(nb_ctx* ctx, const nb_volume* ref, int k, int q, nb_index** out)
{
    if (!ctx || !ref || !out) return nb_fail(ctx, NB_EINVAL, "nb_index_build: bad arguments");
    { const int rc = index_check_args(ctx, ref, k, q); if (rc) return rc; }
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    nb_index* ix = (nb_index*)calloc(1, sizeof(nb_index));
    if (!ix) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    ix->k = k; ix->q = q; ix->n_hash = 1ull << (2 * k);
    cudaEventRecord(ctx->ev0, st);
    const int rc = index_build_range(ctx, ref, k, q, 0, ix->n_hash, &ix->d_tbl, &ix->d_off, &ix->n_off, &ix->n_distinct);
    if (rc) { free(ix); return rc; }
    cudaEventRecord(ctx->ev1, st);
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->tm.index_ms, ctx->ev0, ctx->ev1);
    *out = ix;
    return NB_OK;
}

// ---- the index of one volume built by several GPUs together: every GPU the CSR of one slice of the hash space
struct nb_index_shard {
    uint32_t* d_tbl;        // [h_n + 1] starts from 0
    uint32_t* d_off;
    uint64_t  h_n, n_off;
};

__global__ void k_shard_export(const uint32_t* __restrict__ tbl, uint64_t n, uint32_t base, uint32_t* __restrict__ dst)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = tbl[i] + base;
}

extern "C" int nb_index_build_shard(nb_ctx* ctx, const nb_volume* ref, int k, int q, int shard, int nshards,
                                    nb_index_shard** out, uint64_t* n_offsets, uint64_t* n_distinct)
{
    if (!ctx || !ref || !out || !n_offsets || !n_distinct || nshards < 1 || shard < 0 || shard >= nshards)
        return nb_fail(ctx, NB_EINVAL, "nb_index_build_shard: bad arguments");
    { const int rc = index_check_args(ctx, ref, k, q); if (rc) return rc; }
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_hash = 1ull << (2 * k), slice = (n_hash + nshards - 1) / nshards;
    const uint64_t h_lo = slice * (uint64_t)shard;
    const uint64_t h_n = h_lo >= n_hash ? 0 : (h_lo + slice <= n_hash ? slice : n_hash - h_lo);
    nb_index_shard* sh = (nb_index_shard*)calloc(1, sizeof(nb_index_shard));
    if (!sh) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    sh->h_n = h_n;
    cudaEventRecord(ctx->ev0, st);
    uint64_t nd = 0;
    const int rc = index_build_range(ctx, ref, k, q, h_lo, h_n, &sh->d_tbl, &sh->d_off, &sh->n_off, &nd);
    if (rc) { free(sh); return rc; }
    cudaEventRecord(ctx->ev1, st);
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->tm.index_ms, ctx->ev0, ctx->ev1);
    *out = sh; *n_offsets = sh->n_off; *n_distinct = nd;
    return NB_OK;
}

// Writes the shard into the table all GPUs assemble: starts + base into d_tbl_dst[0, h_n), positions into d_off_dst[0, n_off)
// (both device pointers of the caller, e.g. torch tensors that NCCL all-gathers afterwards).  Frees the shard.
extern "C" int nb_index_shard_export(nb_ctx* ctx, nb_index_shard* sh, uint64_t base, uint32_t* d_tbl_dst, uint32_t* d_off_dst)
{
    if (!ctx || !sh || !d_tbl_dst || !d_off_dst) return nb_fail(ctx, NB_EINVAL, "nb_index_shard_export: bad arguments");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (sh->h_n) {
        unsigned g = (unsigned)((sh->h_n + 255) / 256);
        if (g > NB_SMS * 32) g = NB_SMS * 32;
        k_shard_export<<<g, 256, 0, st>>>(sh->d_tbl, sh->h_n, (uint32_t)base, d_tbl_dst);
        NB_LAUNCH_CHECK(ctx);
    }
    if (sh->n_off) NB_CUDA(ctx, cudaMemcpyAsync(d_off_dst, sh->d_off, sizeof(uint32_t) * sh->n_off, cudaMemcpyDeviceToDevice, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    nb_dev_release(ctx, sh->d_tbl); nb_dev_release(ctx, sh->d_off);
    free(sh);
    return NB_OK;
}

// An index over arrays the caller owns (and keeps alive until nb_index_free, which does not free them).
extern "C" int nb_index_adopt(nb_ctx* ctx, int k, int q, uint32_t* d_tbl, uint32_t* d_off, uint64_t n_off, uint64_t n_distinct, nb_index** out)
{
    if (!ctx || !d_tbl || !d_off || !out || k < 4 || k > 15) return nb_fail(ctx, NB_EINVAL, "nb_index_adopt: bad arguments");
    nb_index* ix = (nb_index*)calloc(1, sizeof(nb_index));
    if (!ix) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    ix->k = k; ix->q = q; ix->n_hash = 1ull << (2 * k); ix->d_tbl = d_tbl; ix->d_off = d_off; ix->n_off = n_off; ix->n_distinct = n_distinct;
    ix->external = 1;
    *out = ix;
    return NB_OK;
}

extern "C" void nb_index_free(nb_ctx* ctx, nb_index* ix)
{
    if (!ix) return;
    if (ix->external) { }       // the arrays belong to the caller (nb_index_adopt)
    else if (ctx) { cudaSetDevice(ctx->device); nb_dev_release(ctx, ix->d_tbl); nb_dev_release(ctx, ix->d_off); }
    else { cudaFree(ix->d_tbl); cudaFree(ix->d_off); }
    free(ix);
}

extern "C" int nb_index_stats(nb_ctx* ctx, const nb_index* ix, uint64_t* n_offsets, uint64_t* n_distinct)
{
    if (!ix) return nb_fail(ctx, NB_EINVAL, "nb_index_stats: NULL index");
    if (n_offsets) *n_offsets = ix->n_off;
    if (n_distinct) *n_distinct = ix->n_distinct;
    return NB_OK;
}

extern "C" int nb_index_download(nb_ctx* ctx, const nb_index* ix, uint32_t* starts, uint32_t* offsets)
{
    if (!ctx || !ix) return nb_fail(ctx, NB_EINVAL, "nb_index_download: bad arguments");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (starts) NB_CUDA(ctx, cudaMemcpy(starts, ix->d_tbl, sizeof(uint32_t) * (ix->n_hash + 1), cudaMemcpyDeviceToHost));
    if (offsets && ix->n_off) NB_CUDA(ctx, cudaMemcpy(offsets, ix->d_off, sizeof(uint32_t) * ix->n_off, cudaMemcpyDeviceToHost));
    return NB_OK;
}
