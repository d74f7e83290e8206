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
             uint32_t* __restrict__ cnt, uint32_t* __restrict__ tbl, uint32_t* __restrict__ offs)
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
            uint32_t h = (uint32_t)(v >> (64 - 2 * i - 2 * k)) & mask;
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

extern "C" int nb_index_build(nb_ctx* ctx, const nb_volume* ref, int k, int q, nb_index** out)
{
    if (!ctx || !ref || !out) return nb_fail(ctx, NB_EINVAL, "nb_index_build: bad arguments");
    if (k < 4 || k > 15) return nb_fail(ctx, NB_ELIMIT, "kmer size %d unsupported (4..15; reference HashBits = 30)", k);
    if (q < 1) return nb_fail(ctx, NB_EINVAL, "kmer count cutoff must be >= 1");
    if (ref->db_size >= (1ull << 32)) return nb_fail(ctx, NB_ELIMIT, "a volume of %llu bases cannot be indexed: index positions are 32-bit (oc2mkdb cuts volumes at 2 Gbp)", (unsigned long long)ref->db_size);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    nb_index* ix = (nb_index*)calloc(1, sizeof(nb_index));
    if (!ix) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    ix->k = k; ix->q = q; ix->n_hash = 1ull << (2 * k);
    const uint64_t nh = ix->n_hash;
    const uint32_t ntiles = (uint32_t)((nh + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t *d_cnt = NULL, *d_invalid = NULL, *d_tile = NULL, *d_tiled = NULL, *d_long = NULL, *d_nlong = NULL;
    uint64_t* d_tot = NULL;
    const uint64_t inv_words = ((ref->db_size + 31) >> 5) + 2;
    cudaEventRecord(ctx->ev0, st);
    NbDevBuf B(ctx);
    {
        int rc;
        if ((rc = B.get(&d_cnt, nh + 8))) return rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&ix->d_tbl, sizeof(uint32_t) * (nh + 8)))) return rc;
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
    k_scan_kmers<false><<<grid_scan, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, k, d_cnt, NULL, NULL);
    NB_LAUNCH_CHECK(ctx);
    k_mask_reduce<<<ntiles, SCAN_THREADS, 0, st>>>(d_cnt, nh, (uint32_t)q, d_tile, d_tiled);
    NB_LAUNCH_CHECK(ctx);
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(d_tile, d_tiled, ntiles, d_tot);
    NB_LAUNCH_CHECK(ctx);
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(d_cnt, nh, d_tile, ix->d_tbl);
    NB_LAUNCH_CHECK(ctx);
    uint64_t tot[2];
    NB_CUDA(ctx, cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    ix->n_off = tot[0];
    ix->n_distinct = tot[1];
    {
        int rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&ix->d_off, sizeof(uint32_t) * (ix->n_off + 16)))) return rc;
        if ((rc = B.get(&d_long, ix->n_off / LONG_RUN + 16))) return rc;
    }
    k_scan_kmers<true><<<grid_scan, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, k, d_cnt, ix->d_tbl, ix->d_off);
    NB_LAUNCH_CHECK(ctx);
    unsigned grid_h = (unsigned)((nh + IDX_THREADS - 1) / IDX_THREADS);
    if (grid_h > NB_SMS * 64) grid_h = NB_SMS * 64;
    k_sort_short_runs<<<grid_h, IDX_THREADS, 0, st>>>(ix->d_tbl, nh, ix->d_off, d_long, d_nlong);
    NB_LAUNCH_CHECK(ctx);
    uint32_t n_long = 0;
    NB_CUDA(ctx, cudaMemcpyAsync(&n_long, d_nlong, sizeof n_long, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_long) {
        unsigned g = n_long < NB_SMS * 16 ? n_long : NB_SMS * 16;
        k_sort_long_runs<<<g, 128, 0, st>>>(ix->d_tbl, ix->d_off, d_long, n_long);
        NB_LAUNCH_CHECK(ctx);
    }
    cudaEventRecord(ctx->ev1, st);
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->tm.index_ms, ctx->ev0, ctx->ev1);
    *out = ix;
    return NB_OK;
}

extern "C" void nb_index_free(nb_ctx* ctx, nb_index* ix)
{
    if (!ix) return;
    if (ctx) { cudaSetDevice(ctx->device); nb_dev_release(ctx, ix->d_tbl); nb_dev_release(ctx, ix->d_off); }
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
