// nb_index.cu -- k-mer index build on the GPU.
//
// Replaces build_lookup_table (lookup_table/lookup_table.c:149-164): same content
// (every k-mer occurring <= q times, its positions ascending), different
// construction and layout:
//   reference: dense u64 table (count<<34|start) + u64 offset list, built by two
//              serial scans and a 4-pass LSD radix sort of (hash<<34|pos) words;
//   here:      dense u32 CSR row-pointer table tbl[4^k + 1] + u32 offsets.
// Counting k-mers straight into a 4 GiB table is one random 32-byte sector of HBM per k-mer and pass (the first
// version of this file: 39 ms for 184 M k-mers).  Now the (hash, position) pairs are first partitioned by the top bits
// of the hash -- one or two passes of at most 256 bins, tiles staged in shared memory and written out as runs per bin --
// into buckets of 2^14 consecutive hash values, so that the counters of a bucket live in shared memory:
//   k_part_count / k_part_scatter   (x 1 or 2)   pairs grouped by bucket (order inside a bucket arbitrary)
//   k_bucket_count                  per bucket: count per hash in shared memory, cut-off > q, kept / distinct totals
//   k_bucket_fill                   per bucket: counts again, exclusive scan, the bucket's slice of tbl (every entry of
//                                   the table is written exactly once, coalesced), positions placed run by run, short
//                                   runs put in ascending order on the spot, long ones (> 24) listed for k_sort_long_runs
// HBM bytes per indexed base: the pair array (8 B) is written and read twice by the partition passes and read
// twice more by the bucket kernels; table 4 B x 4^k written once; offsets 4 B.
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

#define IDX_LB 14                       // hash values per bucket = 2^IDX_LB: one u32 counter each = 64 KB of shared memory
#define IDX_BUCKET (1u << IDX_LB)
#define IDX_TILE 4096                   // pairs per tile of a partition pass
#define IDX_BINS 256                    // bins of a partition pass (at most)
#define IDX_FILL_THREADS 512
#define LONG_RUN 24

struct IdxGeom {
    uint32_t h_lo, h_n;                 // this build's slice of the hash space
    int      k;
    int      sh1, nb1;                  // pass 1: bin = hr >> sh1 (nb1 bins)
    int      sh2, nb2;                  // pass 2: bin = (hr >> sh2) & (nb2 - 1); nb2 = 1: no second pass
    uint32_t n_buckets;
};

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = lane < nw ? s_warp[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const uint32_t r = s_warp[wid] + inc - v;
    if (total) *total = s_warp[32];
    __syncthreads();
    return r;
}

// The k-mers of 16 consecutive start positions (chunk c of the volume) that are valid and lie in the slice: relative hash
// values into hr[], bit i of the result set for position 16 c + i.
__device__ __forceinline__ uint32_t idx_chunk_kmers(const uint32_t* __restrict__ words, const uint32_t* __restrict__ invalid, uint64_t db_size,
                                                    uint64_t c, const IdxGeom& G, uint32_t hr[16])
{
    const uint32_t mask = (uint32_t)((1ull << (2 * G.k)) - 1);
    const uint64_t v = nb_load_be64(words, c);
    const uint32_t bad = __ldg(invalid + (c >> 1)) >> ((c & 1) * 16);
    const uint64_t p0 = c << 4;
    uint32_t ok = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t h = ((uint32_t)(v >> (64 - 2 * i - 2 * G.k)) & mask) - G.h_lo;
        hr[i] = h;
        if (!((bad >> i) & 1) && p0 + i < db_size && h < G.h_n) ok |= 1u << i;
    }
    return ok;
}

// ---- partition pass 1, over the packed reads: bin sizes
__global__ void __launch_bounds__(IDX_THREADS)
k_part1_count(const uint32_t* __restrict__ words, uint64_t db_size, const uint32_t* __restrict__ invalid, IdxGeom G, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_cnt[IDX_BINS];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t nchunk = (db_size + 15) >> 4;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t hr[16];
        const uint32_t ok = idx_chunk_kmers(words, invalid, db_size, c, G, hr);
#pragma unroll
        for (int i = 0; i < 16; ++i) if ((ok >> i) & 1) atomicAdd(&s_cnt[hr[i] >> G.sh1], 1u);
    }
    __syncthreads();
    if (s_cnt[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_cnt[threadIdx.x]);
}

// Exclusive scan of n bin sizes (n <= 65536) into base[n + 1] and the scatter cursors; tiles of IDX_TILE pairs per bin
// into tile_first[n + 1] (the work list of the next partition pass).  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024)
k_part_scan(const uint32_t* __restrict__ hist, uint32_t n, uint32_t* __restrict__ base, uint32_t* __restrict__ cursor, uint32_t* __restrict__ tile_first)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t i0 = threadIdx.x * per, i1 = i0 + per < n ? i0 + per : n;
    uint32_t s = 0, t = 0;
    for (uint32_t i = i0; i < i1; ++i) { s += hist[i]; t += (hist[i] + IDX_TILE - 1) / IDX_TILE; }
    uint32_t tot_s = 0, tot_t = 0;
    uint32_t bs = block_scan_excl(s, s_warp, &tot_s), bt = block_scan_excl(t, s_warp, &tot_t);
    for (uint32_t i = i0; i < i1; ++i) {
        base[i] = bs; cursor[i] = bs;
        if (tile_first) tile_first[i] = bt;
        bs += hist[i]; bt += (hist[i] + IDX_TILE - 1) / IDX_TILE;
    }
    if (threadIdx.x == 0) { base[n] = tot_s; if (tile_first) tile_first[n] = tot_t; }
}

// One tile of a partition pass, the pairs already in registers (NPER per thread, valid ones flagged in `ok`): rank inside
// the bin, space for the tile's share of every bin taken from the bin's cursor (the order of tiles inside a bin is
// arbitrary), the pairs staged bin by bin in shared memory and written out as runs.
template <int NPER>
__device__ __forceinline__ void idx_scatter_tile(const uint32_t (&hr)[NPER], const uint32_t (&pos)[NPER], uint32_t ok, int sh, uint32_t binmask,
                                                 uint32_t* __restrict__ cursor, uint2* __restrict__ dst,
                                                 uint32_t* s_cnt, uint32_t* s_loff, uint32_t* s_goff, uint32_t* s_warp, uint2* s_stage)
{
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t rank[NPER];
#pragma unroll
    for (int i = 0; i < NPER; ++i) if ((ok >> i) & 1) rank[i] = atomicAdd(&s_cnt[(hr[i] >> sh) & binmask], 1u);
    __syncthreads();
    {
        const uint32_t c = s_cnt[threadIdx.x];
        uint32_t total;
        const uint32_t lo = block_scan_excl(c, s_warp, &total);
        s_loff[threadIdx.x] = lo;
        s_goff[threadIdx.x] = c ? atomicAdd(cursor + threadIdx.x, c) : 0u;
        if (threadIdx.x == 0) s_loff[IDX_BINS] = total;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPER; ++i) if ((ok >> i) & 1) s_stage[s_loff[(hr[i] >> sh) & binmask] + rank[i]] = make_uint2(hr[i], pos[i]);
    __syncthreads();
    const uint32_t total = s_loff[IDX_BINS];
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint2 e = s_stage[i];
        const uint32_t b = (e.x >> sh) & binmask;
        dst[s_goff[b] + (i - s_loff[b])] = e;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(IDX_THREADS)
k_part1_scatter(const uint32_t* __restrict__ words, uint64_t db_size, const uint32_t* __restrict__ invalid, IdxGeom G,
                uint32_t* __restrict__ cursor, uint2* __restrict__ dst)
{
    __shared__ uint32_t s_cnt[IDX_BINS], s_loff[IDX_BINS + 1], s_goff[IDX_BINS], s_warp[33];
    __shared__ uint2 s_stage[IDX_TILE];
    const uint64_t nchunk = (db_size + 15) >> 4;
    const uint64_t ntile = (nchunk + IDX_THREADS - 1) / IDX_THREADS;
    for (uint64_t t = blockIdx.x; t < ntile; t += gridDim.x) {
        const uint64_t c = t * IDX_THREADS + threadIdx.x;
        uint32_t hr[16], pos[16], ok = 0;
        if (c < nchunk) ok = idx_chunk_kmers(words, invalid, db_size, c, G, hr);
#pragma unroll
        for (int i = 0; i < 16; ++i) pos[i] = (uint32_t)(c << 4) + i;
        idx_scatter_tile<16>(hr, pos, ok, G.sh1, 0xffffffffu, cursor, dst, s_cnt, s_loff, s_goff, s_warp, s_stage);
    }
}

// ---- partition pass 2, over the pairs of every pass-1 bin: tile -> (bin, range) through tile_first
__device__ __forceinline__ bool idx_tile_range(uint32_t tile, const uint32_t* __restrict__ tile_first, const uint32_t* __restrict__ base1, int nb1,
                                               uint32_t* bin, uint32_t* lo, uint32_t* hi)
{
    if (tile >= tile_first[nb1]) return false;
    int a = 0, b = nb1;                             // last bin whose first tile is <= tile
    while (b - a > 1) { const int m = (a + b) >> 1; if (tile_first[m] <= tile) a = m; else b = m; }
    *bin = (uint32_t)a;
    const uint32_t s = base1[a] + (tile - tile_first[a]) * IDX_TILE, e = base1[a + 1];
    *lo = s; *hi = s + IDX_TILE < e ? s + IDX_TILE : e;
    return true;
}

__global__ void __launch_bounds__(IDX_THREADS)
k_part2_count(const uint2* __restrict__ src, const uint32_t* __restrict__ tile_first, const uint32_t* __restrict__ base1, IdxGeom G,
              uint32_t max_tiles, uint32_t* __restrict__ hist2)
{
    __shared__ uint32_t s_cnt[IDX_BINS];
    for (uint32_t tile = blockIdx.x; tile < max_tiles; tile += gridDim.x) {
        uint32_t bin, lo, hi;
        if (!idx_tile_range(tile, tile_first, base1, G.nb1, &bin, &lo, &hi)) break;
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_cnt[(src[i].x >> G.sh2) & (uint32_t)(G.nb2 - 1)], 1u);
        __syncthreads();
        if (threadIdx.x < (unsigned)G.nb2 && s_cnt[threadIdx.x]) atomicAdd(hist2 + (size_t)bin * G.nb2 + threadIdx.x, s_cnt[threadIdx.x]);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(IDX_THREADS)
k_part2_scatter(const uint2* __restrict__ src, const uint32_t* __restrict__ tile_first, const uint32_t* __restrict__ base1, IdxGeom G,
                uint32_t max_tiles, uint32_t* __restrict__ cursor2, uint2* __restrict__ dst)
{
    __shared__ uint32_t s_cnt[IDX_BINS], s_loff[IDX_BINS + 1], s_goff[IDX_BINS], s_warp[33];
    __shared__ uint2 s_stage[IDX_TILE];
    for (uint32_t tile = blockIdx.x; tile < max_tiles; tile += gridDim.x) {
        uint32_t bin, lo, hi;
        if (!idx_tile_range(tile, tile_first, base1, G.nb1, &bin, &lo, &hi)) break;
        uint32_t hr[16], pos[16], ok = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t x = lo + i * IDX_THREADS + threadIdx.x;
            hr[i] = 0; pos[i] = 0;
            if (x < hi) { const uint2 e = src[x]; hr[i] = e.x; pos[i] = e.y; ok |= 1u << i; }
        }
        idx_scatter_tile<16>(hr, pos, ok, G.sh2, (uint32_t)(G.nb2 - 1), cursor2 + (size_t)bin * G.nb2, dst, s_cnt, s_loff, s_goff, s_warp, s_stage);
    }
}

// ---- per bucket: counts in shared memory, the cut-off, totals
// The counters of a bucket are swept as uint4 (a bucket holds ~2 800 pairs for 16 384 counters at E. coli size: the sweeps,
// not the pairs, are the work).  Entry i lives in 16-byte slot v ^ ((v >> 3) & 7), v = i >> 2: where a thread owns 32
// consecutive entries (8 slots, the scan of k_bucket_fill) the 8 lanes of a quarter warp are on 8 different bank groups,
// and a warp reading consecutive slots still is.
__device__ __forceinline__ uint32_t idx_slot(uint32_t v) { return v ^ ((v >> 3) & 7u); }
__device__ __forceinline__ uint32_t idx_entry(uint32_t i) { return (idx_slot(i >> 2) << 2) | (i & 3u); }

__global__ void __launch_bounds__(IDX_FILL_THREADS)
k_bucket_count(const uint2* __restrict__ pairs, const uint32_t* __restrict__ bbase, IdxGeom G, uint32_t q,
               uint32_t* __restrict__ bucket_kept, uint32_t* __restrict__ bucket_distinct)
{
    extern __shared__ __align__(16) uint32_t s_a[];             // [IDX_BUCKET]
    __shared__ uint32_t s_red[2];
    uint4* s4 = reinterpret_cast<uint4*>(s_a);
    for (uint32_t b = blockIdx.x; b < G.n_buckets; b += gridDim.x) {
        const uint32_t lo = bbase[b], hi = bbase[b + 1];
        if (lo == hi) { if (threadIdx.x == 0) { bucket_kept[b] = 0; bucket_distinct[b] = 0; } continue; }
        for (uint32_t v = threadIdx.x; v < IDX_BUCKET / 4; v += blockDim.x) s4[v] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x < 2) s_red[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_a[idx_entry(pairs[i].x & (IDX_BUCKET - 1))], 1u);
        __syncthreads();
        uint32_t kept = 0, dist = 0;
        for (uint32_t v = threadIdx.x; v < IDX_BUCKET / 4; v += blockDim.x) {
            const uint4 c = s4[v];                              // (any four entries: only totals are wanted)
            if (c.x && c.x <= q) { kept += c.x; ++dist; }
            if (c.y && c.y <= q) { kept += c.y; ++dist; }
            if (c.z && c.z <= q) { kept += c.z; ++dist; }
            if (c.w && c.w <= q) { kept += c.w; ++dist; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { kept += __shfl_xor_sync(0xffffffffu, kept, d); dist += __shfl_xor_sync(0xffffffffu, dist, d); }
        if ((threadIdx.x & 31) == 0 && kept) { atomicAdd(&s_red[0], kept); atomicAdd(&s_red[1], dist); }
        __syncthreads();
        if (threadIdx.x == 0) { bucket_kept[b] = s_red[0]; bucket_distinct[b] = s_red[1]; }
        __syncthreads();
    }
}

// Exclusive scan of the buckets' kept counts (in place) and the totals: out[0] = indexed positions, out[1] = distinct k-mers.
__global__ void __launch_bounds__(1024)
k_bucket_scan(uint32_t* __restrict__ bucket_kept, const uint32_t* __restrict__ bucket_distinct, uint32_t n, uint64_t* __restrict__ out)
{
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_d;
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t i0 = threadIdx.x * per, i1 = i0 + per < n ? i0 + per : n;
    uint32_t s = 0;
    unsigned long long d = 0;
    for (uint32_t i = i0; i < i1; ++i) { s += bucket_kept[i]; d += bucket_distinct[i]; }
    if (threadIdx.x == 0) s_d = 0;
    uint32_t tot = 0;
    uint32_t bs = block_scan_excl(s, s_warp, &tot);
    for (uint32_t i = i0; i < i1; ++i) { const uint32_t c = bucket_kept[i]; bucket_kept[i] = bs; bs += c; }
    if (d) atomicAdd(&s_d, d);
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = tot; out[1] = s_d; bucket_kept[n] = tot; }
}

// ---- per bucket: the slice of the table, the positions run by run
// tbl[h] = start of the run of hash h (tbl[h_n] = number of positions): thread t owns the 32 hash values [32 t, 32 t + 32) of
// the bucket in the scan and later puts in order the runs of two and more positions among them (a bit mask it kept from
// the scan: no sweep over the counters for that); the starts go to the table in one coalesced sweep of 16-byte stores.
__global__ void __launch_bounds__(IDX_FILL_THREADS)
k_bucket_fill(const uint2* __restrict__ pairs, const uint32_t* __restrict__ bbase, const uint32_t* __restrict__ kbase, IdxGeom G, uint32_t q,
              uint32_t* __restrict__ tbl, uint32_t* __restrict__ offs, uint32_t* __restrict__ long_list, uint32_t* __restrict__ n_long)
{
    extern __shared__ __align__(16) uint32_t s_a[];             // [IDX_BUCKET] counts, then run starts / fill cursors, then run ends
    __shared__ uint32_t s_drop[IDX_BUCKET / 32];                // bit: the k-mer occurs more than q times
    __shared__ uint32_t s_warp[33];
    uint4* s4 = reinterpret_cast<uint4*>(s_a);
    static_assert(IDX_BUCKET / IDX_FILL_THREADS == 32, "a thread owns one 32-entry row of the bucket");
    for (uint32_t b = blockIdx.x; b < G.n_buckets; b += gridDim.x) {
        const uint32_t lo = bbase[b], hi = bbase[b + 1], kb = kbase[b], kept = kbase[b + 1] - kb;
        const uint32_t h0 = b << IDX_LB;
        const uint32_t nh = G.h_n - h0 < IDX_BUCKET ? G.h_n - h0 : IDX_BUCKET;
        if (b == G.n_buckets - 1 && threadIdx.x == 0) tbl[G.h_n] = kbase[G.n_buckets];
        if (kept == 0) {                                        // nothing kept: the slice is flat
            for (uint32_t v = threadIdx.x; v < IDX_BUCKET / 4; v += blockDim.x) {
                uint32_t* t4 = tbl + (size_t)h0 + 4u * v;
                if (4u * v + 4u <= nh) *reinterpret_cast<uint4*>(t4) = make_uint4(kb, kb, kb, kb);
                else for (uint32_t x = 0; x < 4; ++x) if (4u * v + x < nh) t4[x] = kb;
            }
            continue;
        }
        for (uint32_t v = threadIdx.x; v < IDX_BUCKET / 4; v += blockDim.x) s4[v] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_a[idx_entry(pairs[i].x & (IDX_BUCKET - 1))], 1u);
        __syncthreads();
        uint32_t multi = 0;                                     // my entries whose run has two or more positions
        {
            uint4 c[8];
            uint32_t drop = 0, sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = s4[idx_slot(8u * threadIdx.x + j)];
                uint32_t* e = &c[j].x;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (e[x] > q) { drop |= 1u << (4 * j + x); e[x] = 0; }
                    sum += e[x];
                    if (e[x] >= 2) multi |= 1u << (4 * j + x);
                }
            }
            s_drop[threadIdx.x] = drop;
            uint32_t run = block_scan_excl(sum, s_warp, NULL);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 st;
                st.x = run; run += c[j].x; st.y = run; run += c[j].y; st.z = run; run += c[j].z; st.w = run; run += c[j].w;
                s4[idx_slot(8u * threadIdx.x + j)] = st;        // start of the run = where its first position goes
            }
        }
        __syncthreads();
        // the bucket's slice of the table: consecutive lanes, consecutive 16 bytes
        for (uint32_t v = threadIdx.x; v < IDX_BUCKET / 4; v += blockDim.x) {
            const uint4 st = s4[idx_slot(v)];
            uint32_t* t4 = tbl + (size_t)h0 + 4u * v;
            if (4u * v + 4u <= nh) *reinterpret_cast<uint4*>(t4) = make_uint4(kb + st.x, kb + st.y, kb + st.z, kb + st.w);
            else {
                const uint32_t v4[4] = {st.x, st.y, st.z, st.w};
                for (uint32_t x = 0; x < 4; ++x) if (4u * v + x < nh) t4[x] = kb + v4[x];
            }
        }
        __syncthreads();                                        // the starts are out before they turn into cursors
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const uint2 e = pairs[i];
            const uint32_t hl = e.x & (IDX_BUCKET - 1);
            if ((s_drop[hl >> 5] >> (hl & 31)) & 1) continue;
            offs[kb + atomicAdd(&s_a[idx_entry(hl)], 1u)] = e.y;      // the cursor of the run ends up as its end
        }
        __syncthreads();
        // ascending positions inside every run (the reference gets them from scan order + a stable sort)
        while (multi) {
            const uint32_t j = (uint32_t)__ffs((int)multi) - 1;
            multi &= multi - 1;
            const uint32_t i = 32u * threadIdx.x + j;
            const uint32_t en = s_a[idx_entry(i)], st = i ? s_a[idx_entry(i - 1)] : 0u;    // the run before mine ends where mine starts
            const uint32_t n = en - st;
            if (n > LONG_RUN) { long_list[atomicAdd(n_long, 1u)] = h0 + i; continue; }
            uint32_t* a = offs + kb + st;
            for (uint32_t x = 1; x < n; ++x) {
                const uint32_t v = a[x];
                uint32_t y = x;
                while (y > 0 && a[y - 1] > v) { a[y] = a[y - 1]; --y; }
                a[y] = v;
            }
        }
        __syncthreads();
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
    IdxGeom G;
    memset(&G, 0, sizeof G);
    G.h_lo = (uint32_t)h_lo; G.h_n = (uint32_t)h_n; G.k = k;
    G.n_buckets = (uint32_t)((h_n + IDX_BUCKET - 1) >> IDX_LB);
    if (G.n_buckets == 0) G.n_buckets = 1;
    {
        int tb = 0;                                             // bits of the bucket number
        while ((1ull << tb) < G.n_buckets) ++tb;
        const int t1 = tb < 8 ? tb : 8, t2 = tb - t1;           // tb <= 16
        G.sh1 = IDX_LB + t2; G.nb1 = 1 << t1;
        G.sh2 = IDX_LB; G.nb2 = 1 << t2;
    }
    const uint32_t nbk = (uint32_t)G.nb1 * (uint32_t)G.nb2;     // >= n_buckets; bucket = bin1 * nb2 + bin2
    uint32_t *d_invalid = NULL, *d_hist1 = NULL, *d_base1 = NULL, *d_cur1 = NULL, *d_tile1 = NULL, *d_hist2 = NULL, *d_base2 = NULL, *d_cur2 = NULL;
    uint32_t *d_kept = NULL, *d_dist = NULL, *d_long = NULL, *d_nlong = NULL, *d_tbl = NULL, *d_off = NULL;
    uint2 *d_pa = NULL, *d_pb = NULL;
    uint64_t* d_tot = NULL;
    const uint64_t inv_words = ((ref->db_size + 31) >> 5) + 2;
    NbDevBuf B(ctx);
    {
        int rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&d_tbl, sizeof(uint32_t) * (h_n + 8)))) return rc;
        if ((rc = B.get(&d_invalid, inv_words))) return rc;
        if ((rc = B.get(&d_hist1, IDX_BINS))) return rc;
        if ((rc = B.get(&d_base1, IDX_BINS + 1))) return rc;
        if ((rc = B.get(&d_cur1, IDX_BINS))) return rc;
        if ((rc = B.get(&d_tile1, IDX_BINS + 1))) return rc;
        if ((rc = B.get(&d_hist2, nbk))) return rc;
        if ((rc = B.get(&d_base2, nbk + 1))) return rc;
        if ((rc = B.get(&d_cur2, nbk))) return rc;
        if ((rc = B.get(&d_kept, nbk + 1))) return rc;
        if ((rc = B.get(&d_dist, nbk))) return rc;
        if ((rc = B.get(&d_tot, 2))) return rc;
        if ((rc = B.get(&d_nlong, 1))) return rc;
        if ((rc = B.get(&d_pa, ref->db_size + 16))) return rc;
        if (G.nb2 > 1 && (rc = B.get(&d_pb, ref->db_size + 16))) return rc;
    }
    NB_CUDA(ctx, cudaMemsetAsync(d_invalid, 0, sizeof(uint32_t) * inv_words, st));
    NB_CUDA(ctx, cudaMemsetAsync(d_hist1, 0, sizeof(uint32_t) * IDX_BINS, st));
    NB_CUDA(ctx, cudaMemsetAsync(d_hist2, 0, sizeof(uint32_t) * nbk, st));
    NB_CUDA(ctx, cudaMemsetAsync(d_nlong, 0, sizeof(uint32_t), st));

    const uint32_t* words = reinterpret_cast<const uint32_t*>(ref->d_pac);
    const uint64_t nchunk = (ref->db_size + 15) >> 4;
    const uint64_t ntile1 = (nchunk + IDX_THREADS - 1) / IDX_THREADS;
    unsigned grid1 = (unsigned)(ntile1 < NB_SMS * 8 ? ntile1 : NB_SMS * 8);
    if (grid1 == 0) grid1 = 1;
    if (ref->nseq) {
        k_mark_invalid<<<(unsigned)((ref->nseq + 255) / 256), 256, 0, st>>>(ref->d_off, ref->d_len, ref->nseq, k, d_invalid);
        NB_LAUNCH_CHECK(ctx);
    }
    k_part1_count<<<grid1, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, G, d_hist1);
    NB_LAUNCH_CHECK(ctx);
    k_part_scan<<<1, 1024, 0, st>>>(d_hist1, (uint32_t)G.nb1, d_base1, d_cur1, d_tile1);
    NB_LAUNCH_CHECK(ctx);
    k_part1_scatter<<<grid1, IDX_THREADS, 0, st>>>(words, ref->db_size, d_invalid, G, d_cur1, d_pa);
    NB_LAUNCH_CHECK(ctx);
    const uint2* d_pairs = d_pa;
    const uint32_t* d_bbase = d_base1;
    if (G.nb2 > 1) {
        // every pass-1 bin is cut into tiles of IDX_TILE pairs: at most db_size / IDX_TILE + nb1 of them
        const uint64_t max_tiles = ref->db_size / IDX_TILE + (uint64_t)G.nb1 + 1;
        unsigned grid2 = (unsigned)(max_tiles < NB_SMS * 8 ? max_tiles : NB_SMS * 8);
        k_part2_count<<<grid2, IDX_THREADS, 0, st>>>(d_pa, d_tile1, d_base1, G, (uint32_t)max_tiles, d_hist2);
        NB_LAUNCH_CHECK(ctx);
        k_part_scan<<<1, 1024, 0, st>>>(d_hist2, nbk, d_base2, d_cur2, NULL);
        NB_LAUNCH_CHECK(ctx);
        k_part2_scatter<<<grid2, IDX_THREADS, 0, st>>>(d_pa, d_tile1, d_base1, G, (uint32_t)max_tiles, d_cur2, d_pb);
        NB_LAUNCH_CHECK(ctx);
        d_pairs = d_pb; d_bbase = d_base2;
    }
    const size_t bsm = sizeof(uint32_t) * IDX_BUCKET;
    NB_CUDA(ctx, cudaFuncSetAttribute(k_bucket_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    NB_CUDA(ctx, cudaFuncSetAttribute(k_bucket_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    const unsigned gridb = G.n_buckets < NB_SMS * 3 ? G.n_buckets : NB_SMS * 3;
    k_bucket_count<<<gridb, IDX_FILL_THREADS, bsm, st>>>(d_pairs, d_bbase, G, (uint32_t)q, d_kept, d_dist);
    NB_LAUNCH_CHECK(ctx);
    k_bucket_scan<<<1, 1024, 0, st>>>(d_kept, d_dist, G.n_buckets, d_tot);
    NB_LAUNCH_CHECK(ctx);
    uint64_t tot[2];
    NB_CUDA(ctx, cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    {
        int rc;
        if ((rc = nb_dev_alloc(ctx, (void**)&d_off, sizeof(uint32_t) * (tot[0] + 16)))) return rc;
        if ((rc = B.get(&d_long, tot[0] / LONG_RUN + 16))) return rc;
    }
    k_bucket_fill<<<gridb, IDX_FILL_THREADS, bsm, st>>>(d_pairs, d_bbase, d_kept, G, (uint32_t)q, d_tbl, d_off, d_long, d_nlong);
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
