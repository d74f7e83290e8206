// nb_map.cu -- candidate finding on the GPU (oc2pmov CAN mode).
//
// Replaces find_candidates (word_finder/word_finder.c:364-412) for every
// (read, strand) "unit" of a query volume against one indexed volume:
//
//   k_probe         one CTA per unit: the packed read is staged into shared
//                   memory with a TMA bulk copy, every z-th k-mer is looked up
//                   in the CSR index, hits below soff_max are counted
//                   (word_finder.c:66-82,107-138; lookup_table.c:176-190).
//   k_build_blocks  one CTA per unit: hits are gathered as 64-bit keys
//                   (block | probe | in-block offset), bitonic-sorted, cut into
//                   scoring blocks, the <= 40 seeds per block selected
//                   (fill_one_seed, word_finder.c:85-104), the time-dependent
//                   pair score reproduced from the hit order (:103) and the
//                   blocks passing the static gates (:394-395) listed in
//                   first-touch order.
//   k_candidates    one warp per unit: the order-dependent part -- seed voting
//                   (scoring_seeds :141-168), DDF-consistent seed gathering with
//                   block zeroing (:249-307), chaining (chain_dp.c:37-159) and
//                   candidate selection (:318-358).
//
// The reference's per-thread dense ScoringBlock table over the whole volume
// (252 B x volume/2000) becomes a per-unit sorted list of touched blocks.
#include <algorithm>
#include <chrono>
#include <vector>

#include "nb_common.cuh"

#define BLK_SEEDS 40
#define PROBE_THREADS 128
#define BUILD_THREADS 256
#define SORT_CAP 4096           // keys sorted in shared memory; larger units use a global scratch
#define CAND_WARPS 4

struct MapDev {
    // index + reference volume
    const uint32_t* tbl;
    const uint32_t* offs;
    const uint64_t* ref_off;
    const uint32_t* ref_len;
    uint32_t        ref_nseq;
    // query volume
    const uint8_t*  rd_pac;
    const uint64_t* rd_off;
    const uint32_t* rd_len;
    uint64_t        read_lo;
    uint32_t        n_units;
    int             k, z, bs, s, a_cut;
    int             same_volume;    // query ids fall inside the reference volume (word_finder.c:121-127)
    uint32_t        stage_bytes;    // k_probe: shared memory of the launch; a longer read is read from global memory
    // per unit
    const uint64_t* probe_base;     // [n_units + 1]
    uint32_t*       unit_hits;      // [n_units]
    const uint64_t* hit_base;       // [n_units + 1]
    const uint64_t* scratch_base;   // [n_units] offset into sort_scratch or ~0
    uint32_t*       unit_nblk;
    uint32_t*       unit_nelig;
    // per probe
    uint32_t*       probe_lo;
    uint32_t*       probe_n;
    // per hit (indexed from hit_base[u])
    uint32_t*       seed_kid;       // 1-based probe number of an accepted seed
    uint16_t*       seed_off;       // in-block offset
    uint32_t*       blk_id;
    uint32_t*       blk_seg;        // index (relative to hit_base) of the block's first seed
    uint16_t*       blk_cnt;        // current score (mutable in k_candidates)
    uint32_t*       elig_tmp;
    uint32_t*       elig_idx;
    int32_t*        ch_q;           // chain seeds
    int32_t*        ch_s;
    int32_t*        dp_f;
    int32_t*        dp_p;
    int32_t*        dp_t;
    int32_t*        dp_v;
    unsigned long long* sort_scratch;
    // output
    nb_candidate*   out;
    uint32_t*       out_unit;
    uint32_t*       out_seq;
    uint32_t*       unit_ncand;     // [n_units + 1] candidates emitted per unit; after k_cand_scan: where the unit's candidates start
    nb_candidate*   out_sorted;     // the candidates by unit, emission order inside (FWD before REV of a read: units 2 r, 2 r + 1)
    int             qid_add, sid_add;   // global ids (pm_worker.c:164-167)
    unsigned long long* counters;   // [0] candidates, [1] seeds, [2] blocks, [3] eligible
};

// ------------------------------------------------------------------ k_probe

__device__ __forceinline__ uint32_t smem_kmer_at(const uint32_t* words, uint32_t p, int k)
{
    uint32_t a = words[p >> 4], b = words[(p >> 4) + 1];
    a = __byte_perm(a, 0, 0x0123);
    b = __byte_perm(b, 0, 0x0123);
    uint64_t v = ((uint64_t)a << 32) | b;
    int sh = 64 - 2 * (int)(p & 15) - 2 * k;
    return (uint32_t)(v >> sh) & (uint32_t)((1ull << (2 * k)) - 1);
}

__global__ void __launch_bounds__(PROBE_THREADS)
k_probe(MapDev P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_hits;
    const uint32_t u = blockIdx.x;
    const uint64_t r = P.read_lo + (u >> 1);
    const int strand = u & 1;
    const uint32_t L = P.rd_len[r];
    const uint64_t off = P.rd_off[r];
    const int k = P.k, z = P.z;
    const uint32_t nprobe = (uint32_t)(P.probe_base[u + 1] - P.probe_base[u]);
    if (threadIdx.x == 0) s_hits = 0;
    if (nprobe == 0) {
        if (threadIdx.x == 0) P.unit_hits[u] = 0;
        return;
    }
    // stage the packed read: 16-byte aligned window [byte0, byte0 + nbytes)
    const uint64_t byte0 = (off >> 2) & ~15ull;
    const uint32_t nbytes = (uint32_t)(((((off + L + 3) >> 2) - byte0) + 8 + 15) & ~15ull);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(smem_raw);
    if (nbytes <= P.stage_bytes) {
        if (threadIdx.x == 0) nb_mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            nb_mbar_expect_tx(&bar, nbytes);
            nb_tma_load_1d(smem_raw, P.rd_pac + byte0, nbytes, &bar);
        }
        nb_mbar_wait(&bar, 0);
    } else {
        // a read beyond the stage (> ~800 kb): its k-mers come straight from global memory (16-byte aligned window; the
        // volume ends in a zero pad, so the second word of the last k-mer exists)
        words = reinterpret_cast<const uint32_t*>(P.rd_pac + byte0);
        __syncthreads();
    }
    const uint32_t rel0 = (uint32_t)(off - (byte0 << 2));

    uint64_t soff_max = ~0ull;
    if (P.same_volume && (u >> 1) + P.read_lo < P.ref_nseq) soff_max = P.ref_off[r];

    const uint64_t pb = P.probe_base[u];
    uint32_t local = 0;
    for (uint32_t i = threadIdx.x; i < nprobe; i += PROBE_THREADS) {
        uint32_t h;
        if (strand == 0) h = smem_kmer_at(words, rel0 + i * z, k);
        else             h = nb_revcomp_kmer(smem_kmer_at(words, rel0 + L - i * z - k, k), k);
        uint32_t lo = __ldg(P.tbl + h), hi = __ldg(P.tbl + h + 1);
        uint32_t n = hi - lo;
        if (n && soff_max != ~0ull) {
            // positions ascend inside a run: hits below soff_max are a prefix
            uint32_t a = 0, b = n;
            while (a < b) {
                uint32_t m = (a + b) >> 1;
                if ((uint64_t)__ldg(P.offs + lo + m) < soff_max) a = m + 1; else b = m;
            }
            n = a;
        }
        P.probe_lo[pb + i] = lo;
        P.probe_n[pb + i] = n;
        local += n;
    }
    // warp-aggregated add into the CTA total
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_hits, local);
    __syncthreads();
    if (threadIdx.x == 0) P.unit_hits[u] = s_hits;
}

// ----------------------------------------------------------- k_build_blocks

__device__ __forceinline__ uint32_t cta_exclusive_scan_u32(uint32_t v, uint32_t* warp_sums, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? warp_sums[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(BUILD_THREADS)
k_build_blocks(MapDev P)
{
    __shared__ unsigned long long s_keys[SORT_CAP];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_nelig;
    const uint32_t u = blockIdx.x;
    const uint32_t n = P.unit_hits[u];
    if (threadIdx.x == 0) { P.unit_nblk[u] = 0; P.unit_nelig[u] = 0; s_nelig = 0; }
    if (n == 0) return;
    const uint64_t hb = P.hit_base[u];
    const uint64_t pb = P.probe_base[u];
    const uint32_t nprobe = (uint32_t)(P.probe_base[u + 1] - pb);
    const uint32_t bs = (uint32_t)P.bs;
    uint32_t Ppad = 32;
    while (Ppad < n) Ppad <<= 1;
    unsigned long long* keys = (n <= SORT_CAP) ? s_keys : (P.sort_scratch + P.scratch_base[u]);

    // 1. gather hits as keys, probe-major (= the order the reference visits them)
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nprobe; base += BUILD_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t cnt = i < nprobe ? P.probe_n[pb + i] : 0;
        uint32_t ex = cta_exclusive_scan_u32(cnt, s_warp, &s_total);
        if (cnt) {
            uint32_t lo = P.probe_lo[pb + i];
            uint32_t dst = carry + ex;
            for (uint32_t t = 0; t < cnt; ++t) {
                uint32_t pos = __ldg(P.offs + lo + t);
                uint32_t b = pos / bs;
                keys[dst + t] = ((unsigned long long)b << 40) | ((unsigned long long)i << 16) | (pos - b * bs);     // block 24 | probe 24 | offset 16 bits
            }
        }
        carry += s_total;
        __syncthreads();
    }
    for (uint32_t i = n + threadIdx.x; i < Ppad; i += BUILD_THREADS) keys[i] = ~0ull;
    __syncthreads();

    // 2. bitonic sort: (block, probe, offset) ascending
    for (uint32_t kk = 2; kk <= Ppad; kk <<= 1)
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < Ppad; i += BUILD_THREADS) {
                uint32_t l = i ^ j;
                if (l > i) {
                    bool up = (i & kk) == 0;
                    unsigned long long x = keys[i], y = keys[l];
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }

    // 3. cut into blocks; the thread owning a block head selects its seeds
    uint32_t nblk = 0;
    uint32_t seeds_local = 0;
    for (uint32_t base = 0; base < n; base += BUILD_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t head = 0;
        if (i < n) head = (i == 0) || ((uint32_t)(keys[i] >> 40) != (uint32_t)(keys[i - 1] >> 40));
        uint32_t ex = cta_exclusive_scan_u32(head, s_warp, &s_total);
        if (head) {
            const uint32_t j = nblk + ex;
            const uint32_t b = (uint32_t)(keys[i] >> 40);
            uint32_t acc = 0, last_probe = 0xffffffffu;
            for (uint32_t t = i; t < n && acc < BLK_SEEDS; ++t) {
                unsigned long long key = keys[t];
                if ((uint32_t)(key >> 40) != b) break;
                uint32_t probe = (uint32_t)(key >> 16) & 0xffffffu;
                if (probe == last_probe) continue;          // one seed per probe per block (:91)
                P.seed_kid[hb + i + acc] = probe + 1;
                P.seed_off[hb + i + acc] = (uint16_t)(key & 0xffffu);
                last_probe = probe;
                ++acc;
            }
            P.blk_id[hb + j] = b;
            P.blk_seg[hb + j] = i;
            P.blk_cnt[hb + j] = (uint16_t)acc;
            seeds_local += acc;
        }
        nblk += s_total;
        __syncthreads();
    }
    __syncthreads();

    // 4. static gates + first-touch order of the eligible blocks
    const int s = P.s;
    for (uint32_t j = threadIdx.x; j < nblk; j += BUILD_THREADS) {
        const uint32_t cnt = P.blk_cnt[hb + j];
        if ((int)cnt < s) continue;
        const uint32_t seg = P.blk_seg[hb + j];
        uint32_t pair = cnt;
        if (j > 0 && P.blk_id[hb + j - 1] + 1 == P.blk_id[hb + j]) {
            // score of the left neighbour as of this block's last accepted insertion (:103):
            // its seeds whose probe number is <= this block's last probe number
            const uint32_t last_kid = P.seed_kid[hb + seg + cnt - 1];
            const uint32_t pseg = P.blk_seg[hb + j - 1], pcnt = P.blk_cnt[hb + j - 1];
            for (uint32_t t = 0; t < pcnt; ++t) pair += P.seed_kid[hb + pseg + t] <= last_kid;
        }
        if ((int)pair >= 2 * s) P.elig_tmp[hb + atomicAdd(&s_nelig, 1u)] = j;
    }
    __syncthreads();
    const uint32_t ne = s_nelig;
    for (uint32_t e = threadIdx.x; e < ne; e += BUILD_THREADS) {
        const uint32_t j = P.elig_tmp[hb + e];
        const uint32_t seg = P.blk_seg[hb + j];
        const unsigned long long key =
            ((unsigned long long)P.seed_kid[hb + seg] << 32) | ((unsigned long long)P.blk_id[hb + j] * bs + P.seed_off[hb + seg]);
        uint32_t rank = 0;
        for (uint32_t o = 0; o < ne; ++o) {
            const uint32_t jo = P.elig_tmp[hb + o];
            const uint32_t sego = P.blk_seg[hb + jo];
            const unsigned long long ko =
                ((unsigned long long)P.seed_kid[hb + sego] << 32) | ((unsigned long long)P.blk_id[hb + jo] * bs + P.seed_off[hb + sego]);
            rank += ko < key;
        }
        P.elig_idx[hb + rank] = j;
    }
    if (threadIdx.x == 0) {
        P.unit_nblk[u] = nblk;
        P.unit_nelig[u] = ne;
        atomicAdd(P.counters + 2, (unsigned long long)nblk);
        atomicAdd(P.counters + 3, (unsigned long long)ne);
    }
    for (int d = 16; d > 0; d >>= 1) seeds_local += __shfl_xor_sync(0xffffffffu, seeds_local, d);
    if ((threadIdx.x & 31) == 0 && seeds_local) atomicAdd(P.counters + 1, (unsigned long long)seeds_local);
}

// ------------------------------------------------------------- k_candidates

// word_finder.c:145: int / (int * float) in single precision, compared in double
__device__ __forceinline__ bool ddf_close(int dloc, int dkid, float window)
{
    float ratio = __fdiv_rn((float)dloc, __fmul_rn((float)dkid, window));
    return fabs(__dsub_rn((double)ratio, 1.0)) < 0.25;
}

__device__ __forceinline__ int ilog2_u32(uint32_t v) { return 31 - __clz(v); }

// first index in blk_id[0..nblk) with id >= b
__device__ __forceinline__ uint32_t blk_lower_bound(const uint32_t* ids, uint32_t nblk, int64_t b)
{
    uint32_t lo = 0, hi = nblk;
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if ((int64_t)ids[m] < b) lo = m + 1; else hi = m;
    }
    return lo;
}

struct UnitView {
    uint64_t  hb;
    uint32_t  nblk;
    uint32_t* blk_id;
    uint32_t* blk_seg;
    uint16_t* blk_cnt;
    uint32_t* seed_kid;
    uint16_t* seed_off;
};

// chain_dp.c:37-159 executed by one lane; returns through *best / *last the first and
// last chain in (score desc, qoff, soff) order and whether some chain contains the seed.
struct Chain { int score; int qbeg, qend, sbeg, send; };

__device__ void chain_lane0(const int32_t* cq, const int32_t* cs, int n, int32_t* f, int32_t* p, int32_t* t, int32_t* v,
                            int k, int min_cnt, int seed_q, int seed_s, int* n_out, Chain* best, Chain* last,
                            bool* any_contains)
{
    const int max_dist = 5000, bw = 500, max_skip = 25, min_sc = 30;
    *n_out = 0;
    *any_contains = false;
    for (int i = 0; i < n; ++i) { f[i] = 0; p[i] = -1; t[i] = 0; v[i] = 0; }
    int st = 0;
    for (int i = 0; i < n; ++i) {
        const int ri = cs[i], qi = cq[i];
        int best_j = -1, best_f = k, n_skip = 0;
        while (st < i && ri - cs[st] > max_dist) ++st;
        for (int j = i - 1; j >= st; --j) {
            const int sj = cs[j], qj = cq[j];
            if (ri <= sj || qi <= qj || qi - qj > max_dist) continue;
            const int dr = ri - sj, dq = qi - qj;
            const int dd = dr > dq ? dr - dq : dq - dr;
            if (dd > bw) continue;
            const int md = dq < dr ? dq : dr;
            int sc = md < k ? md : k;
            const int lg = dd ? ilog2_u32((uint32_t)dd) : 0;
            sc -= (int)__dmul_rn(__dmul_rn((double)dd, 0.01), (double)k) + (lg >> 1);
            sc += f[j];
            if (sc > best_f) {
                best_f = sc; best_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = best_f; p[i] = best_j;
        v[i] = (best_j >= 0 && v[best_j] > best_f) ? v[best_j] : best_f;
    }
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    // Chain ends in (score desc, index asc) order.  The list is produced lazily by
    // repeated selection so that no extra array is needed: `v` is no longer read after
    // the peak walk, so the peak index / score of each end are parked in v/t.
    // t[i]: 0 = chain end candidate, 1 = has successor.  Encode ends: v[i] = peak f, and
    // remember the peak index in a second pass array: reuse f? f is still needed (f[j] at
    // cut points), so park the peak index in t as -(idx+2) (t==0/1 meaning stays distinct).
    int n_u = 0;
    for (int i = 0; i < n; ++i) {
        if (t[i] == 0 && v[i] >= min_sc) {
            int j = i;
            while (j >= 0 && f[j] < v[j]) j = p[j];
            if (j < 0) j = i;
            t[i] = -(j + 2);
            ++n_u;
        }
    }
    if (!n_u) return;
    // `used` marks live in v from here on (v's DP role is over): v[i] = 1 once seed i is on an emitted walk.
    // Peak indices are kept in t (negative codes) until their chain end is consumed.
    // First clear v, but the peak scores are f[peak], still available.
    for (int i = 0; i < n; ++i) v[i] = 0;
    int n_v = 0, kk = 0;
    bool have_best = false;
    for (int it = 0; it < n_u; ++it) {
        // select the not-yet-consumed end with the largest (score, then smallest peak index)
        int sel = -1, sel_sc = 0, sel_j = 0;
        for (int i = 0; i < n; ++i) {
            if (t[i] < -1) {
                int j = -t[i] - 2;
                int sc = f[j];
                if (sel < 0 || sc > sel_sc || (sc == sel_sc && j < sel_j)) { sel = i; sel_sc = sc; sel_j = j; }
            }
        }
        t[sel] = 0;     // consumed
        const int n_v0 = n_v, k0 = kk;
        int j = sel_j;
        const int qend = cq[j] + k, send = cs[j] + k;
        int first;
        do { first = j; ++n_v; v[j] = 1; j = p[j]; } while (j >= 0 && v[j] == 0);
        bool emit = false; int score = 0;
        if (j < 0) { emit = true; score = sel_sc; }
        else if (sel_sc - f[j] >= min_sc) { emit = true; score = sel_sc - f[j]; }
        if (emit && n_v - n_v0 >= min_cnt) {
            Chain c; c.score = score; c.qbeg = cq[first]; c.sbeg = cs[first]; c.qend = qend; c.send = send;
            // order (chain_dp.c:26-33): score desc, qoff (= qend) asc, soff (= send) asc
            bool better = !have_best || c.score > best->score ||
                          (c.score == best->score && (c.qend < best->qend || (c.qend == best->qend && c.send < best->send)));
            bool worse = !have_best || c.score < last->score ||
                         (c.score == last->score && (c.qend > last->qend || (c.qend == last->qend && c.send >= last->send)));
            if (better) *best = c;
            if (worse) *last = c;
            have_best = true;
            if (seed_q >= c.qbeg && seed_q < c.qend && seed_s >= c.sbeg && seed_s < c.send) *any_contains = true;
            ++kk;
        }
        if (k0 == kk) n_v = n_v0;
    }
    *n_out = kk;
}

__global__ void __launch_bounds__(CAND_WARPS * 32)
k_candidates(MapDev P)
{
    __shared__ int s_loc[CAND_WARPS][2 * BLK_SEEDS];
    __shared__ int s_kid[CAND_WARPS][2 * BLK_SEEDS];
    __shared__ int s_votes[CAND_WARPS][2 * BLK_SEEDS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t u = blockIdx.x * CAND_WARPS + w;
    if (u >= P.n_units) return;
    const uint32_t nelig = P.unit_nelig[u];
    if (nelig == 0) return;                                     // (unit_ncand was cleared)
    const uint64_t hb = P.hit_base[u];
    UnitView U;
    U.hb = hb; U.nblk = P.unit_nblk[u];
    U.blk_id = P.blk_id + hb; U.blk_seg = P.blk_seg + hb; U.blk_cnt = P.blk_cnt + hb;
    U.seed_kid = P.seed_kid + hb; U.seed_off = P.seed_off + hb;
    int32_t* cq = P.ch_q + hb; int32_t* cs = P.ch_s + hb;
    int32_t* dpf = P.dp_f + hb; int32_t* dpp = P.dp_p + hb; int32_t* dpt = P.dp_t + hb; int32_t* dpv = P.dp_v + hb;
    const uint32_t* elig = P.elig_idx + hb;
    int* loc = s_loc[w]; int* kid = s_kid[w]; int* votes = s_votes[w];

    const uint64_t r = P.read_lo + (u >> 1);
    const int qdir = u & 1;
    const int qsize = (int)P.rd_len[r];
    const int bs = P.bs, z = P.z, s = P.s, k = P.k;
    const float window = (float)z;
    uint32_t emitted = 0;

    for (uint32_t e = 0; e < nelig; ++e) {
        const uint32_t j = elig[e];
        if ((int)U.blk_cnt[j] < s) continue;                    // current score gate (:394)
        // ---- pool the seeds of the left neighbour (if it still has a score) and of this block
        int n = 0, shift = 0;
        int64_t pool_start = (int64_t)bs * U.blk_id[j];
        const bool has_prev = j > 0 && U.blk_id[j - 1] + 1 == U.blk_id[j] && U.blk_cnt[j - 1] > 0;
        if (has_prev) {
            const int pc = U.blk_cnt[j - 1];
            const uint32_t ps = U.blk_seg[j - 1];
            for (int i = lane; i < pc; i += 32) { kid[i] = (int)U.seed_kid[ps + i]; loc[i] = U.seed_off[ps + i]; }
            n = pc; shift = bs;
            pool_start -= bs;
        }
        {
            const int cc = U.blk_cnt[j];
            const uint32_t sg = U.blk_seg[j];
            for (int i = lane; i < cc; i += 32) { kid[n + i] = (int)U.seed_kid[sg + i]; loc[n + i] = (int)U.seed_off[sg + i] + shift; }
            n += cc;
        }
        for (int i = lane; i < n; i += 32) votes[i] = 0;
        __syncwarp();
        // ---- pairwise DDF voting (scoring_seeds :144-145); lanes own rows i
        for (int i = lane; i + 1 < n; i += 32) {
            int last = kid[i], mine = 0;
            const int ki = kid[i], li = loc[i];
            for (int jj = i + 1; jj < n; ++jj) {
                const int kj = kid[jj];
                if (last == kj) continue;
                const int dk = kj - ki, dl = loc[jj] - li;
                if (dk > 0 && dl > 0 && dl < qsize && ddf_close(dl, dk, window)) {
                    ++mine;
                    atomicAdd(&votes[jj], 1);
                    last = kj;
                }
            }
            if (mine) atomicAdd(&votes[i], mine);
        }
        __syncwarp();
        // ---- winner, tie rule and representative seed (:147-167), lane 0
        int rep = -1;
        if (lane == 0) {
            int best = 0, besti = 0, ties = 0;
            for (int i = 0; i < n; ++i) {
                if (votes[i] > best) { best = votes[i]; besti = i; ties = 0; }
                else if (votes[i] == best) ++ties;
            }
            if (best >= 5) {
                if (ties == best) rep = besti;
                else {
                    int r_loc = 0, r_idx = -1;
                    for (int jj = 0; jj < besti; ++jj) {
                        const int dk = kid[besti] - kid[jj], dl = loc[besti] - loc[jj];
                        if (dk > 0 && dl > 0 && dl < qsize && ddf_close(dl, dk, window))
                            if (r_loc == 0) { r_loc = loc[jj]; r_idx = jj; }
                    }
                    if (r_loc == 0) { r_loc = loc[besti]; r_idx = besti; }
                    for (int jj = besti + 1; jj < n; ++jj) {
                        const int dk = kid[jj] - kid[besti], dl = loc[jj] - loc[besti];
                        if (dk > 0 && dl > 0 && dl <= qsize && ddf_close(dl, dk, window))
                            if (r_loc == 0) { r_loc = loc[jj]; r_idx = jj; }
                    }
                    rep = r_idx;
                }
                if (votes[rep] < 2 * s) rep = -1;               // :228
            }
        }
        rep = __shfl_sync(0xffffffffu, rep, 0);
        if (rep < 0) continue;
        const int r_loc = loc[rep], r_kid = kid[rep];
        __syncwarp();

        // ---- geometry of the seed (:230-245)
        int64_t seed_t = r_loc + pool_start;
        const int seed_q = (r_kid - 1) * z;
        const int seed_b = (int)(seed_t / bs);
        // pdb_offset_to_id: last read starting at or before seed_t
        uint32_t tid;
        {
            uint32_t lo = 0, hi = P.ref_nseq;
            while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if ((int64_t)P.ref_off[m] <= seed_t) lo = m; else hi = m; }
            tid = lo;
        }
        const int tsize = (int)P.ref_len[tid];
        const int64_t tstart = (int64_t)P.ref_off[tid], tend = tstart + tsize;
        const int seed_s = (int)(seed_t - tstart);
        int L = seed_s < seed_q ? seed_s : seed_q;
        int b_lo = seed_b - L / bs - 1;
        if (b_lo < 0) b_lo = 0;
        const int tr = tsize - seed_s, qr = qsize - seed_q;
        L = tr < qr ? tr : qr;
        const int b_hi = seed_b + (L + bs - 1) / bs;

        // ---- gather DDF-consistent seeds left and right of the seed (:249-307)
        int n_ch = 0, support = 0;
        const uint32_t j_lo = blk_lower_bound(U.blk_id, U.nblk, b_lo);
        uint32_t j_seed = j_lo;     // index of the seed's block (it exists: the seed lives there)
        for (int side = 0; side < 2; ++side) {
            uint32_t jb = side == 0 ? j_lo : j_seed;
            if (side == 1) {
                if (lane == 0) { cq[n_ch] = seed_q; cs[n_ch] = seed_s; }
                ++n_ch;
            }
            for (; jb < U.nblk; ++jb) {
                const int b = (int)U.blk_id[jb];
                if (side == 0 ? b > seed_b : b > b_hi) break;
                if (b == seed_b) j_seed = jb;
                const int cnt = U.blk_cnt[jb];
                if (!cnt) continue;
                const uint32_t sg = U.blk_seg[jb];
                int rel = 0;
                for (int base = 0; base < cnt; base += 32) {
                    const int t = base + lane;
                    bool ok = false;
                    int qo = 0, so = 0;
                    if (t < cnt) {
                        const int64_t to = (int64_t)bs * b + U.seed_off[sg + t];
                        qo = ((int)U.seed_kid[sg + t] - 1) * z;
                        if (side == 0) {
                            if (to >= tstart) {
                                so = (int)(to - tstart);
                                if (so < seed_s && qo < seed_q) {
                                    double d = __dsub_rn(__ddiv_rn((double)(seed_s - so), (double)(seed_q - qo)), 1.0);
                                    ok = fabs(d) < 0.25;
                                }
                            }
                        } else {
                            if (to < tend && to >= tstart) {
                                so = (int)(to - tstart);
                                if (so > seed_s && qo > seed_q) {
                                    double d = __dsub_rn(__ddiv_rn((double)(so - seed_s), (double)(qo - seed_q)), 1.0);
                                    ok = fabs(d) < 0.25;
                                }
                            }
                        }
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int dst = n_ch + __popc(m & ((1u << lane) - 1));
                        cq[dst] = qo; cs[dst] = so;
                    }
                    const int c = __popc(m);
                    n_ch += c; rel += c;
                }
                if (b != seed_b && __ddiv_rn((double)rel, (double)cnt) >= 0.4) {
                    __syncwarp();
                    if (lane == 0) U.blk_cnt[jb] = 0;       // :273,305
                }
                support += rel;
            }
            __syncwarp();
        }
        __syncwarp();
        // ---- order chain seeds by (soff, qoff): rank sort through the dp scratch
        for (int i = lane; i < n_ch; i += 32) {
            const int si = cs[i], qi = cq[i];
            int rank = 0;
            for (int o = 0; o < n_ch; ++o) {
                const int so = cs[o], qo = cq[o];
                rank += (so < si) || (so == si && (qo < qi || (qo == qi && o < i)));
            }
            dpf[rank] = qi; dpp[rank] = si;
        }
        __syncwarp();
        for (int i = lane; i < n_ch; i += 32) { cq[i] = dpf[i]; cs[i] = dpp[i]; }
        __syncwarp();

        // ---- chain and pick (:312-357), lane 0
        int take = 0;
        Chain pick;
        pick.score = 0; pick.qbeg = pick.qend = pick.sbeg = pick.send = 0;
        if (lane == 0) {
            int n_chain; Chain best, last; bool any;
            chain_lane0(cq, cs, n_ch, dpf, dpp, dpt, dpv, k, s, seed_q, seed_s, &n_chain, &best, &last, &any);
            if (n_chain) {
                const bool best_has = seed_q >= best.qbeg && seed_q < best.qend && seed_s >= best.sbeg && seed_s < best.send;
                if (best_has) { pick = best; take = 1; }
                else if (any) { pick = last; take = 1; }        // reference emits the last chain scanned (:337-339)
                else if (best.qend - best.qbeg >= 5000) { pick = best; take = 1; }
            }
        }
        take = __shfl_sync(0xffffffffu, take, 0);
        if (!take) continue;
        pick.qbeg = __shfl_sync(0xffffffffu, pick.qbeg, 0);
        pick.qend = __shfl_sync(0xffffffffu, pick.qend, 0);
        pick.sbeg = __shfl_sync(0xffffffffu, pick.sbeg, 0);
        pick.send = __shfl_sync(0xffffffffu, pick.send, 0);
        // ---- zero every block the chain spans (clear_block_scores :171-181)
        {
            const int64_t zb_lo = (pick.sbeg + tstart) / bs, zb_hi = (pick.send + tstart) / bs;
            const uint32_t ja = blk_lower_bound(U.blk_id, U.nblk, zb_lo);
            for (uint32_t jb = ja + lane; jb < U.nblk && (int64_t)U.blk_id[jb] <= zb_hi; jb += 32) U.blk_cnt[jb] = 0;
        }
        __syncwarp();
        const bool keep = (pick.send - pick.sbeg >= P.a_cut) || (pick.qend - pick.qbeg >= P.a_cut);
        if (keep && lane == 0) {
            const unsigned long long slot = atomicAdd(P.counters + 0, 1ull);
            nb_candidate c;
            c.qid = (int)(u >> 1) + (int)P.read_lo;     // local id in the query volume
            c.sid = (int)tid;
            c.qdir = qdir; c.sdir = 0; c.score = support;
            c.qbeg = pick.qbeg; c.qend = pick.qend; c.qsize = qsize;
            c.sbeg = pick.sbeg; c.send = pick.send; c.ssize = tsize;
            c.qoff = seed_q; c.soff = seed_s;
            P.out[slot] = c;
            P.out_unit[slot] = u;
            P.out_seq[slot] = emitted;
        }
        if (keep) ++emitted;
    }
    if (lane == 0) P.unit_ncand[u] = emitted;
}

// The candidates leave k_candidates in the order the warps got their output slots.  What the caller wants is per read, FWD
// strand before REV, emission order inside (pm_worker.c:163-186): an exclusive scan of the per-unit counts (one CTA) and a
// scatter put them there, with the global read ids added, so that the host copies them out as they are.
__global__ void __launch_bounds__(1024)
k_cand_scan(uint32_t* __restrict__ unit_ncand, uint32_t n_units)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t per = (n_units + 1023) / 1024;
    const uint32_t i0 = threadIdx.x * per, i1 = i0 + per < n_units ? i0 + per : n_units;
    uint32_t sum = 0;
    for (uint32_t i = i0; i < i1; ++i) sum += unit_ncand[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    uint32_t run = s_warp[wid] + inc - sum;
    for (uint32_t i = i0; i < i1; ++i) { const uint32_t c = unit_ncand[i]; unit_ncand[i] = run; run += c; }
    if (threadIdx.x == 0) unit_ncand[n_units] = s_warp[32];
}

__global__ void k_cand_place(MapDev P, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nb_candidate c = P.out[i];
    c.qid += P.qid_add; c.sid += P.sid_add;
    P.out_sorted[P.unit_ncand[P.out_unit[i]] + P.out_seq[i]] = c;
}

// ------------------------------------------------------------------- host

// Candidates of the reads [read_lo, read_hi) appended to `acc`.  *too_big: the seed hits of the range need more scratch than
// hit_budget allows and the range holds more than one read -- nothing was appended, the caller splits the range.
static int nb_map_range(nb_ctx* ctx, const nb_index* idx, const nb_volume* ref, int ref_start_id,
                        const nb_volume* reads, int read_start_id, uint64_t read_lo, uint64_t read_hi,
                        const nb_map_options* opt, uint64_t hit_budget, std::vector<nb_candidate>& acc, bool* too_big)
{
    *too_big = false;
    cudaStream_t st = ctx->stream;
    const bool trace = getenv("NB_TRACE") != NULL;
    std::chrono::steady_clock::time_point tr0 = std::chrono::steady_clock::now();
    double tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto lap = [&](int i) { auto t1 = std::chrono::steady_clock::now(); tr[i] += std::chrono::duration<double, std::milli>(t1 - tr0).count(); tr0 = t1; };
    const uint64_t nreads = read_hi - read_lo;
    if (nreads == 0) return NB_OK;
    if (nreads * 2 >= (1ull << 31)) return nb_fail(ctx, NB_ELIMIT, "too many reads in one call");
    const uint32_t U = (uint32_t)(nreads * 2);
    const int k = opt->kmer_size, z = opt->scan_window;

    // ---- host: probe layout
    std::vector<uint64_t> probe_base(U + 1);
    uint64_t np = 0, qbases = 0;
    for (uint32_t u = 0; u < U; ++u) {
        probe_base[u] = np;
        uint32_t L = reads->h_len[read_lo + (u >> 1)];
        uint32_t n = L >= (uint32_t)k ? (L - k) / z + 1 : 0;
        if (n > 0xffffffu) return nb_fail(ctx, NB_ELIMIT, "read %llu needs %u probes (> 2^24 - 1)", (unsigned long long)(read_lo + (u >> 1)), n);
        np += n;
        qbases += L;
    }
    probe_base[U] = np;
    const uint32_t max_len = reads->max_len;
    size_t probe_smem = ((size_t)max_len / 4 + 64 + 15) & ~(size_t)15;
    if (probe_smem > 200 * 1024) probe_smem = 200 * 1024;          // longer reads are probed from global memory
    if (getenv("NB_MAP_STAGE")) probe_smem = std::min<size_t>(probe_smem, (strtoull(getenv("NB_MAP_STAGE"), NULL, 10) + 15) & ~(size_t)15);      // (tests)
    // sort keys of k_build_blocks: block id and probe number of a hit in 24 bits each
    if (((uint64_t)ref->db_size + (uint64_t)opt->block_size - 1) / (uint64_t)opt->block_size >= (1ull << 24))
        return nb_fail(ctx, NB_ELIMIT, "block size %d gives this volume more than 2^24 blocks", opt->block_size);

    NbDevBuf S(ctx);
    MapDev P;
    memset(&P, 0, sizeof P);
    P.tbl = idx->d_tbl; P.offs = idx->d_off;
    P.ref_off = ref->d_off; P.ref_len = ref->d_len; P.ref_nseq = (uint32_t)ref->nseq;
    P.rd_pac = reads->d_pac; P.rd_off = reads->d_off; P.rd_len = reads->d_len;
    P.read_lo = read_lo; P.n_units = U; P.stage_bytes = (uint32_t)probe_smem;
    P.k = k; P.z = z; P.bs = opt->block_size; P.s = opt->block_score_cutoff; P.a_cut = opt->align_size_cutoff;
    // word_finder.c:121-127: the self-hit cut applies when the query's global id lies in the reference volume
    {
        int64_t g_lo = (int64_t)read_lo + read_start_id, g_hi = (int64_t)read_hi + read_start_id;
        P.same_volume = (g_lo >= ref_start_id && g_hi <= (int64_t)ref_start_id + (int64_t)ref->nseq) ? 1 : 0;
        if (!P.same_volume && g_hi > ref_start_id && g_lo < (int64_t)ref_start_id + (int64_t)ref->nseq)
            return nb_fail(ctx, NB_EINVAL, "query id range partially overlaps the reference volume");
    }
    int rc;
    uint64_t* d_probe_base; uint64_t* d_hit_base; uint64_t* d_scratch_base;
    if ((rc = S.get(&d_probe_base, U + 1))) return rc;
    if ((rc = S.get(&d_hit_base, U + 1))) return rc;
    if ((rc = S.get(&d_scratch_base, U))) return rc;
    if ((rc = S.get(&P.unit_hits, U))) return rc;
    if ((rc = S.get(&P.unit_nblk, U))) return rc;
    if ((rc = S.get(&P.unit_nelig, U))) return rc;
    if ((rc = S.get(&P.probe_lo, np))) return rc;
    if ((rc = S.get(&P.probe_n, np))) return rc;
    if ((rc = S.get(&P.counters, 8))) return rc;
    if ((rc = S.get(&P.unit_ncand, (uint64_t)U + 1))) return rc;
    NB_CUDA(ctx, cudaMemsetAsync(P.unit_ncand, 0, sizeof(uint32_t) * ((size_t)U + 1), st));
    P.qid_add = read_start_id; P.sid_add = ref_start_id;
    P.probe_base = d_probe_base; P.hit_base = d_hit_base; P.scratch_base = d_scratch_base;
    NB_CUDA(ctx, cudaMemsetAsync(P.counters, 0, 8 * sizeof(unsigned long long), st));
    NB_CUDA(ctx, cudaMemcpyAsync(d_probe_base, probe_base.data(), sizeof(uint64_t) * (U + 1), cudaMemcpyHostToDevice, st));

    lap(0);
    cudaEventRecord(ctx->ev0, st);
    if (probe_smem > 48 * 1024)
        NB_CUDA(ctx, cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem));
    k_probe<<<U, PROBE_THREADS, probe_smem, st>>>(P);
    NB_LAUNCH_CHECK(ctx);

    // ---- host: hit layout (one D2H of U counters, the exclusive scan is trivial work)
    std::vector<uint32_t> unit_hits(U);
    NB_CUDA(ctx, cudaMemcpyAsync(unit_hits.data(), P.unit_hits, sizeof(uint32_t) * U, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    lap(1);
    std::vector<uint64_t> hit_base(U + 1), scratch_base(U);
    uint64_t H = 0, scratch = 0;
    for (uint32_t u = 0; u < U; ++u) {
        hit_base[u] = H;
        H += unit_hits[u];
        scratch_base[u] = ~0ull;
        if (unit_hits[u] > SORT_CAP) {
            uint64_t pp = 32;
            while (pp < unit_hits[u]) pp <<= 1;
            scratch_base[u] = scratch;
            scratch += pp;
        }
    }
    hit_base[U] = H;
    if (H + scratch > hit_budget && nreads > 1) { *too_big = true; return NB_OK; }
    NB_CUDA(ctx, cudaMemcpyAsync(d_hit_base, hit_base.data(), sizeof(uint64_t) * (U + 1), cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemcpyAsync(d_scratch_base, scratch_base.data(), sizeof(uint64_t) * U, cudaMemcpyHostToDevice, st));
    if ((rc = S.get(&P.seed_kid, H))) return rc;
    if ((rc = S.get(&P.seed_off, H))) return rc;
    if ((rc = S.get(&P.blk_id, H))) return rc;
    if ((rc = S.get(&P.blk_seg, H))) return rc;
    if ((rc = S.get(&P.blk_cnt, H))) return rc;
    if ((rc = S.get(&P.elig_tmp, H))) return rc;
    if ((rc = S.get(&P.elig_idx, H))) return rc;
    if ((rc = S.get(&P.ch_q, H))) return rc;
    if ((rc = S.get(&P.ch_s, H))) return rc;
    if ((rc = S.get(&P.dp_f, H))) return rc;
    if ((rc = S.get(&P.dp_p, H))) return rc;
    if ((rc = S.get(&P.dp_t, H))) return rc;
    if ((rc = S.get(&P.dp_v, H))) return rc;
    if ((rc = S.get(&P.sort_scratch, scratch))) return rc;

    lap(2);
    k_build_blocks<<<U, BUILD_THREADS, 0, st>>>(P);
    NB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev1, st);

    unsigned long long cnts[8];
    NB_CUDA(ctx, cudaMemcpyAsync(cnts, P.counters, sizeof cnts, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    lap(3);
    const uint64_t E = cnts[3];
    if ((rc = S.get(&P.out, E))) return rc;
    if ((rc = S.get(&P.out_unit, E))) return rc;
    if ((rc = S.get(&P.out_seq, E))) return rc;

    k_candidates<<<(U + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, st>>>(P);
    NB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev2, st);
    NB_CUDA(ctx, cudaMemcpyAsync(cnts, P.counters, sizeof cnts, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    lap(4);
    const uint64_t C = cnts[0];
    if (C >= (1ull << 32)) return nb_fail(ctx, NB_ELIMIT, "too many candidates in one piece");
    std::vector<uint32_t> cbase((size_t)U + 1);
    const size_t a0 = acc.size();
    acc.resize(a0 + C);
    cudaEventRecord(ctx->ev3, st);
    if (C) {
        if ((rc = S.get(&P.out_sorted, C))) return rc;
        k_cand_scan<<<1, 1024, 0, st>>>(P.unit_ncand, U);
        NB_LAUNCH_CHECK(ctx);
        k_cand_place<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(P, (uint32_t)C);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(ctx, cudaMemcpyAsync(acc.data() + a0, P.out_sorted, sizeof(nb_candidate) * C, cudaMemcpyDeviceToHost, st));
        NB_CUDA(ctx, cudaMemcpyAsync(cbase.data(), P.unit_ncand, sizeof(uint32_t) * ((size_t)U + 1), cudaMemcpyDeviceToHost, st));
    }
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev2); ctx->tm.map_ms += ms;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->tm.map_probe_ms += ms;
    ctx->tm.d2h_bytes += C * sizeof(nb_candidate) + U * 8;
    ctx->tm.h2d_bytes += (U + 1) * 16 + U * 8;
    lap(5);
    // ---- host: -n truncation of the reads that have more candidates than that (pm_worker.c:168-171, order of pm_worker.c:15-24)
    uint64_t nres = C;
    if (C) {
        bool over = false;
        for (uint32_t u = 0; u < U && !over; u += 2) over = cbase[u + 2] - cbase[u] > (uint32_t)opt->num_candidates;
        if (over) {
            std::vector<nb_candidate> kept;
            kept.reserve(C);
            std::vector<nb_candidate> one;
            for (uint32_t u = 0; u < U; u += 2) {
                const nb_candidate* f = acc.data() + a0 + cbase[u];
                const size_t m = cbase[u + 2] - cbase[u];
                if (m <= (size_t)opt->num_candidates) { kept.insert(kept.end(), f, f + m); continue; }
                one.assign(f, f + m);
                std::stable_sort(one.begin(), one.end(), [](const nb_candidate& a, const nb_candidate& b) {
                    if (a.score != b.score) return a.score > b.score;
                    if (a.qdir != b.qdir) return a.qdir < b.qdir;
                    if (a.sid != b.sid) return a.sid < b.sid;
                    if (a.qoff != b.qoff) return a.qoff < b.qoff;
                    return a.soff < b.soff;
                });
                kept.insert(kept.end(), one.begin(), one.begin() + opt->num_candidates);
            }
            nres = kept.size();
            acc.resize(a0 + nres);
            if (nres) memcpy(acc.data() + a0, kept.data(), sizeof(nb_candidate) * nres);
        }
    }
    lap(6);
    if (trace) fprintf(stderr, "[nb_map_range] wall ms: layout+alloc %.2f probe+sync %.2f hit layout+alloc %.2f blocks+sync %.2f candidates+sync %.2f download %.2f grouping %.2f\n",
                       tr[0], tr[1], tr[2], tr[3], tr[4], tr[5], tr[6]);
    ctx->mc.n_units += U; ctx->mc.n_probes += np; ctx->mc.n_hits += H; ctx->mc.n_seeds += cnts[1];
    ctx->mc.n_blocks += cnts[2]; ctx->mc.n_eligible += cnts[3]; ctx->mc.n_candidates += nres; ctx->mc.query_bases += qbases * 2;
    return NB_OK;
}

// The reads of the range are mapped in pieces: a piece holds at most kProbeChunk probes (the probe arrays are sized by it)
// and its seed hits (13 scratch arrays of one entry per hit) have to fit a share of the free device memory -- a piece that
// does not is cut in two, down to single reads.
extern "C" int nb_map_volume(nb_ctx* ctx, const nb_index* idx, const nb_volume* ref, int ref_start_id,
                             const nb_volume* reads, int read_start_id, uint64_t read_lo, uint64_t read_hi,
                             const nb_map_options* opt, nb_candidate** out, uint64_t* n_out)
{
    if (!ctx || !idx || !ref || !reads || !opt || !out || !n_out) return nb_fail(ctx, NB_EINVAL, "nb_map_volume: bad arguments");
    *out = NULL; *n_out = 0;
    if (read_hi > reads->nseq || read_lo > read_hi) return nb_fail(ctx, NB_EINVAL, "read range out of bounds");
    if (opt->kmer_size != idx->k) return nb_fail(ctx, NB_EINVAL, "kmer size %d differs from the index (%d)", opt->kmer_size, idx->k);
    if (opt->block_size < 1 || opt->block_size > 32767) return nb_fail(ctx, NB_ELIMIT, "block size must be in 1..32767 (reference uses short offsets)");
    if (opt->scan_window < 1) return nb_fail(ctx, NB_EINVAL, "scan window must be >= 1");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->mc, 0, sizeof ctx->mc);
    ctx->tm.map_ms = 0; ctx->tm.map_probe_ms = 0;
    const uint64_t kProbeChunk = getenv("NB_MAP_PROBES") ? strtoull(getenv("NB_MAP_PROBES"), NULL, 10) : (96ull << 20);
    uint64_t hit_budget;
    {
        size_t fr = 0, tot = 0;
        NB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
        hit_budget = (uint64_t)(fr / 2) / 64;               // ~52 bytes of scratch per hit, half of what is free
        if (getenv("NB_MAP_HITS")) hit_budget = strtoull(getenv("NB_MAP_HITS"), NULL, 10);      // (tests)
        if (hit_budget < 1024) hit_budget = 1024;
    }
    std::vector<nb_candidate> acc;
    std::vector<std::pair<uint64_t, uint64_t>> todo;        // stack of read ranges, lowest reads on top
    {
        std::vector<std::pair<uint64_t, uint64_t>> fwd;
        const int k = opt->kmer_size, z = opt->scan_window;
        uint64_t lo = read_lo, np = 0;
        for (uint64_t r = read_lo; r < read_hi; ++r) {
            const uint32_t L = reads->h_len[r];
            const uint64_t n = L >= (uint32_t)k ? 2ull * ((L - k) / z + 1) : 0;
            if (np + n > kProbeChunk && r > lo) { fwd.push_back({lo, r}); lo = r; np = 0; }
            np += n;
        }
        if (lo < read_hi) fwd.push_back({lo, read_hi});
        todo.assign(fwd.rbegin(), fwd.rend());
    }
    while (!todo.empty()) {
        const auto rg = todo.back();
        todo.pop_back();
        bool too_big = false;
        const int rc = nb_map_range(ctx, idx, ref, ref_start_id, reads, read_start_id, rg.first, rg.second, opt, hit_budget, acc, &too_big);
        if (rc) return rc;
        if (too_big) {
            const uint64_t mid = rg.first + (rg.second - rg.first) / 2;
            todo.push_back({mid, rg.second});
            todo.push_back({rg.first, mid});
        }
    }
    nb_candidate* res = (nb_candidate*)malloc(sizeof(nb_candidate) * (acc.size() ? acc.size() : 1));
    if (!res) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    if (!acc.empty()) memcpy(res, acc.data(), sizeof(nb_candidate) * acc.size());
    *out = res;
    *n_out = acc.size();
    return NB_OK;
}
