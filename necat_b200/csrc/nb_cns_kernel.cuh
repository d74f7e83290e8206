// nb_cns_kernel.cuh -- consensus of one template from its accepted alignments.
//
// Replaces get_cns_tags / build_backbone / consensus_backbone_segment
// (tasc/align_tags.c:23-71, tasc/cns_aux.c:22-217) and the segment loops of
// tasc/cbcns.c:108-264.  The reference materialises one 24-byte AlignTag per
// alignment column (~115 k per template), introsorts them and groups them into a
// graph of (t_pos, delta, base) nodes with weighted links to predecessor nodes.
//
// Here no tag array is ever written or sorted, and no kernel is a chain over a whole template except the
// best-path recurrence itself (which is one by definition):
//   k_cns_prep   one warp per template: coverage of every position, the coverage runs that get a consensus
//                (cbcns.c:129-135), the run table.
//   k_cns_ckpt   one warp per accepted alignment: the alignment's cursor (column, query position, node of the
//                previous column) at every span boundary it crosses -- a span is CNS_SPAN template positions --
//                and the number of its columns inside each span.  With these, the spans of a template are
//                independent pieces of work.
//   k_cns_links  one warp per (template, span), taken from a queue: tiles of CNS_TILE positions; for every
//                alignment that touches the tile (in acceptance order) the warp reads the alignment's next op
//                bytes with lane = column, derives (t_pos, delta, base, predecessor) of each column with
//                ballots/popcounts, and adds the alignment's weight to the matching link of that position in a
//                shared-memory slot table.  Columns of one alignment hit distinct (t_pos, delta) keys and the
//                alignments are processed one after the other, so every link weight is the sum in acceptance
//                order -- the order add_one_align was called (cns_aux.c:48).  At the end of a tile every
//                position's links are written to HBM sorted by (delta, base, predecessor), the weight already
//                reduced by the coverage penalty (cns_aux.c:164), the predecessor node resolved to a link index.
//   k_cns_fix    the predecessor indices of the first position of every span (its predecessor position belongs
//                to another warp's span).
//   k_cns_dp     the best-path recurrence (cns_aux.c:157-184): G lanes per template hold a window of G links of the
//                dense stream and take its delta levels one per trip (lane = link of the level); node scores in a
//                shared-memory ring; per level one 16-byte record with the back pointers of its nodes.  A template
//                whose predecessor falls out of the score ring is redone by k_cns_dp_serial (scores in HBM).
//   k_cns_walk   the walk over the back pointers (cns_aux.c:187-212), level records staged through shared
//                memory a window at a time.
#pragma once
#include "nb_common.cuh"
#include "nb_extend.cuh"

#define CNS_MAX_ALN 320         // >= MAX_EXAMINED_CAN (consensus_aux.h:15)
#define CNS_WARPS 4
#ifndef CNS_TILE
#define CNS_TILE 64             // template positions per tile of k_cns_links (multiple of 32)
#endif
#define CNS_SPAN_DEFAULT 1024   // template positions per unit of k_cns_links (multiple of CNS_TILE)
#define CNS_SLOTS 8             // links of one position kept in the slot table (more: the overflow chain)
#define CNS_OV 4096             // overflow links per tile (a pool in HBM per resident warp: crowded positions are rare)
#define CNS_NOPRED 0xffffffffu
#define CNS_PRED_OPEN 0xC0000000u   // | node id wanted at the previous position: resolved by k_cns_fix
#define CNS_TILE_ROUNDS (CNS_TILE / 32)

// link key: delta(8) << 24 | node-start << 23 | base(3) << 20 | level-start << 19 | pred base(3) << 16 |
//           run-first << 15 | run-last << 14 | pred rank(11)
#define CNS_K_NODE     (1u << 23)
#define CNS_K_LEVEL    (1u << 19)
#define CNS_K_RUNFIRST (1u << 15)      // on the first link of the first position of a coverage run
#define CNS_K_RUNLAST  (1u << 14)      // on the first link of the last position of a coverage run
#define CNS_K_RANK(k)  ((k) & 0x7ffu)
#define CNS_K_SORTMASK 0xff77ffffu     // the bits links are ordered by (flags are added after sorting)

struct CnsAln {                 // one accepted alignment, in acceptance order
    const uint8_t* ops;
    uint64_t qseq_off;          // volume offset of the query read
    int32_t  aln_size, toff, tend;
    int32_t  qstart;            // first query base, in query-strand coordinates
    int32_t  qlen, qdir;
    double   weight;
    int32_t  drop;              // 1 = a run of >= 255 insertions: the alignment produces no tags (align_tags.c:40-44; found by k_extend)
    uint32_t ckpt_first;        // first of its span-boundary checkpoints
    uint32_t unit_first;        // first unit (span) of its template
    uint32_t pad_;
};

struct CnsTemplate {
    uint32_t aln_first, n_aln;
    int32_t  T;
    uint32_t seg_first;         // first slot in the segment / run arrays
    uint32_t seg_cap;
    uint32_t out_cap;
    uint64_t out_base;          // into the sequence pool
    uint64_t pos_base;          // into pos[]    (T + 1 records)
    uint64_t link_base;         // into link[]   (link_cap records)
    uint64_t level_base;        // into bp / rev (level_cap slots)
    uint32_t link_cap, level_cap;
    uint32_t unit_first;        // first unit (span) of this template: ceil(T / span) units
    uint32_t pad_;
};

struct CnsUnit { uint32_t tpl, span; };      // one piece of work of k_cns_links

struct __align__(16) CnsCkpt {  // cursor of an alignment at a span boundary b: describes column x - 1, x = first column at position b
    uint32_t x, qpos;
    uint32_t last_node;         // delta << 3 | base of column x - 1
    uint32_t pad_;
};

struct __align__(16) CnsPos {   // one template position
    uint32_t start;             // first link (relative to link_base)
    int32_t  cov;               // tag-producing alignments covering it (cns_aux.c:96-98)
    uint32_t flags;             // 1 = in a scored run, 2 = first of run, 4 = last of run
    uint32_t n;                 // number of links
};

struct __align__(16) CnsLink {  // one link = (node, predecessor node), sorted by key inside a position
    uint32_t key;
    uint32_t pred;              // index of the predecessor node's first link (relative to link_base), CNS_NOPRED = no scored predecessor
    double   w;                 // summed weight minus the coverage penalty; k_cns_dp_serial overwrites the first link of a node with the node score
};

struct __align__(16) CnsLevel { // one (position, delta) level of the best-path recurrence, written by k_cns_dp for k_cns_walk
    uint16_t code[6];           // per base: back pointer = predecessor rank << 3 | predecessor base, rank 0 = none (only the bases that have a node are written)
    uint32_t prev_first_level;  // level index of (position - 1, delta 0): the walk needs no second load per step
};

struct CnsSegment {             // one consensus stretch of >= min_size bases
    int32_t left, right;        // coverage run [left, right)
    int32_t cns_from, cns_to;   // cns_aux.c:214-216
    uint32_t len;
    uint32_t pad_;
    uint64_t seq_off;           // into the sequence pool (base codes 0..3)
};

struct CnsRun {                 // one coverage run of a template: k_cns_prep writes [from, to), k_cns_dp the rest
    int32_t from, to;
    int32_t g_level, g_b, g_t, g_delta;     // best node of the run (cns_aux.c:177-184), g_level < 0: none
    uint32_t first_level;       // level of (from, delta 0)
    uint32_t pad_;
};

struct __align__(16) CnsCursor {    // per alignment cursor of k_cns_links (describes column x-1), kept in HBM/L1 per resident warp
    uint32_t x, qpos;
    int32_t  last_pos;
    uint32_t last_node;             // delta << 3 | base of column x-1; 0xffff = none
};

struct CnsParams {
    const uint8_t*     pac;
    const CnsTemplate* tpl;             // the templates of this launch (a slice of the chunk)
    const CnsTemplate* tpl0;            // all templates of the chunk: CnsUnit::tpl indexes these
    const CnsAln*      aln;
    uint32_t           n_tpl, n_aln;
    CnsSegment*        seg;
    uint32_t*          tpl_nseg;        // [n_tpl]
    uint8_t*           seq_pool;
    const CnsUnit*     unit;            // work list of k_cns_links, longest templates first
    uint32_t           n_units;
    int                span;            // positions per unit
    uint32_t*          span_cols;       // [n_units] alignment columns inside the span = capacity of its link region
    uint2*             span_links;      // [n_units] (first link, number of links) written by k_cns_links
    CnsCkpt*           ckpt;
    uint32_t*          next_unit;       // work queue of k_cns_links
    CnsPos*            pos;
    CnsLink*           link;
    CnsCursor*         cursor;          // k_cns_links: [resident warps][CNS_MAX_ALN]
    uint4*             ovpool;          // k_cns_links: [resident warps][CNS_OV] overflow links (key, next, weight)
    CnsLevel*          bp;              // per level: 5 x 11-bit back pointers + where the previous position's levels start
    uint8_t*           rev;
    CnsRun*            runs;
    uint32_t*          tpl_nrun;        // [n_tpl]
    uint32_t*          redo;            // [n_tpl] templates k_cns_dp hands to k_cns_dp_serial
    uint32_t*          n_redo;
    int                min_cov, min_size;
    int                flush_slow;      // k_cns_links: 1 = always take the list-walking flush (tests)
    int                dp_serial;       // 1 = every template goes to k_cns_dp_serial (min_cov < 1: runs may hold positions without links; tests)
    uint32_t*          error_flag;      // 1 = level capacity, 2 = link pool / link capacity, 4 = output capacity
    unsigned long long* counters;       // [0] levels, [1] columns (tags), [2] links, [3] links that went through the overflow pool
};

__device__ __forceinline__ int cns_qbase(const uint8_t* __restrict__ pac, uint64_t qseq_off, int qlen, int qdir, uint32_t qpos)
{
    const uint64_t p = qdir ? qseq_off + (uint64_t)(qlen - 1 - (int)qpos) : qseq_off + qpos;
    const int b = (__ldg(pac + (p >> 2)) >> (((~p) & 3) << 1)) & 3;
    return qdir ? 3 - b : b;
}

// rank of a base in the reference's tag order (ASCII: '-' < 'A' < 'C' < 'G' < 'T')
__device__ __forceinline__ uint32_t cns_char_rank(uint32_t b) { return b == 4 ? 0u : b + 1u; }

// ------------------------------------------------------------------ coverage, runs
__global__ void __launch_bounds__(128)
k_cns_prep(CnsParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nwarp = gridDim.x * (blockDim.x >> 5);
    for (uint32_t ti = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ti < P.n_tpl; ti += nwarp) {
        const CnsTemplate TP = P.tpl[ti];
        const CnsAln* AL = P.aln + TP.aln_first;
        const int A = (int)TP.n_aln, T = TP.T;
        CnsPos* pos = P.pos + TP.pos_base;
        CnsRun* runs = P.runs + TP.seg_first;
        for (int t = lane; t <= T; t += 32) { CnsPos z; z.start = 0; z.cov = 0; z.flags = 0; z.n = 0; pos[t] = z; }
        __syncwarp();
        // an insertion run of >= 255 drops the whole alignment (align_tags.c:40-44; flagged by k_extend)
        for (int a = lane; a < A; a += 32)
            if (!AL[a].drop) { atomicAdd(&pos[AL[a].toff].cov, 1); atomicAdd(&pos[AL[a].tend].cov, -1); }
        __syncwarp();
        {
            int carry = 0;
            for (int base = 0; base < T; base += 32) {
                const int t = base + lane;
                int v = t < T ? pos[t].cov : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
                v += carry;
                if (t < T) pos[t].cov = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncwarp();
        // coverage runs that get a consensus (cbcns.c:129-135)
        uint32_t nrun = 0;
        int i = 0;
        while (i < T) {
            for (;;) {
                const int t = i + lane;
                const uint32_t low = __ballot_sync(0xffffffffu, t < T && pos[t].cov < P.min_cov);
                if (low == 0xffffffffu) { i += 32; continue; }
                i += __ffs(~low) - 1;
                break;
            }
            if (i > T) i = T;
            int j = i + 1;
            for (;;) {
                const int t = j + lane;
                const uint32_t hi = __ballot_sync(0xffffffffu, t < T && pos[t].cov >= P.min_cov);
                if (hi == 0xffffffffu) { j += 32; continue; }
                j += __ffs(~hi) - 1;
                break;
            }
            if (i < T && (double)(j - i) >= __dmul_rn((double)P.min_size, 0.85)) {
                for (int t = i + lane; t < j; t += 32) pos[t].flags = 1u | (t == i ? 2u : 0u) | (t == j - 1 ? 4u : 0u);
                if (lane == 0 && nrun < TP.seg_cap) {
                    CnsRun r; r.from = i; r.to = j; r.g_level = -1; r.g_b = -1; r.g_t = 0; r.g_delta = 0; r.first_level = 0; r.pad_ = 0;
                    runs[nrun] = r;
                }
                ++nrun;
            }
            i = j;
        }
        if (lane == 0) {
            if (nrun > TP.seg_cap) { atomicOr(P.error_flag, 4u); nrun = TP.seg_cap; }
            P.tpl_nrun[ti] = nrun;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ span-boundary checkpoints
__global__ void __launch_bounds__(128)
k_cns_ckpt(CnsParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1, le_mask = lt_mask | (1u << lane);
    const uint32_t nwarp = gridDim.x * (blockDim.x >> 5);
    const int span = P.span;
    for (uint32_t ai = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ai < P.n_aln; ai += nwarp) {
        const CnsAln al = P.aln[ai];
        if (al.drop) continue;
        const uint8_t* __restrict__ ops = al.ops;
        const uint32_t n = (uint32_t)al.aln_size;
        int sp = al.toff / span;                        // span of the columns seen so far
        int b = (sp + 1) * span;                        // next boundary
        uint32_t x = 0, qp = (uint32_t)al.qstart, last_d = 0xffffu >> 3, x_b = 0, nck = 0;      // last_d: delta of column x - 1
        int last_pos = al.toff - 1;
        while (x < n) {
            const uint32_t c = x + lane;
            const bool ok = c < n;
            const int op = ok ? ops[c] : OP_TONLY;
            const uint32_t tmask = __ballot_sync(0xffffffffu, ok && op != OP_QONLY);
            const uint32_t qmask = __ballot_sync(0xffffffffu, ok && op != OP_TONLY);
            const int ps = last_pos + __popc(tmask & le_mask);
            const uint32_t cross = __ballot_sync(0xffffffffu, ok && ps >= b);
            const int nv = n - x < 32 ? (int)(n - x) : 32;
            // delta of column x + l1 (l1 < nv)
            auto delta_of = [&](int l1) -> uint32_t {
                const uint32_t tl1 = tmask & ((2u << l1) - 1);                              // t-consuming columns up to l1
                return tl1 == 0 ? last_d + (uint32_t)l1 + 1u : (uint32_t)(l1 - (31 - __clz(tl1)));
            };
            if (cross) {
                // column x + L is the first one at position b; its checkpoint describes column x + L - 1
                const int L = __ffs(cross) - 1;
                uint32_t d1, q1;                            // delta and query index of column x + L - 1
                int op1;
                if (L > 0) { d1 = delta_of(L - 1); op1 = __shfl_sync(0xffffffffu, op, L - 1); q1 = qp + __popc(qmask & ((1u << (L - 1)) - 1)); }
                else { d1 = last_d; op1 = x > 0 ? ops[x - 1] : OP_TONLY; q1 = qp - 1; }
                int bb = 4;
                if (op1 != OP_TONLY) bb = cns_qbase(P.pac, al.qseq_off, al.qlen, al.qdir, q1);
                if (lane == 0) {
                    CnsCkpt ck; ck.x = x + (uint32_t)L; ck.qpos = qp + __popc(qmask & ((1u << L) - 1)); ck.last_node = (d1 << 3) | (uint32_t)bb; ck.pad_ = 0;
                    P.ckpt[al.ckpt_first + nck] = ck;
                    atomicAdd(P.span_cols + al.unit_first + sp, x + (uint32_t)L - x_b);
                }
                x_b = x + (uint32_t)L; ++nck; ++sp; b += span;
            }
            last_d = delta_of(nv - 1);
            last_pos += __popc(tmask);
            qp += __popc(qmask);
            x += (uint32_t)nv;
        }
        if (lane == 0) atomicAdd(P.span_cols + al.unit_first + sp, n - x_b);
    }
}

// ------------------------------------------------------------------ links
struct __align__(16) CnsTileShared {
    double   w[CNS_TILE][CNS_SLOTS];
    uint32_t key[CNS_TILE][CNS_SLOTS];
    uint32_t cnt[CNS_TILE];
    uint32_t ov_head[CNS_TILE];
    uint32_t ord[CNS_TILE];             // flush: slot of the i-th link in key order, 4 bits each
    uint32_t carry_key[CNS_SLOTS];      // keys of the last position of the previous tile (carry_cnt of them; 0xffff: read them from HBM)
    uint16_t aidx[CNS_MAX_ALN];
    uint32_t ov_used;
    uint32_t carry_cnt;
    uint32_t pad_[2];
};

__global__ void __launch_bounds__(CNS_WARPS * 32)
k_cns_links(CnsParams P)
{
    extern __shared__ __align__(16) uint8_t cns_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1, le_mask = lt_mask | (1u << lane);
    CnsTileShared& S = reinterpret_cast<CnsTileShared*>(cns_smem)[wid];
    CnsCursor* CU = P.cursor + ((size_t)blockIdx.x * CNS_WARPS + wid) * CNS_MAX_ALN;
    uint4* OV = P.ovpool + ((size_t)blockIdx.x * CNS_WARPS + wid) * CNS_OV;     // x = key, y = next, (z, w) = weight
    const bool kFlushSlow = P.flush_slow != 0;
    const int span = P.span;
    unsigned long long n_cols_total = 0, n_links_total = 0, n_ov_total = 0;

    for (;;) {
        uint32_t ui = 0;
        if (lane == 0) ui = atomicAdd(P.next_unit, 1u);
        ui = __shfl_sync(0xffffffffu, ui, 0);
        if (ui >= P.n_units) break;
        const CnsUnit U = P.unit[ui];
        const CnsTemplate TP = P.tpl0[U.tpl];
        const CnsAln* __restrict__ AL = P.aln + TP.aln_first;
        const int A = (int)TP.n_aln, T = TP.T;
        CnsPos* pos = P.pos + TP.pos_base;
        CnsLink* link = P.link + TP.link_base;
        const int s0 = (int)U.span * span, s1 = (s0 + span < T) ? s0 + span : T;
        uint32_t fail = 0;
        // ---- link region of this span: behind the regions of the template's earlier spans
        uint32_t lb = 0;
        for (uint32_t j = lane; j < U.span; j += 32) lb += P.span_cols[TP.unit_first + j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) lb += __shfl_xor_sync(0xffffffffu, lb, d);
        const uint32_t cap = P.span_cols[TP.unit_first + U.span];
        // ---- nothing to do where no position of the span is in a scored run (k_cns_prep left start = n = 0)
        {
            bool any = false;
            for (int t = s0 + lane; t < s1; t += 32) any |= (pos[t].flags & 1u) != 0;
            if (!__any_sync(0xffffffffu, any)) {
                if (lane == 0) P.span_links[TP.unit_first + U.span] = make_uint2(lb, 0u);
                continue;
            }
        }
        // ---- the alignments that touch the span, in acceptance order, and their cursors at its start
        int n_ov = 0;
        for (int a0 = 0; a0 < A; a0 += 32) {
            const int a = a0 + lane;
            bool ov = false;
            int toff = 0;
            if (a < A) { toff = AL[a].toff; ov = !AL[a].drop && toff < s1 && AL[a].tend > s0; }
            const uint32_t m = __ballot_sync(0xffffffffu, ov);
            if (ov) {
                const int idx = n_ov + __popc(m & lt_mask);
                S.aidx[idx] = (uint16_t)a;
                CnsCursor c;
                if (toff >= s0) { c.x = 0; c.qpos = (uint32_t)AL[a].qstart; c.last_pos = toff - 1; c.last_node = 0xffff; }
                else {
                    const CnsCkpt ck = P.ckpt[AL[a].ckpt_first + (uint32_t)((int)U.span - toff / span - 1)];
                    c.x = ck.x; c.qpos = ck.qpos; c.last_pos = s0 - 1; c.last_node = ck.last_node;
                }
                CU[idx] = c;
            }
            n_ov += __popc(m);
        }
        __syncwarp();
        // up to 32 alignments: their extent stays in registers (lane = alignment of the compacted list)
        int r_toff = 0x7fffffff, r_tend = 0;
        if (lane < n_ov) { const CnsAln& al = AL[S.aidx[lane]]; r_toff = al.toff; r_tend = al.tend; }

        uint32_t ebase = 0;                     // links written so far (relative to lb)
        uint32_t prev_start = 0, prev_cnt = 0;  // links of the last position of the previous tile
        for (int t0 = s0; t0 < s1 && !fail; t0 += CNS_TILE) {
            const int t1 = (t0 + CNS_TILE < s1) ? t0 + CNS_TILE : s1;
            uint4 tile_pos[CNS_TILE_ROUNDS];    // my positions of the tile: t0 + 32 r + lane
            uint32_t tile_run[CNS_TILE_ROUNDS]; // bit l of word r: position t0 + 32 r + l is in a scored run
#pragma unroll
            for (int r = 0; r < CNS_TILE_ROUNDS; ++r) {
                const int t = t0 + 32 * r + lane;
                tile_pos[r] = t < t1 ? *reinterpret_cast<const uint4*>(pos + t) : make_uint4(0, 0, 0, 0);
                tile_run[r] = __ballot_sync(0xffffffffu, tile_pos[r].z & 1u);
                S.cnt[32 * r + lane] = 0; S.ov_head[32 * r + lane] = 0xffffu;
            }
            if (lane == 0) S.ov_used = 0;
            uint32_t any_run = 0;
#pragma unroll
            for (int r = 0; r < CNS_TILE_ROUNDS; ++r) any_run |= tile_run[r];
            __syncwarp();
            if (any_run) {
            // ---- accumulate: the alignments that touch the tile, in acceptance order
            for (int g0 = 0; g0 < n_ov; g0 += 32) {
                if (g0 > 0 || n_ov > 32) {
                    r_toff = 0x7fffffff; r_tend = 0;
                    if (g0 + lane < n_ov) { const CnsAln& al = AL[S.aidx[g0 + lane]]; r_toff = al.toff; r_tend = al.tend; }
                }
                uint32_t amask = __ballot_sync(0xffffffffu, r_toff < t1 && r_tend > t0);
                while (amask) {
                    const int ci = g0 + __ffs(amask) - 1;
                    amask &= amask - 1;
                    const CnsAln al = AL[S.aidx[ci]];
                    const CnsCursor cu = CU[ci];
                    const uint8_t* __restrict__ ops = al.ops;
                    const uint32_t n = (uint32_t)al.aln_size;
                    const uint64_t qso = al.qseq_off;
                    const int qlen = al.qlen, qdir = al.qdir;
                    const double w = al.weight;
                    uint32_t x = cu.x, qp = cu.qpos, last_node = cu.last_node;
                    int last_pos = cu.last_pos;
                    int op = (x + lane < n) ? ops[x + lane] : OP_TONLY;
                    while (x < n) {
                        const uint32_t c = x + lane;
                        const bool ok = c < n;
                        const int op_nx = (c + 32 < n) ? ops[c + 32] : OP_TONLY;       // the next chunk, should this one be taken whole
                        // packed query bases [qp, qp + 32): 9 bytes, one per lane
                        long long lo = qdir ? (long long)qso + (qlen - 1 - (int)(qp + 31)) : (long long)qso + qp;
                        if (lo < 0) lo = 0;
                        const long long pac_lo = lo >> 2;
                        const uint32_t pacb = lane <= 8 ? __ldg(P.pac + pac_lo + lane) : 0u;
                        const uint32_t tmask = __ballot_sync(0xffffffffu, ok && op != OP_QONLY);
                        const uint32_t qmask = __ballot_sync(0xffffffffu, ok && op != OP_TONLY);
                        const uint32_t tle = tmask & le_mask;
                        const int k = __popc(tle);
                        const int ps = last_pos + k;
                        const int delta = (k == 0) ? (int)(last_node >> 3) + lane + 1 : lane - (31 - __clz(tle));
                        const uint32_t late = __ballot_sync(0xffffffffu, ok && ps >= t1);    // columns of later tiles
                        const int nv = n - x < 32 ? (int)(n - x) : 32;
                        const int n_take = late ? __ffs(late) - 1 : nv;
                        if (n_take == 0) break;
                        const bool take = lane < n_take;
                        int b = 4;
                        {
                            const uint32_t q = qp + __popc(qmask & lt_mask);
                            const long long p = qdir ? (long long)qso + (qlen - 1 - (int)q) : (long long)qso + q;
                            const bool want = take && op != OP_TONLY;
                            const uint32_t byte = __shfl_sync(0xffffffffu, pacb, want ? (int)((p >> 2) - pac_lo) : 0);
                            if (want) { b = (byte >> (((~p) & 3) << 1)) & 3; if (qdir) b = 3 - b; }
                        }
                        const uint32_t node = ((uint32_t)delta << 3) | (uint32_t)b;
                        uint32_t pnode = __shfl_up_sync(0xffffffffu, node, 1);
                        if (lane == 0) pnode = last_node;
                        const int pl = ps - t0;
                        bool ins = take && pl >= 0;
                        if (ins) {                                      // position inside a scored run?
                            uint32_t word = tile_run[0];
#pragma unroll
                            for (int r = 1; r < CNS_TILE_ROUNDS; ++r) if ((pl >> 5) == r) word = tile_run[r];
                            ins = (word >> (pl & 31)) & 1;
                        }
                        uint32_t key = 0;
                        int e = -1, oe = 0xffff;
                        if (ins) {
                            uint32_t rank, pb;
                            if (pnode == 0xffff) { rank = 0; pb = 7; }
                            else { pb = pnode & 7; rank = 1 + (delta == 0 ? (pnode >> 3) * 5u : 0u) + cns_char_rank(pb); }
                            key = ((uint32_t)delta << 24) | ((uint32_t)b << 20) | (pb << 16) | rank;
                            const uint4 k0 = *reinterpret_cast<const uint4*>(&S.key[pl][0]);
                            const uint4 k1 = *reinterpret_cast<const uint4*>(&S.key[pl][4]);
                            const uint32_t have = S.cnt[pl];
                            if (have > 7 && k1.w == key) e = 7;
                            if (have > 6 && k1.z == key) e = 6;
                            if (have > 5 && k1.y == key) e = 5;
                            if (have > 4 && k1.x == key) e = 4;
                            if (have > 3 && k0.w == key) e = 3;
                            if (have > 2 && k0.z == key) e = 2;
                            if (have > 1 && k0.y == key) e = 1;
                            if (have > 0 && k0.x == key) e = 0;
                            if (e < 0 && have > CNS_SLOTS) { oe = (int)S.ov_head[pl]; while (oe != 0xffff && OV[oe].x != key) oe = (int)OV[oe].y; }
                        }
                        __syncwarp();                                   // every lane has searched before any lane inserts
                        if (ins) {
                            if (e >= 0) S.w[pl][e] = __dadd_rn(S.w[pl][e], w);                 // acceptance order (cns_aux.c:48)
                            else if (oe != 0xffff) { const uint4 v = OV[oe]; const double nw = __dadd_rn(__hiloint2double((int)v.w, (int)v.z), w); OV[oe].z = (uint32_t)__double2loint(nw); OV[oe].w = (uint32_t)__double2hiint(nw); }
                            else {
                                const uint32_t slot = atomicAdd(&S.cnt[pl], 1u);
                                if (slot < CNS_SLOTS) { S.key[pl][slot] = key; S.w[pl][slot] = w; }
                                else {
                                    const uint32_t ne = atomicAdd(&S.ov_used, 1u);
                                    if (ne < CNS_OV) OV[ne] = make_uint4(key, atomicExch(&S.ov_head[pl], ne), (uint32_t)__double2loint(w), (uint32_t)__double2hiint(w));
                                }
                            }
                            ++n_cols_total;
                        }
                        const int lastl = n_take - 1;
                        last_pos = __shfl_sync(0xffffffffu, ps, lastl);
                        last_node = __shfl_sync(0xffffffffu, node, lastl);
                        qp += __popc(qmask & ((n_take >= 32) ? 0xffffffffu : ((1u << n_take) - 1)));
                        x += (uint32_t)n_take;
                        __syncwarp();
                        if (n_take < 32) break;
                        op = op_nx;
                    }
                    if (lane == 0) { CnsCursor c; c.x = x; c.qpos = qp; c.last_pos = last_pos; c.last_node = last_node; CU[ci] = c; }
                    __syncwarp();
                }
            }
            }
            n_ov_total += S.ov_used;
            if (S.ov_used > CNS_OV) {
                fail |= 2; break; }
            // ---- flush: lane = position (CNS_TILE_ROUNDS of them); its links go to HBM in key order, with the index of the
            //      predecessor node's first link resolved here so that k_cns_dp needs no search.  (Loops are kept rolled on
            //      purpose: the warps of an SM are at different places of this kernel, its code has to stay in the
            //      instruction cache.)
            uint32_t carry = 0;                 // links of the tile before my round
            // number of links of row q (n of them) whose node id (delta << 4 | base) is below `want`
            auto row_less = [&](int q, uint32_t n, uint32_t want) -> uint32_t {
                const uint4 a0 = *reinterpret_cast<const uint4*>(&S.key[q][0]);
                const uint4 a1 = *reinterpret_cast<const uint4*>(&S.key[q][4]);
                uint32_t less = (n > 0 && (a0.x >> 20) < want) + (n > 1 && (a0.y >> 20) < want) + (n > 2 && (a0.z >> 20) < want) + (n > 3 && (a0.w >> 20) < want)
                              + (n > 4 && (a1.x >> 20) < want) + (n > 5 && (a1.y >> 20) < want) + (n > 6 && (a1.z >> 20) < want) + (n > 7 && (a1.w >> 20) < want);
                if (n > CNS_SLOTS) for (int e = (int)S.ov_head[q]; e != 0xffff; e = (int)OV[e].y) less += (OV[e].x >> 20) < want;
                return less;
            };
#pragma unroll 1
            for (int r = 0; r < CNS_TILE_ROUNDS; ++r) {
                const int pl = 32 * r + lane;
                const int t = t0 + pl;
                const uint4 tp = t < t1 ? *reinterpret_cast<const uint4*>(pos + t) : make_uint4(0, 0, 0, 0);
                const int cnt = (int)S.cnt[pl];
                int off = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, off, d); if (lane >= d) off += o; }
                const int round_total = __shfl_sync(0xffffffffu, off, 31);
                off -= cnt;
                if (ebase + carry + (uint32_t)round_total > cap) { fail |= 2; break; }
                const uint32_t my_start = lb + ebase + carry + (uint32_t)off;
                const uint32_t tflag = tp.z;
                if (t < t1) { uint4 pr = tp; pr.x = my_start; pr.w = (uint32_t)cnt; *reinterpret_cast<uint4*>(pos + t) = pr; }
                const double pen = __dmul_rn(0.2, (double)(int)tp.y);               // 0.4 * 0.5 * coverage (cns_aux.c:164)
                // links of the previous position: the lane below, the last lane of the previous round, or the last
                // position of the previous tile (its keys were kept in S.carry_key, or are read back from HBM)
                uint32_t pv_start = __shfl_up_sync(0xffffffffu, my_start, 1);
                if (lane == 0) pv_start = prev_start;
                const bool pv_in_run = (tflag & 1) && !(tflag & 2);                 // the previous position belongs to the same run
                const bool pv_open = pv_in_run && pl == 0 && t0 == s0;              // ... and to another span: k_cns_fix
                const uint32_t flag_run = ((tflag & 2) ? CNS_K_RUNFIRST : 0u) | ((tflag & 4) ? CNS_K_RUNLAST : 0u);
                // 1. a row of up to CNS_SLOTS links: sort the keys in registers (keys of a position are distinct), put them
                //    back in key order, remember which slot (weight) each came from
                const bool small = cnt <= CNS_SLOTS && !kFlushSlow;
                if (cnt > 0 && small) {
                    unsigned long long kv[CNS_SLOTS];
                    {
                        const uint4 a0 = *reinterpret_cast<const uint4*>(&S.key[pl][0]);
                        const uint4 a1 = *reinterpret_cast<const uint4*>(&S.key[pl][4]);
                        const uint32_t kk[CNS_SLOTS] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int j = 0; j < CNS_SLOTS; ++j) kv[j] = j < cnt ? (((unsigned long long)kk[j] << 4) | (unsigned)j) : ~0ull;
                    }
#define CNS_CSWAP(a, b) { const unsigned long long lo_ = kv[a] < kv[b] ? kv[a] : kv[b], hi_ = kv[a] < kv[b] ? kv[b] : kv[a]; kv[a] = lo_; kv[b] = hi_; }
                    CNS_CSWAP(0, 1) CNS_CSWAP(2, 3) CNS_CSWAP(4, 5) CNS_CSWAP(6, 7)
                    CNS_CSWAP(0, 2) CNS_CSWAP(1, 3) CNS_CSWAP(4, 6) CNS_CSWAP(5, 7)
                    CNS_CSWAP(1, 2) CNS_CSWAP(5, 6)
                    CNS_CSWAP(0, 4) CNS_CSWAP(1, 5) CNS_CSWAP(2, 6) CNS_CSWAP(3, 7)
                    CNS_CSWAP(2, 4) CNS_CSWAP(3, 5)
                    CNS_CSWAP(1, 2) CNS_CSWAP(3, 4) CNS_CSWAP(5, 6)
#undef CNS_CSWAP
                    uint32_t ord = 0;
#pragma unroll
                    for (int j = 0; j < CNS_SLOTS; ++j) ord |= ((uint32_t)kv[j] & 15u) << (4 * j);
                    *reinterpret_cast<uint4*>(&S.key[pl][0]) = make_uint4((uint32_t)(kv[0] >> 4), (uint32_t)(kv[1] >> 4), (uint32_t)(kv[2] >> 4), (uint32_t)(kv[3] >> 4));
                    *reinterpret_cast<uint4*>(&S.key[pl][4]) = make_uint4((uint32_t)(kv[4] >> 4), (uint32_t)(kv[5] >> 4), (uint32_t)(kv[6] >> 4), (uint32_t)(kv[7] >> 4));
                    S.ord[pl] = ord;
                }
                __syncwarp();
                // 2. my links in key order: resolve the predecessor, write
                const uint32_t ord = S.ord[pl];
                const int nslot = cnt < CNS_SLOTS ? cnt : CNS_SLOTS;
                uint32_t lastkey = 0;
#pragma unroll 1
                for (int i = 0; i < cnt; ++i) {
                    uint32_t bk; double bw;
                    if (small) { bk = S.key[pl][i]; bw = S.w[pl][(ord >> (4 * i)) & 15u]; }
                    else {
                        // crowded position (or the test switch): selection by increasing key over slots + overflow chain
                        bk = 0xffffffffu; bw = 0.0;
                        for (int j = 0; j < nslot; ++j) {
                            const uint32_t kk = S.key[pl][j];
                            if (kk < bk && (i == 0 || kk > lastkey)) { bk = kk; bw = S.w[pl][j]; }
                        }
                        for (int e = (int)S.ov_head[pl]; e != 0xffff; e = (int)OV[e].y) {
                            const uint4 v = OV[e];
                            if (v.x < bk && (i == 0 || v.x > lastkey)) { bk = v.x; bw = __hiloint2double((int)v.w, (int)v.z); }
                        }
                    }
                    const uint32_t rank = CNS_K_RANK(bk), pb = (bk >> 16) & 7, d = bk >> 24;
                    uint32_t pred = CNS_NOPRED;
                    if (rank != 0) {
                        if (d > 0) {
                            // node (d - 1, pb) of this position: its first link = number of links of this position below it
                            pred = my_start + row_less(pl, (uint32_t)cnt, ((d - 1) << 4) | pb);
                        } else if (pv_in_run) {
                            const uint32_t want = (((rank - 1 - cns_char_rank(pb)) / 5u) << 4) | pb;
                            if (pv_open) pred = CNS_PRED_OPEN | want;
                            else if (pl > 0) pred = pv_start + row_less(pl - 1, S.cnt[pl - 1], want);
                            else if (S.carry_cnt <= CNS_SLOTS) {
                                uint32_t less = 0;
                                for (uint32_t q = 0; q < S.carry_cnt; ++q) less += (S.carry_key[q] >> 20) < want;
                                pred = pv_start + less;
                            } else {
                                uint32_t less = 0;
                                for (uint32_t q = 0; q < prev_cnt; ++q) less += ((link[prev_start + q].key >> 20) & 0xff7u) < want;
                                pred = pv_start + less;
                            }
                        }
                    }
                    // bit 23: first link of its node, bit 19: first link of its delta level (spares k_cns_dp the key compares)
                    const uint32_t fl_node = (i == 0 || (bk >> 20) != (lastkey >> 20)) ? CNS_K_NODE : 0u;
                    const uint32_t fl_level = (i == 0 || (bk >> 24) != (lastkey >> 24)) ? CNS_K_LEVEL : 0u;
                    CnsLink L; L.key = bk | fl_node | fl_level | (i == 0 ? flag_run : 0u); L.pred = pred; L.w = __dsub_rn(bw, pen);
                    link[my_start + i] = L;
                    lastkey = bk;
                }
                __syncwarp();
                prev_start = __shfl_sync(0xffffffffu, my_start, 31);
                prev_cnt = (uint32_t)__shfl_sync(0xffffffffu, cnt, 31);
                carry += (uint32_t)round_total;
            }
            {
                // the keys of the tile's last position stay behind for the first position of the next tile
                const int lp = t1 - t0 - 1;
                const uint32_t lc = S.cnt[lp];
                uint32_t kq = 0;
                if (lane < CNS_SLOTS) kq = S.key[lp][lane];
                __syncwarp();
                if (lane < CNS_SLOTS) S.carry_key[lane] = kq;
                if (lane == 0) S.carry_cnt = lc <= CNS_SLOTS ? lc : 0xffffu;
                const int lr = lp >> 5, ll = lp & 31;
                (void)lr;
                prev_start = __shfl_sync(0xffffffffu, prev_start, 0);       // (already uniform)
                (void)ll;
            }
            ebase += carry;
            __syncwarp();
        }
        if (lane == 0) P.span_links[TP.unit_first + U.span] = make_uint2(lb, fail ? 0u : ebase);
        n_links_total += ebase;
        if (fail && lane == 0) atomicOr(P.error_flag, fail);
        __syncwarp();
    }
    for (int d = 16; d > 0; d >>= 1) n_cols_total += __shfl_xor_sync(0xffffffffu, n_cols_total, d);
    if (lane == 0) { atomicAdd(P.counters + 1, n_cols_total); atomicAdd(P.counters + 2, n_links_total); atomicAdd(P.counters + 3, n_ov_total); }
}

// The first position of a span (not of the template): the predecessors of its delta-0 links are nodes of the last
// position of the previous span, whose links another warp placed (in its own region: the spans of a template are not
// adjacent in memory).  One thread per unit.
__global__ void __launch_bounds__(128)
k_cns_fix(CnsParams P)
{
    const uint32_t ui = blockIdx.x * blockDim.x + threadIdx.x;
    if (ui >= P.n_units) return;
    const CnsUnit U = P.unit[ui];
    if (U.span == 0) return;
    const CnsTemplate TP = P.tpl0[U.tpl];
    const CnsPos* pos = P.pos + TP.pos_base;
    CnsLink* link = P.link + TP.link_base;
    const int b = (int)U.span * P.span;
    const CnsPos cur = pos[b], prv = pos[b - 1];
    for (uint32_t i = 0; i < cur.n; ++i) {
        const uint32_t pred = link[cur.start + i].pred;
        if (pred == CNS_NOPRED || (pred & 0xF0000000u) != CNS_PRED_OPEN) continue;
        const uint32_t want = pred & 0xfffu;
        uint32_t less = 0;
        for (uint32_t q = 0; q < prv.n; ++q) less += ((link[prv.start + q].key >> 20) & 0xff7u) < want;
        // as if the previous position's links were stored right before this span's: link index - pred is then the dense
        // distance k_cns_dp needs, everywhere (inside a span the links of neighbouring positions ARE adjacent)
        link[cur.start + i].pred = cur.start - (prv.n - less);
    }
}

// ------------------------------------------------------------------ best path
// The recurrence (cns_aux.c:157-184) is a chain over the delta LEVELS of a template: the links of one level only need
// node scores of earlier levels, so a level is the unit of parallel work and the number of levels (2.5 per position)
// times the duration of one level step is what the longest template costs.  G lanes own a template (32 / G templates
// per warp, all in lock step, one level per trip):
//   * the lanes hold a WINDOW of G consecutive links of the dense stream k_cns_links / k_cns_fix left (key, first link
//     of the predecessor node, weight; inside a span the links of neighbouring positions are adjacent, so link index -
//     pred is the dense distance back to the predecessor node).  Two ballots over the keys give the level and node
//     starts of the whole window; the levels in it are then taken one after the other without touching memory for
//     links again.  The last level start of a window is where the next window begins (that level may be cut off by the
//     window's end), known as soon as the window is loaded -- so the next window is fetched into registers right then
//     and has landed by the time the levels of this one are done.  No staging ring, no control words, no plan pass;
//   * node scores live in a ring in shared memory indexed by the dense number of the node's first link (a predecessor
//     is a node of this or of the previous position: the ring misses only where two neighbouring positions hold more
//     than SR links; such a template is handed to k_cns_dp_serial);
//   * the links of a node are adjacent lanes: their maximum by shuffles (as many rounds as the longest node of the
//     level needs, none in the common level of single-link nodes), ties between equal scores to the lower predecessor
//     rank = the link the reference meets first (cns_aux.c:166); every node stores its own back pointer into the level
//     record; the best node of a run is tracked per lane and reduced when the run has ended (ties to the lowest dense
//     number = the first one in the reference's traversal order);
//   * a level of more than G links (a window that holds no second level start) is walked link by link by the group's
//     first lane, straight from HBM.
template <int G, int SR>
__global__ void __launch_bounds__(128)
k_cns_dp(CnsParams P)
{
    extern __shared__ __align__(16) uint8_t dp_smem[];
    constexpr int TPW = 32 / G;
    constexpr uint32_t GM = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & (G - 1), gsh = lane & ~(G - 1);          // my lane inside the group; the group's first lane
    const uint32_t slot = (blockIdx.x * 4 + wid) * TPW + (uint32_t)(lane / G);
    double* SC = reinterpret_cast<double*>(dp_smem) + (size_t)(wid * TPW + lane / G) * SR;
    bool done = slot >= P.n_tpl;
    const uint32_t ti = done ? 0 : slot;
    const CnsTemplate TP = P.tpl[ti];
    const uint4* __restrict__ link = reinterpret_cast<const uint4*>(P.link + TP.link_base);
    uint4* bp = reinterpret_cast<uint4*>(P.bp + TP.level_base);
    uint16_t* bp16 = reinterpret_cast<uint16_t*>(bp);
    CnsRun* my_runs = P.runs + TP.seg_first;
    const uint2* __restrict__ spans = P.span_links + TP.unit_first;
    const int n_spans = (TP.T + P.span - 1) / P.span;
    const uint32_t n_runs = done ? 0 : P.tpl_nrun[ti];
    bool redo = false;
    if (!done && P.dp_serial) { redo = true; done = true; }
    auto next_span = [&](int& k, uint32_t& a, uint32_t& a_end) {
        while (a == a_end && k < n_spans) {
            ++k;
            if (k < n_spans) { const uint2 s = spans[k]; a = s.x; a_end = s.x + s.y; }
        }
    };
    // the window: lane gl holds link x + gl of span k (x_end: end of the span); sq = dense number of link x
    int k = -1, nk = -1;
    uint32_t x = 0, x_end = 0, nx = 0, nx_end = 0, sq = 0;
    uint32_t a = 0, top = 0;            // the next level starts at window lane a; the window's last level start is lane top
    uint32_t lv = 1, nd = 1, p0m = 0, rfm = 0, rlm = 0;     // level / node starts of the window (and, one group per warp: position starts, run flags)
    bool primed = false;
    uint4 e = make_uint4(0, 0, 0, 0), en = make_uint4(0, 0, 0, 0);
    if (!done) {
        next_span(nk, nx, nx_end);
        if (nk < n_spans && nx + (uint32_t)gl < nx_end) en = link[nx + gl];
    }
    uint32_t level = 0, cur_fl = 0, prev_fl = 0, run_started = 0, run_emitted = 0, fail = 0;
    int t = 0;
    bool p_last = false, fin = false;
    // best node of the current run among the nodes this lane closed
    double g_best = -1.0;
    uint32_t g_x = 0xffffffffu, g_lb = 0xffffffffu, g_fl = 0;
    int g_t = 0;
    for (;;) {
        // ---- 1. a group whose window holds no complete level any more moves on to the window it fetched
        const bool adv = !done && a >= top;
        if (adv) {
            sq += top;
            k = nk; x = nx; x_end = nx_end; e = en; a = 0;
            if (k >= n_spans) { done = true; fin = p_last; }
        }
        // ---- 2. level and node starts of the windows (a group that did not move gets what it had; with one group per warp
        //         the votes are only taken when it moved, and three more of them spare every level a shuffle)
        if (G < 32 || adv || !primed) {
            const bool valid = !done && x + (uint32_t)gl < x_end;
            const uint32_t lvb = __ballot_sync(full, valid && (e.x & CNS_K_LEVEL));
            const uint32_t ndb = __ballot_sync(full, valid && (e.x & CNS_K_NODE));
            const uint32_t nvalid = done ? 0u : (x_end - x < (uint32_t)G ? x_end - x : (uint32_t)G);
            const uint32_t sent = nvalid < (uint32_t)G ? (1u << nvalid) : 0u;   // the span ends inside the window, and its last level with it
            lv = ((lvb >> gsh) & GM) | sent; nd = ((ndb >> gsh) & GM) | sent;
            if (G == 32) {
                p0m = __ballot_sync(full, valid && (e.x & CNS_K_LEVEL) && (e.x >> 24) == 0);
                rfm = __ballot_sync(full, valid && (e.x & CNS_K_RUNFIRST));
                rlm = __ballot_sync(full, valid && (e.x & CNS_K_RUNLAST));
            }
            primed = true;
        }
        top = 31u - (uint32_t)__clz((int)(lv | 1u));
        const bool lng = !done && top == 0;                 // one level fills the window (or runs past it)
        // ---- 3. groups that moved fetch the window after this one: it begins at the last level start of this one
        if (adv && !done && !lng) {
            nk = k; nx = x + top; nx_end = x_end;
            if (nx == x_end) next_span(nk, nx, nx_end);
            if (nk < n_spans && nx + (uint32_t)gl < nx_end) en = link[nx + gl];
        }
        const bool lev = !done && !lng;                     // (then a < top: a complete level starts at lane a)
        // ---- 4. position / run bookkeeping from the first link of the level (uniform inside a group)
        bool newpos, runfirst, runlast;
        if (G == 32) { newpos = (lev || lng) && ((p0m >> a) & 1u); runfirst = (rfm >> a) & 1u; runlast = (rlm >> a) & 1u; }
        else {
            const uint32_t k0 = __shfl_sync(full, e.x, gsh + (int)a);
            newpos = (lev || lng) && (k0 >> 24) == 0;       // a position starts with its delta-0 level
            runfirst = (k0 & CNS_K_RUNFIRST) != 0; runlast = (k0 & CNS_K_RUNLAST) != 0;
        }
        // one vote for everything the groups of the warp have to agree on: anybody still at work, a run to close, a long level
        const uint32_t ev = __reduce_or_sync(full, ((done && !fin) ? 0u : 1u) | ((fin || (newpos && p_last)) ? 2u : 0u) | (lng ? 4u : 0u));
        if (!(ev & 1u)) break;
        if (ev & 2u) {
            // the best node of the run that ended (cns_aux.c:177-184): highest score, first in traversal order
            double mx = g_best;
#pragma unroll
            for (int d = G >> 1; d > 0; d >>= 1) { const double o = __shfl_xor_sync(full, mx, d, G); mx = o > mx ? o : mx; }
            uint32_t mnx = g_best == mx ? g_x : 0xffffffffu;
#pragma unroll
            for (int d = G >> 1; d > 0; d >>= 1) { const uint32_t o = __shfl_xor_sync(full, mnx, d, G); mnx = o < mnx ? o : mnx; }
            const uint32_t win = (__ballot_sync(full, g_best == mx && g_x == mnx) >> gsh) & GM;
            const int wl = gsh + (win ? __ffs((int)win) - 1 : 0);
            const uint32_t r_lb = __shfl_sync(full, g_lb, wl), r_fl = __shfl_sync(full, g_fl, wl);
            const int r_t = __shfl_sync(full, g_t, wl);
            if (fin || (newpos && p_last)) {
                if (gl == 0 && run_emitted < n_runs) {
                    CnsRun r = my_runs[run_emitted];                // from and first_level were written at the run's start
                    r.to = t + 1;
                    if (r_lb == 0xffffffffu) { r.g_level = -1; r.g_b = -1; r.g_t = 0; r.g_delta = 0; }
                    else { r.g_level = (int)(r_lb >> 3); r.g_b = (int)(r_lb & 7u); r.g_t = r_t; r.g_delta = (int)((r_lb >> 3) - r_fl); }
                    my_runs[run_emitted] = r;
                }
                ++run_emitted;
                p_last = false; fin = false;
            }
        }
        if (newpos) {
            prev_fl = cur_fl; cur_fl = level;
            if (runfirst) {
                t = run_started < n_runs ? my_runs[run_started].from : 0;
                g_best = -1.0; g_x = 0xffffffffu; g_lb = 0xffffffffu;
                if (gl == 0 && run_started < n_runs) my_runs[run_started].first_level = level;
                ++run_started;
            } else ++t;
            p_last = runlast;
        }
        __syncwarp();                                       // the scores of the previous trip are in the ring
        // ---- 5. a level of more than G links: the group's first lane walks it
        if (ev & 4u) {
            uint32_t used = 0;
            if (lng && gl == 0) {
                bool have = false;
                uint32_t cb = 0, nfirst = 0, bkey = 0, xi = x, si = sq;
                double best = -1.0;
                auto close_node = [&]() {
                    SC[nfirst & (SR - 1)] = best;
                    const uint32_t lb = level * 8 + cb;
                    if (level < TP.level_cap) bp16[lb] = (uint16_t)((CNS_K_RANK(bkey) << 3) | ((bkey >> 16) & 7u)); else fail = 1;
                    if (best > g_best) { g_best = best; g_x = nfirst; g_lb = lb; g_t = t; g_fl = cur_fl; }
                };
                for (; xi < x_end; ++xi, ++si) {
                    const uint4 q = link[xi];
                    if (xi != x && (q.x & CNS_K_LEVEL)) break;
                    if (q.x & CNS_K_NODE) {
                        if (have) close_node();
                        have = true; cb = (q.x >> 20) & 7u; nfirst = si; best = -1.0; bkey = 0;
                    }
                    double s2 = __hiloint2double((int)q.w, (int)q.z);
                    if (q.y != CNS_NOPRED) {
                        const uint32_t d = xi - q.y;
                        if (d >= (uint32_t)SR) redo = true;
                        s2 = __dadd_rn(s2, SC[(si - d) & (SR - 1)]);
                    }
                    if (s2 > -1.0 && (s2 > best || (s2 == best && CNS_K_RANK(q.x) < CNS_K_RANK(bkey)))) { best = s2; bkey = q.x; }
                }
                if (have) close_node();
                if (level < TP.level_cap) bp[level].w = prev_fl; else fail = 1;
                used = xi - x;
            }
            used = __shfl_sync(full, used, gsh);
            if (lng) {
                ++level;
                sq += used;
                nk = k; nx = x + used; nx_end = x_end;
                if (nx == x_end) next_span(nk, nx, nx_end);
                en = make_uint4(0, 0, 0, 0);
                if (nk < n_spans && nx + (uint32_t)gl < nx_end) en = link[nx + gl];
                a = 0; top = 0;                             // the next trip moves on to that window
            }
            __syncwarp();
        }
        // ---- 6. the level at window lane a: lanes [a, a + n)
        const uint32_t n = lev ? (uint32_t)__ffs((int)(lv >> (a + 1))) : 0u;
        const bool act = lev && (uint32_t)gl >= a && (uint32_t)gl < a + n;
        const uint32_t key = e.x, pred = e.y;
        double sc = -1.0;
        uint32_t bkey = 0;
        if (act) {
            // never-scored predecessors count 0 (cns_aux.c:63-70)
            double rs = 0.0;
            if (pred != CNS_NOPRED) {
                const uint32_t d = x + (uint32_t)gl - pred;
                if (d >= (uint32_t)SR) redo = true;
                rs = SC[(sq + (uint32_t)gl - d) & (SR - 1)];
            }
            const double s1 = __dadd_rn(__hiloint2double((int)e.w, (int)e.z), rs);
            if (s1 > -1.0) { sc = s1; bkey = key; }         // a link has to score above -1 to count (cns_aux.c:160,170)
        }
        // the longest node of the level decides the rounds (all groups of the warp run the same number)
        int rounds = 0;
        {
            uint32_t z = lev ? ((~nd) >> a) & ((1u << n) - 1u) : 0u;     // window lanes of the level that continue a node
            while (z) { z &= z >> 1; ++rounds; }
        }
        if (G < 32) rounds = __reduce_max_sync(full, rounds);
        if (rounds) {
            const int node_end = act ? gl + __ffs((int)(nd >> (gl + 1))) : 0;      // the next node start above me (the level's end is one)
            const double msc = sc;
            const uint32_t mkey = bkey;
            for (int r = 1; r <= rounds; ++r) {
                const double osc = __shfl_down_sync(full, msc, r, G);
                const uint32_t okey = __shfl_down_sync(full, mkey, r, G);
                if (gl + r < node_end && (osc > sc || (osc == sc && CNS_K_RANK(okey) < CNS_K_RANK(bkey)))) { sc = osc; bkey = okey; }
            }
        }
        if (act && (key & CNS_K_NODE)) {
            SC[(sq + (uint32_t)gl) & (SR - 1)] = sc;
            const uint32_t lb = level * 8 + ((key >> 20) & 7u);
            if (level < TP.level_cap) bp16[lb] = (uint16_t)((CNS_K_RANK(bkey) << 3) | ((bkey >> 16) & 7u)); else fail = 1;
            if (sc > g_best) { g_best = sc; g_x = sq + (uint32_t)gl; g_lb = lb; g_t = t; g_fl = cur_fl; }
            if ((uint32_t)gl == a && level < TP.level_cap) bp[level].w = prev_fl;
        }
        if (lev) { ++level; a += n; }
    }
    redo = ((__ballot_sync(full, redo) >> gsh) & GM) != 0;
    const bool gfail = ((__ballot_sync(full, fail != 0) >> gsh) & GM) != 0;
    if (gl == 0 && slot < P.n_tpl) {
        if (redo) P.redo[atomicAdd(P.n_redo, 1u)] = ti;
        else if (!P.dp_serial) {
            if (gfail) atomicOr(P.error_flag, 1u);
            atomicAdd(P.counters + 0, (unsigned long long)level);
        }
    }
}

// The same recurrence for the templates k_cns_dp could not finish (a predecessor further back than its score ring) and
// for min_cov < 1 (runs may then hold positions without links): one thread per template, over the position records,
// node scores kept in the weight field of the first link of each node in HBM.
__global__ void __launch_bounds__(128)
k_cns_dp_serial(CnsParams P)
{
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= *P.n_redo) return;
    const uint32_t ti = P.redo[ri];
    const CnsTemplate TP = P.tpl[ti];
    CnsLink* link = P.link + TP.link_base;
    CnsLevel* bp = P.bp + TP.level_base;
    CnsRun* my_runs = P.runs + TP.seg_first;
    const CnsPos* pos = P.pos + TP.pos_base;
    const uint32_t n_runs = P.tpl_nrun[ti];
    uint32_t level = 0, cur_fl = 0, prev_fl = 0, fail = 0;
    for (uint32_t r = 0; r < n_runs; ++r) {
        const int from = my_runs[r].from, to = my_runs[r].to;
        double g_best = -1.0;
        int g_level = -1, g_b = -1, g_t = 0, g_delta = 0;
        my_runs[r].first_level = level;
        uint32_t pv_end = 0;                        // end of the links of the last position that had any
        for (int t = from; t < to; ++t) {
            const CnsPos ps = pos[t];
            if (ps.n == 0) continue;
            prev_fl = cur_fl; cur_fl = level;
            bool have_node = false, chosen = false;
            uint32_t cur_b = 0, node_first = 0, best_key = 0;
            double best = -1.0;
            auto close_node = [&]() {
                const double ns = chosen ? best : -1.0;
                link[node_first].w = ns;
                if (level < TP.level_cap) bp[level].code[cur_b] = (uint16_t)(chosen ? (CNS_K_RANK(best_key) << 3) | ((best_key >> 16) & 7u) : 0u);
                if (ns > g_best) { g_best = ns; g_level = (int)level; g_b = (int)cur_b; g_t = t; g_delta = (int)(level - cur_fl); }
            };
            for (uint32_t i = 0; i < ps.n; ++i) {
                const uint32_t x = ps.start + i;
                const uint4 e = *reinterpret_cast<const uint4*>(link + x);
                const uint32_t key = e.x, pred = e.y;
                if (key & CNS_K_NODE) {
                    if (have_node) close_node();
                    if ((key & CNS_K_LEVEL) && i > 0) {
                        if (level < TP.level_cap) bp[level].prev_first_level = prev_fl; else fail |= 1;
                        ++level;
                    }
                    have_node = true; cur_b = (key >> 20) & 7u; node_first = x; best = -1.0; chosen = false;
                }
                double sc = __hiloint2double((int)e.w, (int)e.z);
                // a predecessor in the previous position: pred counts back from this position's first link as if that
                // position's links were stored right before it (they are, except across a span boundary: k_cns_fix)
                if (pred != CNS_NOPRED) sc = __dadd_rn(sc, link[pred >= ps.start ? pred : pv_end - (ps.start - pred)].w);
                const bool better = chosen ? (sc > best || (sc == best && CNS_K_RANK(key) < CNS_K_RANK(best_key))) : (sc > -1.0);
                if (better) { best = sc; best_key = key; chosen = true; }
            }
            if (have_node) close_node();
            if (level < TP.level_cap) bp[level].prev_first_level = prev_fl; else fail |= 1;
            ++level;
            pv_end = ps.start + ps.n;
        }
        CnsRun rr = my_runs[r];
        rr.g_level = g_level; rr.g_b = g_b; rr.g_t = g_t; rr.g_delta = g_delta;
        my_runs[r] = rr;
    }
    if (fail) atomicOr(P.error_flag, fail);
    atomicAdd(P.counters + 0, (unsigned long long)level);
}

// ------------------------------------------------------------------ walk-back of the runs (cns_aux.c:187-212)
// One warp per template.  The walk is a pointer chase over the level records, always towards smaller level indices and
// never more than the levels of two positions at a time: the warp stages a window of CNS_WALK_WIN records in shared
// memory, its first lane walks inside it and collects the emitted bases in shared memory, the warp writes them out.
#define CNS_WALK_WIN 512
#define CNS_WALK_OUT 512

__global__ void __launch_bounds__(128)
k_cns_walk(CnsParams P)
{
    __shared__ __align__(16) uint4 s_win[4][CNS_WALK_WIN];
    __shared__ uint8_t s_out[4][CNS_WALK_OUT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t ti = blockIdx.x * 4 + wid;
    if (ti >= P.n_tpl) return;
    uint4* W = s_win[wid];
    uint8_t* O = s_out[wid];
    const CnsTemplate TP = P.tpl[ti];
    const CnsLevel* bp = P.bp + TP.level_base;
    uint8_t* rev = P.rev + TP.level_base;
    const uint32_t nrun = P.tpl_nrun[ti];
    uint32_t nseg = 0, out_used = 0, fail = 0;
    for (uint32_t k = 0; k < nrun; ++k) {
        const CnsRun R = P.runs[TP.seg_first + k];
        if (R.g_level < 0) continue;
        int cs = R.g_level, cb = R.g_b, ct = R.g_t, cd = R.g_delta;
        uint8_t* out = rev + R.first_level;                 // a run emits at most one base per level
        uint32_t n = 0;
        int cns_from = 0;
        bool finished = false;
        while (!finished) {
            // window [wlo, cs]
            const int wlo = cs + 1 > CNS_WALK_WIN ? cs + 1 - CNS_WALK_WIN : 0;
            for (int i = lane; i <= cs - wlo; i += 32) W[i] = *reinterpret_cast<const uint4*>(bp + wlo + i);
            __syncwarp();
            uint32_t no = 0;
            if (lane == 0) {
                while (cs >= wlo && no < CNS_WALK_OUT) {
                    const uint4 lv = W[cs - wlo];                                           // five 16-bit codes (x, y, z), prev_first_level (w)
                    const uint32_t word = cb < 2 ? lv.x : (cb < 4 ? lv.y : lv.z);
                    const uint32_t code = (word >> ((cb & 1) << 4)) & 0xffffu;
                    const uint32_t pb = code & 7, prank = code >> 3;           // predecessor base, predecessor rank (0: none)
                    if (prank == 0) { finished = true; break; }     // best_p_t_pos == -1: this node is not emitted
                    const int emit = cb;
                    int ns, nt, nd;
                    bool stop_after = false;
                    if (cd > 0) { ns = cs - 1; nt = ct; nd = cd - 1; }
                    else {
                        nt = ct - 1; nd = (int)((prank - 1 - cns_char_rank(pb)) / 5u);     // rank = 1 + delta * 5 + base rank
                        if (nt < R.from) { stop_after = true; ns = 0; }     // predecessor outside the run: never scored
                        else ns = (int)lv.w + nd;
                    }
                    cns_from = nt;
                    if (emit != 4) O[no++] = (uint8_t)emit;
                    if (stop_after) { finished = true; break; }
                    cs = ns; cb = (int)pb; ct = nt; cd = nd;
                }
            }
            __syncwarp();
            no = __shfl_sync(0xffffffffu, no, 0);
            finished = __shfl_sync(0xffffffffu, (int)finished, 0) != 0;
            cs = __shfl_sync(0xffffffffu, cs, 0); cb = __shfl_sync(0xffffffffu, cb, 0);
            ct = __shfl_sync(0xffffffffu, ct, 0); cd = __shfl_sync(0xffffffffu, cd, 0);
            cns_from = __shfl_sync(0xffffffffu, cns_from, 0);
            for (uint32_t i = lane; i < no; i += 32) out[n + i] = O[i];
            n += no;
            __syncwarp();
        }
        if ((int)n >= P.min_size) {
            if (nseg < TP.seg_cap && out_used + n <= TP.out_cap) {
                uint8_t* dst = P.seq_pool + TP.out_base + out_used;
                __syncwarp();
                for (uint32_t i = lane; i < n; i += 32) dst[i] = out[n - 1 - i];
                if (lane == 0) {
                    CnsSegment sg;
                    sg.left = R.from; sg.right = R.to; sg.cns_from = cns_from; sg.cns_to = R.g_t + 1;
                    sg.len = n; sg.pad_ = 0; sg.seq_off = TP.out_base + out_used;
                    P.seg[TP.seg_first + nseg] = sg;
                }
                ++nseg; out_used += n;
            } else fail |= 4;
        }
    }
    if (lane == 0) {
        if (fail) atomicOr(P.error_flag, fail);
        P.tpl_nseg[ti] = nseg;
    }
}
