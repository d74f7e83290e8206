// nb_cns_kernel.cuh -- consensus of one template from its accepted alignments.
//
// Replaces get_cns_tags / build_backbone / consensus_backbone_segment
// (tasc/align_tags.c:23-71, tasc/cns_aux.c:22-217) and the segment loops of
// tasc/cbcns.c:108-264.  The reference materialises one 24-byte AlignTag per
// alignment column (~115 k per template), introsorts them and groups them into a
// graph of (t_pos, delta, base) nodes with weighted links to predecessor nodes.
//
// Here no tag array is ever written or sorted.  One warp owns a template and
// sweeps it in tiles of 32 template positions:
//   phase A  for every accepted alignment (in acceptance order) the warp reads the
//            alignment's next op bytes with lane = column, derives (t_pos, delta,
//            base, predecessor) of each column with ballots/popcounts, and adds the
//            alignment's weight to the matching link of that position in a
//            shared-memory pool (per-position lists).  Columns of one alignment hit
//            distinct (t_pos, delta) keys, alignments are processed one after the
//            other, so every link weight is the sum in acceptance order -- the
//            order add_one_align was called (cns_aux.c:48).
//            At the end of a tile every position's links are written to HBM sorted by
//            (delta, base, predecessor): k_cns_links.
//   phase B  k_cns_dp: the best-path recurrence (cns_aux.c:157-184) is a chain over
//            template positions with ~4 links each -- no intra-template parallelism
//            worth a warp -- so it runs one THREAD per template, tens of thousands
//            of templates in flight; per (t_pos, delta) level an 11-bit back pointer
//            per node is packed into one u64.  At the end of a coverage run the same
//            thread walks the back pointers (cns_aux.c:187-212) and emits the bases.
#pragma once
#include "nb_common.cuh"
#include "nb_extend.cuh"

#define CNS_MAX_ALN 320         // >= MAX_EXAMINED_CAN (consensus_aux.h:15)
#define CNS_WARPS 4
#define CNS_POOL 768            // links per tile (shared-memory pool; ~115 used on average)
#define CNS_NOPRED 0xffffffffu
#define CNS_FLUSH_REGS 8        // links of one position sorted in registers at the flush (more: the list-walking flush)

struct CnsAln {                 // one accepted alignment, in acceptance order
    const uint8_t* ops;
    uint64_t qseq_off;          // volume offset of the query read
    int32_t  aln_size, toff, tend;
    int32_t  qstart;            // first query base, in query-strand coordinates
    int32_t  qlen, qdir;
    double   weight;
    int32_t  drop;              // 1 = a run of >= 255 insertions: the alignment produces no tags (align_tags.c:40-44; found by k_extend)
    int32_t  pad_;
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
};

struct __align__(16) CnsPos {   // one template position
    uint32_t start;             // first link (relative to link_base); pos[T].start = total
    int32_t  cov;               // tag-producing alignments covering it (cns_aux.c:96-98)
    uint32_t flags;             // 1 = in a scored run, 2 = first of run, 4 = last of run
    uint32_t first_level;       // level index of (t, delta 0), written by k_cns_dp
};

struct __align__(16) CnsLink {  // one link = (node, predecessor node), sorted by key inside a position
    uint32_t key;               // delta(8) << 24 | node-start << 23 | base(3) << 20 | level-start << 19 | pred base(3) << 16 | pred rank(16)
    uint32_t pred;              // index of the predecessor node's first link, CNS_NOPRED = no scored predecessor
    double   w;                 // summed weight; k_cns_dp overwrites the first link of a node with the node score
};

struct __align__(16) CnsLevel { // one (position, delta) level of the best-path recurrence, written by k_cns_dp for k_cns_walk
    unsigned long long packed;  // 5 x 11-bit back pointers, one per base: predecessor delta << 3 | predecessor base, 7 = none
    uint32_t prev_first_level;  // level index of (position - 1, delta 0): the walk needs no second load per step
    uint32_t pad_;
};

struct CnsSegment {             // one consensus stretch of >= min_size bases
    int32_t left, right;        // coverage run [left, right)
    int32_t cns_from, cns_to;   // cns_aux.c:214-216
    uint32_t len;
    uint32_t pad_;
    uint64_t seq_off;           // into the sequence pool (base codes 0..3)
};

struct __align__(16) CnsCursor {    // per alignment cursor of k_cns_links (describes column x-1), kept in HBM/L1 per resident warp
    uint32_t x, qpos;
    int32_t  last_pos;
    uint16_t last_node;             // delta << 3 | base of column x-1; 0xffff = none
    uint8_t  valid, pad_;
};

struct CnsParams {
    const uint8_t*     pac;
    const CnsTemplate* tpl;
    const CnsAln*      aln;
    uint32_t           n_tpl;
    CnsSegment*        seg;
    uint32_t*          tpl_nseg;        // [n_tpl]
    uint8_t*           seq_pool;
    uint32_t*          next_tpl;        // work queue of k_cns_links
    CnsPos*            pos;
    CnsLink*           link;
    CnsCursor*         cursor;          // k_cns_links: [resident warps][CNS_MAX_ALN]
    CnsLevel*          bp;              // per level: 5 x 11-bit back pointers + where the previous position's levels start
    uint8_t*           rev;
    int                min_cov, min_size;
    int                lanes;           // k_cns_dp: templates per warp (the other lanes idle)
    int                walk_lanes;      // k_cns_walk: likewise
    int                flush_slow;      // k_cns_links: 1 = always take the list-walking flush (tests)
    int                links_multi;     // k_cns_links: 1 = reload / store the alignment state every tile even for <= 32 alignments (tests)
    uint32_t*          error_flag;      // 1 = level capacity, 2 = link pool / link capacity, 4 = output capacity
    unsigned long long* counters;       // [0] levels, [1] columns (tags), [2] links
};

struct __align__(16) CnsWarpShared {
    // links of the current tile: per-position singly linked lists
    uint32_t e_key[CNS_POOL];
    double   e_w[CNS_POOL];
    uint16_t e_next[CNS_POOL];
    uint32_t head[32];
    uint32_t pool_used;
};

__device__ __forceinline__ int cns_qbase(const uint8_t* __restrict__ pac, uint64_t qseq_off, int qlen, int qdir, uint32_t qpos)
{
    const uint64_t p = qdir ? qseq_off + (uint64_t)(qlen - 1 - (int)qpos) : qseq_off + qpos;
    const int b = (__ldg(pac + (p >> 2)) >> (((~p) & 3) << 1)) & 3;
    return qdir ? 3 - b : b;
}

// rank of a base in the reference's tag order (ASCII: '-' < 'A' < 'C' < 'G' < 'T')
__device__ __forceinline__ uint32_t cns_char_rank(uint32_t b) { return b == 4 ? 0u : b + 1u; }

// ------------------------------------------------------------------ phase A
__global__ void __launch_bounds__(CNS_WARPS * 32)
k_cns_links(CnsParams P)
{
    extern __shared__ __align__(16) uint8_t cns_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1, le_mask = lt_mask | (1u << lane);
    CnsWarpShared& S = reinterpret_cast<CnsWarpShared*>(cns_smem)[wid];
    CnsCursor* CU = P.cursor + ((size_t)blockIdx.x * CNS_WARPS + wid) * CNS_MAX_ALN;
    const bool kFlushSlow = P.flush_slow != 0;
    unsigned long long n_cols_total = 0, n_links_total = 0;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(P.next_tpl, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= P.n_tpl) break;
        const CnsTemplate TP = P.tpl[ti];
        const CnsAln* AL = P.aln + TP.aln_first;
        const int A = (int)TP.n_aln, T = TP.T;
        CnsPos* pos = P.pos + TP.pos_base;
        CnsLink* link = P.link + TP.link_base;
        uint32_t fail = 0;

        // ---- 0. cursors; an insertion run of >= 255 drops the whole alignment (align_tags.c:40-44; flagged by k_extend)
        for (int a = lane; a < A; a += 32) {
            CnsCursor c; c.x = 0; c.qpos = (uint32_t)AL[a].qstart; c.last_pos = AL[a].toff - 1; c.last_node = 0xffff;
            c.valid = (uint8_t)!AL[a].drop; c.pad_ = 0;
            CU[a] = c;
        }
        __syncwarp();
        // ---- 1. coverage
        for (int t = lane; t <= T; t += 32) { CnsPos z; z.start = 0; z.cov = 0; z.flags = 0; z.first_level = 0; pos[t] = z; }
        __syncwarp();
        for (int a = lane; a < A; a += 32)
            if (CU[a].valid) { atomicAdd(&pos[AL[a].toff].cov, 1); atomicAdd(&pos[AL[a].tend].cov, -1); }
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
        // ---- 2. coverage runs that get a consensus (cbcns.c:129-135)
        {
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
                }
                i = j;
            }
        }
        __syncwarp();

        // ---- 3. tile sweep: links of 32 positions at a time, alignments in acceptance order
        uint32_t r_x = 0, r_qp = 0, r_ln = 0xffff, r_ql = 0;       // lane a < 32: cursor and description of alignment a
        int r_lp = 0, r_n = 0, r_toff = 0, r_tend = 0;
        const uint8_t* r_ops = NULL;
        uint64_t r_qso = 0;
        double r_w = 0.0;
        bool r_valid = false;
        // Up to 32 alignments: their state stays in these registers for the whole template.  More (long templates): the
        // alignments are taken 32 at a time, every tile loads a group's state from HBM / L1 (coalesced, lane = alignment)
        // and stores the cursors back afterwards.
        const bool multi = A > 32 || P.links_multi;
        auto load_group = [&](int g0) {
            const int a = g0 + lane;
            r_valid = false; r_toff = 0; r_tend = 0;
            if (a < A) {
                const CnsAln al = AL[a];
                const CnsCursor cu = CU[a];
                r_ops = al.ops; r_n = al.aln_size; r_toff = al.toff; r_tend = al.tend; r_qso = al.qseq_off;
                r_ql = (uint32_t)al.qlen | ((uint32_t)al.qdir << 31); r_w = al.weight;
                r_valid = cu.valid != 0;
                r_x = cu.x; r_qp = cu.qpos; r_lp = cu.last_pos; r_ln = cu.last_node;
            }
        };
        if (!multi) load_group(0);
        uint32_t ebase = 0;                     // links written so far
        uint32_t prev_start = 0, prev_cnt = 0;  // links of the last position of the previous tile
        uint32_t prev_nid[CNS_FLUSH_REGS];      // ... and their node ids in key order, when the fast flush left them (prev_fast)
        bool prev_fast = false;
#pragma unroll
        for (int j = 0; j < CNS_FLUSH_REGS; ++j) prev_nid[j] = 0xfffu;
        for (int t0 = 0; t0 < T && !fail; t0 += 32) {
            const int t1 = (t0 + 32 < T) ? t0 + 32 : T;
            const uint32_t tile_flag = (t0 + lane < T) ? pos[t0 + lane].flags : 0;
            const uint32_t tile_any = __ballot_sync(0xffffffffu, tile_flag & 1);
            if (lane == 0) S.pool_used = 0;
            S.head[lane] = 0xffffu;
            __syncwarp();
            // Columns [x, ..) of one alignment that fall on this tile: derive (position, delta, base, predecessor) of
            // each column with ballots, add the alignment's weight to the matching link.  cur0 / cur1 are the op bytes
            // of columns x + lane and x + 32 + lane when `pre` says they were fetched ahead of time.
            auto sweep = [&](const uint8_t* __restrict__ ops, const int n, const uint64_t qso, const int qlen, const int qdir, const double w,
                             uint32_t& x, uint32_t& qp, int& last_pos, uint32_t& last_node, const bool pre, const int cur0, const int cur1,
                             const uint32_t cur_pac, const long long pac_lo) {
                for (int it = 0; x < (uint32_t)n; ++it) {
                    const uint32_t c = x + lane;
                    const bool ok = c < (uint32_t)n;
                    int op;
                    if (pre && it == 0) op = cur0;
                    else if (pre && it == 1) op = cur1;
                    else op = ok ? ops[c] : OP_TONLY;
                    const uint32_t tmask = __ballot_sync(0xffffffffu, ok && op != OP_QONLY);
                    const uint32_t qmask = __ballot_sync(0xffffffffu, ok && op != OP_TONLY);
                    const uint32_t vmask = __ballot_sync(0xffffffffu, ok);
                    const uint32_t tle = tmask & le_mask;
                    const int k = __popc(tle);
                    const int ps = last_pos + k;
                    const int delta = (k == 0) ? (int)(last_node >> 3) + lane + 1 : lane - (31 - __clz(tle));
                    const uint32_t late = __ballot_sync(0xffffffffu, ok && ps >= t1);    // columns of later tiles
                    const int n_take = late ? __ffs(late) - 1 : __popc(vmask);
                    if (n_take == 0) break;
                    const bool take = lane < n_take;
                    int b = 4;
                    if (pre && it < 2) {
                        // lanes 0..16 hold the 17 bytes of packed query bases around the cursor (fetched with the ops)
                        const uint32_t q = qp + __popc(qmask & lt_mask);
                        const long long p = qdir ? (long long)qso + (qlen - 1 - (int)q) : (long long)qso + q;
                        const bool want = take && op != OP_TONLY;
                        const uint32_t byte = __shfl_sync(0xffffffffu, cur_pac, want ? (int)((p >> 2) - pac_lo) : 0);
                        if (want) { b = (byte >> (((~p) & 3) << 1)) & 3; if (qdir) b = 3 - b; }
                    } else if (take && op != OP_TONLY) b = cns_qbase(P.pac, qso, qlen, qdir, qp + __popc(qmask & lt_mask));
                    const uint32_t node = ((uint32_t)delta << 3) | (uint32_t)b;
                    uint32_t pnode = __shfl_up_sync(0xffffffffu, node, 1);
                    if (lane == 0) pnode = last_node;
                    const int pl = ps - t0;
                    const bool ins = take && pl >= 0 && ((tile_any >> (pl & 31)) & 1);   // position inside a scored run
                    uint32_t key = 0;
                    int e = 0xffff;
                    if (ins) {
                        uint32_t rank, pb;
                        if (pnode == 0xffff) { rank = 0; pb = 7; }
                        else { pb = pnode & 7; rank = 1 + (delta == 0 ? (pnode >> 3) * 5u : 0u) + cns_char_rank(pb); }
                        key = ((uint32_t)delta << 24) | ((uint32_t)b << 20) | (pb << 16) | rank;
                        e = (int)S.head[pl];
                        while (e != 0xffff && S.e_key[e] != key) e = S.e_next[e];
                    }
                    __syncwarp();
                    if (ins) {
                        if (e != 0xffff) S.e_w[e] = __dadd_rn(S.e_w[e], w);        // acceptance order (cns_aux.c:48)
                        else {
                            const uint32_t ne = atomicAdd(&S.pool_used, 1u);
                            if (ne < CNS_POOL) {
                                S.e_key[ne] = key; S.e_w[ne] = w;
                                S.e_next[ne] = (uint16_t)atomicExch(&S.head[pl], ne);
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
                }
            };
            // cursor and description of an alignment live in the registers of lane a (mod 32); the op bytes and query bases
            // of the next alignment are fetched while the current one is processed
            for (int g0 = 0; g0 < A; g0 += 32) {
                if (multi) load_group(g0);
                uint32_t amask = __ballot_sync(0xffffffffu, r_valid && r_toff < t1 && r_tend > t0);
                int a = amask ? __ffs(amask) - 1 : -1;
                int pre0 = OP_TONLY, pre1 = OP_TONLY;
                uint32_t pre_pac = 0;
                long long pre_lo = 0;
                auto fetch = [&](int an) {
                    const uint32_t fx = __shfl_sync(0xffffffffu, r_x, an);
                    const uint32_t fn = (uint32_t)__shfl_sync(0xffffffffu, r_n, an);
                    const uint8_t* fo = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, (unsigned long long)r_ops, an));
                    const uint32_t c = fx + lane;
                    pre0 = c < fn ? fo[c] : OP_TONLY;
                    pre1 = c + 32 < fn ? fo[c + 32] : OP_TONLY;
                    // packed query bases [qp, qp + 64) of that alignment: 17 bytes, one per lane
                    const uint32_t fq = __shfl_sync(0xffffffffu, r_qp, an), fl = __shfl_sync(0xffffffffu, r_ql, an);
                    const long long fso = (long long)__shfl_sync(0xffffffffu, (unsigned long long)r_qso, an);
                    long long lo = (fl >> 31) ? fso + ((int)(fl & 0x7fffffffu) - 1 - (int)(fq + 63)) : fso + fq;
                    if (lo < 0) lo = 0;
                    pre_lo = lo >> 2;
                    pre_pac = lane <= 16 ? __ldg(P.pac + pre_lo + lane) : 0u;
                };
                if (a >= 0) fetch(a);
                while (a >= 0) {
                    amask &= amask - 1;
                    const int nxt = amask ? __ffs(amask) - 1 : -1;
                    uint32_t x = __shfl_sync(0xffffffffu, r_x, a), qp = __shfl_sync(0xffffffffu, r_qp, a);
                    int last_pos = __shfl_sync(0xffffffffu, r_lp, a);
                    uint32_t last_node = __shfl_sync(0xffffffffu, r_ln, a);
                    const uint8_t* ops = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, (unsigned long long)r_ops, a));
                    const int n = __shfl_sync(0xffffffffu, r_n, a);
                    const uint64_t qso = __shfl_sync(0xffffffffu, (unsigned long long)r_qso, a);
                    const uint32_t ql = __shfl_sync(0xffffffffu, r_ql, a);
                    const double w = __shfl_sync(0xffffffffu, r_w, a);
                    const int cur0 = pre0, cur1 = pre1;
                    const uint32_t cur_pac = pre_pac;
                    const long long cur_lo = pre_lo;
                    if (nxt >= 0) fetch(nxt);
                    sweep(ops, n, qso, (int)(ql & 0x7fffffffu), (int)(ql >> 31), w, x, qp, last_pos, last_node, true, cur0, cur1, cur_pac, cur_lo);
                    if (lane == a) { r_x = x; r_qp = qp; r_lp = last_pos; r_ln = last_node; }
                    a = nxt;
                }
                if (multi) {
                    if (g0 + lane < A) {
                        CnsCursor c; c.x = r_x; c.qpos = r_qp; c.last_pos = r_lp; c.last_node = (uint16_t)r_ln; c.valid = (uint8_t)r_valid; c.pad_ = 0;
                        CU[g0 + lane] = c;
                    }
                    __syncwarp();
                }
            }
            if (S.pool_used > CNS_POOL) { fail |= 2; break; }
            // ---- flush: lane = position; its links go to HBM in key order, with the index of the
            //      predecessor node's first link resolved here so that k_cns_dp needs no search
            // one walk over my position's list: up to CNS_FLUSH_REGS links go to registers as key << 16 | pool slot
            unsigned long long kv[CNS_FLUSH_REGS];
            int cnt = 0;
            int e_walk = (int)S.head[lane];
#pragma unroll
            for (int j = 0; j < CNS_FLUSH_REGS; ++j) {
                kv[j] = ~0ull;
                if (e_walk != 0xffff) { kv[j] = ((unsigned long long)S.e_key[e_walk] << 16) | (unsigned)e_walk; e_walk = S.e_next[e_walk]; ++cnt; }
            }
            const bool crowded = e_walk != 0xffff;                  // more links than registers: the slow flush below
            for (; e_walk != 0xffff; e_walk = S.e_next[e_walk]) ++cnt;
            const bool any_crowded = __any_sync(0xffffffffu, crowded) || kFlushSlow;
            int off = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, off, d); if (lane >= d) off += o; }
            const int tile_total = __shfl_sync(0xffffffffu, off, 31);
            off -= cnt;
            if (ebase + (uint32_t)tile_total > TP.link_cap) { fail |= 2; break; }
            const uint32_t my_start = ebase + (uint32_t)off;
            if (t0 + lane < T) pos[t0 + lane].start = my_start;
            // links of the previous position: the lane below, or the last position of the previous tile (in HBM)
            uint32_t pv_start = __shfl_up_sync(0xffffffffu, my_start, 1);
            if (lane == 0) pv_start = prev_start;
            const bool pv_in_run = (tile_flag & 1) && !(tile_flag & 2);            // the previous position belongs to the same run
            if (!any_crowded) {
                // ---- fast flush: sort the registers (keys of a position are distinct), publish the node ids of the
                //      sorted links so that the next lane can locate its predecessors, resolve, write
#define CNS_CSWAP(a, b) { const unsigned long long lo_ = kv[a] < kv[b] ? kv[a] : kv[b], hi_ = kv[a] < kv[b] ? kv[b] : kv[a]; kv[a] = lo_; kv[b] = hi_; }
                CNS_CSWAP(0, 1) CNS_CSWAP(2, 3) CNS_CSWAP(4, 5) CNS_CSWAP(6, 7)
                CNS_CSWAP(0, 2) CNS_CSWAP(1, 3) CNS_CSWAP(4, 6) CNS_CSWAP(5, 7)
                CNS_CSWAP(1, 2) CNS_CSWAP(5, 6)
                CNS_CSWAP(0, 4) CNS_CSWAP(1, 5) CNS_CSWAP(2, 6) CNS_CSWAP(3, 7)
                CNS_CSWAP(2, 4) CNS_CSWAP(3, 5)
                CNS_CSWAP(1, 2) CNS_CSWAP(3, 4) CNS_CSWAP(5, 6)
#undef CNS_CSWAP
                uint32_t nid[CNS_FLUSH_REGS];                       // delta << 4 | base of each sorted link; 0xfff = none
#pragma unroll
                for (int j = 0; j < CNS_FLUSH_REGS; ++j) nid[j] = (uint32_t)(kv[j] >> 36) & 0xfffu;
                // (the 512 bytes reuse the list pointers: every lane has its whole list in registers by now)
                uint4* sorted_nid = reinterpret_cast<uint4*>(S.e_next);
                __syncwarp();
                {
                    uint4 pk;
                    pk.x = nid[0] | (nid[1] << 16); pk.y = nid[2] | (nid[3] << 16); pk.z = nid[4] | (nid[5] << 16); pk.w = nid[6] | (nid[7] << 16);
                    sorted_nid[lane] = pk;
                }
                __syncwarp();
                uint32_t pn[CNS_FLUSH_REGS];                        // the previous position's sorted node ids
                {
                    const uint4 pk = sorted_nid[lane > 0 ? lane - 1 : 0];
                    pn[0] = pk.x & 0xffff; pn[1] = pk.x >> 16; pn[2] = pk.y & 0xffff; pn[3] = pk.y >> 16;
                    pn[4] = pk.z & 0xffff; pn[5] = pk.z >> 16; pn[6] = pk.w & 0xffff; pn[7] = pk.w >> 16;
                }
                const bool lane0_regs = prev_fast;
                if (lane == 0 && lane0_regs) {
#pragma unroll
                    for (int j = 0; j < CNS_FLUSH_REGS; ++j) pn[j] = prev_nid[j];
                }
#pragma unroll
                for (int i = 0; i < CNS_FLUSH_REGS; ++i) {
                    if (i < cnt) {
                        const uint32_t bk = (uint32_t)(kv[i] >> 16);
                        const int be = (int)(kv[i] & 0xffff);
                        const uint32_t rank = bk & 0xffff, pb = (bk >> 16) & 7, d = bk >> 24;
                        uint32_t pred = CNS_NOPRED;
                        if (rank != 0) {
                            if (d > 0) {
                                // node (d - 1, pb) of this position: its first link = number of links of this position below it
                                const uint32_t want = ((d - 1) << 4) | pb;
                                uint32_t less = 0;
#pragma unroll
                                for (int j = 0; j < CNS_FLUSH_REGS; ++j) less += nid[j] < want;
                                pred = my_start + less;
                            } else if (pv_in_run) {
                                const uint32_t want = (((rank - 1 - cns_char_rank(pb)) / 5u) << 4) | pb;
                                uint32_t less = 0;
                                if (lane > 0 || lane0_regs) {
#pragma unroll
                                    for (int j = 0; j < CNS_FLUSH_REGS; ++j) less += pn[j] < want;
                                } else {
                                    for (uint32_t q = 0; q < prev_cnt; ++q) less += ((link[prev_start + q].key >> 20) & 0xff7u) < want;
                                }
                                pred = pv_start + less;
                            }
                        }
                        // bit 23: first link of its node, bit 19: first link of its delta level (spares k_cns_dp the key compares)
                        const uint32_t lastnid = i > 0 ? nid[i > 0 ? i - 1 : 0] : 0xffffu;
                        const uint32_t fl_node = (i == 0 || nid[i] != lastnid) ? (1u << 23) : 0u;
                        const uint32_t fl_level = (i == 0 || (nid[i] >> 4) != (lastnid >> 4)) ? (1u << 19) : 0u;
                        CnsLink L; L.key = bk | fl_node | fl_level; L.pred = pred; L.w = S.e_w[be];
                        link[my_start + i] = L;
                    }
                }
#pragma unroll
                for (int j = 0; j < CNS_FLUSH_REGS; ++j) prev_nid[j] = __shfl_sync(0xffffffffu, nid[j], 31);
                prev_fast = true;
            } else {
            prev_fast = false;
            uint32_t lastkey = 0;
            for (int i = 0; i < cnt; ++i) {        // selection by increasing key (keys of a position are distinct)
                uint32_t bk = 0xffffffffu; int be = 0xffff;
                for (int e = (int)S.head[lane]; e != 0xffff; e = S.e_next[e]) {
                    const uint32_t kk = S.e_key[e];
                    if (kk < bk && (i == 0 || kk > lastkey)) { bk = kk; be = e; }
                }
                const uint32_t rank = bk & 0xffff, pb = (bk >> 16) & 7, d = bk >> 24;
                uint32_t pred = CNS_NOPRED;
                if (rank != 0) {
                    if (d > 0) {
                        // node (d - 1, pb) of this position: its first link = number of links of this position below it
                        const uint32_t want = ((d - 1) << 4) | pb;
                        uint32_t less = 0;
                        for (int e = (int)S.head[lane]; e != 0xffff; e = S.e_next[e]) less += (S.e_key[e] >> 20) < want;
                        pred = my_start + less;
                    } else if (pv_in_run) {
                        const uint32_t want = (((rank - 1 - cns_char_rank(pb)) / 5u) << 4) | pb;
                        uint32_t less = 0;
                        if (lane > 0) {
                            for (int e = (int)S.head[lane - 1]; e != 0xffff; e = S.e_next[e]) less += (S.e_key[e] >> 20) < want;
                        } else {
                            for (uint32_t q = 0; q < prev_cnt; ++q) less += ((link[prev_start + q].key >> 20) & 0xff7u) < want;
                        }
                        pred = pv_start + less;
                    }
                }
                // bit 23: first link of its node, bit 19: first link of its delta level (spares k_cns_dp the key compares)
                const uint32_t fl_node = (i == 0 || (bk >> 20) != (lastkey >> 20)) ? (1u << 23) : 0u;
                const uint32_t fl_level = (i == 0 || (bk >> 24) != (lastkey >> 24)) ? (1u << 19) : 0u;
                CnsLink L; L.key = bk | fl_node | fl_level; L.pred = pred; L.w = S.e_w[be];
                link[my_start + i] = L;
                lastkey = bk;
            }
            }
            prev_start = __shfl_sync(0xffffffffu, my_start, 31);
            prev_cnt = (uint32_t)__shfl_sync(0xffffffffu, cnt, 31);
            ebase += (uint32_t)tile_total;
            __syncwarp();
        }
        if (lane == 0) pos[T].start = ebase;
        n_links_total += ebase;
        if (fail && lane == 0) atomicOr(P.error_flag, fail);
        __syncwarp();
    }
    for (int d = 16; d > 0; d >>= 1) n_cols_total += __shfl_xor_sync(0xffffffffu, n_cols_total, d);
    if (lane == 0) { atomicAdd(P.counters + 1, n_cols_total); atomicAdd(P.counters + 2, n_links_total); }
}

// ------------------------------------------------------------------ phase B
// One run of a template, recorded by k_cns_dp for k_cns_walk.
struct CnsRun {
    int32_t from, to;           // coverage run [from, to)
    int32_t g_level, g_b, g_t, g_delta;     // best node of the run (cns_aux.c:177-184), g_level < 0: none
};

// One template per WARP.  The recurrence is a chain over template positions -- and, inside a position, over its
// delta levels -- so the only parallelism inside a template is across the links of one level (1 to ~10).  That is
// still what decides the kernel's duration: the longest template of a job has ~60 k positions, and a single thread
// walking its ~280 k links and positions one after the other takes longer than all other templates together.  Here
// the lanes of the warp take the links of the current position (lane = link), so a position costs one pass per delta
// level instead of one iteration per link:
//   * links are streamed through a 256-link ring in shared memory with cp.async, 128 links ahead of their use;
//   * position records are loaded 32 at a time (lane = position), one chunk ahead, and handed out by shuffles;
//   * node scores live in a 256-entry ring in shared memory indexed by the first link of the node; a predecessor is
//     at most two positions back, so the ring misses only where two neighbouring positions hold more than SC_RING
//     links -- those positions also spill their scores into link[].w (where the reference-shaped kernel kept them);
//   * the links of a node are reduced with shuffles (max score, ties to the lower predecessor rank), the five 11-bit
//     back pointers of a level are OR-reduced into one u64, the best node of a run is tracked per lane and reduced
//     when the run ends (ties to the lowest link index = the first one in the reference's traversal order).
// A position with more than FAST_MAX links is handed to lane 0, which scores it link by link.
#define CNS_DP_LINK_RING 256

__device__ __forceinline__ void cns_cp_async16(void* dst_smem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(nb_smem_addr(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cns_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cns_cp_async_wait(int pending)      // at most `pending` (0..3) groups stay in flight
{
    if (pending <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 3;" ::: "memory");
}

__device__ __forceinline__ uint32_t cns_bp_code(uint32_t key)       // back pointer of a node whose best link has this key
{
    const uint32_t rank = key & 0xffff, pb = (key >> 16) & 7;
    if (rank == 0) return 7u;
    return (((key >> 24) == 0 ? (rank - 1 - cns_char_rank(pb)) / 5u : 0u) << 3) | pb;
}

template <int FAST_MAX, int SC_RING>
__global__ void __launch_bounds__(128)
k_cns_dp(CnsParams P, CnsRun* __restrict__ runs, uint32_t* __restrict__ tpl_nrun)
{
    __shared__ __align__(16) uint4 s_links[4][CNS_DP_LINK_RING];
    __shared__ double s_score[4][SC_RING];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t ti = blockIdx.x * 4 + wid;
    if (ti >= P.n_tpl) return;
    uint4* LK = s_links[wid];
    double* SC = s_score[wid];
    const uint32_t lt_mask = (1u << lane) - 1, le_mask = lt_mask | (1u << lane);
    const CnsTemplate TP = P.tpl[ti];
    const int T = TP.T;
    CnsPos* pos = P.pos + TP.pos_base;
    CnsLink* link = P.link + TP.link_base;
    CnsLevel* bp = P.bp + TP.level_base;
    CnsRun* my_runs = runs + TP.seg_first;
    uint32_t nrun = 0, fail = 0, level = 0;
    uint32_t prev_fl = 0;                       // first level of the previous scored position
    // best node of the current run, per lane
    double g_best = -1.0;
    int g_level = -1, g_b = -1, g_t = 0, g_delta = 0, run_from = 0;
    uint32_t g_x = 0xffffffffu;
    // link staging: batches of 32 links, [.., fill) issued, [.., landed) known to have arrived
    uint32_t fill = 0, landed = 0;
    auto issue_batch = [&]() {
        cns_cp_async16(LK + ((fill + lane) & (CNS_DP_LINK_RING - 1)), link + fill + lane);     // the pool has slack past the end
        cns_cp_async_commit();
        fill += 32;
    };
    auto load_pos = [&](int tbase) {
        const int idx = tbase + lane;
        return *reinterpret_cast<const uint4*>(pos + (idx <= T ? idx : T));
    };
    uint4 pc = load_pos(0), pn = load_pos(32);
    for (int t0 = 0; t0 < T; t0 += 32) {
        // lane j describes position t0 + j: first link, number of links, flags, whether its scores must be spilled
        const uint32_t n0 = __shfl_sync(0xffffffffu, pn.x, 0), n1 = __shfl_sync(0xffffffffu, pn.x, 1);
        uint32_t e1 = __shfl_down_sync(0xffffffffu, pc.x, 1), e2 = __shfl_down_sync(0xffffffffu, pc.x, 2);
        if (lane == 31) { e1 = n0; e2 = n1; }
        if (lane == 30) e2 = n0;
        const bool inside = t0 + lane < T;
        const uint32_t my_n = inside ? e1 - pc.x : 0u;
        const uint32_t my_info = (my_n & 0xffffu) | ((inside ? pc.z & 7u : 0u) << 16) | ((e2 - pc.x > (uint32_t)SC_RING ? 1u : 0u) << 20);
        const int jn = T - t0 < 32 ? T - t0 : 32;
        for (int j = 0; j < jn; ++j) {
            const int t = t0 + j;
            const uint32_t s = __shfl_sync(0xffffffffu, pc.x, j);
            const uint32_t info = __shfl_sync(0xffffffffu, my_info, j);
            const uint32_t n = info & 0xffffu, fl = (info >> 16) & 7u;
            const bool spill = (info >> 20) & 1u;
            if (fl & 2) { g_best = -1.0; g_level = -1; g_x = 0xffffffffu; run_from = t; }
            if ((fl & 1) && n > 0) {
                const int cov = __shfl_sync(0xffffffffu, (int)pc.y, j);
                const double pen = __dmul_rn(0.2, (double)cov);                 // 0.4 * 0.5 * coverage (cns_aux.c:164)
                if (lane == 0) pos[t].first_level = level;
                const uint32_t cur_fl = level;
                const uint32_t e = s + n;
                if (n <= (uint32_t)FAST_MAX) {
                    // ---- staging: the links of this position must have landed; keep 128 links ahead
                    if ((s & ~31u) > fill) { cns_cp_async_wait(0); fill = s & ~31u; landed = fill; }      // after a position that bypassed the ring
                    while (fill < e + 128 && fill + 32 - (s & ~31u) <= CNS_DP_LINK_RING) issue_batch();
                    if (e > landed) {
                        const uint32_t pending = (fill - ((e + 31) & ~31u)) >> 5;
                        cns_cp_async_wait((int)pending);
                        landed = fill - 32 * (pending > 3 ? 3 : pending);
                        __syncwarp();
                    }
                    const bool act = (uint32_t)lane < n;
                    const uint32_t x = s + lane;
                    uint4 lr = make_uint4(0, CNS_NOPRED, 0, 0);
                    if (act) lr = LK[x & (CNS_DP_LINK_RING - 1)];
                    const uint32_t key = lr.x, pred = lr.y;
                    const double base = __dsub_rn(__hiloint2double((int)lr.w, (int)lr.z), pen);
                    const bool isstart = act && ((key >> 23) & 1);
                    uint32_t lf = __ballot_sync(0xffffffffu, act && ((key >> 19) & 1));
                    while (lf) {                                                // one pass per delta level
                        const int lo = __ffs(lf) - 1;
                        lf &= lf - 1;
                        const int hi = lf ? __ffs(lf) - 1 : (int)n;
                        const bool in = lane >= lo && lane < hi;
                        double sc = base;
                        if (in && pred != CNS_NOPRED)                           // never-scored predecessors count 0 (cns_aux.c:63-70)
                            sc = __dadd_rn(base, (x - pred < (uint32_t)SC_RING) ? SC[pred & (SC_RING - 1)] : link[pred].w);
                        const uint32_t nf = __ballot_sync(0xffffffffu, in && isstart);
                        const uint32_t above = nf & ~le_mask;
                        const int nend = above ? __ffs(above) - 1 : hi;         // my node ends here (meaningful in its first lane)
                        const int len = (in && isstart) ? nend - lane : 0;
                        const int maxlen = __reduce_max_sync(0xffffffffu, len);
                        double best = sc;
                        uint32_t bkey = key;
                        bool chosen = in && sc > -1.0;
                        for (int k = 1; k < maxlen; ++k) {                      // the other links of the node, one shuffle distance at a time
                            const double osc = __shfl_down_sync(0xffffffffu, sc, k);
                            const uint32_t okey = __shfl_down_sync(0xffffffffu, key, k);
                            if (lane + k < nend && osc > -1.0 &&
                                (!chosen || osc > best || (osc == best && (okey & 0xffff) < (bkey & 0xffff)))) { best = osc; bkey = okey; chosen = true; }
                        }
                        unsigned long long contrib = 0;
                        if (in && isstart) {
                            const double ns = chosen ? best : -1.0;
                            SC[x & (SC_RING - 1)] = ns;
                            if (spill) link[x].w = ns;
                            const uint32_t b = (key >> 20) & 7;
                            contrib = (unsigned long long)(chosen ? cns_bp_code(bkey) : 7u) << (11 * b);
                            if (ns > g_best) { g_best = ns; g_level = (int)level; g_b = (int)b; g_t = t; g_delta = (int)(key >> 24); g_x = x; }
                        }
                        const uint32_t plo = __reduce_or_sync(0xffffffffu, (uint32_t)contrib);
                        const uint32_t phi = __reduce_or_sync(0xffffffffu, (uint32_t)(contrib >> 32));
                        if (lane == 0) {
                            if (level < TP.level_cap) { CnsLevel L; L.packed = ((unsigned long long)phi << 32) | plo; L.prev_first_level = prev_fl; L.pad_ = 0; bp[level] = L; }
                            else fail |= 1;
                        }
                        ++level;
                        __syncwarp();
                    }
                } else {
                    // ---- a crowded position: lane 0 scores it link by link, straight from HBM
                    if (lane == 0) {
                        bool have_node = false, have_level = false, chosen = false;
                        uint32_t cur_d = 0, cur_b = 0, node_first = 0, best_key = 0;
                        double best = -1.0;
                        unsigned long long packed = 0;
                        auto close_node = [&]() {
                            const double ns = chosen ? best : -1.0;
                            SC[node_first & (SC_RING - 1)] = ns;
                            if (spill) link[node_first].w = ns;
                            packed |= (unsigned long long)(chosen ? cns_bp_code(best_key) : 7u) << (11 * cur_b);
                            if (ns > g_best) { g_best = ns; g_level = (int)level; g_b = (int)cur_b; g_t = t; g_delta = (int)cur_d; g_x = node_first; }
                        };
                        auto put_level = [&]() {
                            if (level < TP.level_cap) { CnsLevel L; L.packed = packed; L.prev_first_level = prev_fl; L.pad_ = 0; bp[level] = L; }
                            else fail |= 1;
                        };
                        for (uint32_t x = s; x < e; ++x) {
                            const uint4 lr = *reinterpret_cast<const uint4*>(link + x);
                            const uint32_t key = lr.x, pred = lr.y;
                            if ((key >> 23) & 1) {
                                if (have_node) close_node();
                                if ((key >> 19) & 1) {
                                    if (have_level) { put_level(); ++level; }
                                    have_level = true; packed = 0;
                                }
                                have_node = true; cur_b = (key >> 20) & 7; cur_d = key >> 24; node_first = x; best = -1.0; best_key = 0; chosen = false;
                            }
                            double sc = __dsub_rn(__hiloint2double((int)lr.w, (int)lr.z), pen);
                            if (pred != CNS_NOPRED)
                                sc = __dadd_rn(sc, (x - pred < (uint32_t)SC_RING) ? SC[pred & (SC_RING - 1)] : link[pred].w);
                            const bool better = chosen ? (sc > best || (sc == best && (key & 0xffff) < (best_key & 0xffff))) : (sc > -1.0);
                            if (better) { best = sc; best_key = key; chosen = true; }
                        }
                        if (have_node) close_node();
                        if (have_level) { put_level(); ++level; }
                    }
                    level = __shfl_sync(0xffffffffu, level, 0);
                    __syncwarp();
                }
                prev_fl = cur_fl;
            }
            if (fl & 4) {
                // ---- the run ends at t: its best node (cns_aux.c:177-184) = highest score, first in traversal order
                double mx = g_best;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, mx, d); mx = o > mx ? o : mx; }
                const uint32_t cand = __ballot_sync(0xffffffffu, g_best == mx);
                const uint32_t mnx = __reduce_min_sync(0xffffffffu, g_best == mx ? g_x : 0xffffffffu);
                const uint32_t win = __ballot_sync(0xffffffffu, g_best == mx && g_x == mnx) & cand;
                const int wl = __ffs(win) - 1;
                const int r_level = __shfl_sync(0xffffffffu, g_level, wl), r_b = __shfl_sync(0xffffffffu, g_b, wl);
                const int r_t = __shfl_sync(0xffffffffu, g_t, wl), r_delta = __shfl_sync(0xffffffffu, g_delta, wl);
                if (lane == 0 && nrun < TP.seg_cap) {
                    CnsRun r; r.from = run_from; r.to = t + 1; r.g_level = r_level; r.g_b = r_b; r.g_t = r_t; r.g_delta = r_delta;
                    my_runs[nrun] = r;
                }
                ++nrun;
            }
        }
        pc = pn;
        pn = load_pos(t0 + 64);
    }
    cns_cp_async_wait(0);
    if (lane == 0) {
        if (fail) atomicOr(P.error_flag, fail);
        if (nrun > TP.seg_cap) { atomicOr(P.error_flag, 4u); nrun = TP.seg_cap; }
        tpl_nrun[ti] = nrun;
        atomicAdd(P.counters + 0, (unsigned long long)level);
    }
}

// Walk-back of the runs of one template per thread (cns_aux.c:187-212).
__global__ void __launch_bounds__(128)
k_cns_walk(CnsParams P, const CnsRun* __restrict__ runs, const uint32_t* __restrict__ tpl_nrun)
{
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(gt & 31) >= P.walk_lanes) return;
    const uint32_t ti = (gt >> 5) * (uint32_t)P.walk_lanes + (gt & 31);
    if (ti >= P.n_tpl) return;
    const CnsTemplate TP = P.tpl[ti];
    const CnsPos* pos = P.pos + TP.pos_base;
    const CnsLevel* bp = P.bp + TP.level_base;
    uint8_t* rev = P.rev + TP.level_base;
    const uint32_t nrun = tpl_nrun[ti];
    uint32_t nseg = 0, out_used = 0, fail = 0;
    for (uint32_t k = 0; k < nrun; ++k) {
        const CnsRun R = runs[TP.seg_first + k];
        if (R.g_level < 0) continue;
        int cs = R.g_level, cb = R.g_b, ct = R.g_t, cd = R.g_delta;
        uint8_t* out = rev + pos[R.from].first_level;        // a run emits at most one base per level
        uint32_t n = 0;
        int cns_from = 0;
        for (;;) {
            const uint4 lv = *reinterpret_cast<const uint4*>(bp + cs);             // packed (x, y), prev_first_level (z)
            const uint32_t code = (uint32_t)((((unsigned long long)lv.y << 32) | lv.x) >> (11 * cb)) & 0x7ff;
            const uint32_t pb = code & 7;
            if (pb == 7) break;                         // best_p_t_pos == -1: this node is not emitted
            const int emit = cb;
            int ns, nt, nd;
            bool stop_after = false;
            if (cd > 0) { ns = cs - 1; nt = ct; nd = cd - 1; }
            else {
                nt = ct - 1; nd = (int)(code >> 3);
                if (nt < R.from) { stop_after = true; ns = 0; }     // predecessor outside the run: never scored
                else ns = (int)lv.z + nd;
            }
            cns_from = nt;
            if (emit != 4) { out[n] = (uint8_t)emit; ++n; }
            if (stop_after) break;
            cs = ns; cb = (int)pb; ct = nt; cd = nd;
        }
        if ((int)n >= P.min_size) {
            if (nseg < TP.seg_cap && out_used + n <= TP.out_cap) {
                uint8_t* dst = P.seq_pool + TP.out_base + out_used;
                for (uint32_t i = 0; i < n; ++i) dst[i] = out[n - 1 - i];
                CnsSegment sg;
                sg.left = R.from; sg.right = R.to; sg.cns_from = cns_from; sg.cns_to = R.g_t + 1;
                sg.len = n; sg.pad_ = 0; sg.seq_off = TP.out_base + out_used;
                P.seg[TP.seg_first + nseg] = sg;
                ++nseg; out_used += n;
            } else fail |= 4;
        }
    }
    if (fail) atomicOr(P.error_flag, fail);
    P.tpl_nseg[ti] = nseg;
}
