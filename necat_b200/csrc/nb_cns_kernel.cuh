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

struct CnsAln {                 // one accepted alignment, in acceptance order
    const uint8_t* ops;
    uint64_t qseq_off;          // volume offset of the query read
    int32_t  aln_size, toff, tend;
    int32_t  qstart;            // first query base, in query-strand coordinates
    int32_t  qlen, qdir;
    double   weight;
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
    unsigned long long* bp;             // per level: 5 x 11-bit back pointers
    uint8_t*           rev;
    int                min_cov, min_size;
    int                lanes;           // k_cns_dp / k_cns_walk: templates per warp (the other lanes idle)
    uint32_t*          error_flag;      // 1 = level capacity, 2 = link pool / link capacity, 4 = output capacity
    unsigned long long* counters;       // [0] levels, [1] columns (tags), [2] links
};

struct CnsWarpShared {
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

        // ---- 0. an insertion run of >= 255 drops the whole alignment (align_tags.c:40-44); lane = column
        for (int a = 0; a < A; ++a) {
            const uint8_t* ops = AL[a].ops;
            const int n = AL[a].aln_size;
            int carry = 0, bad = 0;
            for (int base = 0; base < n && !bad; base += 32) {
                const int c = base + lane;
                const uint32_t ins = __ballot_sync(0xffffffffu, c < n && ops[c] == OP_QONLY);
                if (ins == 0xffffffffu) { carry += 32; if (carry >= 255) bad = 1; continue; }
                const int lead = __ffs(~ins) - 1;                  // ones at the low end continue the carried run
                if (carry + lead >= 255) bad = 1;
                carry = __clz(~ins);                               // ones at the high end (inner runs are < 32)
            }
            if (lane == 0) {
                CnsCursor c; c.x = 0; c.qpos = (uint32_t)AL[a].qstart; c.last_pos = AL[a].toff - 1; c.last_node = 0xffff;
                c.valid = (uint8_t)!bad; c.pad_ = 0;
                CU[a] = c;
            }
        }
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
        uint32_t ebase = 0;                     // links written so far
        uint32_t prev_start = 0, prev_cnt = 0;  // links of the last position of the previous tile
        for (int t0 = 0; t0 < T && !fail; t0 += 32) {
            const int t1 = (t0 + 32 < T) ? t0 + 32 : T;
            const uint32_t tile_flag = (t0 + lane < T) ? pos[t0 + lane].flags : 0;
            const uint32_t tile_any = __ballot_sync(0xffffffffu, tile_flag & 1);
            if (lane == 0) S.pool_used = 0;
            S.head[lane] = 0xffffu;
            __syncwarp();
            for (int a = 0; a < A; ++a) {
                if (AL[a].toff >= t1 || AL[a].tend <= t0) continue;
                const CnsCursor cu = CU[a];
                if (!cu.valid) continue;
                const uint8_t* ops = AL[a].ops;
                const int n = AL[a].aln_size;
                const uint64_t qso = AL[a].qseq_off;
                const int qlen = AL[a].qlen, qdir = AL[a].qdir;
                const double w = AL[a].weight;
                uint32_t x = cu.x, qp = cu.qpos;
                int last_pos = cu.last_pos;
                uint32_t last_node = cu.last_node;
                while (x < (uint32_t)n) {
                    const uint32_t c = x + lane;
                    const bool ok = c < (uint32_t)n;
                    const int op = ok ? ops[c] : OP_TONLY;
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
                    if (take && op != OP_TONLY) b = cns_qbase(P.pac, qso, qlen, qdir, qp + __popc(qmask & lt_mask));
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
                if (lane == 0) { CnsCursor c; c.x = x; c.qpos = qp; c.last_pos = last_pos; c.last_node = (uint16_t)last_node; c.valid = 1; c.pad_ = 0; CU[a] = c; }
                __syncwarp();
            }
            if (S.pool_used > CNS_POOL) { fail |= 2; break; }
            // ---- flush: lane = position; its links go to HBM in key order, with the index of the
            //      predecessor node's first link resolved here so that k_cns_dp needs no search
            int cnt = 0;
            for (int e = (int)S.head[lane]; e != 0xffff; e = S.e_next[e]) ++cnt;
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

// Flat loop: every iteration either advances to the next template position or scores one link, so the
// templates sharing a warp stay converged although their link counts differ.  Per link: one 16-byte load of
// the link and one 8-byte load of the predecessor's score.
__global__ void __launch_bounds__(128)
k_cns_dp(CnsParams P, CnsRun* __restrict__ runs, uint32_t* __restrict__ tpl_nrun)
{
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(gt & 31) >= P.lanes) return;
    const uint32_t ti = (gt >> 5) * (uint32_t)P.lanes + (gt & 31);
    if (ti >= P.n_tpl) return;
    const CnsTemplate TP = P.tpl[ti];
    const int T = TP.T;
    CnsPos* pos = P.pos + TP.pos_base;
    CnsLink* link = P.link + TP.link_base;
    unsigned long long* bp = P.bp + TP.level_base;
    CnsRun* my_runs = runs + TP.seg_first;
    uint32_t nrun = 0, fail = 0, level = 0;
    double g_best = -1.0;
    int g_level = -1, g_b = -1, g_t = 0, g_delta = 0, run_from = 0;
    int t = -1;
    uint32_t e = 0, fl = 0;                     // links [.., e) of position t
    double pen = 0.0;
    bool have_node = false, have_level = false;
    uint32_t cur_d = 0, cur_b = 0, node_first = 0, best_rank = 0, code = 7;
    double best = -1.0;
    bool chosen = false;
    unsigned long long packed = 0;
    uint32_t x = 0;
    // the link and position records are streamed: each is loaded one iteration before it is used
    uint4 pr_next = *reinterpret_cast<const uint4*>(pos);           // pos[t + 1]
    uint4 pr_next2 = *reinterpret_cast<const uint4*>(pos + 1);      // pos[t + 2]
    uint4 lr_next = *reinterpret_cast<const uint4*>(link);          // link[x]
    for (;;) {
        if (x == e) {
            // ---- finish the position, move to the next one
            if (have_node) {
                const double ns = chosen ? best : -1.0;
                link[node_first].w = ns;                                // node score lives in the node's first link
                packed |= (unsigned long long)(chosen ? code : 7u) << (11 * cur_b);
                if (ns > g_best) { g_best = ns; g_level = (int)level; g_b = (int)cur_b; g_t = t; g_delta = (int)cur_d; }
                have_node = false;
            }
            if (have_level) {
                if (level < TP.level_cap) bp[level] = packed; else fail |= 1;
                ++level; have_level = false;
            }
            if (t >= 0 && (fl & 4)) {                                   // the run ends at t
                if (nrun < TP.seg_cap) {
                    CnsRun r; r.from = run_from; r.to = t + 1; r.g_level = g_level; r.g_b = g_b; r.g_t = g_t; r.g_delta = g_delta;
                    my_runs[nrun] = r;
                }
                ++nrun;
            }
            ++t;
            if (t >= T) break;
            const uint4 pr = pr_next;                                   // start, cov, flags, -
            pr_next = pr_next2;
            pr_next2 = *reinterpret_cast<const uint4*>(pos + (t + 2 <= T ? t + 2 : T));
            fl = pr.z;
            e = pr_next.x;                                              // pos[t + 1].start
            if (fl & 2) { g_best = -1.0; g_level = -1; run_from = t; }
            if (fl & 1) { pen = __dmul_rn(0.2, (double)(int)pr.y); pos[t].first_level = level; }   // 0.4 * 0.5 * coverage (cns_aux.c:164)
            continue;
        }
        // ---- one link
        const uint4 lr = lr_next;
        lr_next = *reinterpret_cast<const uint4*>(link + x + 1);        // capacity leaves room past the last link
        const uint32_t key = lr.x, pred = lr.y;
        const double w = __hiloint2double((int)lr.w, (int)lr.z);
        const uint32_t d = key >> 24, b = (key >> 20) & 7;
        if ((key >> 23) & 1) {                                          // a new node starts: close the previous one
            if (have_node) {
                const double ns = chosen ? best : -1.0;
                link[node_first].w = ns;
                packed |= (unsigned long long)(chosen ? code : 7u) << (11 * cur_b);
                if (ns > g_best) { g_best = ns; g_level = (int)level; g_b = (int)cur_b; g_t = t; g_delta = (int)cur_d; }
            }
            if ((key >> 19) & 1) {                                      // ... and a new delta level
                if (have_level) {
                    if (level < TP.level_cap) bp[level] = packed; else fail |= 1;
                    ++level;
                }
                have_level = true; packed = 0;
            }
            have_node = true; cur_b = b; cur_d = d; node_first = x; best = -1.0; best_rank = 0; code = 7; chosen = false;
        }
        const uint32_t rank = key & 0xffff, pb = (key >> 16) & 7;
        double sc = __dsub_rn(w, pen);
        if (pred != CNS_NOPRED) sc = __dadd_rn(sc, link[pred].w);      // never-scored predecessors count 0 (cns_aux.c:63-70)
        const bool better = chosen ? (sc > best || (sc == best && rank < best_rank)) : (sc > -1.0);
        if (better) {
            best = sc; best_rank = rank; chosen = true;
            code = rank == 0 ? 7u : (((d == 0 ? (rank - 1 - cns_char_rank(pb)) / 5u : 0u) << 3) | pb);
        }
        ++x;
    }
    if (fail) atomicOr(P.error_flag, fail);
    if (nrun > TP.seg_cap) { atomicOr(P.error_flag, 4u); nrun = TP.seg_cap; }
    tpl_nrun[ti] = nrun;
    atomicAdd(P.counters + 0, (unsigned long long)level);
}

// Walk-back of the runs of one template per thread (cns_aux.c:187-212).
__global__ void __launch_bounds__(128)
k_cns_walk(CnsParams P, const CnsRun* __restrict__ runs, const uint32_t* __restrict__ tpl_nrun)
{
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(gt & 31) >= P.lanes) return;
    const uint32_t ti = (gt >> 5) * (uint32_t)P.lanes + (gt & 31);
    if (ti >= P.n_tpl) return;
    const CnsTemplate TP = P.tpl[ti];
    const CnsPos* pos = P.pos + TP.pos_base;
    const unsigned long long* bp = P.bp + TP.level_base;
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
            const uint32_t code = (uint32_t)(bp[cs] >> (11 * cb)) & 0x7ff;
            const uint32_t pb = code & 7;
            if (pb == 7) break;                         // best_p_t_pos == -1: this node is not emitted
            const int emit = cb;
            int ns, nt, nd;
            bool stop_after = false;
            if (cd > 0) { ns = cs - 1; nt = ct; nd = cd - 1; }
            else {
                nt = ct - 1; nd = (int)(code >> 3);
                if (nt < R.from) { stop_after = true; ns = 0; }     // predecessor outside the run: never scored
                else ns = (int)pos[nt].first_level + nd;
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
