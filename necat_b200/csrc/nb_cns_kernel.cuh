// nb_cns_kernel.cuh -- consensus of one template from its accepted alignments.
//
// Replaces get_cns_tags / build_backbone / consensus_backbone_segment
// (tasc/align_tags.c:23-71, tasc/cns_aux.c:22-217) and the segment loops of
// tasc/cbcns.c:108-264.  The reference materialises one 24-byte AlignTag per
// alignment column (~115 k per template), introsorts them and groups them into a
// graph of (t_pos, delta, base) nodes.  Here no tag is ever written: one warp
// walks the template left to right with every accepted alignment's cursor in
// lock step ("pile-up walk").  At each (t_pos, delta) step the alignments that
// have a column there are grouped by (base, predecessor node) with
// __match_any_sync -- exactly the link groups of build_base_links -- their
// weights are summed in acceptance order (the order add_one_align was called),
// the <= 5 nodes of the step take their best predecessor (cns_aux.c:164-180) and
// an 11-bit back pointer per node is packed into one u64 per step.  The walk-back
// then follows those pointers.
#pragma once
#include "nb_common.cuh"
#include "nb_extend.cuh"

#define CNS_MAX_ALN 320         // >= MAX_EXAMINED_CAN (consensus_aux.h:15)
#define CNS_WARPS 4
#define CNS_TABLE 96            // link groups per step

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
    uint32_t seg_first;         // first slot in the segment arrays
    uint32_t seg_cap;
    uint64_t out_base;          // into the sequence pool
    uint32_t out_cap;
    uint32_t pad_;
};

struct CnsSegment {             // one consensus stretch of >= min_size bases
    int32_t left, right;        // coverage run [left, right)
    int32_t cns_from, cns_to;   // cns_aux.c:214-216
    uint32_t len;
    uint32_t pad_;
    uint64_t seq_off;           // into the sequence pool (base codes 0..3)
};

struct CnsParams {
    const uint8_t*     pac;
    const CnsTemplate* tpl;
    const CnsAln*      aln;
    uint32_t           n_tpl;
    CnsSegment*        seg;
    uint32_t*          tpl_nseg;        // [n_tpl]
    uint8_t*           seq_pool;
    uint32_t*          next_tpl;        // work queue
    int32_t*           cov;             // per warp slot: [maxT + 2]
    uint32_t*          pos_first;       // per warp slot: [maxT + 2]
    unsigned long long* bp;             // per warp slot: [cap_steps]
    uint8_t*           rev;             // per warp slot: [cap_steps]
    uint32_t           maxT, cap_steps;
    int                min_cov, min_size;
    uint32_t*          error_flag;      // 1 = step capacity exceeded, 2 = link table full, 4 = output capacity
    unsigned long long* counters;       // [0] steps, [1] columns (tags)
};

struct CnsWarpShared {
    // per alignment dynamic state
    uint32_t x[CNS_MAX_ALN];            // next column
    uint32_t qpos[CNS_MAX_ALN];         // next query base
    double   pscore[CNS_MAX_ALN];       // score of the node of the previous column
    uint16_t prev[CNS_MAX_ALN];         // previous column's node: delta << 3 | base ; 0xffff = none
    uint8_t  valid[CNS_MAX_ALN];
    int32_t  tpos[CNS_MAX_ALN];         // template position of the column at the cursor
    // link groups of the current step
    uint32_t e_key[CNS_TABLE];          // base << 16 | pred_rank
    double   e_w[CNS_TABLE];
    double   e_ps[CNS_TABLE];
    uint32_t n_entries;
    double   node_score[5];
    uint32_t node_present;
};

__device__ __forceinline__ int cns_qbase(const uint8_t* __restrict__ pac, const CnsAln& a, uint32_t qpos)
{
    const uint64_t p = a.qdir ? a.qseq_off + (uint64_t)(a.qlen - 1 - (int)qpos) : a.qseq_off + qpos;
    const int b = (__ldg(pac + (p >> 2)) >> (((~p) & 3) << 1)) & 3;
    return a.qdir ? 3 - b : b;
}

// rank of a base in the reference's tag order (ASCII: '-' < 'A' < 'C' < 'G' < 'T')
__device__ __forceinline__ uint32_t cns_char_rank(int b) { return b == 4 ? 0u : (uint32_t)b + 1u; }

__global__ void __launch_bounds__(CNS_WARPS * 32)
k_consensus(CnsParams P)
{
    extern __shared__ __align__(16) uint8_t cns_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    CnsWarpShared& S = reinterpret_cast<CnsWarpShared*>(cns_smem)[wid];
    const size_t slot = (size_t)blockIdx.x * CNS_WARPS + wid;
    int32_t* cov = P.cov + slot * (size_t)(P.maxT + 2);
    uint32_t* pos_first = P.pos_first + slot * (size_t)(P.maxT + 2);
    unsigned long long* bp = P.bp + slot * (size_t)P.cap_steps;
    uint8_t* rev = P.rev + slot * (size_t)P.cap_steps;
    unsigned long long n_steps_total = 0, n_cols_total = 0;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(P.next_tpl, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= P.n_tpl) break;
        const CnsTemplate TP = P.tpl[ti];
        const CnsAln* AL = P.aln + TP.aln_first;
        const int A = (int)TP.n_aln, T = TP.T;
        const int nchunk = (A + 31) >> 5;
        uint32_t nseg = 0, out_used = 0;

        // ---- 0. an insertion run of >= 255 drops the whole alignment (align_tags.c:40-44)
        for (int a = lane; a < A; a += 32) {
            const CnsAln al = AL[a];
            int run = 0, bad = 0;
            for (int x = 0; x < al.aln_size; ++x) {
                const int op = al.ops[x];
                run = (op == OP_QONLY) ? run + 1 : 0;
                if (run >= 255) { bad = 1; break; }
            }
            S.valid[a] = (uint8_t)!bad;
            S.x[a] = 0; S.qpos[a] = (uint32_t)al.qstart; S.prev[a] = 0xffff; S.pscore[a] = 0.0; S.tpos[a] = al.toff;
        }
        // ---- 1. coverage[t] = number of tag-producing alignments covering t (cns_aux.c:96-98)
        for (int t = lane; t <= T; t += 32) cov[t] = 0;
        __syncwarp();
        for (int a = lane; a < A; a += 32)
            if (S.valid[a]) { atomicAdd(cov + AL[a].toff, 1); atomicAdd(cov + AL[a].tend, -1); }
        __syncwarp();
        {
            int carry = 0;
            for (int base = 0; base < T; base += 32) {
                const int t = base + lane;
                int v = t < T ? cov[t] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
                v += carry;
                if (t < T) cov[t] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncwarp();

        // ---- 2. coverage runs (cbcns.c:129-135)
        int i = 0;
        while (i < T) {
            // skip positions below min_cov, then take the run at or above it (warp-parallel scans)
            for (;;) {
                const int t = i + lane;
                const uint32_t low = __ballot_sync(0xffffffffu, t < T && cov[t] < P.min_cov);
                if (low == 0xffffffffu) { i += 32; continue; }
                i += __ffs(~low) - 1;
                break;
            }
            if (i > T) i = T;
            int j = i + 1;
            for (;;) {
                const int t = j + lane;
                const uint32_t hi = __ballot_sync(0xffffffffu, t < T && cov[t] >= P.min_cov);
                if (hi == 0xffffffffu) { j += 32; continue; }
                j += __ffs(~hi) - 1;
                break;
            }
            if ((double)(j - i) >= __dmul_rn((double)P.min_size, 0.85) && i < T) {
                const int from = i, to = j;
                // ---- 3a. bring every cursor to `from`, remembering the node of its last column
                for (int a = lane; a < A; a += 32) {
                    if (!S.valid[a]) continue;
                    const CnsAln al = AL[a];
                    S.pscore[a] = 0.0;
                    if (al.toff >= from || al.tend <= from) continue;
                    uint32_t x = S.x[a], qp = S.qpos[a];
                    uint32_t prev = S.prev[a];
                    int t = S.tpos[a];
                    while (t < from && x < (uint32_t)al.aln_size) {
                        // first column of position t consumes a template base
                        int op = al.ops[x];
                        int b = (op == OP_TONLY) ? 4 : cns_qbase(P.pac, al, qp);
                        if (op != OP_TONLY) ++qp;
                        ++x;
                        int d = 0;
                        prev = (uint32_t)b;
                        while (x < (uint32_t)al.aln_size && al.ops[x] == OP_QONLY) {
                            ++d;
                            b = cns_qbase(P.pac, al, qp); ++qp; ++x;
                            prev = ((uint32_t)d << 3) | (uint32_t)b;
                        }
                        ++t;
                    }
                    S.x[a] = x; S.qpos[a] = qp; S.prev[a] = (uint16_t)prev; S.tpos[a] = t;
                }
                __syncwarp();
                // ---- 3b. forward sweep over [from, to)
                uint32_t step = 0;
                double g_best = -1.0;
                int g_step = -1, g_b = -1, g_t = 0, g_delta = 0;
                bool overflow = false;
                for (int t = from; t < to && !overflow; ++t) {
                    if (lane == 0) pos_first[t] = step;
                    const double cov_pen = __dmul_rn(0.2, (double)cov[t]);      // 0.4 * 0.5 * coverage (cns_aux.c:164)
                    for (int delta = 0;; ++delta) {
                        if (lane == 0) { S.n_entries = 0; S.node_present = 0; }
                        __syncwarp();
                        uint32_t any = 0;
                        // pass 1: group the columns of this step into links
                        for (int c = 0; c < nchunk; ++c) {
                            const int a = (c << 5) + lane;
                            bool part = false;
                            uint32_t key = 0;
                            double w = 0.0, ps = 0.0;
                            if (a < A && S.valid[a]) {
                                const CnsAln& al = AL[a];
                                if (al.toff <= t && t < al.tend) {
                                    const uint32_t x = S.x[a];
                                    if (delta == 0) part = true;
                                    else part = x < (uint32_t)al.aln_size && al.ops[x] == OP_QONLY;
                                    if (part) {
                                        const int op = al.ops[x];
                                        const int b = (op == OP_TONLY) ? 4 : cns_qbase(P.pac, al, S.qpos[a]);
                                        const uint32_t pv = S.prev[a];
                                        uint32_t rank;
                                        if (pv == 0xffff) rank = 0;
                                        else rank = 1 + (delta == 0 ? (pv >> 3) * 5u : 0u) + cns_char_rank((int)(pv & 7));
                                        key = ((uint32_t)b << 16) | rank;
                                        // keep the predecessor's base code for the back pointer
                                        key |= (pv == 0xffff ? 7u : (pv & 7u)) << 20;
                                        w = al.weight;
                                        ps = (pv == 0xffff) ? 0.0 : S.pscore[a];
                                    }
                                }
                            }
                            const uint32_t pm = __ballot_sync(0xffffffffu, part);
                            if (!pm) continue;
                            any |= pm;
                            const uint32_t n_before = S.n_entries;
                            if (part) {
                                const uint32_t grp = __match_any_sync(pm, key);
                                const int leader = __ffs(grp) - 1;
                                // existing link from an earlier chunk?
                                int e = -1;
                                for (uint32_t k = 0; k < n_before; ++k) if (S.e_key[k] == key) { e = (int)k; break; }
                                double acc = e >= 0 ? S.e_w[e] : 0.0;
                                for (uint32_t m = grp; m; m &= m - 1) {
                                    const int src = __ffs(m) - 1;
                                    const double wm = __shfl_sync(grp, w, src);
                                    acc = __dadd_rn(acc, wm);               // acceptance order (cns_aux.c:48)
                                }
                                if (lane == leader) {
                                    if (e < 0) {
                                        e = (int)atomicAdd(&S.n_entries, 1u);
                                        if (e < CNS_TABLE) { S.e_key[e] = key; S.e_ps[e] = ps; }
                                    }
                                    if (e < CNS_TABLE) S.e_w[e] = acc;
                                }
                            }
                            __syncwarp();
                        }
                        if (!any) break;            // no column at this delta: position finished
                        __syncwarp();
                        const uint32_t ne = S.n_entries;
                        if (ne > CNS_TABLE) { if (lane == 0) atomicOr(P.error_flag, 2u); overflow = true; break; }
                        if (step >= P.cap_steps) { if (lane == 0) atomicOr(P.error_flag, 1u); overflow = true; break; }
                        // pass 2: the <= 5 nodes take their best link (cns_aux.c:157-184); lane b owns base b
                        uint32_t code = 7;           // back pointer: pd << 3 | pb ; pb == 7 = none
                        double best = -1.0;
                        bool present = false;
                        if (lane < 5) {
                            bool chosen = false;
                            uint32_t best_rank = 0;
                            for (uint32_t k = 0; k < ne; ++k) {
                                const uint32_t key = S.e_key[k];
                                if (((key >> 16) & 0xf) != (uint32_t)lane) continue;
                                present = true;
                                const uint32_t rank = key & 0xffff;
                                double sc = __dsub_rn(S.e_w[k], cov_pen);
                                if (rank != 0) sc = __dadd_rn(sc, S.e_ps[k]);
                                const bool better = chosen ? (sc > best || (sc == best && rank < best_rank)) : (sc > -1.0);
                                if (better) {
                                    best = sc; best_rank = rank; chosen = true;
                                    if (rank == 0) code = 7;
                                    else {
                                        const uint32_t pb = (key >> 20) & 7;
                                        const uint32_t pd = delta == 0 ? (rank - 1 - cns_char_rank((int)pb)) / 5u : 0u;
                                        code = (pd << 3) | pb;
                                    }
                                }
                            }
                            if (present) S.node_score[lane] = best;
                        }
                        const uint32_t pres = __ballot_sync(0xffffffffu, present) & 0x1f;
                        unsigned long long packed = 0;
#pragma unroll
                        for (int b = 0; b < 5; ++b) {
                            const uint32_t cb = __shfl_sync(0xffffffffu, code, b);
                            const double sb = __shfl_sync(0xffffffffu, best, b);
                            packed |= (unsigned long long)(cb & 0x7ff) << (11 * b);
                            if (((pres >> b) & 1) && sb > g_best) { g_best = sb; g_step = (int)step; g_b = b; g_t = t; g_delta = delta; }
                        }
                        if (lane == 0) bp[step] = packed;
                        __syncwarp();
                        // pass 3: advance the cursors that had a column here
                        for (int c = 0; c < nchunk; ++c) {
                            const int a = (c << 5) + lane;
                            if (a < A && S.valid[a]) {
                                const CnsAln& al = AL[a];
                                if (al.toff <= t && t < al.tend) {
                                    const uint32_t x = S.x[a];
                                    bool part = delta == 0 ? true : (x < (uint32_t)al.aln_size && al.ops[x] == OP_QONLY);
                                    if (part) {
                                        const int op = al.ops[x];
                                        int b = 4;
                                        if (op != OP_TONLY) { b = cns_qbase(P.pac, al, S.qpos[a]); S.qpos[a] += 1; }
                                        S.x[a] = x + 1;
                                        S.prev[a] = (uint16_t)(((uint32_t)delta << 3) | (uint32_t)b);
                                        S.pscore[a] = S.node_score[b];
                                        ++n_cols_total;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        ++step;
                    }
                }
                n_steps_total += step;
                for (int a = lane; a < A; a += 32)
                    if (S.valid[a] && AL[a].toff < to && AL[a].tend > from) S.tpos[a] = AL[a].tend < to ? AL[a].tend : to;
                __syncwarp();
                // ---- 3c. walk back from the best node (cns_aux.c:187-212)
                if (!overflow && g_step >= 0) {
                    int cs = g_step, cb = g_b, ct = g_t, cd = g_delta;
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
                            if (nt < from) { stop_after = true; ns = 0; }   // predecessor outside the run: never scored
                            else ns = (int)pos_first[nt] + nd;
                        }
                        cns_from = nt;
                        if (emit != 4) { if (lane == 0) rev[n] = (uint8_t)emit; ++n; }
                        if (stop_after) break;
                        cs = ns; cb = (int)pb; ct = nt; cd = nd;
                    }
                    __syncwarp();
                    if ((int)n >= P.min_size) {
                        if (nseg < TP.seg_cap && out_used + n <= TP.out_cap) {
                            uint8_t* dst = P.seq_pool + TP.out_base + out_used;
                            for (uint32_t k = lane; k < n; k += 32) dst[k] = rev[n - 1 - k];
                            if (lane == 0) {
                                CnsSegment sg;
                                sg.left = from; sg.right = to; sg.cns_from = cns_from; sg.cns_to = g_t + 1;
                                sg.len = n; sg.pad_ = 0; sg.seq_off = TP.out_base + out_used;
                                P.seg[TP.seg_first + nseg] = sg;
                            }
                            ++nseg; out_used += n;
                        } else if (lane == 0) atomicOr(P.error_flag, 4u);
                    }
                }
                __syncwarp();
            }
            i = j;
        }
        if (lane == 0) P.tpl_nseg[ti] = nseg;
        __syncwarp();
    }
    for (int d = 16; d > 0; d >>= 1) n_cols_total += __shfl_xor_sync(0xffffffffu, n_cols_total, d);
    if (lane == 0) { atomicAdd(P.counters + 0, n_steps_total); atomicAdd(P.counters + 1, n_cols_total); }
}
