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
//   phase B  positions of the tile are then scored in order: lanes = links of the
//            position; per delta level the <= 5 nodes take their best link
//            (cns_aux.c:157-184) and an 11-bit back pointer per node is packed into
//            one u64 per (t_pos, delta) step.
// At the end of a coverage run the walk-back follows the back pointers
// (cns_aux.c:187-212).
#pragma once
#include "nb_common.cuh"
#include "nb_extend.cuh"

#define CNS_MAX_ALN 320         // >= MAX_EXAMINED_CAN (consensus_aux.h:15)
#define CNS_WARPS 2
#define CNS_POOL 1024           // links per tile
#define CNS_NODES 160           // nodes of one position

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
    uint8_t*           flag;            // per warp slot: [maxT + 2]  1 = in a scored run, 2 = first of run, 4 = last of run
    unsigned long long* bp;             // per warp slot: [cap_steps]
    uint8_t*           rev;             // per warp slot: [cap_steps]
    uint32_t           maxT, cap_steps;
    int                min_cov, min_size;
    uint32_t*          error_flag;      // 1 = step capacity, 2 = link pool / node list full, 4 = output capacity
    unsigned long long* counters;       // [0] steps, [1] columns (tags)
};

struct CnsWarpShared {
    // per alignment cursor (describes column x-1; warp-uniform reads)
    uint32_t x[CNS_MAX_ALN];
    uint32_t qpos[CNS_MAX_ALN];
    int32_t  last_pos[CNS_MAX_ALN];
    uint16_t last_node[CNS_MAX_ALN];    // delta << 3 | base of column x-1; 0xffff = none
    uint8_t  valid[CNS_MAX_ALN];
    // links of the current tile: per-position singly linked lists
    uint32_t e_key[CNS_POOL];           // delta(8) << 24 | base(3) << 20 | pred base(3) << 16 | pred rank(16)
    double   e_w[CNS_POOL];
    uint16_t e_next[CNS_POOL];
    uint32_t head[32];
    uint32_t pool_used;
    // node scores of the previous / current position
    uint16_t n_key[2][CNS_NODES];       // delta << 3 | base
    double   n_score[2][CNS_NODES];
    uint32_t n_cnt[2];
};

__device__ __forceinline__ int cns_qbase(const uint8_t* __restrict__ pac, uint64_t qseq_off, int qlen, int qdir, uint32_t qpos)
{
    const uint64_t p = qdir ? qseq_off + (uint64_t)(qlen - 1 - (int)qpos) : qseq_off + qpos;
    const int b = (__ldg(pac + (p >> 2)) >> (((~p) & 3) << 1)) & 3;
    return qdir ? 3 - b : b;
}

// rank of a base in the reference's tag order (ASCII: '-' < 'A' < 'C' < 'G' < 'T')
__device__ __forceinline__ uint32_t cns_char_rank(uint32_t b) { return b == 4 ? 0u : b + 1u; }

__device__ __forceinline__ double cns_node_lookup(const CnsWarpShared& S, int buf, uint32_t nkey)
{
    const uint32_t n = S.n_cnt[buf];
    for (uint32_t i = 0; i < n; ++i) if (S.n_key[buf][i] == nkey) return S.n_score[buf][i];
    return 0.0;         // never scored (outside the run): cns_aux.c:63-70 initialises scores to 0
}

__global__ void __launch_bounds__(CNS_WARPS * 32)
k_consensus(CnsParams P)
{
    extern __shared__ __align__(16) uint8_t cns_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1, le_mask = lt_mask | (1u << lane);
    CnsWarpShared& S = reinterpret_cast<CnsWarpShared*>(cns_smem)[wid];
    const size_t slot = (size_t)blockIdx.x * CNS_WARPS + wid;
    int32_t* cov = P.cov + slot * (size_t)(P.maxT + 2);
    uint32_t* pos_first = P.pos_first + slot * (size_t)(P.maxT + 2);
    uint8_t* flag = P.flag + slot * (size_t)(P.maxT + 2);
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
        uint32_t nseg = 0, out_used = 0;
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
                S.valid[a] = (uint8_t)!bad;
                S.x[a] = 0; S.qpos[a] = (uint32_t)AL[a].qstart; S.last_pos[a] = AL[a].toff - 1; S.last_node[a] = 0xffff;
            }
        }
        // ---- 1. coverage[t] = number of tag-producing alignments covering t (cns_aux.c:96-98)
        for (int t = lane; t <= T; t += 32) { cov[t] = 0; flag[t] = 0; }
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
        // ---- 2. coverage runs that get a consensus (cbcns.c:129-135)
        {
            int i = 0;
            while (i < T) {
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
                if (i < T && (double)(j - i) >= __dmul_rn((double)P.min_size, 0.85)) {
                    for (int t = i + lane; t < j; t += 32) flag[t] = (uint8_t)(1 | (t == i ? 2 : 0) | (t == j - 1 ? 4 : 0));
                }
                i = j;
            }
        }
        __syncwarp();

        // ---- 3. sweep the template tile by tile
        uint32_t step = 0;
        double g_best = -1.0;
        int g_step = -1, g_b = -1, g_t = 0, g_delta = 0, run_from = 0;
        int cur = 0;                            // node buffer of the current position
        if (lane == 0) { S.n_cnt[0] = 0; S.n_cnt[1] = 0; }
        for (int t0 = 0; t0 < T && !fail; t0 += 32) {
            const int t1 = (t0 + 32 < T) ? t0 + 32 : T;
            const uint32_t tile_flag = (t0 + lane < T) ? flag[t0 + lane] : 0;
            const uint32_t tile_any = __ballot_sync(0xffffffffu, tile_flag & 1);
            // ---------------- phase A: links of the tile, alignments in acceptance order
            if (lane == 0) S.pool_used = 0;
            S.head[lane] = 0xffffu;
            __syncwarp();
            for (int a = 0; a < A; ++a) {
                if (!S.valid[a]) continue;
                if (AL[a].toff >= t1 || AL[a].tend <= t0) continue;
                const uint8_t* ops = AL[a].ops;
                const int n = AL[a].aln_size;
                const uint64_t qso = AL[a].qseq_off;
                const int qlen = AL[a].qlen, qdir = AL[a].qdir;
                const double w = AL[a].weight;
                uint32_t x = S.x[a], qp = S.qpos[a];
                int last_pos = S.last_pos[a];
                uint32_t last_node = S.last_node[a];
                while (x < (uint32_t)n) {
                    const uint32_t c = x + lane;
                    const bool ok = c < (uint32_t)n;
                    const int op = ok ? ops[c] : OP_TONLY;
                    const uint32_t tmask = __ballot_sync(0xffffffffu, ok && op != OP_QONLY);
                    const uint32_t qmask = __ballot_sync(0xffffffffu, ok && op != OP_TONLY);
                    const uint32_t vmask = __ballot_sync(0xffffffffu, ok);
                    const uint32_t tle = tmask & le_mask;
                    const int k = __popc(tle);
                    const int pos = last_pos + k;
                    const int delta = (k == 0) ? (int)(last_node >> 3) + lane + 1 : lane - (31 - __clz(tle));
                    const uint32_t late = __ballot_sync(0xffffffffu, ok && pos >= t1);   // columns of later tiles
                    const int n_take = late ? __ffs(late) - 1 : __popc(vmask);
                    if (n_take == 0) break;
                    const bool take = lane < n_take;
                    int b = 4;
                    if (take && op != OP_TONLY) b = cns_qbase(P.pac, qso, qlen, qdir, qp + __popc(qmask & lt_mask));
                    const uint32_t node = ((uint32_t)delta << 3) | (uint32_t)b;
                    uint32_t pnode = __shfl_up_sync(0xffffffffu, node, 1);
                    if (lane == 0) pnode = last_node;
                    const int pl = pos - t0;
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
                    last_pos = __shfl_sync(0xffffffffu, pos, lastl);
                    last_node = __shfl_sync(0xffffffffu, node, lastl);
                    qp += __popc(qmask & ((n_take >= 32) ? 0xffffffffu : ((1u << n_take) - 1)));
                    x += (uint32_t)n_take;
                    __syncwarp();
                    if (n_take < 32) break;
                }
                if (lane == 0) { S.x[a] = x; S.qpos[a] = qp; S.last_pos[a] = last_pos; S.last_node[a] = (uint16_t)last_node; }
                __syncwarp();
            }
            if (S.pool_used > CNS_POOL) { fail |= 2; break; }
            if (!tile_any) continue;
            // ---------------- phase B: score the positions of the tile in order; lanes = links of the position
            for (int t = t0; t < t1 && !fail; ++t) {
                const uint32_t fl = __shfl_sync(0xffffffffu, tile_flag, t - t0);
                if (!(fl & 1)) continue;
                if (fl & 2) {                       // a run starts: nothing scored so far
                    if (lane == 0) { S.n_cnt[0] = 0; S.n_cnt[1] = 0; }
                    g_best = -1.0; g_step = -1; step = 0; run_from = t;
                    __syncwarp();
                }
                const int prev = cur ^ 1;
                if (lane == 0) { S.n_cnt[cur] = 0; pos_first[t] = step; }
                // hand the links of this position to the lanes (two per lane at most)
                int e0 = 0xffff, e1 = 0xffff, cnt = 0;
                for (int e = (int)S.head[t - t0]; e != 0xffff; e = S.e_next[e], ++cnt) {
                    if (cnt == lane) e0 = e;
                    if (cnt == lane + 32) e1 = e;
                }
                if (cnt > 64) { fail |= 2; break; }
                const double pen = __dmul_rn(0.2, (double)cov[t]);             // 0.4 * 0.5 * coverage (cns_aux.c:164)
                int maxd = 0;
                uint32_t k0 = 0, k1 = 0;
                double w0 = 0, w1 = 0;
                if (e0 != 0xffff) { k0 = S.e_key[e0]; w0 = S.e_w[e0]; maxd = (int)(k0 >> 24); }
                if (e1 != 0xffff) { k1 = S.e_key[e1]; w1 = S.e_w[e1]; maxd = max(maxd, (int)(k1 >> 24)); }
                for (int d = 16; d > 0; d >>= 1) maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, d));
                __syncwarp();
                for (int d = 0; d <= maxd && !fail; ++d) {
                    if (step >= P.cap_steps) { fail |= 1; break; }
                    double nsc[5]; uint32_t nrank[5], ncode[5]; bool npres[5], nchosen[5];
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) { nsc[bb] = -1.0; nrank[bb] = 0; ncode[bb] = 7; npres[bb] = false; nchosen[bb] = false; }
                    // each link of this delta level gets its score; two passes for lanes holding two links
                    for (int half = 0; half < 2; ++half) {
                        const int e = half ? e1 : e0;
                        const uint32_t key = half ? k1 : k0;
                        const bool act = e != 0xffff && (int)(key >> 24) == d;
                        const uint32_t am = __ballot_sync(0xffffffffu, act);
                        if (!am) continue;
                        const uint32_t b = (key >> 20) & 7, pb = (key >> 16) & 7, rank = key & 0xffff;
                        double sc = 0.0;
                        if (act) {
                            sc = __dsub_rn(half ? w1 : w0, pen);
                            if (rank != 0) {
                                double ps;
                                if (d == 0) ps = cns_node_lookup(S, prev, (uint32_t)(((rank - 1 - cns_char_rank(pb)) / 5u) << 3) | pb);
                                else ps = cns_node_lookup(S, cur, (uint32_t)((d - 1) << 3) | pb);
                                sc = __dadd_rn(sc, ps);
                            }
                        }
                        // per base: best link = largest score, ties to the smallest predecessor rank; needs score > -1
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) {
                            uint32_t gm = __ballot_sync(0xffffffffu, act && b == (uint32_t)bb);
                            if (!gm) continue;
                            npres[bb] = true;
                            for (; gm; gm &= gm - 1) {
                                const int src = __ffs(gm) - 1;
                                const double s2 = __shfl_sync(0xffffffffu, sc, src);
                                const uint32_t key2 = __shfl_sync(0xffffffffu, key, src);
                                const uint32_t r2 = key2 & 0xffff, pb2 = (key2 >> 16) & 7;
                                const bool better = nchosen[bb] ? (s2 > nsc[bb] || (s2 == nsc[bb] && r2 < nrank[bb])) : (s2 > -1.0);
                                if (better) {
                                    nsc[bb] = s2; nrank[bb] = r2; nchosen[bb] = true;
                                    ncode[bb] = r2 == 0 ? 7u : (((d == 0 ? (r2 - 1 - cns_char_rank(pb2)) / 5u : 0u) << 3) | pb2);
                                }
                            }
                        }
                    }
                    unsigned long long packed = 0;
                    // record the nodes of this level, the back pointers, the run's best node (cns_aux.c:177-184)
                    uint32_t n0 = S.n_cnt[cur];
                    __syncwarp();
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) {
                        if (!npres[bb]) continue;
                        const double ns = nchosen[bb] ? nsc[bb] : -1.0;
                        packed |= (unsigned long long)(nchosen[bb] ? ncode[bb] : 7u) << (11 * bb);
                        if (n0 < CNS_NODES) {
                            if (lane == 0) { S.n_key[cur][n0] = (uint16_t)((d << 3) | bb); S.n_score[cur][n0] = ns; }
                        } else fail |= 2;
                        ++n0;
                        if (ns > g_best) { g_best = ns; g_step = (int)step; g_b = bb; g_t = t; g_delta = d; }
                    }
                    if (lane == 0) { S.n_cnt[cur] = n0 < CNS_NODES ? n0 : CNS_NODES; bp[step] = packed; }
                    __syncwarp();
                    ++step;
                }
                cur ^= 1;
                // ---------------- the run ends here: walk back from its best node (cns_aux.c:187-212)
                if ((fl & 4) && !fail) {
                    n_steps_total += step;
                    if (g_step >= 0) {
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
                                if (nt < run_from) { stop_after = true; ns = 0; }   // predecessor outside the run: never scored
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
                                    sg.left = run_from; sg.right = t + 1; sg.cns_from = cns_from; sg.cns_to = g_t + 1;
                                    sg.len = n; sg.pad_ = 0; sg.seq_off = TP.out_base + out_used;
                                    P.seg[TP.seg_first + nseg] = sg;
                                }
                                ++nseg; out_used += n;
                            } else fail |= 4;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (fail && lane == 0) atomicOr(P.error_flag, fail);
        if (lane == 0) P.tpl_nseg[ti] = nseg;
        __syncwarp();
    }
    for (int d = 16; d > 0; d >>= 1) n_cols_total += __shfl_xor_sync(0xffffffffu, n_cols_total, d);
    if (lane == 0) { atomicAdd(P.counters + 0, n_steps_total); atomicAdd(P.counters + 1, n_cols_total); }
}
