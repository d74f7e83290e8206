// nb_extend.cu -- block-wise Myers bit-vector extension, up to 16 candidates in flight per warp.
//
// Replaces onc_align / oca_extend / Edlib_align (gapped_align/oc_aligner.c:157-451,
// gapped_align/edlib_ex.c:733-800).  Per 512-base block the reference runs a banded
// prefix search, a second banded global pass that stores the band, and a walk-back.
// Inside the band all cells that can lie on an optimal path are exact, so the
// result equals: best prefix distance d* over target prefixes (first minimum), fail
// if d* > (int)(min(qn,tn)*error*1.1); walk back from (qn, j*) preferring up, then
// left, then diagonal (edlib_ex.c:463,500,540).  Here:
//   * one unbanded forward pass, 32-bit words, lane w owns query rows [32w, 32w+32); the
//     horizontal carry travels lane to lane by shuffle along the anti-diagonal
//     (lane w works on column s-w at step s); two 512-row blocks side by side in the two
//     half-warps;
//   * per cell only two bits are kept for the walk-back: Pv (cell = upper + 1) and
//     Ph (cell = left + 1), and only for the words within +-48 rows of the block's diagonal
//     (4 words per column; the walk that leaves that band makes the pass run again around
//     the diagonal it reached);
//   * the walk-backs of a warp's candidates run lane-parallel (lane = candidate), one cell
//     per trip without a branch, and find the last run of M matches
//     (oc_aligner.c:222-240) on the fly.
// Integer/bitwise work only: no tensor cores.
#include "nb_extend.cuh"

struct SeqView {
    const uint8_t* pac;
    uint64_t off;
    int len;
    int strand;
};

__device__ __forceinline__ int seq_base(const SeqView& v, int idx)
{
    const uint64_t p = v.strand ? v.off + (uint64_t)(v.len - 1 - idx) : v.off + (uint64_t)idx;
    const int b = (__ldg(v.pac + (p >> 2)) >> (((~p) & 3) << 1)) & 3;
    return v.strand ? 3 - b : b;
}

// ---------------------------------------------------------------------------------------------------------------------
// A warp keeps EXT_SLOTS candidates in flight, lane = slot (the lanes beyond have none).  The pieces of a candidate's state machine that are one
// dependent chain per candidate -- cutting the next block, the walk-back over the (Pv, Ph) bits, the bookkeeping
// between blocks -- run with every lane on its own candidate; the forward passes, which have 16 to 26 independent
// 32-row words per DP column, are run by the whole warp, two candidates at a time in its two halves.
//
// Only the words of a DP column that lie within +-48 rows (at least) of the block's diagonal are kept for the walk-back:
// 4 words per column instead of 16-26 (8 words, +-112 rows around the middle between the diagonal and the end cell, for
// the last block of a side, whose two sides differ by up to 30 %).  A walk that leaves the band (the path of a 512-column
// block of 15 %-error reads wanders ~10 rows) makes the forward pass of that block run again with the band centred on
// the diagonal the walk has reached, and continues from where it stopped.
#ifndef EXT_SLOTS
#define EXT_SLOTS 16          // candidates in flight per warp: more of them make the per-lane phases denser, fewer keep the chain of one
#endif                      // candidate (forward passes of all slots, then its walk-back) short and the launch's tail with it
#define EXT_BAND_WORDS 8        // at most

enum {
    X_STATE, X_TI, X_QOFF_LO, X_QOFF_HI, X_TOFF_LO, X_TOFF_HI, X_QLEN, X_QSTRAND, X_TLEN, X_OPS_LO, X_OPS_HI, X_LCAP, X_QS, X_TS,
    X_PHASE, X_QORG, X_TORG, X_QTOT, X_TTOT, X_DIR, X_OUTB, X_OUTDIR, X_QIDX, X_TIDX, X_NCOL, X_QBLK, X_TBLK, X_LAST,
    X_BEST, X_BESTJ, X_L, X_LKEEP, X_ACNT, X_SHIFT, X_NBW,
    X_WI, X_WJ, X_WR, X_WFLAGS, X_WACNT, X_WQCNT, X_WTCNT, X_WM, X_WACC,
    X_NF
};
enum { ST_IDLE = 0, ST_BLOCK = 1, ST_WALK = 2, ST_FINISH = 3 };

struct HalfScratch {
    uint8_t q[EXT_MAX_BLK + 32];        // the block's bases start at q[qa]: 16-base words are unpacked at 16-byte boundaries
    uint8_t t[EXT_MAX_BLK + 96];        // likewise; lane 0 of the forward pass reads up to 31 columns past the block
};

struct WarpScratch {
    HalfScratch h[2];
    uint32_t    pe[4][32];              // forward pass: match mask of lane l's 32 query rows for symbol c at pe[c][l]
    uint2       sb[2][28];              // forward pass: per 32 steps the (+1, -1) bits of the last query row, per half-warp
    int32_t     f[X_NF][EXT_SLOTS];     // state of the candidates in flight, one field per row: f[field][slot]
};

#define XF(field, slot) S.f[field][slot]

// Forward pass over one block, WIDTH lanes per block (16: two blocks per warp, 32: one).  Lane `lid` of the group
// owns query rows [32 lid, 32 lid + 32) and works on column s - lid at step s.  What travels from lane to lane
// with one shuffle per step is the horizontal delta of the row above (bit 31: +1, bit 0: -1) together with the
// symbol of that column (bits 7..8, so that it is the byte offset of the symbol's row in S.pe) -- the neighbour needs
// exactly that column next.  The loop has no branch: a lane computes on whatever arrives before its first column and
// after its last one, its state is set when its first column comes up, and what it stores for columns that do not
// exist is never read.
//   * (Pv, Ph) of word-row w and column c is kept where c is within the band of row w: mat[w * 32 nbw + c - cl(w)],
//     cl(w) = 32 (w - nbw / 2 + 1) - 16 - shift, 32 nbw columns per row -- a lane writes consecutive addresses, and so
//     reads the walk-back when it follows a diagonal.
//   * the score of the last query row is not tracked in the loop: its +1 / -1 bits are collected 32 steps at a time
//     and the first minimum over the columns is found afterwards, 32 steps per lane.
// Returns the best prefix distance and its column in every lane of the group.
template <int WIDTH>
__device__ __forceinline__ void myers_forward(WarpScratch& S, const uint8_t* __restrict__ sq, const uint8_t* __restrict__ st, int qblk, int tblk,
                                              uint2* __restrict__ mat, int shift, int nbw, int lane, int* best_out, int* bestj_out)
{
    const int lid = lane & (WIDTH - 1), half = WIDTH < 32 ? lane >> 4 : 0;
    const int nw = (qblk + 31) >> 5;
    {
        uint32_t pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0;
        if (lid < nw) {
            const int r0 = lid << 5;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                if (r0 + r < qblk) {
                    const int c = sq[r0 + r];
                    const uint32_t bit = 1u << r;
                    pe0 |= c == 0 ? bit : 0; pe1 |= c == 1 ? bit : 0; pe2 |= c == 2 ? bit : 0; pe3 |= c == 3 ? bit : 0;
                }
            }
        }
        S.pe[0][lane] = pe0; S.pe[1][lane] = pe1; S.pe[2][lane] = pe2; S.pe[3][lane] = pe3;
    }
    const int lw = qblk > 0 ? (qblk - 1) >> 5 : 0, lb = (qblk - 1) & 31;
    int nsteps = qblk > 0 ? tblk + nw - 1 : 0;
    if (WIDTH < 32) { const int o = __shfl_xor_sync(0xffffffffu, nsteps, WIDTH); nsteps = nsteps > o ? nsteps : o; }
    const int W = 32 * nbw;
    const int cl = 32 * (lid - (nbw >> 1) + 1) - 16 - shift;
    uint2* __restrict__ row = mat + (size_t)lid * W;
    int rel = lid < nw ? -lid - cl : 0x40000000;            // column - cl(lid) of the step; never in [0, W) for a lane without rows
    const char* pe_base = reinterpret_cast<const char*>(&S.pe[0][lane]);
    uint32_t Pv = 0xffffffffu, Mv = 0, cout = 0, accP = 0, accM = 0;
    __syncwarp();
    int s = 0;
    const int nchunk = (nsteps + 31) >> 5;
    // one step of the anti-diagonal sweep
#define EXT_STEP(RESET)                                                                                           \
    {                                                                                                             \
        uint32_t cin = __shfl_up_sync(0xffffffffu, cout, 1, WIDTH);                                               \
        if (lid == 0) cin = 0x80000000u | ((uint32_t)st[s] << 7);       /* row 0 of the matrix: D[0][j] = j */    \
        const uint32_t symoff = cin & 0x180u;                                                                     \
        uint32_t Eq = *reinterpret_cast<const uint32_t*>(pe_base + symoff);                                       \
        const uint32_t Xv = Eq | Mv;                                                                              \
        Eq |= cin & 1u;                                                                                           \
        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;                                                         \
        const uint32_t Ph = Mv | ~(Xh | Pv);                                                                      \
        const uint32_t Mh = Pv & Xh;                                                                              \
        cout = (Ph & 0x80000000u) | symoff | (Mh >> 31);                                                          \
        accP = __funnelshift_r(accP, Ph >> lb, 1);                                                                \
        accM = __funnelshift_r(accM, Mh >> lb, 1);                                                                \
        const uint32_t Phs = (Ph << 1) | (cin >> 31);                                                             \
        const uint32_t Mhs = (Mh << 1) | (cin & 1u);                                                              \
        Pv = Mhs | ~(Xv | Phs);                                                                                   \
        Mv = Phs & Xv;                                                                                            \
        if ((unsigned)rel < (unsigned)W) row[rel] = make_uint2(Pv, Ph);                                           \
        if (RESET && s == lid - 1) { Pv = 0xffffffffu; Mv = 0; }        /* my first column is next */             \
        ++s; ++rel;                                                                                               \
    }
    for (int c = 0; c < nchunk; ++c) {
        const int n = nsteps - 32 * c < 32 ? nsteps - 32 * c : 32;
        if (c == 0 || n < 32) {
            // the first chunk (lanes meet their first column: lid <= 31) and the last, partial one
#pragma unroll 4
            for (int b = 0; b < n; ++b) EXT_STEP(true)
        } else {
#pragma unroll 8
            for (int b = 0; b < 32; ++b) EXT_STEP(false)
        }
        if (lid == lw) S.sb[half][c] = make_uint2(accP >> (32 - n), accM >> (32 - n));     // bit b: step 32 c + b
    }
#undef EXT_STEP
    __syncwarp();
    // ---- first minimum of qblk + sum of (+1, -1) over the columns: lane = 32 steps
    int best = 0x7fffffff, bests = 0x7fffffff, carry = 0;
    for (int c0 = 0; c0 < nchunk; c0 += WIDTH) {
        const int c = c0 + lid;
        uint32_t bp = 0, bm = 0;
        if (c < nchunk) { const uint2 v = S.sb[half][c]; bp = v.x; bm = v.y; }
        // steps of lane lw that are columns of the block: [lw, lw + tblk)
        const int lo = lw - 32 * c, hi = lw + tblk - 32 * c;
        uint32_t vm = 0xffffffffu;
        if (lo > 0) vm = lo >= 32 ? 0u : vm << lo;
        if (hi < 32) vm = hi <= 0 ? 0u : vm & (0xffffffffu >> (32 - hi));
        bp &= vm; bm &= vm;
        const int tot = __popc(bp) - __popc(bm);
        int pre = tot;                                                  // inclusive sum over the lanes of the group
#pragma unroll
        for (int d = 1; d < WIDTH; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, pre, d, WIDTH); if (lid >= d) pre += o; }
        int run = qblk + carry + pre - tot, lmin = 0x7fffffff, lpos = 0;
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
            run += (int)((bp >> b) & 1u) - (int)((bm >> b) & 1u);
            if (((vm >> b) & 1u) && run < lmin) { lmin = run; lpos = b; }
        }
        if (lmin < best) { best = lmin; bests = 32 * c + lpos; }
        carry += __shfl_sync(0xffffffffu, pre, WIDTH - 1, WIDTH);
    }
#pragma unroll
    for (int d = WIDTH >> 1; d > 0; d >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, d, WIDTH), os = __shfl_xor_sync(0xffffffffu, bests, d, WIDTH);
        if (ob < best || (ob == best && os < bests)) { best = ob; bests = os; }
    }
    *best_out = best;
    *bestj_out = best == 0x7fffffff ? -1 : bests - lw;
}

// ---- whole-warp pieces ------------------------------------------------------------------------------------------------

// Unpacks the bases of the block of candidate `slot` (both reads) into H: one 32-bit word of 16 bases per lane and
// trip, written as one 16-byte store.  Base x of the block lands at q[qa + x] (t[ta + x]).
__device__ __noinline__ int ext_stage(WarpScratch& S, HalfScratch& H, const ExtParams& P, int slot, int lane)
{
    const int dir = XF(X_DIR, slot);
    int qa = 0, ta = 0;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const uint8_t* pac = side ? P.pac : P.q_pac;
        const uint64_t off = side ? (((uint64_t)(uint32_t)XF(X_TOFF_HI, slot) << 32) | (uint32_t)XF(X_TOFF_LO, slot))
                                  : (((uint64_t)(uint32_t)XF(X_QOFF_HI, slot) << 32) | (uint32_t)XF(X_QOFF_LO, slot));
        const int len = side ? XF(X_TLEN, slot) : XF(X_QLEN, slot);
        const int strand = side ? 0 : XF(X_QSTRAND, slot);
        const int blk = side ? XF(X_TBLK, slot) : XF(X_QBLK, slot);
        const int s0 = (side ? XF(X_TORG, slot) : XF(X_QORG, slot)) + dir * (side ? XF(X_TIDX, slot) : XF(X_QIDX, slot));   // sequence index of x = 0
        uint8_t* dst = side ? H.t : H.q;
        // physical base index of x: p0 + g x
        const long long p0 = strand ? (long long)off + (len - 1 - s0) : (long long)off + s0;
        const int g = strand ? -dir : dir;
        const long long plo = g > 0 ? p0 : p0 - (blk - 1), phi = g > 0 ? p0 + (blk - 1) : p0;
        const int a = g > 0 ? (int)(p0 & 15) : 15 - (int)(p0 & 15);
        const long long w0 = g > 0 ? (p0 >> 4) : (p0 >> 4);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(pac);
        for (long long w = (plo >> 4) + lane; w <= (phi >> 4); w += 32) {
            const uint32_t v = __ldg(words + w);
            // byte i of v holds bases 4 i .. 4 i + 3, first base in its top bits
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t x = (v >> (8 * i)) & 0xffu;
                uint32_t r = (x >> 6) | (((x >> 4) & 3u) << 8) | (((x >> 2) & 3u) << 16) | ((x & 3u) << 24);
                if (strand) r ^= 0x03030303u;
                o[i] = r;
            }
            uint4 st4;
            long long idx;
            if (g > 0) { st4 = make_uint4(o[0], o[1], o[2], o[3]); idx = (w - w0) * 16; }
            else {
                st4 = make_uint4(__byte_perm(o[3], 0, 0x0123), __byte_perm(o[2], 0, 0x0123), __byte_perm(o[1], 0, 0x0123), __byte_perm(o[0], 0, 0x0123));
                idx = (w0 - w) * 16;
            }
            *reinterpret_cast<uint4*>(dst + idx) = st4;
        }
        if (side) ta = a; else qa = a;
    }
    __syncwarp();
    return qa | (ta << 8);
}

// End of the right side: assemble the result (oc_aligner.c:385-451).
__device__ __noinline__ void ext_finish(WarpScratch& S, const ExtParams& P, int slot, int lane)
{
    const uint64_t ops_base = ((uint64_t)(uint32_t)XF(X_OPS_HI, slot) << 32) | (uint32_t)XF(X_OPS_LO, slot);
    uint8_t* ops = P.ops + ops_base;
    const int R = XF(X_NCOL, slot), lcap = XF(X_LCAP, slot), L = XF(X_L, slot), left_keep = XF(X_LKEEP, slot), acnt = XF(X_ACNT, slot);
    int QS = XF(X_QS, slot), TS = XF(X_TS, slot);
    uint8_t* rbase = ops + (left_keep ? lcap - acnt : lcap);
    int rstart = 0, right_keep = R;
    if (!left_keep) {                                   // :385-415
        int a2 = 0, q2 = 0, t2 = 0, m2 = 0;
        for (int x = 0; x < R; ++x) {
            const int op = rbase[x];
            q2 += op != OP_TONLY; t2 += op != OP_QONLY;
            m2 = op == OP_MATCH ? m2 + 1 : 0;
            ++a2;
            if (m2 == 8) break;
        }
        if (m2 == 8) { rstart = a2 - 8; QS += q2 - 8; TS += t2 - 8; right_keep = R - rstart; }
        else right_keep = 0;
    }
    // counts over the kept parts
    const uint8_t* fin = left_keep ? ops + (lcap - L) : rbase + rstart;
    const int n_left = left_keep, n_all = left_keep + right_keep;
    int lq = 0, lt = 0, rq = 0, rt = 0, nm = 0;
    int ins_carry = 0, long_ins = 0;                    // a run of >= 255 query-only columns (align_tags.c:40-44)
    for (int base = 0; base < n_all; base += 32) {
        const int x = base + lane;
        const int op = x < n_all ? fin[x] : OP_MATCH;
        if (x < n_all) {
            const int cq = op != OP_TONLY, ct = op != OP_QONLY;
            if (x < n_left) { lq += cq; lt += ct; } else { rq += cq; rt += ct; }
            nm += op == OP_MATCH;
        }
        const uint32_t ins = __ballot_sync(0xffffffffu, x < n_all && op == OP_QONLY);
        if (ins == 0xffffffffu) { ins_carry += 32; if (ins_carry >= 255) long_ins = 1; }
        else {
            if (ins_carry + __ffs(~ins) - 1 >= 255) long_ins = 1;       // ones at the low end continue the carried run
            ins_carry = __clz(~ins);                                    // ones at the high end (inner runs are < 32)
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        lq += __shfl_xor_sync(0xffffffffu, lq, d); lt += __shfl_xor_sync(0xffffffffu, lt, d);
        rq += __shfl_xor_sync(0xffffffffu, rq, d); rt += __shfl_xor_sync(0xffffffffu, rt, d);
        nm += __shfl_xor_sync(0xffffffffu, nm, d);
    }
    __syncwarp();
    if (lane == 0) {
        nb_alignment A;
        A.qoff = QS - lq; A.qend = QS + rq; A.toff = TS - lt; A.tend = TS + rt;
        A.aln_size = n_all; A.n_match = nm;
        A.ok = n_all >= P.min_align_size;
        A.long_ins = long_ins;
        A.ops_offset = (uint64_t)(fin - P.ops);
        P.results[XF(X_TI, slot)] = A;
        XF(X_STATE, slot) = ST_IDLE;
    }
    __syncwarp();
}

// ---- per-lane pieces (slot = lane) ---------------------------------------------------------------------------------------

struct LaneCtx {            // the views a lane needs of its candidate
    SeqView qv, tv;
    uint8_t* ops;           // the candidate's region of the pool
};

__device__ __forceinline__ LaneCtx lane_ctx(WarpScratch& S, const ExtParams& P, int lane)
{
    LaneCtx C;
    C.qv.pac = P.q_pac; C.qv.off = ((uint64_t)(uint32_t)XF(X_QOFF_HI, lane) << 32) | (uint32_t)XF(X_QOFF_LO, lane);
    C.qv.len = XF(X_QLEN, lane); C.qv.strand = XF(X_QSTRAND, lane);
    C.tv.pac = P.pac; C.tv.off = ((uint64_t)(uint32_t)XF(X_TOFF_HI, lane) << 32) | (uint32_t)XF(X_TOFF_LO, lane);
    C.tv.len = XF(X_TLEN, lane); C.tv.strand = 0;
    C.ops = P.ops + (((uint64_t)(uint32_t)XF(X_OPS_HI, lane) << 32) | (uint32_t)XF(X_OPS_LO, lane));
    return C;
}

// Starts a side of the extension: columns go to ops[out_base + x * out_dir].
__device__ __forceinline__ void lp_begin_side(WarpScratch& S, int lane, int phase, int q_origin, int q_total, int t_origin, int t_total,
                                              int dir, int out_base, int out_dir)
{
    XF(X_PHASE, lane) = phase; XF(X_QORG, lane) = q_origin; XF(X_QTOT, lane) = q_total; XF(X_TORG, lane) = t_origin; XF(X_TTOT, lane) = t_total;
    XF(X_DIR, lane) = dir; XF(X_OUTB, lane) = out_base; XF(X_OUTDIR, lane) = out_dir; XF(X_QIDX, lane) = 0; XF(X_TIDX, lane) = 0; XF(X_NCOL, lane) = 0;
}

// Cuts the next block of the current side.  False: nothing left on one of the reads.
__device__ __forceinline__ bool lp_next_block(WarpScratch& S, const ExtParams& P, int lane)
{
    const int bs = P.block_size;
    const int qleft = XF(X_QTOT, lane) - XF(X_QIDX, lane), tleft = XF(X_TTOT, lane) - XF(X_TIDX, lane);
    int qblk, tblk, last;
    if (qleft < bs + 100 || tleft < bs + 100) {             // oc_aligner.c:127-137
        qblk = (int)__dmul_rn((double)tleft, 1.3); qblk = qblk < qleft ? qblk : qleft;
        tblk = (int)__dmul_rn((double)qleft, 1.3); tblk = tblk < tleft ? tblk : tleft;
        last = 1;
    } else { qblk = tblk = bs; last = 0; }
    if (qblk == 0 || tblk == 0) return false;
    XF(X_QBLK, lane) = qblk; XF(X_TBLK, lane) = tblk; XF(X_LAST, lane) = last; XF(X_WFLAGS, lane) = 0;
    {
        const int d = qblk - tblk, wide = d > 32 || d < -32;
        XF(X_NBW, lane) = wide ? 8 : 4; XF(X_SHIFT, lane) = wide ? d / 2 : 0;
    }
    XF(X_STATE, lane) = ST_BLOCK;
    return true;
}

// The current side has ended.  Left side: re-anchor at the first run of 8 matches counted from the seed
// (oc_aligner.c:329-367) and start the right side; right side: the candidate waits for ext_finish.
__device__ __noinline__ void lp_end_side(WarpScratch& S, const ExtParams& P, int lane)
{
    if (XF(X_PHASE, lane) == 0) {
        const uint8_t* ops = P.ops + (((uint64_t)(uint32_t)XF(X_OPS_HI, lane) << 32) | (uint32_t)XF(X_OPS_LO, lane));
        const int L = XF(X_NCOL, lane), lcap = XF(X_LCAP, lane);
        int acnt = 0, qcnt = 0, tcnt = 0, m = 0;
        for (int x = 0; x < L; ++x) {
            const int op = ops[lcap - 1 - x];
            qcnt += op != OP_TONLY; tcnt += op != OP_QONLY;
            m = op == OP_MATCH ? m + 1 : 0;
            ++acnt;
            if (m == 8) break;
        }
        int QS = XF(X_QS, lane), TS = XF(X_TS, lane), left_keep = 0;
        if (m == 8) { QS -= qcnt; TS -= tcnt; left_keep = L - acnt; }
        else acnt = 0;
        XF(X_QS, lane) = QS; XF(X_TS, lane) = TS; XF(X_L, lane) = L; XF(X_LKEEP, lane) = left_keep; XF(X_ACNT, lane) = acnt;
        // final alignment = ops[lcap - L, lcap - acnt) ++ right extension written from lcap - acnt
        lp_begin_side(S, lane, 1, QS, XF(X_QLEN, lane) - QS, TS, XF(X_TLEN, lane) - TS, +1, left_keep ? lcap - acnt : lcap, +1);
        if (lp_next_block(S, P, lane)) return;
    }
    XF(X_STATE, lane) = ST_FINISH;
}

// Fetches candidate ti into my slot and cuts its first block.
__device__ __forceinline__ void lp_start_task(WarpScratch& S, const ExtParams& P, uint32_t ti, int lane)
{
    const ExtTask T = P.tasks[ti];
    const uint64_t qo = P.q_off[T.qid], to = P.seq_off[T.sid];
    XF(X_TI, lane) = (int)ti;
    XF(X_QOFF_LO, lane) = (int)(uint32_t)qo; XF(X_QOFF_HI, lane) = (int)(uint32_t)(qo >> 32);
    XF(X_TOFF_LO, lane) = (int)(uint32_t)to; XF(X_TOFF_HI, lane) = (int)(uint32_t)(to >> 32);
    XF(X_QLEN, lane) = (int)P.q_len[T.qid]; XF(X_QSTRAND, lane) = T.qdir; XF(X_TLEN, lane) = (int)P.seq_len[T.sid];
    XF(X_OPS_LO, lane) = (int)(uint32_t)T.ops_base; XF(X_OPS_HI, lane) = (int)(uint32_t)(T.ops_base >> 32);
    XF(X_LCAP, lane) = (int)T.lcap; XF(X_QS, lane) = T.qoff; XF(X_TS, lane) = T.soff;
    XF(X_L, lane) = 0; XF(X_LKEEP, lane) = 0; XF(X_ACNT, lane) = 0;
    // left of the seed, read right to left (oc_aligner.c:313-327); column x lives at lcap - 1 - x
    lp_begin_side(S, lane, 0, T.qoff - 1, T.qoff, T.soff - 1, T.soff, -1, (int)(T.lcap ? T.lcap - 1 : 0), -1);
    if (!lp_next_block(S, P, lane)) lp_end_side(S, P, lane);
}

// After the forward pass of my block: walk back from (qblk, j*) preferring up, then left, then diagonal
// (edlib_ex.c:463,500,540), keep the part up to the last run of M matches (oc_aligner.c:213-262), then cut the next
// block or end the side.  The walk reads the band of (Pv, Ph) words the forward pass left in `mat`; its columns go,
// last to first, to `rev` (four op bytes per word).
__device__ __forceinline__ void lp_walk_post(WarpScratch& S, const ExtParams& P, const uint2* __restrict__ mat, uint32_t* __restrict__ rev, int lane)
{
    const LaneCtx C = lane_ctx(S, P, lane);
    const int qblk = XF(X_QBLK, lane), tblk = XF(X_TBLK, lane), last = XF(X_LAST, lane), best = XF(X_BEST, lane), bestj = XF(X_BESTJ, lane);
    const int dir = XF(X_DIR, lane);
    const int q0 = XF(X_QORG, lane) + dir * XF(X_QIDX, lane), t0 = XF(X_TORG, lane) + dir * XF(X_TIDX, lane);     // sequence index of base 0 of the block
    const int mn = qblk < tblk ? qblk : tblk;
    const int k0 = (int)__dmul_rn(__dmul_rn((double)mn, P.error), 1.1);         // edlib_ex.c:751
    const bool failed = best > k0;
    int done = last;
    if (failed && qblk > 30 && tblk > 30) done = 1;                             // oc_aligner.c:220 (qfae = tfae = 0)
    const int M = done ? P.tail_match_len : 8;
    int asize = 0, found = 0, acnt = 0, qcnt = 0, tcnt = 0;
    // bases of the two reads through one cached 16-base word each
    uint32_t qw = 0, tw = 0;
    long long qwi = -1, twi = -1;
    auto cached_base = [](const SeqView& v, int idx, uint32_t& w, long long& wi) -> int {
        const uint64_t p = v.strand ? v.off + (uint64_t)(v.len - 1 - idx) : v.off + (uint64_t)idx;
        const long long i = (long long)(p >> 4);
        if (i != wi) { w = __ldg(reinterpret_cast<const uint32_t*>(v.pac) + i); wi = i; }
        const int k = (int)(p & 15);
        const int b = (w >> (8 * (k >> 2) + 2 * (3 - (k & 3)))) & 3;
        return v.strand ? 3 - b : b;
    };
    if (!failed) {
        int i = qblk, j = bestj + 1, r = 0, m = 0;
        uint32_t acc = 0;
        if (XF(X_WFLAGS, lane) & 2) {           // continue after the band was moved
            i = XF(X_WI, lane); j = XF(X_WJ, lane); r = XF(X_WR, lane); found = XF(X_WFLAGS, lane) & 1;
            acnt = XF(X_WACNT, lane); qcnt = XF(X_WQCNT, lane); tcnt = XF(X_WTCNT, lane); m = XF(X_WM, lane); acc = (uint32_t)XF(X_WACC, lane);
        }
        const int shift = XF(X_SHIFT, lane), nbw = XF(X_NBW, lane), W = 32 * nbw;
        const int c0 = 32 * ((nbw >> 1) - 1) + 16 + shift;              // column - cl(w) = column - 32 w + c0
        const int band_lo = P.band_test ? (W >> 1) - 32 : 0, band_hi = P.band_test ? (W >> 1) + 32 : W;
        bool moved = false;
        // One cell per trip, no branch inside a trip (the lanes of the warp walk different candidates: whatever one of them
        // branches on, all of them pay for): the two stored bits select the move, the bases are compared on every trip.
        auto book = [&](int op) {
            acc |= (uint32_t)op << (8 * (r & 3));
            if ((r & 3) == 3) { rev[r >> 2] = acc; acc = 0; }
            ++r;
            const int nf = found ^ 1;
            qcnt += nf & (op != OP_TONLY); tcnt += nf & (op != OP_QONLY); acnt += nf;
            m = op == OP_MATCH ? m + 1 : 0;
            found |= (m == M);
        };
        while (i > 0 && j > 0) {
            const int wi = (i - 1) >> 5;
            const int rel = j - 1 - 32 * wi + c0;
            if ((unsigned)(rel - band_lo) >= (unsigned)(band_hi - band_lo)) { moved = true; break; }
            const uint2 v = mat[wi * W + rel];
            const int bit = (i - 1) & 31;
            const uint32_t up = (v.x >> bit) & 1u;                      // cell = upper + 1: query base against a gap
            const uint32_t lf = ((v.y >> bit) & 1u) & (up ^ 1u);        // cell = left + 1: gap against a template base
            const uint32_t dg = (up | lf) ^ 1u;
            const int qb = cached_base(C.qv, q0 + dir * (i - 1), qw, qwi), tb = cached_base(C.tv, t0 + dir * (j - 1), tw, twi);
            const int op = up ? OP_QONLY : (lf ? OP_TONLY : (qb == tb ? OP_MATCH : OP_MISMATCH));
            i -= (int)(up | dg); j -= (int)(lf | dg);
            book(op);
        }
        if (!moved) {
            for (; j > 0; --j) book(OP_TONLY);                         // row 0 of the matrix
            for (; i > 0; --i) book(OP_QONLY);                         // column 0
        }
        if (moved) {
            // the walk left the band: the forward pass runs again with the band on this diagonal
            XF(X_WI, lane) = i; XF(X_WJ, lane) = j; XF(X_WR, lane) = r; XF(X_WFLAGS, lane) = 2 | found;
            XF(X_WACNT, lane) = acnt; XF(X_WQCNT, lane) = qcnt; XF(X_WTCNT, lane) = tcnt; XF(X_WM, lane) = m; XF(X_WACC, lane) = (int)acc;
            XF(X_SHIFT, lane) = i - j;
            XF(X_STATE, lane) = ST_BLOCK;
            atomicAdd(P.counters + 3, 1ull);
            return;
        }
        if (r & 3) rev[r >> 2] = acc;
        asize = r;
    }
    uint8_t* out = C.ops + XF(X_OUTB, lane);
    const int out_dir = XF(X_OUTDIR, lane), ncol = XF(X_NCOL, lane);
    const int kcut = asize - acnt;
    if (!found || kcut < 1) {
        // exact-match prefix only, then stop (:243-253)
        int n = 0;
        while (n < mn && cached_base(C.qv, q0 + dir * n, qw, qwi) == cached_base(C.tv, t0 + dir * n, tw, twi)) ++n;
        for (int x = 0; x < n; ++x) out[(long long)(ncol + x) * out_dir] = OP_MATCH;
        XF(X_NCOL, lane) = ncol + n;
        lp_end_side(S, P, lane);
        return;
    }
    const int keep = kcut + (done ? M : 0);
    {
        // out[(ncol + x) * out_dir] = rev[asize - 1 - x], x < keep: in address order d[u], u < keep, that is a copy of
        // rev[asize - keep + u] (left side) or of rev[asize - 1 - u] (right side); four bytes at a time where d is aligned
        const uint8_t* rb = reinterpret_cast<const uint8_t*>(rev);
        uint8_t* d = out_dir > 0 ? out + ncol : out - (ncol + keep - 1);
        const int sidx = out_dir > 0 ? asize - 1 : asize - keep, sdir = out_dir > 0 ? -1 : 1;
        int u = 0;
        while (u < keep && ((uintptr_t)(d + u) & 3)) { d[u] = rb[sidx + sdir * u]; ++u; }
        for (; u + 4 <= keep; u += 4) {
            const int lo = sdir > 0 ? sidx + u : sidx - u - 3;                          // the four source bytes are rev[lo .. lo + 3]
            const uint32_t w0 = rev[lo >> 2], w1 = rev[(lo >> 2) + 1];
            uint32_t w = __funnelshift_r(w0, w1, 8 * (lo & 3));
            if (sdir < 0) w = __byte_perm(w, 0, 0x0123);
            *reinterpret_cast<uint32_t*>(d + u) = w;
        }
        for (; u < keep; ++u) d[u] = rb[sidx + sdir * u];
    }
    XF(X_QIDX, lane) += qblk - qcnt;            // qfae = qblk
    XF(X_TIDX, lane) += (bestj + 1) - tcnt;     // tfae = j*
    XF(X_NCOL, lane) = ncol + keep;
    if (done || !lp_next_block(S, P, lane)) lp_end_side(S, P, lane);
}

#ifndef EXT_MIN_CTAS
#define EXT_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(EXT_WARPS * 32, EXT_MIN_CTAS)
k_extend(ExtParams P)
{
    extern __shared__ __align__(16) uint8_t ext_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(ext_smem)[wid];
    const size_t wslot = ((size_t)blockIdx.x * EXT_WARPS + wid) * EXT_SLOTS;
    uint2* mat = P.mat + (wslot + (lane & (EXT_SLOTS - 1))) * EXT_MAT_STRIDE;          // of my slot
    uint32_t* rev = reinterpret_cast<uint32_t*>(P.rev + (wslot + (lane & (EXT_SLOTS - 1))) * EXT_REV_STRIDE);
    unsigned long long nb = 0, nc = 0, nwd = 0;
    // A candidate is a chain of block alignments, and a warp runs the forward passes of its slots one pair after the other: the
    // chain of a 60 kb candidate in a warp of 16 takes half as long as the whole launch of one GPU, and longer than the launch
    // of one GPU of eight.  So the longest candidates (the front of the task list) have a queue of their own, served by a few
    // warps that keep only two in flight; when it is empty those warps join the others.
    bool long_warp = (uint32_t)(blockIdx.x * EXT_WARPS + wid) < P.n_long_warps;
    bool has = lane < (long_warp ? 2 : P.slots);        // this lane owns a slot
    if (lane < EXT_SLOTS) XF(X_STATE, lane) = ST_IDLE;
    __syncwarp();
    bool drained = false;
    for (;;) {
        // ---- idle slots take candidates from the queue
        const uint32_t idle = __ballot_sync(0xffffffffu, has && XF(X_STATE, lane) == ST_IDLE);
        if (idle && !drained) {
            uint32_t base = 0, lim = 0, moved = 0;
            if (lane == 0) {
                const uint32_t want = (uint32_t)__popc(idle);
                int q = long_warp ? 1 : 0;
                base = atomicAdd(P.next_task + q, want); lim = q ? P.n_long : P.n_tasks;
                if (base >= lim) { q ^= 1; moved = 1; base = atomicAdd(P.next_task + q, want); lim = q ? P.n_long : P.n_tasks; }    // that queue is empty: the other one
            }
            base = __shfl_sync(0xffffffffu, base, 0); lim = __shfl_sync(0xffffffffu, lim, 0); moved = __shfl_sync(0xffffffffu, moved, 0);
            if (has && XF(X_STATE, lane) == ST_IDLE) {
                const uint32_t ti = base + (uint32_t)__popc(idle & lt_mask);
                if (ti < lim) lp_start_task(S, P, ti, lane);
            }
            if (base >= lim) drained = true;                                    // both are
            if (moved && long_warp) { long_warp = false; has = lane < P.slots; }     // from now on a warp like the others
        }
        __syncwarp();
        // ---- candidates whose right side has ended: result record, whole warp
        uint32_t fin = __ballot_sync(0xffffffffu, has && XF(X_STATE, lane) == ST_FINISH);
        while (fin) {
            const int slot = __ffs(fin) - 1;
            fin &= fin - 1;
            ext_finish(S, P, slot, lane);
        }
        const uint32_t blk = __ballot_sync(0xffffffffu, has && XF(X_STATE, lane) == ST_BLOCK);
        if (!blk) {
            if (drained && __ballot_sync(0xffffffffu, has && XF(X_STATE, lane) != ST_IDLE) == 0) break;
            continue;
        }
        // ---- forward passes: side by side in the two half-warps when both blocks fit 16 lanes
        uint32_t small = blk & __ballot_sync(0xffffffffu, has && XF(X_QBLK, lane) <= 512), big = blk & ~small;
        while (small | big) {
            int sa, sb = -1;
            if (small) {
                sa = __ffs(small) - 1; small &= small - 1;
                if (small) { sb = __ffs(small) - 1; small &= small - 1; }
            } else { sa = __ffs(big) - 1; big &= big - 1; }
            const int qb0 = XF(X_QBLK, sa), tb0 = XF(X_TBLK, sa);
            const int a0 = ext_stage(S, S.h[0], P, sa, lane), a1 = sb >= 0 ? ext_stage(S, S.h[1], P, sb, lane) : 0;
            const int qa0 = a0 & 255, ta0 = a0 >> 8, qa1 = a1 & 255, ta1 = a1 >> 8;
            int best, bestj;
            if (qb0 <= 512) {
                const int hh = lane >> 4;
                const int sl = hh ? sb : sa;
                const int qb = sl >= 0 ? XF(X_QBLK, sl) : 0, tb = sl >= 0 ? XF(X_TBLK, sl) : 0;
                myers_forward<16>(S, S.h[hh].q + (hh ? qa1 : qa0), S.h[hh].t + (hh ? ta1 : ta0), qb, tb,
                                  P.mat + (wslot + (sl >= 0 ? sl : 0)) * EXT_MAT_STRIDE, sl >= 0 ? XF(X_SHIFT, sl) : 0, sl >= 0 ? XF(X_NBW, sl) : 4, lane, &best, &bestj);
                __syncwarp();
                if ((lane & 15) == 0 && sl >= 0) { XF(X_BEST, sl) = best; XF(X_BESTJ, sl) = bestj; XF(X_STATE, sl) = ST_WALK; }
            } else {
                myers_forward<32>(S, S.h[0].q + qa0, S.h[0].t + ta0, qb0, tb0, P.mat + (wslot + sa) * EXT_MAT_STRIDE, XF(X_SHIFT, sa), XF(X_NBW, sa), lane, &best, &bestj);
                __syncwarp();
                if (lane == 0) { XF(X_BEST, sa) = best; XF(X_BESTJ, sa) = bestj; XF(X_STATE, sa) = ST_WALK; }
            }
            if (lane == 0) {
                nb += 1; nc += (unsigned long long)tb0; nwd += (unsigned long long)tb0 * ((qb0 + 31) >> 5);
                if (sb >= 0) { const int q1 = XF(X_QBLK, sb), t1 = XF(X_TBLK, sb); nb += 1; nc += (unsigned long long)t1; nwd += (unsigned long long)t1 * ((q1 + 31) >> 5); }
            }
            __syncwarp();
        }
        // ---- walk-backs and bookkeeping: every lane on its own candidate
        if (has && XF(X_STATE, lane) == ST_WALK) lp_walk_post(S, P, mat, rev, lane);
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(P.counters + 0, nb);
        atomicAdd(P.counters + 1, nc);
        atomicAdd(P.counters + 2, nwd);
    }
}

// Launches the persistent extension kernel over d_tasks.  d_ops must hold every task's region.
int nb_run_extension(nb_ctx* ctx, const nb_volume* queries, const nb_volume* templates, const ExtTask* d_tasks, uint32_t n_tasks,
                     nb_alignment* d_results, uint8_t* d_ops, double error, int min_align_size, int tail_match_len)
{
    if (n_tasks == 0) return NB_OK;
    cudaStream_t st = ctx->stream;
    static const int kCtas = getenv("NB_EXT_CTAS") ? atoi(getenv("NB_EXT_CTAS")) : EXT_MIN_CTAS;
    int ctas_per_sm = kCtas < 1 ? 1 : kCtas;        // x 8 warps x 32 candidates per SM
    if (ctx->ext_ctas > 0 && ctx->ext_ctas < ctas_per_sm) ctas_per_sm = ctx->ext_ctas;
    unsigned grid = NB_SMS * ctas_per_sm;
    // A small launch spreads its candidates over all the warps there are: a candidate is a chain of block alignments, and a
    // warp runs the forward passes of its slots one pair after the other -- with few candidates per warp the chain of each
    // is that much shorter (the later extension rounds of a consensus job are a few hundred candidates).
    int n_slots = (int)((n_tasks + (uint64_t)grid * EXT_WARPS - 1) / ((uint64_t)grid * EXT_WARPS));
    n_slots = (n_slots + 1) & ~1;
    if (n_slots < 2) n_slots = 2;
    if (n_slots > EXT_SLOTS) n_slots = EXT_SLOTS;
    {
        // no more warps than there are candidates to fill their slots
        const unsigned want = (unsigned)((n_tasks + EXT_WARPS * (unsigned)n_slots - 1) / (EXT_WARPS * (unsigned)n_slots));
        if (grid > want) grid = want ? want : 1;
    }
    const size_t slots = (size_t)NB_SMS * (kCtas < 1 ? 1 : kCtas) * EXT_WARPS * EXT_SLOTS;      // candidates in flight (sized for the full grid)
    // the scratch belongs to the job's context: its workers' launches never run side by side (nb_consensus_job's GPU lock)
    nb_ctx* own = (ctx->parent && !ctx->ext_private) ? ctx->parent : ctx;
    if (!own->ext_mat || own->ext_mat_slots < slots) {
        if (own->ext_mat) cudaFree(own->ext_mat);
        own->ext_mat = NULL;
        NB_CUDA(ctx, cudaMalloc(&own->ext_mat, (sizeof(uint2) * EXT_MAT_STRIDE + EXT_REV_STRIDE) * slots));
        own->ext_mat_slots = slots;
    }
    uint2* s_mat = (uint2*)own->ext_mat;
    uint8_t* s_rev = (uint8_t*)own->ext_mat + sizeof(uint2) * EXT_MAT_STRIDE * own->ext_mat_slots;
    NbDevBuf B(ctx);
    uint32_t* d_next = NULL;
    unsigned long long* d_cnt = NULL;
    { int rc; if ((rc = B.get(&d_next, 2))) return rc; if ((rc = B.get(&d_cnt, 4))) return rc; }
    // the queue of the longest candidates: 2 % of a large launch, an eighth of the warps
    static const int kLongQ = getenv("NB_EXT_LONGQ") ? atoi(getenv("NB_EXT_LONGQ")) : 1;
    uint32_t n_long = 0, n_long_warps = 0;
    // (measured, one rank of N on the E. coli-size set: 37.5 -> 33.7 ms at N = 4, 28.9 -> 26.1 at N = 8, nothing at N = 2, and
    //  86.2 -> 88.7 at N = 1, where the launch lasts four times the longest chain anyway: large launches go without)
    if (kLongQ == 1 && n_slots > 2 && n_tasks >= 4u * grid * EXT_WARPS && n_tasks <= 80000u) { n_long = n_tasks / 50; n_long_warps = grid * EXT_WARPS / 8; }
    if (kLongQ == 2 && n_tasks >= 4) { n_long = n_tasks / 4; n_long_warps = grid * EXT_WARPS / 2 ? grid * EXT_WARPS / 2 : 1; }      // (tests: any launch)
    const uint32_t h_next[2] = {n_long, 0u};
    NB_CUDA(ctx, cudaMemcpyAsync(d_next, h_next, sizeof h_next, cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 4, st));
    ExtParams P;
    P.pac = templates->d_pac; P.seq_off = templates->d_off; P.seq_len = templates->d_len;
    P.q_pac = queries->d_pac; P.q_off = queries->d_off; P.q_len = queries->d_len;
    P.tasks = d_tasks; P.n_tasks = n_tasks; P.results = d_results; P.ops = d_ops; P.mat = s_mat; P.rev = s_rev;
    P.next_task = d_next; P.counters = d_cnt;
    P.error = error; P.min_align_size = min_align_size; P.tail_match_len = tail_match_len; P.block_size = 512;
    P.slots = n_slots; P.n_long = n_long; P.n_long_warps = n_long_warps;
    P.band_test = getenv("NB_EXT_BAND_TEST") ? atoi(getenv("NB_EXT_BAND_TEST")) : 0;     // (read at every launch: tests switch it)
    const size_t smem = sizeof(WarpScratch) * EXT_WARPS;
    NB_CUDA(ctx, cudaFuncSetAttribute(k_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEventRecord(ctx->ev2, st);
    k_extend<<<grid, EXT_WARPS * 32, smem, st>>>(P);
    NB_LAUNCH_CHECK(ctx);
    cudaEventRecord(ctx->ev3, st);
    unsigned long long cnt[4];
    NB_CUDA(ctx, cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->cc.n_extensions += n_tasks;
    ctx->cc.n_block_aligns += cnt[0];
    ctx->cc.n_dp_columns += cnt[1];
    ctx->cc.n_dp_word_updates += cnt[2];
    if (getenv("NB_TRACE")) fprintf(stderr, "[nb_run_extension] %u tasks, %llu block passes, %llu band moves\n", n_tasks, cnt[0], cnt[3]);
    if (cnt[2] > ctx->cc.big_ext_word_updates) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
        ctx->cc.big_ext_tasks = n_tasks; ctx->cc.big_ext_block_aligns = cnt[0]; ctx->cc.big_ext_word_updates = cnt[2]; ctx->cc.big_ext_dp_columns = cnt[1];
        ctx->cc.big_ext_usec = (uint64_t)(ms * 1000.0f);
    }
    return NB_OK;
}

// ------------------------------------------------------------------ C-ABI

// gathers the final alignments out of the (sparse) task regions into one dense buffer
__global__ void k_compact_ops(const uint8_t* __restrict__ pool, const nb_alignment* __restrict__ res,
                              const uint64_t* __restrict__ dst_off, uint8_t* __restrict__ dst, uint32_t n)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint8_t* s = pool + res[i].ops_offset;
        uint8_t* d = dst + dst_off[i];
        const int len = res[i].aln_size;
        for (int x = threadIdx.x; x < len; x += blockDim.x) d[x] = s[x];
    }
}

static inline void unpack_pcan(const nb_packed_candidate* p, int* qid, int* sid, int* qdir, int* sdir, int* qoff, int* soff)
{
    const uint32_t w = p->item[0];
    *sdir = (int)(w >> 31); *qdir = (int)((w >> 30) & 1);
    *sid = (int)p->item[1]; *qid = (int)p->item[4];
    if (w & (1u << 29)) { *qoff = (int)p->item[5]; *soff = (int)p->item[2]; }      // gapped_candidate.c:44-50
    else                { *qoff = (int)p->item[6]; *soff = (int)p->item[3]; }
}

extern "C" int nb_extend_candidates(nb_ctx* ctx, const nb_volume* reads, const nb_packed_candidate* cans, uint64_t n,
                                    double error, int min_align_size, int tail_match_len, nb_alignment** out_aln,
                                    uint8_t** out_ops, uint64_t* out_ops_bytes)
{
    if (!ctx || !reads || (!cans && n) || !out_aln || !out_ops || !out_ops_bytes)
        return nb_fail(ctx, NB_EINVAL, "nb_extend_candidates: bad arguments");
    *out_aln = NULL; *out_ops = NULL; *out_ops_bytes = 0;
    if (n == 0) return NB_OK;
    if (n >= (1ull << 31)) return nb_fail(ctx, NB_ELIMIT, "too many candidates in one call");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ExtTask* h_tasks = (ExtTask*)malloc(sizeof(ExtTask) * n);
    if (!h_tasks) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    uint64_t pool = 0;
    for (uint64_t i = 0; i < n; ++i) {
        int qid, sid, qdir, sdir, qoff, soff;
        unpack_pcan(&cans[i], &qid, &sid, &qdir, &sdir, &qoff, &soff);
        if (sdir != 0 || (uint64_t)qid >= reads->nseq || (uint64_t)sid >= reads->nseq ||
            qoff < 0 || soff < 0 || (uint32_t)qoff > reads->h_len[qid] || (uint32_t)soff > reads->h_len[sid]) {
            free(h_tasks);
            return nb_fail(ctx, NB_EINVAL, "candidate %llu is not normalised / out of range", (unsigned long long)i);
        }
        ExtTask& T = h_tasks[i];
        T.qid = (uint32_t)qid; T.sid = (uint32_t)sid; T.qoff = qoff; T.soff = soff; T.qdir = qdir; T.pad_ = 0;
        const uint64_t ql = reads->h_len[qid], sl = reads->h_len[sid];
        const uint64_t mn = (uint64_t)(qoff < soff ? qoff : soff);
        T.lcap = (uint32_t)(qoff + soff);
        T.ops_cap = (uint32_t)(ql + sl + mn + 64);
        T.ops_base = pool;
        pool += (T.ops_cap + 15) & ~15ull;
    }
    ExtTask* d_tasks = NULL; nb_alignment* d_res = NULL; uint8_t* d_ops = NULL;
    NB_CUDA(ctx, cudaMalloc(&d_tasks, sizeof(ExtTask) * n));
    NB_CUDA(ctx, cudaMalloc(&d_res, sizeof(nb_alignment) * n));
    NB_CUDA(ctx, cudaMalloc(&d_ops, pool + 64));
    NB_CUDA(ctx, cudaMemcpyAsync(d_tasks, h_tasks, sizeof(ExtTask) * n, cudaMemcpyHostToDevice, st));
    cudaEventRecord(ctx->ev0, st);
    int rc = nb_run_extension(ctx, reads, reads, d_tasks, (uint32_t)n, d_res, d_ops, error, min_align_size, tail_match_len);
    cudaEventRecord(ctx->ev1, st);
    if (rc) { cudaFree(d_tasks); cudaFree(d_res); cudaFree(d_ops); free(h_tasks); return rc; }
    nb_alignment* h_res = (nb_alignment*)malloc(sizeof(nb_alignment) * n);
    if (!h_res) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    NB_CUDA(ctx, cudaMemcpyAsync(h_res, d_res, sizeof(nb_alignment) * n, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->tm.extend_ms, ctx->ev0, ctx->ev1);
    // compact the ops of the alignments into one exact-size buffer
    uint64_t total = 0;
    uint64_t* h_dst = (uint64_t*)malloc(sizeof(uint64_t) * n);
    if (!h_dst) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    for (uint64_t i = 0; i < n; ++i) { h_dst[i] = total; total += (uint64_t)h_res[i].aln_size; }
    uint8_t* h_ops = (uint8_t*)malloc(total ? total : 1);
    if (!h_ops) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    uint64_t* d_dst = NULL; uint8_t* d_dense = NULL;
    NB_CUDA(ctx, cudaMalloc(&d_dst, sizeof(uint64_t) * n));
    NB_CUDA(ctx, cudaMalloc(&d_dense, total + 16));
    NB_CUDA(ctx, cudaMemcpyAsync(d_dst, h_dst, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
    k_compact_ops<<<NB_SMS * 8, 128, 0, st>>>(d_ops, d_res, d_dst, d_dense, (uint32_t)n);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(ctx, cudaMemcpyAsync(h_ops, d_dense, total, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; ++i) h_res[i].ops_offset = h_dst[i];
    free(h_dst); cudaFree(d_dst); cudaFree(d_dense);
    cudaFree(d_tasks); cudaFree(d_res); cudaFree(d_ops);
    free(h_tasks);
    *out_aln = h_res; *out_ops = h_ops; *out_ops_bytes = total;
    return NB_OK;
}
