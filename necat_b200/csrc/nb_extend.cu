// nb_extend.cu -- block-wise Myers bit-vector extension, one warp per candidate.
//
// Replaces onc_align / oca_extend / Edlib_align (gapped_align/oc_aligner.c:157-451,
// gapped_align/edlib_ex.c:733-800).  Per 512-base block the reference runs a banded
// prefix search, a second banded global pass that stores the band, and a walk-back.
// Inside the band all cells that can lie on an optimal path are exact, so the
// result equals: best prefix distance d* over target prefixes (first minimum), fail
// if d* > (int)(min(qn,tn)*error*1.1); walk back from (qn, j*) preferring up, then
// left, then diagonal (edlib_ex.c:463,500,540).  Here:
//   * one unbanded pass, 32-bit words, lane w owns query rows [32w, 32w+32); the
//     horizontal carry travels lane to lane by shuffle along the anti-diagonal
//     (lane w works on column s-w at step s);
//   * per cell only two bits are kept for the walk-back: Pv (cell = upper + 1) and
//     Ph (cell = left + 1), written as one coalesced uint2 per lane per step into a
//     skewed [step][lane] scratch in HBM/L2;
//   * the walk-back reads that scratch through a 32-column register window and
//     finds the last run of M matches (oc_aligner.c:222-240) on the fly.
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

// Shared memory of one warp.  A warp works on TWO candidates at a time: a 512-row block fills only 16 of the 32
// lanes of the forward pass, so the two half-warps run the forward passes of two different candidates side by
// side; everything else (block loading, walk-back, bookkeeping) is done by the whole warp, one candidate after
// the other.
struct HalfScratch {
    uint8_t q[EXT_MAX_BLK];
    uint8_t t[EXT_MAX_BLK + 64];        // padded: lane 0 of the forward pass reads up to 31 columns past the block
};

// State of one candidate in flight (shared memory: written by lane 0, read by the whole warp).
struct ExtCtx {
    int      active;
    uint32_t ti;
    uint64_t qoff_, toff_;              // volume offsets of the two reads
    int      qlen, qstrand, tlen;
    uint64_t ops_base;
    int      lcap;
    int      QS, TS;                    // extension origin; moved to the start of the kept left part afterwards
    int      phase;                     // 0 = left of the seed, 1 = right of it
    int      q_origin, t_origin, q_total, t_total, dir;
    long long out_base;                 // column x of the current side lives at ops[out_base + x * out_dir]
    int      out_dir;
    int      qidx, tidx, ncol;
    int      qblk, tblk, last;          // current block
    int      best, bestj;               // forward pass result
    int      L, left_keep, acnt;        // left side: columns, kept columns, columns cut at the seed end
};

struct WarpScratch {
    HalfScratch h[2];
    uint8_t     rev[2 * EXT_MAX_BLK];
    ExtCtx      ctx[2];
};

// Forward pass over one block, WIDTH lanes per block (16: two blocks per warp, 32: one).  Lane `lid` of the group
// owns query rows [32 lid, 32 lid + 32) and works on column s - lid at step s.  What travels from lane to lane
// with one shuffle per step is the horizontal delta of the row above (bit 31: +1, bit 30: -1) together with the
// symbol of that column (bits 0..1) -- the neighbour needs exactly that column next; the four match masks stay
// in registers.  Returns the best prefix distance and its column in every lane of the group.
template <int WIDTH>
__device__ __forceinline__ void myers_forward(const uint8_t* __restrict__ sq, const uint8_t* __restrict__ st, int qblk, int tblk,
                                              uint2* __restrict__ mat, int lid, int* best_out, int* bestj_out)
{
    const int nw = (qblk + 31) >> 5;
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
    uint32_t Pv = 0xffffffffu, Mv = 0;
    const int lw = qblk > 0 ? (qblk - 1) >> 5 : 0, lb = (qblk - 1) & 31;
    int score = qblk, best = 0x7fffffff, bestj = -1;
    int nsteps = qblk > 0 ? tblk + nw - 1 : 0;
    if (WIDTH < 32) { const int o = __shfl_xor_sync(0xffffffffu, nsteps, WIDTH); nsteps = nsteps > o ? nsteps : o; }
    const bool mine = lid < nw;
    uint2* mp = mat + lid;                                  // skewed [step][lane] layout, advanced by nw per step
    uint32_t cout = 0;
    for (int s = 0; s < nsteps; ++s, mp += nw) {
        uint32_t cin = __shfl_up_sync(0xffffffffu, cout, 1, WIDTH);
        if (lid == 0) cin = 0x80000000u | (uint32_t)(st[s] & 3);        // row 0 of the matrix: D[0][j] = j
        const int col = s - lid;
        if (mine && (unsigned)col < (unsigned)tblk) {
            const uint32_t tc = cin & 3;
            const uint32_t elo = (tc & 1) ? pe1 : pe0, ehi = (tc & 1) ? pe3 : pe2;
            uint32_t Eq = (tc & 2) ? ehi : elo;
            const uint32_t Xv = Eq | Mv;
            Eq |= (cin >> 30) & 1;
            const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            const uint32_t Ph = Mv | ~(Xh | Pv);
            const uint32_t Mh = Pv & Xh;
            cout = (Ph & 0x80000000u) | ((Mh >> 1) & 0x40000000u) | tc;
            score += (int)((Ph >> lb) & 1) - (int)((Mh >> lb) & 1);     // meaningful in lane lw only
            if (score < best) { best = score; bestj = col; }
            const uint32_t Phs = __funnelshift_l(cin, Ph, 1);           // (Ph << 1) | (hin > 0)
            const uint32_t Mhs = __funnelshift_l(cin << 1, Mh, 1);      // (Mh << 1) | (hin < 0)
            Pv = Mhs | ~(Xv | Phs);
            Mv = Phs & Xv;
            *mp = make_uint2(Pv, Ph);
        }
    }
    *best_out = __shfl_sync(0xffffffffu, best, lw, WIDTH);
    *bestj_out = __shfl_sync(0xffffffffu, bestj, lw, WIDTH);
}

// Walk back from (qblk, jstar).  Columns are produced last to first into `rev`.
// Also reports the last run of M matches (counted from the end, oc_aligner.c:225-240).
//
// The walk reads the (Pv, Ph) scratch through a register window of 32 columns x 32 diagonals: at a refill lane l
// takes the two words of column j - l that hold rows i - l - 16 .. i - l + 15 and turns them into three masks
// indexed by diagonal (bit dd <-> diagonal (i - j) + dd - 16): "cell = upper + 1", "cell = left + 1", "cell
// exists".  Whatever path the walk takes inside the window, the state of the cell (i - k, j - k) is then one bit
// test in lane l0 + k: every iteration resolves a whole run of diagonal moves with three ballots, plus the gap
// move that ended it.
struct WalkInfo { int asize, found, acnt, qcnt, tcnt; };

__device__ __forceinline__ WalkInfo myers_walkback(const uint8_t* __restrict__ sq, const uint8_t* __restrict__ st, uint8_t* __restrict__ rev,
                                                   int qblk, int jstar, const uint2* __restrict__ mat, int lane, int M)
{
    const int nw = (qblk + 31) >> 5;
    int i = qblk, j = jstar, r = 0;
    WalkInfo W; W.asize = 0; W.found = 0; W.acnt = 0; W.qcnt = 0; W.tcnt = 0;
    int m = 0;
    int jw = -(1 << 28), D0 = 0;                    // the window was filled at column jw on diagonal D0
    uint32_t UP = 0, LEFT = 0, VALID = 0;
    while (i > 0 || j > 0) {
        if (i == 0 || j == 0) {                                 // along the matrix border
            const int op = (i == 0) ? OP_TONLY : OP_QONLY;
            if (i == 0) --j; else --i;
            if (lane == 0) rev[r] = (uint8_t)op;
            ++r;
            if (!W.found) {
                W.qcnt += op != OP_TONLY; W.tcnt += op != OP_QONLY; m = 0; ++W.acnt;
            }
            continue;
        }
        int l0 = jw - j, dd = (i - j) - D0 + 16;
        uint32_t bv = 0;
        bool have = (unsigned)l0 < 32u && (unsigned)dd < 32u;
        if (have) { bv = __ballot_sync(0xffffffffu, (VALID >> dd) & 1); have = (bv >> l0) & 1; }
        if (!have) {
            // refill, centred on the diagonal through (i, j): my column is j - lane, its rows i - lane - 16 .. i - lane + 15
            jw = j; D0 = i - j; l0 = 0; dd = 16;
            const int col = j - lane, rc = i - lane;
            int ww = (rc - 17) >> 5;                            // word of row rc - 16 (row x is bit (x - 1) & 31 of word (x - 1) >> 5)
            ww = ww < 0 ? 0 : (ww > nw - 1 ? nw - 1 : ww);
            uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
            if (col >= 1) {
                a = mat[(size_t)(col - 1 + ww) * nw + ww];
                if (ww + 1 < nw) b = mat[(size_t)(col + ww) * nw + ww + 1];
            }
            const int s0 = rc - 17 - 32 * ww;                   // bit of diagonal dd = 0 in the 64-bit value b:a
            if (s0 >= 0) { UP = __funnelshift_r(a.x, b.x, s0); LEFT = __funnelshift_r(a.y, b.y, s0); }     // s0 <= 31
            else if (s0 > -32) { UP = a.x << (-s0); LEFT = a.y << (-s0); }
            else { UP = 0; LEFT = 0; }
            const int first = 17 - rc;                          // rows below 1 do not exist
            VALID = col >= 1 ? (first <= 0 ? 0xffffffffu : (first < 32 ? 0xffffffffu << first : 0u)) : 0u;
            bv = __ballot_sync(0xffffffffu, (VALID >> dd) & 1);
        }
        const uint32_t bu = __ballot_sync(0xffffffffu, (UP >> dd) & 1);
        const uint32_t bl = __ballot_sync(0xffffffffu, (LEFT >> dd) & 1);
        const uint32_t dm = (bv & ~bu & ~bl) >> l0;             // lane l0 + k: the cell (i - k, j - k) is a diagonal move
        const int run = (dm == 0xffffffffu) ? 32 : __ffs(~dm) - 1;
        if (run > 0) {
            // a run of diagonal moves: all its columns at once
            const int k = lane - l0;
            bool mt = false;
            if (k >= 0 && k < run) {
                mt = sq[i - k - 1] == st[j - k - 1];
                rev[r + k] = (uint8_t)(mt ? OP_MATCH : OP_MISMATCH);
            }
            if (!W.found) {
                const uint32_t mm = __ballot_sync(0xffffffffu, mt) >> l0;
                for (int p = 0; p < run; ++p) {
                    ++W.qcnt; ++W.tcnt; ++W.acnt;
                    m = ((mm >> p) & 1) ? m + 1 : 0;
                    if (m == M) { W.found = 1; break; }
                }
            }
            i -= run; j -= run; r += run; l0 += run;
        }
        // the move that ended the run, when its cell is in the window (then it is an up or a left move)
        if (i > 0 && j > 0 && l0 < 32 && ((bv >> l0) & 1)) {
            int op;
            if ((bu >> l0) & 1) { op = OP_QONLY; --i; }
            else { op = OP_TONLY; --j; }
            if (lane == 0) rev[r] = (uint8_t)op;
            ++r;
            if (!W.found) {
                W.qcnt += op != OP_TONLY;
                W.tcnt += op != OP_QONLY;
                m = 0;
                ++W.acnt;
            }
        }
    }
    W.asize = r;
    return W;
}

// ---- the per-candidate state machine (oc_aligner.c:157-286 per side, :303-451 around it)

// Starts a side of the extension: columns go to ops[out_base + x * out_dir].
__device__ __forceinline__ void ext_begin_side(ExtCtx& C, int lane, int phase, int q_origin, int q_total, int t_origin, int t_total,
                                               int dir, long long out_base, int out_dir)
{
    if (lane == 0) {
        C.phase = phase; C.q_origin = q_origin; C.q_total = q_total; C.t_origin = t_origin; C.t_total = t_total; C.dir = dir;
        C.out_base = out_base; C.out_dir = out_dir; C.qidx = 0; C.tidx = 0; C.ncol = 0;
    }
    __syncwarp();
}

// Cuts the next block of the current side and unpacks its bases.  False: nothing left on one of the reads.
__device__ __forceinline__ bool ext_next_block(ExtCtx& C, HalfScratch& H, const ExtParams& P, int lane)
{
    const int bs = P.block_size;
    const int qidx = C.qidx, tidx = C.tidx;
    const int qleft = C.q_total - qidx, tleft = C.t_total - tidx;
    int qblk, tblk, last;
    if (qleft < bs + 100 || tleft < bs + 100) {             // oc_aligner.c:127-137
        qblk = (int)__dmul_rn((double)tleft, 1.3); qblk = qblk < qleft ? qblk : qleft;
        tblk = (int)__dmul_rn((double)qleft, 1.3); tblk = tblk < tleft ? tblk : tleft;
        last = 1;
    } else { qblk = tblk = bs; last = 0; }
    if (qblk == 0 || tblk == 0) return false;
    SeqView qv; qv.pac = P.q_pac; qv.off = C.qoff_; qv.len = C.qlen; qv.strand = C.qstrand;
    SeqView tv; tv.pac = P.pac; tv.off = C.toff_; tv.len = C.tlen; tv.strand = 0;
    const int dir = C.dir, qo = C.q_origin + dir * qidx, to = C.t_origin + dir * tidx;
    __syncwarp();
    for (int x = lane; x < qblk; x += 32) H.q[x] = (uint8_t)seq_base(qv, qo + dir * x);
    for (int x = lane; x < tblk; x += 32) H.t[x] = (uint8_t)seq_base(tv, to + dir * x);
    if (lane == 0) { C.qblk = qblk; C.tblk = tblk; C.last = last; }
    __syncwarp();
    return true;
}

// End of the right side: assemble the result (oc_aligner.c:385-451).
__device__ __forceinline__ void ext_finish(ExtCtx& C, const ExtParams& P, int lane)
{
    uint8_t* ops = P.ops + C.ops_base;
    const int R = C.ncol, lcap = C.lcap, L = C.L, left_keep = C.left_keep, acnt = C.acnt;
    int QS = C.QS, TS = C.TS;
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
        P.results[C.ti] = A;
        C.active = 0;
    }
    __syncwarp();
}

// The current side has ended.  Left side: re-anchor at the first run of 8 matches counted from the seed
// (oc_aligner.c:329-367) and start the right side; right side: finish the candidate.  On return the candidate
// either has a block ready for the forward pass or is no longer active.
__device__ __forceinline__ void ext_end_side(ExtCtx& C, HalfScratch& H, const ExtParams& P, int lane)
{
    __syncwarp();
    if (C.phase == 0) {
        const uint8_t* ops = P.ops + C.ops_base;
        const int L = C.ncol, lcap = C.lcap;
        int acnt = 0, qcnt = 0, tcnt = 0, m = 0;
        for (int x = 0; x < L; ++x) {
            const int op = ops[lcap - 1 - x];
            qcnt += op != OP_TONLY; tcnt += op != OP_QONLY;
            m = op == OP_MATCH ? m + 1 : 0;
            ++acnt;
            if (m == 8) break;
        }
        int QS = C.QS, TS = C.TS, left_keep = 0;
        if (m == 8) { QS -= qcnt; TS -= tcnt; left_keep = L - acnt; }
        else acnt = 0;
        __syncwarp();
        if (lane == 0) { C.QS = QS; C.TS = TS; C.L = L; C.left_keep = left_keep; C.acnt = acnt; }
        // final alignment = ops[lcap - L, lcap - acnt) ++ right extension written from lcap - acnt
        ext_begin_side(C, lane, 1, QS, C.qlen - QS, TS, C.tlen - TS, +1, (long long)(left_keep ? lcap - acnt : lcap), +1);
        if (ext_next_block(C, H, P, lane)) return;
    }
    ext_finish(C, P, lane);
}

// Fetches the next candidate for this slot and prepares its first block.
__device__ __forceinline__ void ext_start_task(ExtCtx& C, HalfScratch& H, const ExtParams& P, uint32_t ti, int lane)
{
    const ExtTask T = P.tasks[ti];
    __syncwarp();
    if (lane == 0) {
        C.active = 1; C.ti = ti;
        C.qoff_ = P.q_off[T.qid]; C.qlen = (int)P.q_len[T.qid]; C.qstrand = T.qdir;
        C.toff_ = P.seq_off[T.sid]; C.tlen = (int)P.seq_len[T.sid];
        C.ops_base = T.ops_base; C.lcap = (int)T.lcap; C.QS = T.qoff; C.TS = T.soff;
        C.L = 0; C.left_keep = 0; C.acnt = 0;
    }
    // left of the seed, read right to left (oc_aligner.c:313-327); column x lives at lcap - 1 - x
    ext_begin_side(C, lane, 0, T.qoff - 1, T.qoff, T.soff - 1, T.soff, -1, (long long)(T.lcap ? T.lcap - 1 : 0), -1);
    if (!ext_next_block(C, H, P, lane)) ext_end_side(C, H, P, lane);
}

// After the forward pass of a block: walk back, keep the part up to the last run of matches (oc_aligner.c:213-262),
// then cut the next block or end the side.
__device__ __forceinline__ void ext_post_block(ExtCtx& C, HalfScratch& H, uint8_t* rev, const ExtParams& P, const uint2* mat, int lane)
{
    const int qblk = C.qblk, tblk = C.tblk, last = C.last, best = C.best, bestj = C.bestj;
    const int mn = qblk < tblk ? qblk : tblk;
    const int k0 = (int)__dmul_rn(__dmul_rn((double)mn, P.error), 1.1);         // edlib_ex.c:751
    const bool failed = best > k0;
    int done = last;
    if (failed && qblk > 30 && tblk > 30) done = 1;                             // oc_aligner.c:220 (qfae = tfae = 0)
    const int M = done ? P.tail_match_len : 8;
    WalkInfo W; W.asize = 0; W.found = 0; W.acnt = 0; W.qcnt = 0; W.tcnt = 0;
    if (!failed) W = myers_walkback(H.q, H.t, rev, qblk, bestj + 1, mat, lane, M);
    __syncwarp();
    uint8_t* out = P.ops + C.ops_base + C.out_base;
    const int out_dir = C.out_dir, ncol = C.ncol;
    const int kcut = W.asize - W.acnt;
    if (!W.found || kcut < 1) {
        // exact-match prefix only, then stop (:243-253)
        int n = 0;
        for (int base = 0; base < mn; base += 32) {
            const int x = base + lane;
            const bool eq = x < mn && H.q[x] == H.t[x];
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            if (bal == 0xffffffffu) { n += 32; continue; }
            n += __ffs(~bal) - 1;
            break;
        }
        for (int x = lane; x < n; x += 32) out[(long long)(ncol + x) * out_dir] = OP_MATCH;
        __syncwarp();
        if (lane == 0) C.ncol = ncol + n;
        ext_end_side(C, H, P, lane);
        return;
    }
    const int keep = kcut + (done ? M : 0);
    for (int x = lane; x < keep; x += 32) out[(long long)(ncol + x) * out_dir] = rev[W.asize - 1 - x];
    __syncwarp();
    if (lane == 0) {
        C.qidx += qblk - W.qcnt;            // qfae = qblk
        C.tidx += (bestj + 1) - W.tcnt;     // tfae = j*
        C.ncol = ncol + keep;
    }
    __syncwarp();
    if (done || !ext_next_block(C, H, P, lane)) ext_end_side(C, H, P, lane);
}

__global__ void __launch_bounds__(EXT_WARPS * 32, 4)
k_extend(ExtParams P)
{
    extern __shared__ __align__(16) uint8_t ext_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& S = reinterpret_cast<WarpScratch*>(ext_smem)[wid];
    const size_t slot = (size_t)blockIdx.x * EXT_WARPS + wid;
    uint2* mat0 = P.mat + slot * 2 * EXT_MAT_STRIDE;
    uint2* mat1 = mat0 + EXT_MAT_STRIDE;
    unsigned long long nb = 0, nc = 0, nwd = 0;
    if (lane == 0) { S.ctx[0].active = 0; S.ctx[1].active = 0; }
    __syncwarp();
    bool drained = false;
    for (;;) {
        // ---- refill the idle slots from the queue
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            while (!S.ctx[h].active && !drained) {
                uint32_t ti = 0;
                if (lane == 0) ti = atomicAdd(P.next_task, 1u);
                ti = __shfl_sync(0xffffffffu, ti, 0);
                if (ti >= P.n_tasks) { drained = true; break; }
                ext_start_task(S.ctx[h], S.h[h], P, ti, lane);          // may finish the candidate right away
            }
        }
        const int a0 = S.ctx[0].active, a1 = S.ctx[1].active;
        if (!a0 && !a1) break;
        // ---- forward passes: side by side in the two half-warps when both blocks fit 16 lanes
        const int q0 = a0 ? S.ctx[0].qblk : 0, t0 = a0 ? S.ctx[0].tblk : 0;
        const int q1 = a1 ? S.ctx[1].qblk : 0, t1 = a1 ? S.ctx[1].tblk : 0;
        int best, bestj;
        if (q0 <= 512 && q1 <= 512) {
            const int hh = lane >> 4;
            myers_forward<16>(S.h[hh].q, S.h[hh].t, hh ? q1 : q0, hh ? t1 : t0, hh ? mat1 : mat0, lane & 15, &best, &bestj);
            __syncwarp();
            if (lane == 0 && a0) { S.ctx[0].best = best; S.ctx[0].bestj = bestj; }
            if (lane == 16 && a1) { S.ctx[1].best = best; S.ctx[1].bestj = bestj; }
        } else {
            if (a0) { myers_forward<32>(S.h[0].q, S.h[0].t, q0, t0, mat0, lane, &best, &bestj); if (lane == 0) { S.ctx[0].best = best; S.ctx[0].bestj = bestj; } }
            __syncwarp();
            if (a1) { myers_forward<32>(S.h[1].q, S.h[1].t, q1, t1, mat1, lane, &best, &bestj); if (lane == 0) { S.ctx[1].best = best; S.ctx[1].bestj = bestj; } }
        }
        if (lane == 0) {
            if (a0) { nb += 1; nc += (unsigned long long)t0; nwd += (unsigned long long)t0 * ((q0 + 31) >> 5); }
            if (a1) { nb += 1; nc += (unsigned long long)t1; nwd += (unsigned long long)t1 * ((q1 + 31) >> 5); }
        }
        __syncwarp();
        // ---- walk-backs and bookkeeping, one candidate after the other, whole warp
        if (a0) ext_post_block(S.ctx[0], S.h[0], S.rev, P, mat0, lane);
        if (a1) ext_post_block(S.ctx[1], S.h[1], S.rev, P, mat1, lane);
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
    static const int kCtas = getenv("NB_EXT_CTAS") ? atoi(getenv("NB_EXT_CTAS")) : 4;
    int ctas_per_sm = kCtas;                // x 8 warps per SM
    if (ctx->ext_ctas > 0 && ctx->ext_ctas < ctas_per_sm) ctas_per_sm = ctx->ext_ctas;
    const unsigned grid = NB_SMS * ctas_per_sm;
    const size_t slots = (size_t)NB_SMS * kCtas * EXT_WARPS * 2;     // two candidates in flight per warp (sized for the full grid)
    if (!ctx->ext_mat || ctx->ext_mat_slots < slots) {
        if (ctx->ext_mat) cudaFree(ctx->ext_mat);
        ctx->ext_mat = NULL;
        NB_CUDA(ctx, cudaMalloc(&ctx->ext_mat, sizeof(uint2) * EXT_MAT_STRIDE * slots));
        ctx->ext_mat_slots = slots;
    }
    uint2* s_mat = (uint2*)ctx->ext_mat;
    NbDevBuf B(ctx);
    uint32_t* d_next = NULL;
    unsigned long long* d_cnt = NULL;
    { int rc; if ((rc = B.get(&d_next, 1))) return rc; if ((rc = B.get(&d_cnt, 4))) return rc; }
    NB_CUDA(ctx, cudaMemsetAsync(d_next, 0, sizeof(uint32_t), st));
    NB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 4, st));
    ExtParams P;
    P.pac = templates->d_pac; P.seq_off = templates->d_off; P.seq_len = templates->d_len;
    P.q_pac = queries->d_pac; P.q_off = queries->d_off; P.q_len = queries->d_len;
    P.tasks = d_tasks; P.n_tasks = n_tasks; P.results = d_results; P.ops = d_ops; P.mat = s_mat;
    P.next_task = d_next; P.counters = d_cnt;
    P.error = error; P.min_align_size = min_align_size; P.tail_match_len = tail_match_len; P.block_size = 512;
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
    if (cnt[2] > ctx->cc.big_ext_word_updates) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
        ctx->cc.big_ext_tasks = n_tasks; ctx->cc.big_ext_block_aligns = cnt[0]; ctx->cc.big_ext_word_updates = cnt[2];
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
