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

struct WarpScratch {
    uint8_t q[EXT_MAX_BLK];
    uint8_t t[EXT_MAX_BLK + 64];        // padded: the forward pass reads one column ahead
    uint8_t rev[2 * EXT_MAX_BLK];
    uint32_t peq[4][32];                // match masks: peq[symbol][lane] covers query rows [32 lane, 32 lane + 32)
};

// Forward pass over one block.  Returns best prefix distance and its column.
__device__ __forceinline__ void myers_forward(WarpScratch& S, int qblk, int tblk, uint2* __restrict__ mat, int lane,
                                              int* best_out, int* bestj_out)
{
    const int nw = (qblk + 31) >> 5;
    // per-lane match masks of the 4 symbols
    uint32_t pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0;
    if (lane < nw) {
        const int r0 = lane << 5;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            if (r0 + r < qblk) {
                const int c = S.q[r0 + r];
                const uint32_t bit = 1u << r;
                pe0 |= c == 0 ? bit : 0; pe1 |= c == 1 ? bit : 0; pe2 |= c == 2 ? bit : 0; pe3 |= c == 3 ? bit : 0;
            }
        }
    }
    S.peq[0][lane] = pe0; S.peq[1][lane] = pe1; S.peq[2][lane] = pe2; S.peq[3][lane] = pe3;
    __syncwarp();
    uint32_t Pv = 0xffffffffu, Mv = 0;
    int hout = 0;
    const int lw = (qblk - 1) >> 5, lb = (qblk - 1) & 31;
    int score = qblk, best = 0x7fffffff, bestj = -1;
    const int nsteps = tblk + nw - 1;
    const bool mine = lane < nw;
    uint2* mp = mat + lane;                                 // skewed [step][lane] layout, advanced by nw per step
    // match mask of the column this lane handles at step 0, fetched one step ahead of its use
    int colp = -lane;
    uint32_t Eq_next = S.peq[S.t[colp >= 0 ? colp : 0] & 3][lane];
    for (int s = 0; s < nsteps; ++s, mp += nw) {
        int hin = __shfl_up_sync(0xffffffffu, hout, 1);
        if (lane == 0) hin = 1;
        const int col = s - lane;
        uint32_t Eq = Eq_next;
        {
            const int cn = col + 1;                          // next step's column (padding of S.t makes the read safe)
            Eq_next = S.peq[S.t[cn >= 0 ? cn : 0] & 3][lane];
        }
        if (mine && (unsigned)col < (unsigned)tblk) {
            const uint32_t neg = hin < 0, pos = hin > 0;
            const uint32_t Xv = Eq | Mv;
            Eq |= neg;
            const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            const uint32_t Ph = Mv | ~(Xh | Pv);
            const uint32_t Mh = Pv & Xh;
            hout = (int)(Ph >> 31) - (int)(Mh >> 31);
            if (lane == lw) {
                score += (int)((Ph >> lb) & 1) - (int)((Mh >> lb) & 1);
                if (score < best) { best = score; bestj = col; }
            }
            const uint32_t Phs = (Ph << 1) | pos, Mhs = (Mh << 1) | neg;
            Pv = Mhs | ~(Xv | Phs);
            Mv = Phs & Xv;
            *mp = make_uint2(Pv, Ph);
        }
    }
    *best_out = __shfl_sync(0xffffffffu, best, lw);
    *bestj_out = __shfl_sync(0xffffffffu, bestj, lw);
}

// Walk back from (qblk, jstar).  Columns are produced last to first into S.rev.
// Also reports the last run of M matches (counted from the end, oc_aligner.c:225-240).
struct WalkInfo { int asize, found, acnt, qcnt, tcnt; };

__device__ __forceinline__ WalkInfo myers_walkback(WarpScratch& S, int qblk, int jstar, const uint2* __restrict__ mat, int lane,
                                                   int M)
{
    const int nw = (qblk + 31) >> 5;
    int i = qblk, j = jstar, r = 0;
    WalkInfo W; W.asize = 0; W.found = 0; W.acnt = 0; W.qcnt = 0; W.tcnt = 0;
    int m = 0;
    // Register window over 32 columns: lane l caches words wa, wa+1 (rows 32 wa + 1 .. 32 wa + 64) of column
    // jw - l, centred on the diagonal through the cell the window was filled at.  Whole diagonal runs are
    // then resolved in one step: lane l tests the cell (i - (l - l0), j - (l - l0)).
    int jw = -1000, wa = 0;
    uint2 pa = make_uint2(0, 0), pb = make_uint2(0, 0);
    while (i > 0 || j > 0) {
        if (i == 0 || j == 0) {                                 // along the matrix border
            const int op = (i == 0) ? OP_TONLY : OP_QONLY;
            if (i == 0) --j; else --i;
            if (lane == 0) S.rev[r] = (uint8_t)op;
            ++r;
            if (!W.found) {
                W.qcnt += op != OP_TONLY; W.tcnt += op != OP_QONLY; m = 0; ++W.acnt;
            }
            continue;
        }
        int l0 = jw - j;
        bool known = false, up = false, left = false;
        int k = 0, ri = 0, cj = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (l0 < 0 || l0 >= 32 || attempt == 1) {
                jw = j; l0 = 0;
                const int col = j - lane, row = i - lane;
                int ww = ((row - 1) - 16) >> 5;
                ww = ww < 0 ? 0 : (ww > nw - 1 ? nw - 1 : ww);
                wa = ww;
                if (col >= 1) {
                    pa = mat[(size_t)(col - 1 + ww) * nw + ww];
                    pb = (ww + 1 < nw) ? mat[(size_t)(col + ww) * nw + ww + 1] : make_uint2(0, 0);
                }
            }
            // my cell on the current diagonal
            k = lane - l0; ri = i - k; cj = j - k;
            known = false; up = false; left = false;
            if (k >= 0 && ri >= 1 && cj >= 1) {
                const int w = (ri - 1) >> 5, b = (ri - 1) & 31, d = w - wa;
                if (d == 0) { known = true; up = (pa.x >> b) & 1; left = (pa.y >> b) & 1; }
                else if (d == 1) { known = true; up = (pb.x >> b) & 1; left = (pb.y >> b) & 1; }
            }
            if ((__ballot_sync(0xffffffffu, known) >> l0) & 1) break;       // the current cell is in the window
        }
        const uint32_t dm = __ballot_sync(0xffffffffu, known && !up && !left) >> l0;
        const int run = (dm == 0xffffffffu) ? 32 : __ffs(~dm) - 1;
        if (run > 0) {
            // a run of diagonal moves: all its columns at once
            bool mt = false;
            if (k >= 0 && k < run) {
                mt = S.q[ri - 1] == S.t[cj - 1];
                S.rev[r + k] = (uint8_t)(mt ? OP_MATCH : OP_MISMATCH);
            }
            if (!W.found) {
                const uint32_t mm = __ballot_sync(0xffffffffu, mt) >> l0;
                for (int p = 0; p < run; ++p) {
                    ++W.qcnt; ++W.tcnt; ++W.acnt;
                    m = ((mm >> p) & 1) ? m + 1 : 0;
                    if (m == M) { W.found = 1; break; }
                }
            }
            i -= run; j -= run; r += run;
            continue;
        }
        // a single up / left move at (i, j)
        const uint32_t km = __ballot_sync(0xffffffffu, known);
        int isup, isleft;
        if ((km >> l0) & 1) {
            isup = __shfl_sync(0xffffffffu, (int)up, l0);
            isleft = __shfl_sync(0xffffffffu, (int)left, l0);
        } else {
            const int w = (i - 1) >> 5, b = (i - 1) & 31;
            const uint2 cur = mat[(size_t)(j - 1 + w) * nw + w];       // outside the window: read it directly
            isup = (cur.x >> b) & 1; isleft = (cur.y >> b) & 1;
        }
        int op;
        if (isup) { op = OP_QONLY; --i; }
        else if (isleft) { op = OP_TONLY; --j; }
        else { op = (S.q[i - 1] == S.t[j - 1]) ? OP_MATCH : OP_MISMATCH; --i; --j; }    // diagonal seen through the direct read
        if (lane == 0) S.rev[r] = (uint8_t)op;
        ++r;
        if (!W.found) {
            W.qcnt += op != OP_TONLY;
            W.tcnt += op != OP_QONLY;
            m = (op == OP_MATCH) ? m + 1 : 0;
            ++W.acnt;
            if (m == M) W.found = 1;
        }
    }
    W.asize = r;
    return W;
}

// oc_aligner.c:157-286.  Extends from an origin in one direction; column x of the
// extension is stored at out[x * out_dir].  Returns the number of columns.
__device__ int extend_side(WarpScratch& S, const ExtParams& P, uint2* mat, int lane, const SeqView& qv, int q_origin,
                           int q_total, const SeqView& tv, int t_origin, int t_total, int dir, uint8_t* out, int out_dir,
                           unsigned long long* n_blocks, unsigned long long* n_cols, unsigned long long* n_words)
{
    int qidx = 0, tidx = 0, ncol = 0;
    const int bs = P.block_size;
    for (;;) {
        const int qleft = q_total - qidx, tleft = t_total - tidx;
        int qblk, tblk, last;
        if (qleft < bs + 100 || tleft < bs + 100) {             // :127-137
            qblk = (int)__dmul_rn((double)tleft, 1.3); qblk = qblk < qleft ? qblk : qleft;
            tblk = (int)__dmul_rn((double)qleft, 1.3); tblk = tblk < tleft ? tblk : tleft;
            last = 1;
        } else { qblk = tblk = bs; last = 0; }
        if (qblk == 0 || tblk == 0) break;
        __syncwarp();
        for (int x = lane; x < qblk; x += 32) S.q[x] = (uint8_t)seq_base(qv, q_origin + dir * (qidx + x));
        for (int x = lane; x < tblk; x += 32) S.t[x] = (uint8_t)seq_base(tv, t_origin + dir * (tidx + x));
        __syncwarp();
        int best, bestj;
        myers_forward(S, qblk, tblk, mat, lane, &best, &bestj);
        __syncwarp();
        if (lane == 0) { *n_blocks += 1; *n_cols += (unsigned long long)tblk; *n_words += (unsigned long long)tblk * ((qblk + 31) >> 5); }
        const int mn = qblk < tblk ? qblk : tblk;
        const int k0 = (int)__dmul_rn(__dmul_rn((double)mn, P.error), 1.1);         // edlib_ex.c:751
        const bool failed = best > k0;
        int done = last;
        if (failed && qblk > 30 && tblk > 30) done = 1;                             // oc_aligner.c:220 (qfae = tfae = 0)
        const int M = done ? P.tail_match_len : 8;
        WalkInfo W; W.asize = 0; W.found = 0; W.acnt = 0; W.qcnt = 0; W.tcnt = 0;
        if (!failed) W = myers_walkback(S, qblk, bestj + 1, mat, lane, M);
        __syncwarp();
        const int kcut = W.asize - W.acnt;
        int keep;
        if (!W.found || kcut < 1) {
            // exact-match prefix only, then stop (:243-253)
            int n = 0;
            const int lim = mn;
            for (int base = 0; base < lim; base += 32) {
                const int x = base + lane;
                const bool eq = x < lim && S.q[x] == S.t[x];
                const uint32_t bal = __ballot_sync(0xffffffffu, eq);
                if (bal == 0xffffffffu) { n += 32; continue; }
                n += __ffs(~bal) - 1;
                break;
            }
            for (int x = lane; x < n; x += 32) out[(ncol + x) * out_dir] = OP_MATCH;
            ncol += n;
            break;
        }
        keep = kcut + (done ? M : 0);
        qidx += qblk - W.qcnt;              // qfae = qblk
        tidx += (bestj + 1) - W.tcnt;       // tfae = j*
        for (int x = lane; x < keep; x += 32) out[(ncol + x) * out_dir] = S.rev[W.asize - 1 - x];
        ncol += keep;
        if (done) break;
    }
    __syncwarp();
    return ncol;
}

__global__ void __launch_bounds__(EXT_WARPS * 32, 4)
k_extend(ExtParams P)
{
    __shared__ WarpScratch s_scratch[EXT_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& S = s_scratch[wid];
    const size_t slot = (size_t)blockIdx.x * EXT_WARPS + wid;
    uint2* mat = P.mat + slot * EXT_MAT_STRIDE;
    unsigned long long nb = 0, nc = 0, nwd = 0;
    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(P.next_task, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= P.n_tasks) break;
        const ExtTask T = P.tasks[ti];
        SeqView qv; qv.pac = P.q_pac; qv.off = P.q_off[T.qid]; qv.len = (int)P.q_len[T.qid]; qv.strand = T.qdir;
        SeqView tv; tv.pac = P.pac; tv.off = P.seq_off[T.sid]; tv.len = (int)P.seq_len[T.sid]; tv.strand = 0;
        uint8_t* ops = P.ops + T.ops_base;
        int QS = T.qoff, TS = T.soff;
        // ---- left of the seed, read right to left (oc_aligner.c:313-327); column x lives at lcap-1-x
        const int L = extend_side(S, P, mat, lane, qv, QS - 1, QS, tv, TS - 1, TS, -1, ops + (T.lcap ? T.lcap - 1 : 0), -1, &nb, &nc, &nwd);
        // first run of 8 matches counted from the seed (:329-345)
        int acnt = 0, qcnt = 0, tcnt = 0, m = 0;
        for (int x = 0; x < L; ++x) {
            const int op = ops[(int)T.lcap - 1 - x];
            qcnt += op != OP_TONLY; tcnt += op != OP_QONLY;
            m = op == OP_MATCH ? m + 1 : 0;
            ++acnt;
            if (m == 8) break;
        }
        int left_keep = 0;
        if (m == 8) { QS -= qcnt; TS -= tcnt; left_keep = L - acnt; }
        else acnt = 0;
        // final alignment = ops[lcap - L, lcap - acnt) ++ right extension written from lcap - acnt
        uint8_t* rbase = ops + (left_keep ? (int)T.lcap - acnt : (int)T.lcap);
        const int R = extend_side(S, P, mat, lane, qv, QS, qv.len - QS, tv, TS, tv.len - TS, +1, rbase, +1, &nb, &nc, &nwd);
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
        const uint8_t* fin = left_keep ? ops + ((int)T.lcap - L) : rbase + rstart;
        const int n_left = left_keep, n_all = left_keep + right_keep;
        int lq = 0, lt = 0, rq = 0, rt = 0, nm = 0;
        for (int x = lane; x < n_all; x += 32) {
            const int op = fin[x];
            const int cq = op != OP_TONLY, ct = op != OP_QONLY;
            if (x < n_left) { lq += cq; lt += ct; } else { rq += cq; rt += ct; }
            nm += op == OP_MATCH;
        }
        for (int d = 16; d > 0; d >>= 1) {
            lq += __shfl_xor_sync(0xffffffffu, lq, d); lt += __shfl_xor_sync(0xffffffffu, lt, d);
            rq += __shfl_xor_sync(0xffffffffu, rq, d); rt += __shfl_xor_sync(0xffffffffu, rt, d);
            nm += __shfl_xor_sync(0xffffffffu, nm, d);
        }
        if (lane == 0) {
            nb_alignment A;
            A.qoff = QS - lq; A.qend = QS + rq; A.toff = TS - lt; A.tend = TS + rt;
            A.aln_size = n_all; A.n_match = nm;
            A.ok = n_all >= P.min_align_size;
            A.ops_offset = (uint64_t)(fin - P.ops);
            P.results[ti] = A;
        }
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
    static const int kCtas = getenv("NB_EXT_CTAS") ? atoi(getenv("NB_EXT_CTAS")) : 4;
    int ctas_per_sm = kCtas;                // x 8 warps per SM
    const unsigned grid = NB_SMS * ctas_per_sm;
    const size_t slots = (size_t)grid * EXT_WARPS;
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
    k_extend<<<grid, EXT_WARPS * 32, 0, st>>>(P);
    NB_LAUNCH_CHECK(ctx);
    unsigned long long cnt[4];
    NB_CUDA(ctx, cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->cc.n_extensions += n_tasks;
    ctx->cc.n_block_aligns += cnt[0];
    ctx->cc.n_dp_columns += cnt[1];
    ctx->cc.n_dp_word_updates += cnt[2];
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
