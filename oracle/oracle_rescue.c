/*
 * oracle_rescue.c -- CPU restatement of the "rescue" branch of cns_extension (oc2cns -r 1,
 * consensus/consensus_aux.c:169-201): when the block-wise extension fails, or leaves more than 200 bases of
 * the candidate's chain unaligned, the pair is realigned with
 *   1. Myers' O(nd) local alignment with path-tip trimming (gapped_align/oc_daligner.c:34-79 over
 *      gapped_align/align.c:382-1041 forward_wave, :1043-1750 reverse_wave, :1754-1945 Local_Alignment), then
 *   2. a global edit-distance alignment of the two ranges it found, bounded by its difference count
 *      (edlib/edlib_wrapper.c:111-255 over the vendored edlib NW + traceback / Hirschberg, edlib/edlib.cpp:887-1361),
 *      trimmed to its first and last run of kMatchSize matches.
 *
 * TEST INFRASTRUCTURE ONLY (see necat_oracle.h).  Written from the algorithms, with different data structures:
 * the waves keep no trace points (only the end points and the difference count are used downstream), and the
 * global alignment uses plain integer DP matrices -- the choices that decide WHICH co-optimal alignment comes out
 * (Hirschberg's first split row, the traceback's up / left / diagonal preference, the 1 MiB switch between the two)
 * are the reference's.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "necat_oracle.h"
#include "oracle_rescue.h"

/* ------------------------------------------------------------------ 1. O(nd) waves */

#define W_TRIM_LEN 15            /* align.c:200 */
#define W_PATH_LEN 60            /* align.c:204 */
#define W_PATH_TOP (1ull << W_PATH_LEN)
#define W_PATH_INT (W_PATH_TOP - 1)
#define W_TRIM_MASK ((1 << W_TRIM_LEN) - 1)
#define W_TRIM_MLAG 200          /* align.c:211 */
#define W_WAVE_LAG 30            /* align.c:212 */

struct ORescue {
    int      path_ave;                       /* align.c:282 */
    int16_t  tip_score[W_TRIM_MASK + 1];     /* score of the last 15 edits, match = +mscore, difference = -dscore */
    int16_t  tip_table[W_TRIM_MASK + 1];     /* that score minus the best score of any proper prefix (>= 0: every suffix non-negative) */
    double   error;
    /* wave vectors, indexed by diagonal - vmin */
    int*      V;
    int*      M;
    uint64_t* T;
    int       vcap, vmin;
    /* scratch of the global alignment */
    int*      dp;
    size_t    dp_cap;
    uint8_t*  ops;
    size_t    ops_cap;
};

ORescue* o_rescue_new(double error)
{
    ORescue* r = (ORescue*)calloc(1, sizeof(ORescue));
    r->error = error;
    /* New_Align_Spec(1 - error, 100, {.25,.25,.25,.25}), align.c:257-301: base bias 50/50 -> Bias_Factor[9] = 1 */
    const double ave_corr = 1.0 - error;
    const double freq_match = 0.5;
    int bias = (int)((freq_match + .025) * 20. - 1.);
    static const double bias_factor[10] = { .690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000 };
    r->path_ave = (int)(W_PATH_LEN * (1. - bias_factor[bias] * (1. - ave_corr)));
    const int mscore = (int)(1000 * bias_factor[bias] * (1. - ave_corr));
    const int dscore = 1000 - mscore;
    for (int p = 0; p <= W_TRIM_MASK; ++p) {            /* set_table, align.c:242-253: first edit in the top bit */
        int s = 0, mx = 0;
        for (int i = W_TRIM_LEN - 1; i >= 0; --i) {
            if (s > mx) mx = s;
            s += ((p >> i) & 1) ? mscore : -dscore;
        }
        r->tip_table[p] = (int16_t)(s - mx);
        r->tip_score[p] = (int16_t)s;
    }
    return r;
}

void o_rescue_free(ORescue* r)
{
    if (!r) return;
    free(r->V); free(r->M); free(r->T); free(r->dp); free(r->ops);
    free(r);
}

/* the vectors cover diagonals [lo, hi] (plus one guard cell on each side) */
static void wave_reserve(ORescue* r, int lo, int hi)
{
    if (r->V && lo - 1 >= r->vmin && hi + 1 < r->vmin + r->vcap) return;
    const int span = hi - lo + 3;
    int ncap = r->vcap ? r->vcap : 1024;
    while (ncap < span * 2) ncap *= 2;
    const int nmin = lo - 1 - (ncap - span) / 2;
    int* V = (int*)malloc(sizeof(int) * (size_t)ncap);
    int* M = (int*)malloc(sizeof(int) * (size_t)ncap);
    uint64_t* T = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)ncap);
    if (r->V) {
        /* keep what is live: the old range that overlaps the new window */
        for (int k = r->vmin; k < r->vmin + r->vcap; ++k) {
            if (k < nmin || k >= nmin + ncap) continue;
            V[k - nmin] = r->V[k - r->vmin]; M[k - nmin] = r->M[k - r->vmin]; T[k - nmin] = r->T[k - r->vmin];
        }
        free(r->V); free(r->M); free(r->T);
    }
    r->V = V; r->M = M; r->T = T; r->vcap = ncap; r->vmin = nmin;
}

typedef struct { int apos, bpos, diffs; } WaveTip;

/*
 * One direction of Local_Alignment from the point on diagonal k0 with anti-diagonal `anti` (a position + b position).
 * dir = +1: forward_wave (align.c:382-1041), dir = -1: reverse_wave (:1043-1750).  a and b carry the sentinel 4 at
 * index -1 and at index len.  Furthest-reaching points are kept as c = 2 y + k (y = b position, k = diagonal), with
 * the match / difference history of the last 60 edits (T) and its number of matches (M); the tip reported is the last
 * furthest point whose two 15-edit windows are "prefix positive" (TABLE / SCORE), or the best point that ran into an
 * end of a sequence.
 */
static WaveTip wave(ORescue* r, const uint8_t* a, const uint8_t* b, int k0, int anti, int dir)
{
    /* reverse: position y means "b[y - 1] is the next base to compare" */
    const uint8_t* as = dir > 0 ? a : a - 1;
    const uint8_t* bs = dir > 0 ? b : b - 1;
    const int INV = dir > 0 ? -1 : INT_MAX;                     /* "no point on this diagonal" */
#define BETTER(x, y) (dir > 0 ? (x) > (y) : (x) < (y))          /* x is further along than y */
    int low = k0, hgh = k0, dif = 0;
    wave_reserve(r, low - 1, hgh + 1);
    int more = 1;
    int aclip = dir > 0 ? INT_MAX : -INT_MAX, bclip = dir > 0 ? -INT_MAX : INT_MAX;
    int besta = anti, trima = anti, lasta = anti, morea = anti;
    int besty = (anti - hgh) >> 1, trimy = besty, morey = besty;
    int trimd = 0, mored = 0, morem = -1;
#define VV(k) r->V[(k) - r->vmin]
#define MM(k) r->M[(k) - r->vmin]
#define TT(k) r->T[(k) - r->vmin]
    /* slide along diagonal k from b position y; returns the new y; records sequence ends */
#define SLIDE(k, y, bvec, mcnt, track)                                                           \
    for (;;) {                                                                                   \
        const int cb__ = bs[(y)];                                                                \
        if (cb__ == 4) { more = 0; if (dir > 0 ? bclip < (k) : bclip > (k)) bclip = (k); break; } \
        const int ca__ = as[(y) + (k)];                                                          \
        if (cb__ != ca__) { if (ca__ == 4) { more = 0; aclip = (k); } break; }                   \
        (y) += dir;                                                                              \
        if (track) { if (((bvec) & W_PATH_TOP) == 0) (mcnt) += 1; (bvec) = ((bvec) << 1) | 1; }  \
    }
    /* ---- wave 0: the start diagonal only (Local_Alignment is called with low == hgh, oc_daligner.c:62-69) */
    {
        const int k = k0;
        int y = (anti - k) >> 1;
        uint64_t bv = 0; int mc = 0;
        SLIDE(k, y, bv, mc, 0)
        const int c = (y << 1) + k;
        if (BETTER(c, besta)) { besta = trima = lasta = c; besty = trimy = y; }
        VV(k) = c; TT(k) = W_PATH_INT; MM(k) = W_PATH_LEN;
    }
    /* a diagonal that ran into the end of a (aclip) or b (bclip) leaves the band; the best such point is remembered */
#define CLIPS(with_dif)                                                                          \
    if (more == 0) {                                                                             \
        if (bs[besty] != 4 && as[besta - besty] != 4) more = 1;                                  \
        if (dir > 0) {                                                                           \
            if (hgh >= aclip) { hgh = aclip - 1; if (morem <= MM(aclip)) { morem = MM(aclip); morea = VV(aclip); morey = (morea - aclip) / 2; if (with_dif) mored = dif; } } \
            if (low <= bclip) { low = bclip + 1; if (morem <= MM(bclip)) { morem = MM(bclip); morea = VV(bclip); morey = (morea - bclip) / 2; if (with_dif) mored = dif; } } \
            aclip = INT_MAX; bclip = -INT_MAX;                                                   \
        } else {                                                                                 \
            if (low <= aclip) { low = aclip + 1; if (morem <= MM(aclip)) { morem = MM(aclip); morea = VV(aclip); morey = (morea - aclip) / 2; if (with_dif) mored = dif; } } \
            if (hgh >= bclip) { hgh = bclip - 1; if (morem <= MM(bclip)) { morem = MM(bclip); morea = VV(bclip); morey = (morea - bclip) / 2; if (with_dif) mored = dif; } } \
            aclip = -INT_MAX; bclip = INT_MAX;                                                   \
        }                                                                                        \
    }
    CLIPS(0)
    /* ---- waves 1, 2, ...: one more difference each */
    while (more && (dir > 0 ? lasta >= besta - W_TRIM_MLAG : lasta <= besta + W_TRIM_MLAG)) {
        low -= 1; hgh += 1;
        wave_reserve(r, low - 1, hgh + 1);
        VV(low) = INV; VV(hgh) = INV;
        dif += 1;
        VV(hgh + 1) = INV; VV(low - 1) = INV;
        /* sweep against the direction of the dependency so that the old values of the neighbours are still there:
           forward goes from hgh down (needs old V[k-1]), reverse from low up (needs old V[k+1]) */
        uint64_t t_side = W_PATH_INT;       /* history of the neighbour already overwritten in this sweep */
        int n_side = W_PATH_LEN;
        int v_side = INV;                   /* its old furthest point */
        int v_self_next = dir > 0 ? VV(hgh) : VV(low);
        for (int k = dir > 0 ? hgh : low; dir > 0 ? k >= low : k <= hgh; k -= dir) {
            const int kn = k - dir;                             /* the neighbour not yet overwritten */
            const int v_self = v_self_next;                     /* old V[k] */
            const int v_far = VV(kn);                           /* old V[kn] */
            v_self_next = v_far;
            int c, m; uint64_t bv;
            /* forward: candidates v_side + 1 (from k + 1), v_far + 1 (from k - 1), v_self + 2 (same diagonal);
               reverse mirrors with - 1 / - 2.  Tie order as in align.c:737-768 / :1391-1422 */
            if (BETTER(v_far, v_self)) {
                if (BETTER(v_side, v_far)) { c = v_side + dir; m = n_side; bv = t_side; }
                else { c = v_far + dir; m = MM(kn); bv = TT(kn); }
            } else {
                if (BETTER(v_side, v_self)) { c = v_side + dir; m = n_side; bv = t_side; }
                else { c = v_self + 2 * dir; m = MM(k); bv = TT(k); }
            }
            if (bv & W_PATH_TOP) m -= 1;                        /* the difference enters the history */
            bv <<= 1;
            int y = (c - k) >> 1;
            SLIDE(k, y, bv, m, 1)
            c = (y << 1) + k;
            if (BETTER(c, besta)) {
                besta = c; besty = y;
                if (m >= r->path_ave) {
                    lasta = c;
                    if (r->tip_table[bv & W_TRIM_MASK] >= 0 &&
                        r->tip_table[(bv >> W_TRIM_LEN) & W_TRIM_MASK] + r->tip_score[bv & W_TRIM_MASK] >= 0) { trima = c; trimy = y; trimd = dif; }
                }
            }
            t_side = TT(k); n_side = MM(k); v_side = v_self;
            VV(k) = c; TT(k) = bv; MM(k) = m;
        }
        CLIPS(1)
        /* drop the diagonals that lag more than WAVE_LAG behind the best point */
        const int lim = dir > 0 ? besta - W_WAVE_LAG : besta + W_WAVE_LAG;
        while (hgh >= low) {
            if (BETTER(lim, VV(hgh))) hgh -= 1;
            else { while (BETTER(lim, VV(low))) low += 1; break; }
        }
    }
    WaveTip tip;
    if (morem >= 0) { tip.apos = morea - morey; tip.bpos = morey; tip.diffs = mored; }
    else { tip.apos = trima - trimy; tip.bpos = trimy; tip.diffs = trimd; }
    return tip;
#undef BETTER
#undef VV
#undef MM
#undef TT
#undef SLIDE
#undef CLIPS
}

/* ocda_go, oc_daligner.c:34-79.  query / target: one byte per base (0..3), sizes as given. */
int o_rescue_local(ORescue* r, const uint8_t* query, int qstart, int qsize, const uint8_t* target, int tstart, int tsize,
                   int min_align_size, int* qb, int* qe, int* tb, int* te, int* diffs)
{
    uint8_t* a = (uint8_t*)malloc((size_t)qsize + 2);
    uint8_t* b = (uint8_t*)malloc((size_t)tsize + 2);
    a[0] = 4; memcpy(a + 1, query, (size_t)qsize); a[qsize + 1] = 4;
    b[0] = 4; memcpy(b + 1, target, (size_t)tsize); b[tsize + 1] = 4;
    const int k0 = qstart - tstart, anti = qstart + tstart;
    const WaveTip f = wave(r, a + 1, b + 1, k0, anti, +1);
    /* the reverse pass starts from the diagonal the forward path started on: the only one there is */
    const WaveTip v = wave(r, a + 1, b + 1, k0, anti, -1);
    free(a); free(b);
    *qb = v.apos; *tb = v.bpos; *qe = f.apos; *te = f.bpos; *diffs = f.diffs + v.diffs;
    const int asize = *qe - *qb, bsize = *te - *tb;
    return asize >= min_align_size && bsize >= min_align_size;
}

/* ------------------------------------------------------------------ 2. global alignment, edlib's way */

#define OP_M 0      /* EDLIB_EDOP_MATCH */
#define OP_I 1      /* EDLIB_EDOP_INSERT: query base against a gap in the target */
#define OP_D 2      /* EDLIB_EDOP_DELETE */
#define OP_X 3      /* EDLIB_EDOP_MISMATCH */

typedef struct { uint8_t* a; size_t n, m; } OpVec;
static void opv_push(OpVec* v, uint8_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (uint8_t*)realloc(v->a, v->m); }
    v->a[v->n++] = x;
}

/* last column of the global DP of q[0..qn) against t[0..tn): col[i] = D[i][tn] (col[0] = tn).  Myers' bit-vector
 * recurrence over 64-row words with the global boundary D[0][j] = j (horizontal delta +1 into the first word); the
 * strides let the same routine run over the reversed sequences (Hirschberg's right half). */
static void nw_last_column(const uint8_t* q, int qn, const uint8_t* t, int tn, int q_step, int t_step, int* col)
{
    const int nw = (qn + 63) / 64;
    uint64_t* peq = (uint64_t*)calloc((size_t)nw * 4 + 1, sizeof(uint64_t));
    uint64_t* pv = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)nw + 1));
    uint64_t* mv = (uint64_t*)calloc((size_t)nw + 1, sizeof(uint64_t));
    for (int i = 0; i < qn; ++i) peq[(size_t)(q[(long)i * q_step] & 3) * nw + (i >> 6)] |= 1ull << (i & 63);
    for (int w = 0; w < nw; ++w) pv[w] = ~0ull;
    for (int j = 0; j < tn; ++j) {
        const uint64_t* eqc = peq + (size_t)(t[(long)j * t_step] & 3) * nw;
        int hin = 1;
        for (int w = 0; w < nw; ++w) {
            uint64_t Eq = eqc[w];
            const uint64_t Pv = pv[w], Mv = mv[w];
            const uint64_t Xv = Eq | Mv;
            if (hin < 0) Eq |= 1ull;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
            Ph <<= 1; Mh <<= 1;
            if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
            pv[w] = Mh | ~(Xv | Ph);
            mv[w] = Ph & Xv;
            hin = hout;
        }
    }
    col[0] = tn;
    for (int i = 0; i < qn; ++i) {
        const uint64_t bit = 1ull << (i & 63);
        col[i + 1] = col[i] + ((pv[i >> 6] & bit) ? 1 : 0) - ((mv[i >> 6] & bit) ? 1 : 0);
    }
    free(peq); free(pv); free(mv);
}

/*
 * obtainAlignmentTraceback, edlib.cpp:887-1090: walk back from (qn, tn); at every cell the reference tries, in this
 * order: up (a query base against a gap: INSERT) if the cell above holds score - 1, then left (DELETE) if the cell
 * to the left holds score - 1, then the diagonal.  Operations are produced last to first and reversed.
 */
static void nw_traceback(ORescue* r, const uint8_t* q, int qn, const uint8_t* t, int tn, OpVec* out)
{
    const size_t W = (size_t)tn + 1;
    const size_t need = ((size_t)qn + 1) * W;
    if (need > r->dp_cap) { free(r->dp); r->dp = (int*)malloc(sizeof(int) * need); r->dp_cap = need; }
    int* D = r->dp;
    for (int j = 0; j <= tn; ++j) D[j] = j;
    for (int i = 1; i <= qn; ++i) {
        int* row = D + (size_t)i * W;
        const int* prev = row - W;
        row[0] = i;
        const int qc = q[i - 1];
        for (int j = 1; j <= tn; ++j) {
            int v = prev[j] + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            const int dg = prev[j - 1] + (qc != t[j - 1]);
            if (dg < v) v = dg;
            row[j] = v;
        }
    }
    const size_t first = out->n;
    int i = qn, j = tn;
    while (i > 0 || j > 0) {
        const int sc = D[(size_t)i * W + j];
        if (i > 0 && D[(size_t)(i - 1) * W + j] + 1 == sc) { opv_push(out, OP_I); --i; }
        else if (j > 0 && D[(size_t)i * W + (j - 1)] + 1 == sc) { opv_push(out, OP_D); --j; }
        else { opv_push(out, q[i - 1] == t[j - 1] ? OP_M : OP_X); --i; --j; }
    }
    for (size_t x = first, y = out->n; x + 1 < y; ++x) { --y; const uint8_t s = out->a[x]; out->a[x] = out->a[y]; out->a[y] = s; }
}

/* test hook: the column routine on its own (checked against the kernel core's strip-mined version) */
void o_nw_last_column(const uint8_t* q, int qn, const uint8_t* t, int tn, int q_step, int t_step, int* col)
{
    nw_last_column(q, qn, t, tn, q_step, t_step, col);
}

/* test hook: the leaf (full-matrix traceback) on its own; ops_out receives qn + tn bytes at most, returns their number */
int o_nw_traceback_ops(const uint8_t* q, int qn, const uint8_t* t, int tn, uint8_t* ops_out)
{
    ORescue* r = o_rescue_new(0.5);
    OpVec v = { NULL, 0, 0 };
    nw_traceback(r, q, qn, t, tn, &v);
    memcpy(ops_out, v.a, v.n);
    const int n = (int)v.n;
    free(v.a);
    o_rescue_free(r);
    return n;
}

static int ceil_div(int x, int y) { return x % y ? x / y + 1 : x / y; }

/* obtainAlignment, edlib.cpp:1106-1161 + obtainAlignmentHirschberg, :1177-1340 */
static int nw_align(ORescue* r, const uint8_t* q, int qn, const uint8_t* t, int tn, int best, OpVec* out)
{
    if (qn == 0 || tn == 0) {
        for (int x = 0; x < qn + tn; ++x) opv_push(out, qn == 0 ? OP_D : OP_I);
        return 1;
    }
    const int blocks = ceil_div(qn, 64);
    const long long data = (long long)(2 * 8 + 4) * blocks * tn + (long long)2 * 4 * tn;
    if (data < 1024 * 1024) { nw_traceback(r, q, qn, t, tn, out); return 1; }
    /* Hirschberg: the optimal path crosses from column lw - 1 to column lw going from row i to row i + 1 for the FIRST
       row i in 0 .. qn - 2 with left[i] + right[i + 1] == best; failing that through the boundary rows -1, then qn - 1 */
    const int lw = tn / 2, rw = tn - lw;
    int* left = (int*)malloc(sizeof(int) * ((size_t)qn + 1));
    int* right = (int*)malloc(sizeof(int) * ((size_t)qn + 1));
    nw_last_column(q, qn, t, lw, 1, 1, left);                            /* left[i + 1] = D(q[0..i], t[0..lw)) */
    nw_last_column(q + qn - 1, qn, t + tn - 1, rw, -1, -1, right);       /* right[x] = D(last x query bases, last rw target bases) */
    int split = -2, ls = 0, rs = 0;
    for (int i = 0; i <= qn - 2; ++i) {
        const int l = left[i + 1], rr = right[qn - 1 - i];
        if (l > best || rr > best) continue;                             /* outside the band the reference stores: never equal */
        if (l + rr == best) { split = i; ls = l; rs = rr; break; }
    }
    if (split == -2 && lw + right[qn] == best) { split = -1; ls = lw; rs = right[qn]; }
    if (split == -2 && left[qn] + rw == best) { split = qn - 1; ls = left[qn]; rs = rw; }
    free(left); free(right);
    if (split == -2) return 0;
    const int ul = split + 1;
    if (!nw_align(r, q, ul, t, lw, ls, out)) return 0;
    return nw_align(r, q + ul, qn - ul, t + lw, rw, rs, out);
}

/*
 * edlib_go(..., tolerance, min_align_size, find_path = TRUE, kMatchSize), edlib_wrapper.c:111-255: global alignment of
 * query[qfrom, qto) with target[tfrom, tto) if its edit distance is <= tolerance; trimmed to the first and the last
 * run of kMatchSize matching columns.  Returns 1 and the gapped strings ('-' gaps, "ACGT") on success.
 */
int o_rescue_global(ORescue* r, const uint8_t* query, int qfrom, int qto, const uint8_t* target, int tfrom, int tto,
                    int tolerance, int min_align_size, int k_match, int* qoff, int* qend, int* toff, int* tend,
                    double* ident_perc, char** qaln, char** taln, int* aln_size)
{
    const uint8_t* q = query + qfrom; const uint8_t* t = target + tfrom;
    const int qn = qto - qfrom, tn = tto - tfrom;
    if (qn <= 0 || tn <= 0) return 0;
    int* col = (int*)malloc(sizeof(int) * ((size_t)qn + 1));
    nw_last_column(q, qn, t, tn, 1, 1, col);
    const int dist = col[qn];
    free(col);
    if (tolerance >= 0 && dist > tolerance) return 0;                    /* numLocations == 0 */
    /* alignLen = endLocations[0] - startLocations[0] = tn - 1 (:165-168) */
    const int align_len = tn - 1;
    const double align_diff = dist / (double)align_len;
    if (align_len < min_align_size) return 0;
    if (align_diff > r->error) return 0;
    OpVec ops = { NULL, 0, 0 };
    if (!nw_align(r, q, qn, t, tn, dist, &ops)) { free(ops.a); return 0; }
    const int n = (int)ops.n;
    /* first run of k_match matches from the left (:197-216) */
    int from = 0, pq = 0, pt = 0, m = 0;
    while (from < n) {
        const int op = ops.a[from];
        m = op == OP_M ? m + 1 : 0;
        if (op != OP_D) ++pq;
        if (op != OP_I) ++pt;
        ++from;
        if (m == k_match) break;
    }
    if (m != k_match) { free(ops.a); return 0; }
    from -= k_match; pq -= k_match; pt -= k_match;
    /* last run from the right (:218-233) */
    int to = n, tq = 0, tt = 0;
    m = 0;
    while (to) {
        const int op = ops.a[to - 1];
        m = op == OP_M ? m + 1 : 0;
        if (op != OP_D) ++tq;
        if (op != OP_I) ++tt;
        --to;
        if (m == k_match) break;
    }
    to += k_match; tq -= k_match; tt -= k_match;
    const int len = to - from;
    char* qa = (char*)malloc((size_t)len + 1);
    char* ta = (char*)malloc((size_t)len + 1);
    int qi = pq, ti = pt, same = 0;
    for (int x = from; x < to; ++x) {
        const int op = ops.a[x];
        qa[x - from] = op == OP_D ? '-' : "ACGT"[q[qi]];
        ta[x - from] = op == OP_I ? '-' : "ACGT"[t[ti]];
        if (op != OP_D) ++qi;
        if (op != OP_I) ++ti;
        same += op == OP_M;
    }
    qa[len] = 0; ta[len] = 0;
    free(ops.a);
    *qoff = qfrom + pq; *qend = qfrom + (qn - tq); *toff = tfrom + pt; *tend = tfrom + (tn - tt);
    *ident_perc = len ? 100.0 * same / len : 0.0;
    *qaln = qa; *taln = ta; *aln_size = len;
    return 1;
}
