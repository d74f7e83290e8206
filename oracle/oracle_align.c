/*
 * oracle_align.c -- CPU restatement of the block-wise edit-distance extension
 * (gapped_align/oc_aligner.c + gapped_align/edlib_ex.c).
 * TEST INFRASTRUCTURE ONLY -- see necat_oracle.h.
 *
 * Restatement notes.  The reference's Edlib_align (edlib_ex.c:733-800) runs a
 * banded Myers bit-vector prefix search (SHW, :106-223), then a second banded
 * global pass that stores the band (:226-370) and walks it back (:383-621).
 * Inside the Ukkonen band every cell that can lie on an alignment of cost <= k is
 * exact, so the result equals:
 *   1. d* = min_j D[qn][j] over target prefixes j = 1..tn, fail if d* > k0 with
 *      k0 = (int)(min(qn, tn) * error * 1.1) (:751); j* = smallest j attaining d*
 *      (:190-199 keeps the first end position);
 *   2. the global alignment of the query with target[0..j*) obtained by walking
 *      back from (qn, j*) preferring "up" (query base against a gap), then
 *      "left" (gap against a target base), then the diagonal (:463,500,540).
 * Here one unbanded bit-vector pass stores the vertical delta vectors of every
 * column; the walk-back derives the three neighbour values from them.
 * o_edlib_align_dp() is the same statement on a plain integer matrix and is
 * used by the tests to cross-check the bit-vector version.
 */
#include "necat_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define O_MAX_WORDS 64      /* edlib_ex_aux.h:18 MaxNumBlocks */
#define O_MAX_COLS  4096    /* edlib_ex_aux.h:17 MaxSeqSize */

struct OAligner {
    double    error;
    uint64_t* pv;       /* [O_MAX_COLS][O_MAX_WORDS] */
    uint64_t* mv;
    int*      bottom;   /* D[qn][j] for j = 1..tn at index j-1 */
    uint8_t*  ops;      /* walk-back scratch */
    char *qabuf, *tabuf;
    uint8_t *qfrag, *tfrag;
    char *qaln[2], *taln[2];    /* [0] = left (reverse) extension, [1] = right */
    int   aln_len[2];
    size_t aln_cap;
    char *q_out, *t_out; size_t out_cap;
};

OAligner* o_aligner_new(double error)
{
    OAligner* a = (OAligner*)calloc(1, sizeof(OAligner));
    a->error = error;
    a->pv = (uint64_t*)malloc(sizeof(uint64_t) * O_MAX_COLS * O_MAX_WORDS);
    a->mv = (uint64_t*)malloc(sizeof(uint64_t) * O_MAX_COLS * O_MAX_WORDS);
    a->bottom = (int*)malloc(sizeof(int) * O_MAX_COLS);
    a->ops = (uint8_t*)malloc(2 * O_MAX_COLS + 64);
    a->qabuf = (char*)malloc(100000); a->tabuf = (char*)malloc(100000);  /* oc_aligner.c:24-25 */
    a->qfrag = (uint8_t*)malloc(O_MAX_COLS + 8); a->tfrag = (uint8_t*)malloc(O_MAX_COLS + 8);
    return a;
}

void o_aligner_free(OAligner* a)
{
    if (!a) return;
    free(a->pv); free(a->mv); free(a->bottom); free(a->ops); free(a->qabuf); free(a->tabuf);
    free(a->qfrag); free(a->tfrag);
    for (int i = 0; i < 2; ++i) { free(a->qaln[i]); free(a->taln[i]); }
    free(a->q_out); free(a->t_out);
    free(a);
}

static const char kDna[] = "ACGT-";

/* Emit the gapped strings for an op list given in walk-back (end to start) order.
 * op: 0 = diagonal, 1 = up (query base / gap), 2 = left (gap / target base). */
static void ops_to_strings(const uint8_t* ops, int nops, const uint8_t* q, const uint8_t* t,
                           char* qaln, char* taln, int* qend, int* tend)
{
    int qi = 0, ti = 0, n = 0;
    for (int x = nops - 1; x >= 0; --x) {
        switch (ops[x]) {
        case 0: qaln[n] = kDna[q[qi++]]; taln[n] = kDna[t[ti++]]; break;
        case 1: qaln[n] = kDna[q[qi++]]; taln[n] = '-'; break;
        default: qaln[n] = '-'; taln[n] = kDna[t[ti++]]; break;
        }
        ++n;
    }
    qaln[n] = 0; taln[n] = 0;
    *qend = qi; *tend = ti;
}

static int popcnt_below(const uint64_t* w, int nbits)
{
    int c = 0, full = nbits >> 6, rem = nbits & 63;
    for (int i = 0; i < full; ++i) c += __builtin_popcountll(w[i]);
    if (rem) c += __builtin_popcountll(w[full] & ((1ull << rem) - 1));
    return c;
}

int o_edlib_align(OAligner* A, const uint8_t* q, int qn, const uint8_t* t, int tn,
                  char* qaln, char* taln, int* qend, int* tend)
{
    qaln[0] = 0; taln[0] = 0; *qend = 0; *tend = 0;
    if (qn <= 0 || tn <= 0 || qn > O_MAX_WORDS * 64 || tn > O_MAX_COLS) return 0;
    const int nw = (qn + 63) >> 6;
    const int k0 = (int)((qn < tn ? qn : tn) * A->error * 1.1);
    uint64_t peq[4][O_MAX_WORDS];
    memset(peq, 0, sizeof peq);
    for (int i = 0; i < qn; ++i) peq[q[i]][i >> 6] |= 1ull << (i & 63);
    uint64_t pv[O_MAX_WORDS], mv[O_MAX_WORDS];
    for (int w = 0; w < nw; ++w) { pv[w] = ~0ull; mv[w] = 0; }
    const int lastw = (qn - 1) >> 6, lastb = (qn - 1) & 63;
    int score = qn, best = -1, best_j = -1;
    for (int j = 0; j < tn; ++j) {
        const uint64_t* eqv = peq[t[j]];
        int hin = 1;                         /* D[0][j] = j */
        for (int w = 0; w < nw; ++w) {
            uint64_t P = pv[w], M = mv[w], E = eqv[w];
            uint64_t xv = E | M;
            if (hin < 0) E |= 1;
            uint64_t xh = (((E & P) + P) ^ P) | E;
            uint64_t ph = M | ~(xh | P);
            uint64_t mh = P & xh;
            if (w == lastw) score += (int)((ph >> lastb) & 1) - (int)((mh >> lastb) & 1);
            int hout = (int)(ph >> 63) - (int)(mh >> 63);
            ph <<= 1; mh <<= 1;
            if (hin < 0) mh |= 1; else if (hin > 0) ph |= 1;
            pv[w] = mh | ~(xv | ph);
            mv[w] = ph & xv;
            hin = hout;
            A->pv[(size_t)j * O_MAX_WORDS + w] = pv[w];
            A->mv[(size_t)j * O_MAX_WORDS + w] = mv[w];
        }
        if (best < 0 || score < best) { best = score; best_j = j; }
    }
    if (best > k0) return 0;
    /* walk back from (qn, best_j + 1) */
    int i = qn, j = best_j + 1, cur = best, nops = 0;
    while (i > 0 || j > 0) {
        if (i == 0) { A->ops[nops++] = 2; --j; continue; }
        if (j == 0) { A->ops[nops++] = 1; --i; continue; }
        const uint64_t* cpv = A->pv + (size_t)(j - 1) * O_MAX_WORDS;
        const uint64_t* cmv = A->mv + (size_t)(j - 1) * O_MAX_WORDS;
        const int w = (i - 1) >> 6, b = (i - 1) & 63;
        int vd = (int)((cpv[w] >> b) & 1) - (int)((cmv[w] >> b) & 1);   /* D[i][j] - D[i-1][j] */
        if (vd == 1) { A->ops[nops++] = 1; --i; --cur; continue; }
        int left;                                                       /* D[i][j-1] */
        if (j == 1) left = i;
        else {
            const uint64_t* lpv = A->pv + (size_t)(j - 2) * O_MAX_WORDS;
            const uint64_t* lmv = A->mv + (size_t)(j - 2) * O_MAX_WORDS;
            left = (j - 1) + popcnt_below(lpv, i) - popcnt_below(lmv, i);
        }
        if (left + 1 == cur) { A->ops[nops++] = 2; --j; cur = left; continue; }
        /* diagonal: D[i-1][j-1] = D[i-1][j] - (horizontal delta), recomputed from the left column */
        int diag;
        if (j == 1) diag = i - 1;
        else {
            const uint64_t* lpv = A->pv + (size_t)(j - 2) * O_MAX_WORDS;
            const uint64_t* lmv = A->mv + (size_t)(j - 2) * O_MAX_WORDS;
            diag = (j - 1) + popcnt_below(lpv, i - 1) - popcnt_below(lmv, i - 1);
        }
        A->ops[nops++] = 0; --i; --j; cur = diag;
    }
    ops_to_strings(A->ops, nops, q, t, qaln, taln, qend, tend);
    return 1;
}

/* Plain-matrix statement of the same function (tests only; O(qn * tn) ints). */
int o_edlib_align_dp(double error, const uint8_t* q, int qn, const uint8_t* t, int tn,
                     char* qaln, char* taln, int* qend, int* tend)
{
    qaln[0] = 0; taln[0] = 0; *qend = 0; *tend = 0;
    if (qn <= 0 || tn <= 0) return 0;
    const int k0 = (int)((qn < tn ? qn : tn) * error * 1.1);
    const int W = tn + 1;
    int* D = (int*)malloc(sizeof(int) * (size_t)(qn + 1) * W);
    for (int j = 0; j <= tn; ++j) D[j] = j;
    for (int i = 1; i <= qn; ++i) {
        D[i * W] = i;
        for (int j = 1; j <= tn; ++j) {
            int d = D[(i - 1) * W + j - 1] + (q[i - 1] != t[j - 1]);
            int u = D[(i - 1) * W + j] + 1, l = D[i * W + j - 1] + 1;
            if (u < d) d = u;
            if (l < d) d = l;
            D[i * W + j] = d;
        }
    }
    int best = -1, bj = 0;
    for (int j = 1; j <= tn; ++j) if (best < 0 || D[qn * W + j] < best) { best = D[qn * W + j]; bj = j; }
    if (best > k0) { free(D); return 0; }
    uint8_t* ops = (uint8_t*)malloc((size_t)qn + tn + 8);
    int i = qn, j = bj, nops = 0;
    while (i > 0 || j > 0) {
        if (i == 0) { ops[nops++] = 2; --j; }
        else if (j == 0) { ops[nops++] = 1; --i; }
        else if (D[(i - 1) * W + j] + 1 == D[i * W + j]) { ops[nops++] = 1; --i; }
        else if (D[i * W + j - 1] + 1 == D[i * W + j]) { ops[nops++] = 2; --j; }
        else { ops[nops++] = 0; --i; --j; }
    }
    ops_to_strings(ops, nops, q, t, qaln, taln, qend, tend);
    free(ops); free(D);
    return 1;
}

/* ------------------------------------------------------------- extension */

static void aln_reserve(OAligner* A, size_t need)
{
    if (need <= A->aln_cap) return;
    A->aln_cap = need * 2 + 1024;
    for (int i = 0; i < 2; ++i) {
        A->qaln[i] = (char*)realloc(A->qaln[i], A->aln_cap);
        A->taln[i] = (char*)realloc(A->taln[i], A->aln_cap);
    }
}

/*
 * oc_aligner.c:157-286: extend from a fixed start in one direction, one block of
 * up to block_size bases per step.  dir = +1 reads query[0], query[1], ...;
 * dir = -1 reads query[0], query[-1], ... (the caller passes the pointer to the
 * first base of the extension).  Output goes to A->qaln[side] / A->taln[side].
 */
static void extend_one_side(OAligner* A, int side, const uint8_t* query, int qsize, const uint8_t* target, int tsize,
                            int block_size, int dir, int tail_match_len)
{
    int qidx = 0, tidx = 0, out = 0;
    const int kMat = 8;                                 /* oc_aligner.c:9 */
    for (;;) {
        /* next block (:111-154) */
        const int qleft = qsize - qidx, tleft = tsize - tidx;
        int qblk, tblk, last;
        if (qleft < block_size + 100 || tleft < block_size + 100) {
            qblk = (int)(tleft * 1.3); if (qblk > qleft) qblk = qleft;
            tblk = (int)(qleft * 1.3); if (tblk > tleft) tblk = tleft;
            last = 1;
        } else { qblk = tblk = block_size; last = 0; }
        if (qblk == 0 || tblk == 0) break;
        for (int i = 0; i < qblk; ++i) A->qfrag[i] = query[dir * (qidx + i)];
        for (int i = 0; i < tblk; ++i) A->tfrag[i] = target[dir * (tidx + i)];
        int qfae, tfae;
        o_edlib_align(A, A->qfrag, qblk, A->tfrag, tblk, A->qabuf, A->tabuf, &qfae, &tfae);

        int done = last;
        if (qblk - qfae > 30 && tblk - tfae > 30) done = 1;         /* :220 */
        const int M = done ? tail_match_len : kMat;
        int asize = (int)strlen(A->qabuf);
        int acnt = 0, qcnt = 0, tcnt = 0, k = asize - 1, m = 0;
        while (k >= 0) {                                             /* :225-240 */
            if (A->qabuf[k] != '-') ++qcnt;
            if (A->tabuf[k] != '-') ++tcnt;
            m = (A->qabuf[k] == A->tabuf[k]) ? m + 1 : 0;
            ++acnt;
            if (m == M) break;
            --k;
        }
        if (m != M || k < 1) {                                       /* :243-253: exact-match prefix only */
            asize = 0;
            for (int i = 0; i < qblk && i < tblk; ++i) {
                if (A->qfrag[i] != A->tfrag[i]) break;
                A->qabuf[asize] = kDna[A->qfrag[i]];
                A->tabuf[asize] = kDna[A->tfrag[i]];
                ++asize;
            }
            done = 1;
        } else {
            asize -= acnt;
            qidx += qfae - qcnt;
            tidx += tfae - tcnt;
            if (done) asize += M;
        }
        aln_reserve(A, (size_t)out + (size_t)asize + 8);
        memcpy(A->qaln[side] + out, A->qabuf, (size_t)asize);
        memcpy(A->taln[side] + out, A->tabuf, (size_t)asize);
        out += asize;
        if (done) break;
    }
    aln_reserve(A, (size_t)out + 8);
    A->aln_len[side] = out;
}

int o_onc_align(OAligner* A, const uint8_t* query, int qstart, int qsize, const uint8_t* target, int tstart, int tsize,
                int block_size, int min_align_size, int tail_match_len, OAlignment* res)
{
    const int kMat = 8;
    int QS = qstart, TS = tstart;
    aln_reserve(A, 64);
    /* left of the seed, read right-to-left (:313-327) */
    extend_one_side(A, 0, query + QS - 1, QS, target + TS - 1, TS, block_size, -1, tail_match_len);
    size_t need = (size_t)A->aln_len[0] + 64;
    if (need > A->out_cap) { A->out_cap = need * 2; A->q_out = (char*)realloc(A->q_out, A->out_cap); A->t_out = (char*)realloc(A->t_out, A->out_cap); }
    int n = 0, rq = 0, rt = 0;
    {
        /* re-anchor at the first run of 8 matches counted from the seed (:329-367) */
        int qc = 0, tc = 0, ac = 0, m = 0;
        const int L = A->aln_len[0];
        for (int i = 0; i < L; ++i) {
            if (A->qaln[0][i] != '-') ++qc;
            if (A->taln[0][i] != '-') ++tc;
            m = (A->qaln[0][i] == A->taln[0][i]) ? m + 1 : 0;
            ++ac;
            if (m == kMat) break;
        }
        if (m == kMat) {
            QS -= qc; TS -= tc;
            for (int i = L; i > ac; --i) {
                char a = A->qaln[0][i - 1], b = A->taln[0][i - 1];
                A->q_out[n] = a; A->t_out[n] = b; ++n;
                if (a != '-') ++rq;
                if (b != '-') ++rt;
            }
        }
    }
    extend_one_side(A, 1, query + QS, qsize - QS, target + TS, tsize - TS, block_size, +1, tail_match_len);
    need = (size_t)n + (size_t)A->aln_len[1] + 64;
    if (need > A->out_cap) { A->out_cap = need * 2; A->q_out = (char*)realloc(A->q_out, A->out_cap); A->t_out = (char*)realloc(A->t_out, A->out_cap); }
    int fq = 0, ft = 0;
    const int L1 = A->aln_len[1];
    if (n == 0) {                                       /* :385-415 */
        int qc = 0, tc = 0, ac = 0, m = 0;
        for (int i = 0; i < L1; ++i) {
            if (A->qaln[1][i] != '-') ++qc;
            if (A->taln[1][i] != '-') ++tc;
            m = (A->qaln[1][i] == A->taln[1][i]) ? m + 1 : 0;
            ++ac;
            if (m == kMat) break;
        }
        if (m == kMat) {
            ac -= kMat; qc -= kMat; tc -= kMat;
            QS += qc; TS += tc;
            for (int i = ac; i < L1; ++i) {
                char a = A->qaln[1][i], b = A->taln[1][i];
                if (a != '-') ++fq;
                if (b != '-') ++ft;
                A->q_out[n] = a; A->t_out[n] = b; ++n;
            }
        }
    } else {                                            /* :416-427 */
        for (int i = 0; i < L1; ++i) {
            char a = A->qaln[1][i], b = A->taln[1][i];
            if (a != '-') ++fq;
            if (b != '-') ++ft;
            A->q_out[n] = a; A->t_out[n] = b; ++n;
        }
    }
    A->q_out[n] = 0; A->t_out[n] = 0;
    res->qoff = QS - rq; res->qend = QS + fq; res->toff = TS - rt; res->tend = TS + ft;
    res->qaln = A->q_out; res->taln = A->t_out; res->aln_size = n;
    int same = 0;
    for (int i = 0; i < n; ++i) same += A->q_out[i] == A->t_out[i];
    res->ident_perc = n ? 100.0 * same / n : 0.0;       /* :289-300 */
    res->ok = n >= min_align_size;
    return res->ok;
}
