/*
 * oracle_cns.c -- CPU restatement of the oc2cns half: candidate regrouping
 * (oc2pcan), per-template overlap selection, tag generation, backbone
 * construction and best-path consensus.
 * TEST INFRASTRUCTURE ONLY -- see necat_oracle.h.
 */
#include "necat_oracle.h"
#include "oracle_rescue.h"

#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int o_cns_tag_order_mode = 0;
int o_cns_partition_size = 100000;

void o_cns_options_default(OCnsOptions* o)
{
    o->min_align_size = 400; o->min_cov = 4; o->max_cov = 12; o->min_size = 500; o->full_consensus = 0;
    o->error = 0.5; o->mapping_ratio = 0.8; o->num_threads = 1; o->rescue_long_indels = 0;
    o->use_fixed_ident_cutoff = 0; o->small_memory = 0;
}

void o_cns_recvec_free(OCnsRecVec* v)
{
    for (size_t i = 0; i < v->n; ++i) free(v->a[i].seq);
    free(v->a);
    memset(v, 0, sizeof *v);
}

static void rec_push(OCnsRecVec* v, const OCnsRecord* r)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 64; v->a = (OCnsRecord*)realloc(v->a, sizeof(OCnsRecord) * v->m); }
    v->a[v->n++] = *r;
}

/* ------------------------------------------------------------------- tags */

/* tasc/align_tags.h:8-16 (+ seq: order of creation, for the canonical tie order) */
typedef struct {
    double  weight;
    int     t_pos, p_t_pos;
    uint8_t delta, p_delta;
    char    q_base, p_q_base;
    int     seq;
} OTag;

typedef struct { OTag* a; size_t n, m; } OTagVec;

static void tag_push(OTagVec* v, const OTag* t)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 4096; v->a = (OTag*)realloc(v->a, sizeof(OTag) * v->m); }
    v->a[v->n++] = *t;
}

/* tasc/align_tags.c:23-71 */
static int tags_from_alignment(const char* qaln, const char* taln, size_t n, int toff, double weight, OTagVec* tags)
{
    int run = 0;
    for (size_t i = 0; i < n; ++i) {            /* an insertion run of >= 255 drops the whole alignment (:40-44) */
        if (qaln[i] != '-') ++run;
        if (taln[i] != '-') run = 0;
        if (run >= 255) return 0;
    }
    OTag t;
    t.weight = weight;
    int j = toff - 1, jj = 0, pj = -1, pjj = 0;
    char pq = '-';
    for (size_t i = 0; i < n; ++i) {
        if (qaln[i] != '-') ++jj;
        if (taln[i] != '-') { ++j; jj = 0; }
        t.t_pos = j; t.p_t_pos = pj; t.delta = (uint8_t)jj; t.p_delta = (uint8_t)pjj;
        t.q_base = qaln[i]; t.p_q_base = pq;
        t.seq = (int)tags->n;
        pj = j; pjj = jj; pq = qaln[i];
        tag_push(tags, &t);
    }
    return 1;
}

/* tasc/align_tags.c:6-18 */
static int tag_lt(const OTag* a, const OTag* b)
{
    if (a->t_pos != b->t_pos) return a->t_pos < b->t_pos;
    if (a->delta != b->delta) return a->delta < b->delta;
    if (a->q_base != b->q_base) return a->q_base < b->q_base;
    if (a->p_t_pos != b->p_t_pos) return a->p_t_pos < b->p_t_pos;
    if (a->p_delta != b->p_delta) return a->p_delta < b->p_delta;
    if (a->p_q_base != b->p_q_base) return a->p_q_base < b->p_q_base;
    return 0;
}

static int tag_cmp_canonical(const void* x, const void* y)
{
    const OTag* a = (const OTag*)x; const OTag* b = (const OTag*)y;
    if (tag_lt(a, b)) return -1;
    if (tag_lt(b, a)) return 1;
    return a->seq < b->seq ? -1 : (a->seq > b->seq);
}

/*
 * Emulation of klib's introsort (klib/ksort.h:166-229) on OTag with tag_lt, so that
 * the order of equal-key tags -- and with it the floating-point summation order
 * in the link weights (tasc/cns_aux.c:48) -- is the one the reference binary
 * produces.  Median-of-three quicksort with an explicit stack, ranges of <= 16
 * left for a final insertion sort, comb sort when the depth budget runs out.
 */
static void tag_insertsort(OTag* s, OTag* t)
{
    for (OTag* i = s + 1; i < t; ++i)
        for (OTag* j = i; j > s && tag_lt(j, j - 1); --j) { OTag x = *j; *j = *(j - 1); *(j - 1) = x; }
}

static void tag_combsort(size_t n, OTag* a)
{
    const double shrink = 1.2473309501039786540366528676643;
    int swapped;
    size_t gap = n;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        swapped = 0;
        for (OTag* i = a; i < a + n - gap; ++i) {
            OTag* j = i + gap;
            if (tag_lt(j, i)) { OTag x = *i; *i = *j; *j = x; swapped = 1; }
        }
    } while (swapped || gap > 2);
    if (gap != 1) tag_insertsort(a, a + n);
}

static void tag_introsort_klib(size_t n, OTag* a)
{
    typedef struct { OTag *l, *r; int depth; } frame_t;
    if (n < 1) return;
    if (n == 2) { if (tag_lt(&a[1], &a[0])) { OTag x = a[0]; a[0] = a[1]; a[1] = x; } return; }
    int d;
    for (d = 2; (1ul << d) < n; ++d) {}
    frame_t* stack = (frame_t*)malloc(sizeof(frame_t) * (sizeof(size_t) * d + 2));
    frame_t* top = stack;
    OTag *s = a, *t = a + (n - 1);
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { tag_combsort((size_t)(t - s) + 1, s); t = s; continue; }
            OTag *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
            if (tag_lt(k, i)) { if (tag_lt(k, j)) k = j; }
            else k = tag_lt(j, i) ? i : j;
            OTag rp = *k;
            if (k != t) { OTag x = *k; *k = *t; *t = x; }
            for (;;) {
                do ++i; while (tag_lt(i, &rp));
                do --j; while (i <= j && tag_lt(&rp, j));
                if (j <= i) break;
                OTag x = *i; *i = *j; *j = x;
            }
            { OTag x = *i; *i = *t; *t = x; }
            if (i - s > t - i) {
                if (i - s > 16) { top->l = s; top->r = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->l = i + 1; top->r = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { free(stack); tag_insertsort(a, a + n); return; }
            --top; s = top->l; t = top->r; d = top->depth;
        }
    }
}

/* --------------------------------------------------------------- backbone */

/* One node of the consensus graph = (t_pos, delta, base); tasc/cns_aux.h:19-31 */
typedef struct {
    int    first_link, n_link;
    int    coverage;
    int    best_pt, best_pd, best_pb;
    double score;
} ONode;

typedef struct { double weight; int pt, pd, pb; } OLink;   /* cns_aux.h:9-15, pb = base code 0..4 */

typedef struct {
    int*   pos_first;   /* [T + 1] first node slot of each template position */
    int*   pos_ndelta;  /* [T] */
    ONode* node;        /* node[(pos_first[t] + delta) * 5 + base] */
    OLink* link;
    size_t n_node_slots, n_link;
    int*   coverage;    /* [T] */
    int    T;
} OBackbone;

static int base_code(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

static ONode* bb_node(const OBackbone* bb, int t, int d, int b)
{
    return bb->node + ((size_t)(bb->pos_first[t] + d) * 5 + b);
}

/* tasc/cns_aux.c:22-124 */
static void backbone_build(OTagVec* tags, int T, OBackbone* bb)
{
    memset(bb, 0, sizeof *bb);
    bb->T = T;
    bb->pos_first = (int*)calloc((size_t)T + 1, sizeof(int));
    bb->pos_ndelta = (int*)calloc((size_t)T + 1, sizeof(int));
    bb->coverage = (int*)calloc((size_t)T + 1, sizeof(int));
    OTag* a = tags->a;
    const size_t n = tags->n;
    if (o_cns_tag_order_mode == 1) tag_introsort_klib(n, a);
    else qsort(a, n, sizeof(OTag), tag_cmp_canonical);
    /* slots: n_delta(t) = largest delta at t + 1 */
    for (size_t i = 0; i < n;) {
        size_t j = i;
        while (j < n && a[j].t_pos == a[i].t_pos) ++j;
        bb->pos_ndelta[a[i].t_pos] = a[j - 1].delta + 1;
        i = j;
    }
    int slots = 0;
    for (int t = 0; t < T; ++t) { bb->pos_first[t] = slots; slots += bb->pos_ndelta[t]; }
    bb->pos_first[T] = slots;
    bb->n_node_slots = (size_t)slots;
    bb->node = (ONode*)malloc(sizeof(ONode) * ((size_t)slots * 5 + 1));
    for (size_t i = 0; i < (size_t)slots * 5; ++i) {
        ONode* nd = bb->node + i;
        nd->first_link = 0; nd->n_link = 0; nd->coverage = 0;
        nd->best_pt = -1; nd->best_pd = 255; nd->best_pb = '.'; nd->score = 0;       /* cns_aux.c:63-70 */
    }
    bb->link = (OLink*)malloc(sizeof(OLink) * (n + 1));
    size_t nl = 0;
    for (size_t i = 0; i < n;) {
        /* node group: same (t_pos, delta, q_base) */
        size_t j = i;
        while (j < n && a[j].t_pos == a[i].t_pos && a[j].delta == a[i].delta && a[j].q_base == a[i].q_base) ++j;
        ONode* nd = bb_node(bb, a[i].t_pos, a[i].delta, base_code(a[i].q_base));
        nd->first_link = (int)nl;
        nd->coverage = (int)(j - i);
        for (size_t x = i; x < j;) {
            size_t y = x + 1;
            while (y < j && a[y].p_t_pos == a[x].p_t_pos && a[y].p_delta == a[x].p_delta && a[y].p_q_base == a[x].p_q_base) ++y;
            OLink* L = bb->link + nl++;
            L->pt = a[x].p_t_pos; L->pd = a[x].p_delta; L->pb = base_code(a[x].p_q_base);
            L->weight = 0;
            for (size_t z = x; z < y; ++z) L->weight += a[z].weight;     /* summed in sorted order (:48) */
            x = y;
        }
        nd->n_link = (int)(nl - (size_t)nd->first_link);
        i = j;
    }
    bb->n_link = nl;
    /* coverage[t] = number of tags with delta 0 at t (:96-98) */
    for (size_t i = 0; i < n; ++i) if (a[i].delta == 0) ++bb->coverage[a[i].t_pos];
}

static void backbone_free(OBackbone* bb)
{
    free(bb->pos_first); free(bb->pos_ndelta); free(bb->node); free(bb->link); free(bb->coverage);
    memset(bb, 0, sizeof *bb);
}

/* tasc/cns_aux.c:127-217.  Returns the consensus of [from, to) as base codes. */
static int backbone_segment(OBackbone* bb, int from, int to, uint8_t** out, int* cns_from, int* cns_to)
{
    ONode* gbest = NULL;
    int gt = 0, gb = -1;
    double gscore = -1.0;
    for (int t = from; t < to; ++t)
        for (int d = 0; d < bb->pos_ndelta[t]; ++d)
            for (int b = 0; b < 5; ++b) {
                ONode* nd = bb_node(bb, t, d, b);
                if (!nd->coverage) continue;
                double best = -1;
                for (int l = 0; l < nd->n_link; ++l) {
                    const OLink* L = bb->link + nd->first_link + l;
                    double sc = L->weight - 0.4 * 0.5 * bb->coverage[t];
                    if (L->pt != -1) sc += bb_node(bb, L->pt, L->pd, L->pb)->score;
                    if (sc > best) { best = sc; nd->best_pt = L->pt; nd->best_pd = L->pd; nd->best_pb = L->pb; }
                }
                nd->score = best;
                if (best > gscore) { gscore = best; gbest = nd; gt = t; gb = b; }
            }
    int cap = 1024, n = 0;
    uint8_t* s = (uint8_t*)malloc((size_t)cap);
    int b = gb, cfrom = 0;
    ONode* nd = gbest;
    while (nd) {
        int cur = b;
        int pt = nd->best_pt;
        if (pt == -1) break;                     /* the first node of the walk is not emitted */
        int pd = nd->best_pd; b = nd->best_pb;
        nd = bb_node(bb, pt, pd, b);
        cfrom = pt;
        if (cur != 4) { if (n == cap) { cap *= 2; s = (uint8_t*)realloc(s, (size_t)cap); } s[n++] = (uint8_t)cur; }
    }
    for (int i = 0; i < n / 2; ++i) { uint8_t x = s[i]; s[i] = s[n - 1 - i]; s[n - 1 - i] = x; }
    *out = s;
    if (cns_from) *cns_from = cfrom;
    if (cns_to) *cns_to = gt + 1;
    return n;
}

/* ------------------------------------------------------- template driver */

static int cand_cns_gt(const void* x, const void* y)      /* gapped_candidate.c:96-126 */
{
    const OPackedCandidate* a = (const OPackedCandidate*)x; const OPackedCandidate* b = (const OPackedCandidate*)y;
    int as = (int)(a->item[0] & ((1u << 29) - 1)), bs = (int)(b->item[0] & ((1u << 29) - 1));
    if (as != bs) return as > bs ? -1 : 1;
    int av = (int)a->item[4], bv = (int)b->item[4];
    if (av != bv) return av < bv ? -1 : 1;
    av = (int)((a->item[0] >> 30) & 1); bv = (int)((b->item[0] >> 30) & 1);
    if (av != bv) return av < bv ? -1 : 1;
    av = (int)a->item[5]; bv = (int)b->item[5];
    if (av != bv) return av < bv ? -1 : 1;
    av = (int)a->item[2]; bv = (int)b->item[2];
    if (av != bv) return av < bv ? -1 : 1;
    return 0;
}

static double cns_weight(double ident)        /* consensus_one_read.c:11-16 */
{
    double e = (100.0 - ident) / 100.0 / 2.0;
    double w = (1.0 - e) * (1.0 - e) + e * e / 3.0;
    if (100.0 - ident <= 1.0e-6) w = 1.0;
    return w;
}

static int full_cov_ovlp(int ql, int qr, int qs, int tl, int tr, int ts, int L, int M)   /* consensus_aux.c:93-112 */
{
    if (ql <= M && qs - qr <= M) return 1;
    if (tl <= M && ts - tr <= M) return 1;
    if (qs - qr <= M) { if (tl > M) return 0; if (qr - ql >= L) return 1; }
    if (ts - tr <= M) { if (ql > M) return 0; if (qr - ql >= L) return 1; }
    return 0;
}

static int mapping_range_ok(int ql, int qr, int qs, int tl, int tr, int ts, int min_size, double ratio)  /* :115-121 */
{
    if (qr - ql >= min_size || tr - tl >= min_size) return 1;
    return (qr - ql >= qs * ratio) || (tr - tl >= ts * ratio);
}

static int good_overlap(int qoff, int qend, int qsize, int soff, int send, int ssize)   /* error_estimate.c:8-30 */
{
    const int M = 200;
    int ql = qoff, qr = qsize - qend, sl = soff, sr = ssize - send;
    return (ql <= M && qr <= M) || (sl <= M && sr <= M) || (qr <= M && sl <= M) || (sr <= M && ql <= M);
}

typedef struct {
    int qid, qoff, qend, qsize, toff, tend, aln_size;
    double ident;
    char *qaln, *taln;
} OPooled;

static int cov_full(const int* cov, int from, int to, int max_cov)   /* consensus_one_read.c:139-145 */
{
    int n = 0;
    for (int i = from; i < to; ++i) if (cov[i] >= max_cov) ++n;
    return n == to - from;
}

static int dbl_gt(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return x > y ? -1 : (x < y);
}

static void emit_record(OCnsRecVec* out, int id, int left, int right, int org, int ncan, int novl, double cutoff,
                        const uint8_t* codes, int n, int is_raw)
{
    OCnsRecord r;
    r.id = id; r.left = left; r.right = right; r.org_size = org; r.num_can = ncan; r.num_ovlps = novl;
    r.ident_cutoff = cutoff; r.seq_len = n; r.is_raw = is_raw;
    r.seq = (char*)malloc((size_t)n + 1);
    for (int i = 0; i < n; ++i) r.seq[i] = "ACGT"[codes[i]];
    r.seq[n] = 0;
    rec_push(out, &r);
}

/*
 * cns_extension, consensus/consensus_aux.c:124-215.  The block-wise extension first; with -r 1, if it failed or left more
 * than 200 bases of the candidate's chain unaligned, the O(nd) local alignment + the global alignment of its range
 * (oracle_rescue.c) -- whose result is taken when both succeed; otherwise the block-wise result if there was one.
 * On success A describes the alignment; *rescued_q / *rescued_t (malloc()ed, to be freed by the caller) hold the strings
 * when they came from the rescue.
 */
static int cns_extension(OAligner* al, ORescue* rs, const OCandidate* c, const uint8_t* query, const uint8_t* target, int T,
                         int min_align_size, int rescue, OAlignment* A, char** rescued_q, char** rescued_t)
{
    *rescued_q = NULL; *rescued_t = NULL;
    const int small = o_onc_align(al, query, (int)c->qoff, (int)c->qsize, target, (int)c->soff, T, 512, min_align_size, 4, A);
    int r = small;
    if (r) {
        const int lhang = A->qoff > (int)c->qbeg ? A->qoff - (int)c->qbeg : 0;
        const int rhang = A->qend < (int)c->qend ? (int)c->qend - A->qend : 0;
        if (lhang + rhang > 200) r = 0;
    }
    if (r) return 1;
    if (rescue) {
        int qb, qe, tb, te, diffs;
        if (o_rescue_local(rs, query, (int)c->qoff, (int)c->qsize, target, (int)c->soff, T, min_align_size, &qb, &qe, &tb, &te, &diffs)) {
            OAlignment R;
            memset(&R, 0, sizeof R);
            if (o_rescue_global(rs, query, qb, qe, target, tb, te, diffs, min_align_size, 4, &R.qoff, &R.qend, &R.toff, &R.tend,
                                &R.ident_perc, &R.qaln, &R.taln, &R.aln_size)) {
                R.ok = 1;
                *A = R; *rescued_q = R.qaln; *rescued_t = R.taln;
                return 1;
            }
        }
    }
    return small;
}

int o_consensus_one_read(const OVolume* reads, const OPackedCandidate* cans_in, size_t n, const OCnsOptions* opt, OCnsRecVec* out)
{
    if ((size_t)opt->min_cov > n) return 0;                              /* :218 */
    OPackedCandidate* cans = (OPackedCandidate*)malloc(sizeof(OPackedCandidate) * n);
    memcpy(cans, cans_in, sizeof(OPackedCandidate) * n);
    qsort(cans, n, sizeof(OPackedCandidate), cand_cns_gt);               /* :257 */
    if (n > 300) n = 300;                                                /* MAX_EXAMINED_CAN */
    const int tid = (int)cans[0].item[1];
    const int T = (int)reads->seq[tid].size;
    uint8_t* target = (uint8_t*)malloc((size_t)T + 1);
    o_extract(reads, (uint64_t)tid, 0, target);
    uint64_t maxq = 0;
    for (size_t i = 0; i < n; ++i) { uint64_t s = reads->seq[cans[i].item[4]].size; if (s > maxq) maxq = s; }
    uint8_t* query = (uint8_t*)malloc(maxq + 1);
    int* cov = (int*)calloc((size_t)T + 1, sizeof(int));
    OAligner* al = o_aligner_new(opt->error);
    ORescue* rs = opt->rescue_long_indels ? o_rescue_new(opt->error) : NULL;
    char *rq = NULL, *rt = NULL;            /* strings of a rescued alignment */
    OTagVec tags = {0, 0, 0};
    /* extended read ids: small linear set (<= 300 entries) */
    int* used = (int*)malloc(sizeof(int) * (n + 1));
    size_t n_used = 0;
#define USED(q) ({ int f__ = 0; for (size_t u__ = 0; u__ < n_used; ++u__) if (used[u__] == (q)) { f__ = 1; break; } f__; })
#define USE(q) do { if (!USED(q)) used[n_used++] = (q); } while (0)

    int num_can = 0, num_ovlps = 0;
    double cutoff;
    size_t next = 0;
    OCandidate c;
    OAlignment A;
    if (opt->use_fixed_ident_cutoff) {                                   /* :276-281 */
        cutoff = 100.0 * (1.0 - opt->error);
    } else {
        /* get_good_overlaps, error_estimate.c:96-184 */
        OPooled pool[50];
        int n_pool = 0, n_ident = 0;
        double ident[15];
        size_t i;
        for (i = 0; i < n && i < 50; ++i) {
            o_unpack_candidate(&c, &cans[i]);
            c.qsize = (int64_t)reads->seq[c.qid].size; c.ssize = T;
            if (USED(c.qid)) continue;
            o_extract(reads, (uint64_t)c.qid, c.qdir, query);
            if (!cns_extension(al, rs, &c, query, target, T, opt->min_align_size, opt->rescue_long_indels, &A, &rq, &rt)) continue;
            OPooled* p = &pool[n_pool++];
            p->qid = c.qid; p->qoff = A.qoff; p->qend = A.qend; p->qsize = (int)c.qsize; p->toff = A.toff; p->tend = A.tend;
            p->ident = A.ident_perc; p->aln_size = A.aln_size;
            p->qaln = (char*)malloc((size_t)A.aln_size + 1); memcpy(p->qaln, A.qaln, (size_t)A.aln_size);
            p->taln = (char*)malloc((size_t)A.aln_size + 1); memcpy(p->taln, A.taln, (size_t)A.aln_size);
            free(rq); free(rt); rq = rt = NULL;
            USE(c.qid);
            if (good_overlap(A.qoff, A.qend, (int)c.qsize, A.toff, A.tend, T)) {
                ident[n_ident++] = A.ident_perc;
                if (n_ident == 15) break;
            }
        }
        next = i;                               /* last_extended_can_id + 1 (:179, consensus_one_read.c:309) */
        if (n_ident < 15) {                     /* get_idents, error_estimate.c:66-93 */
            int m = 0;
            for (int x = 0; x < n_pool; ++x)
                if (good_overlap(pool[x].qoff, pool[x].qend, pool[x].qsize, pool[x].toff, pool[x].tend, T)) { ident[m++] = pool[x].ident; if (m == 15) break; }
            if (m < 15) {
                m = 0;
                for (int x = 0; x < n_pool; ++x)
                    if ((pool[x].qend - pool[x].qoff >= pool[x].qsize * 0.6) || (pool[x].tend - pool[x].toff >= T * 0.6)) { ident[m++] = pool[x].ident; if (m == 15) break; }
            }
            n_ident = m;
        }
        qsort(ident, (size_t)n_ident, sizeof(double), dbl_gt);
        if (n_ident >= 8) n_ident = (int)(n_ident * 0.7);
        if (n_ident < 5) cutoff = 0.0;          /* error_estimate.c:38-43 */
        else {
            double sum = 0.0;
            for (int x = 0; x < n_ident; ++x) sum += ident[x];
            double avg = sum / n_ident, se = 0.0;
            for (int x = 0; x < n_ident; ++x) se += (avg - ident[x]) * (avg - ident[x]);
            se /= n_ident;
            se = sqrt(se);
            cutoff = avg - se * 5;
        }
        /* add_extended_overlaps, consensus_one_read.c:148-183 */
        for (int x = 0; x < n_pool; ++x) {
            OPooled* p = &pool[x];
            if (!(p->ident < cutoff) &&
                mapping_range_ok(p->qoff, p->qend, p->qsize, p->toff, p->tend, T, opt->min_align_size, opt->mapping_ratio)) {
                ++num_ovlps;
                tags_from_alignment(p->qaln, p->taln, (size_t)p->aln_size, p->toff, cns_weight(p->ident), &tags);
                for (int t = p->toff; t < p->tend; ++t) ++cov[t];
                USE(p->qid);
            }
            free(p->qaln); free(p->taln);
        }
        num_can = (int)next;
    }
    /* main loop, consensus_one_read.c:316-372 */
    while (next < n && !cov_full(cov, 0, T, opt->max_cov)) {
        size_t from = next, to = from + 50;
        if (to > n) to = n;
        next = to;
        for (size_t i = from; i < to; ++i) {
            o_unpack_candidate(&c, &cans[i]);
            c.qsize = (int64_t)reads->seq[c.qid].size; c.ssize = T;
            if (USED(c.qid)) continue;
            if (cov_full(cov, (int)c.sbeg, (int)c.send, opt->max_cov)) continue;
            o_extract(reads, (uint64_t)c.qid, c.qdir, query);
            int ok = cns_extension(al, rs, &c, query, target, T, opt->min_align_size, opt->rescue_long_indels, &A, &rq, &rt);
            ++num_can;
            if (ok && !(A.ident_perc < cutoff && !full_cov_ovlp(A.qoff, A.qend, (int)c.qsize, A.toff, A.tend, T, 5000, 100)) &&
                mapping_range_ok(A.qoff, A.qend, (int)c.qsize, A.toff, A.tend, T, opt->min_align_size, opt->mapping_ratio)) {
                ++num_ovlps;
                tags_from_alignment(A.qaln, A.taln, (size_t)A.aln_size, A.toff, cns_weight(A.ident_perc), &tags);
                for (int t = A.toff; t < A.tend; ++t) ++cov[t];
                USE(c.qid);
            }
            free(rq); free(rt); rq = rt = NULL;
        }
    }
    /* consensus, tasc/cbcns.c:108-264 */
    OBackbone bb;
    backbone_build(&tags, T, &bb);
    int corrected = 1;
    if (!opt->full_consensus) {
        /* consensus_broken (:108-163) + raw leftovers (consensus_one_read.c:20-65,92-106) */
        int* iv = (int*)malloc(sizeof(int) * 2 * ((size_t)T / 2 + 2));
        int niv = 0, i = 0;
        while (i < T) {
            while (i < T && bb.coverage[i] < opt->min_cov) ++i;
            int j = i + 1;
            while (j < T && bb.coverage[j] >= opt->min_cov) ++j;
            if (j - i >= opt->min_size * 0.85) {
                uint8_t* s; int len = backbone_segment(&bb, i, j, &s, NULL, NULL);
                if (len >= opt->min_size) {
                    emit_record(out, tid, i, j, T, num_can, num_ovlps, cutoff, s, len, 0);
                    iv[2 * niv] = i; iv[2 * niv + 1] = j; ++niv;
                }
                free(s);
            }
            i = j;
        }
        /* get_raw_intvs */
        int first = 1, open = 1, left = 0;
        const int end_off = T - 1;
        for (int x = 0; x < niv; ++x) {
            int fs = iv[2 * x], fe = iv[2 * x + 1];
            if (first) {
                open = 0; first = 0;
                if (fs > 0) left = 0;
                else { left = fe + 1; continue; }
            }
            int right = fs - 1;
            if (right - left + 1 >= 1000) emit_record(out, tid, left, right + 1, T, num_can, num_ovlps, cutoff, target + left, right + 1 - left, 1);
            if (fe >= end_off) { open = 0; break; }
            left = fe + 1;
        }
        if (open) {
            int right = end_off;
            if (right - left + 1 >= 1000) emit_record(out, tid, left, right + 1, T, num_can, num_ovlps, cutoff, target + left, right + 1 - left, 1);
        }
        free(iv);
    } else {
        /* consensus_unbroken (:172-264) */
        uint8_t* all = (uint8_t*)malloc(1); int nall = 0;
        int* iv = (int*)malloc(sizeof(int) * 4 * ((size_t)T / 2 + 2));
        int niv = 0, i = 0;
        while (i < T) {
            while (i < T && bb.coverage[i] < opt->min_cov) ++i;
            int j = i + 1;
            while (j < T && bb.coverage[j] >= opt->min_cov) ++j;
            if (j - i >= opt->min_size * 0.85) {
                uint8_t* s; int rf, rt;
                int len = backbone_segment(&bb, i, j, &s, &rf, &rt);
                if (len >= opt->min_size) {
                    all = (uint8_t*)realloc(all, (size_t)nall + (size_t)len + 1);
                    memcpy(all + nall, s, (size_t)len);
                    iv[4 * niv] = rf; iv[4 * niv + 1] = rt; iv[4 * niv + 2] = nall; iv[4 * niv + 3] = nall + len; ++niv;
                    nall += len;
                }
                free(s);
            }
            i = j;
        }
        if (niv == 0) corrected = 0;
        else {
            uint8_t* sq = (uint8_t*)malloc((size_t)T + (size_t)nall + 8);
            int ns = 0, last_raw = 0;
            for (int x = 0; x < niv; ++x) {
                int rf = iv[4 * x], rt = iv[4 * x + 1], cf = iv[4 * x + 2], ct = iv[4 * x + 3];
                if (rf > last_raw) { memcpy(sq + ns, target + last_raw, (size_t)(rf - last_raw)); ns += rf - last_raw; }
                memcpy(sq + ns, all + cf, (size_t)(ct - cf)); ns += ct - cf;
                last_raw = rt;
            }
            if (last_raw != T) { memcpy(sq + ns, target + last_raw, (size_t)(T - last_raw)); ns += T - last_raw; }
            emit_record(out, tid, 0, T, T, num_can, num_ovlps, cutoff, sq, ns, 0);
            free(sq);
        }
        free(all); free(iv);
    }
    backbone_free(&bb);
    free(tags.a); free(used); free(cov); free(query); free(target); free(cans);
    o_aligner_free(al);
    o_rescue_free(rs);
    return corrected;
#undef USED
#undef USE
}

/* ------------------------------------------------------------- whole job */

static int pcan_sid_lt(const void* x, const void* y)
{
    uint32_t a = ((const OPackedCandidate*)x)->item[1], b = ((const OPackedCandidate*)y)->item[1];
    return a < b ? -1 : (a > b);
}

int o_cns_job(const OVolume* reads, const OPackedCandidate* cans, size_t n, const OCnsOptions* opt, OCnsRecVec* out)
{
    const int batch = o_cns_partition_size;                     /* oc2pcan -p, default pcan_options.c:10-14 */
    const int nreads = (int)reads->nseq;
    const int nparts = (nreads + batch - 1) / batch;
    for (int p = 0; p < nparts; ++p) {
        const uint32_t lo = (uint32_t)p * batch, hi = lo + batch;
        /* oc2pcan (pcan.c:58-70): every record under its subject and, swapped, under its query */
        size_t m = 0, cap = 1024;
        OPackedCandidate* part = (OPackedCandidate*)malloc(sizeof(OPackedCandidate) * cap);
        for (size_t i = 0; i < n; ++i) {
            uint32_t sid = cans[i].item[1], qid = cans[i].item[4];
            if (m + 2 > cap) { cap *= 2; part = (OPackedCandidate*)realloc(part, sizeof(OPackedCandidate) * cap); }
            if (sid >= lo && sid < hi) part[m++] = cans[i];
            if (qid >= lo && qid < hi) o_swap_roles(&cans[i], &part[m++]);
        }
        if (m == 0) { free(part); continue; }
        qsort(part, m, sizeof(OPackedCandidate), pcan_sid_lt);              /* consensus_one_partition.c:41 */
        const int min_id = (int)part[0].item[1], max_id = (int)part[m - 1].item[1];
        for (size_t i = 0; i < m; ++i)                                      /* :55-72 */
            o_normalise_sdir(&part[i], (uint32_t)reads->seq[part[i].item[4]].size, (uint32_t)reads->seq[part[i].item[1]].size);
        uint8_t* corrected = (uint8_t*)calloc((size_t)nreads + 1, 1);
        for (size_t i = 0; i < m;) {
            size_t j = i + 1;
            while (j < m && part[j].item[1] == part[i].item[1]) ++j;
            if (o_consensus_one_read(reads, part + i, j - i, opt, out)) corrected[part[i].item[1]] = 1;
            i = j;
        }
        /* never-corrected reads of [min_id, max_id) go to raw_out whole (:173-194) */
        for (int id = min_id; id < max_id; ++id) {
            if (corrected[id]) continue;
            int L = (int)reads->seq[id].size;
            uint8_t* s = (uint8_t*)malloc((size_t)L + 1);
            o_extract(reads, (uint64_t)id, 0, s);
            emit_record(out, id, 0, L, L, 0, 0, 0.0, s, L, 1);
            free(s);
        }
        free(corrected);
        free(part);
    }
    return 0;
}

size_t o_format_cns_record(const OVolume* reads, const OCnsRecord* r, char* buf, size_t cap)
{
    const char* name = reads->hdr + reads->seq[r->id].hdr_offset;
    int n;
    if (strncmp(name, "ontsa_id", 8) == 0) n = snprintf(buf, cap, ">%s_", name);      /* cns_seq.h:25-29 */
    else n = snprintf(buf, cap, ">ontsa_id_%d_", r->id);
    n += snprintf(buf + n, cap - (size_t)n, "(%d_%d_%d_%d_%d_%d_%lf)\n", r->left, r->right, r->seq_len, r->org_size,
                  r->num_can, r->num_ovlps, r->ident_cutoff);
    if ((size_t)n + (size_t)r->seq_len + 2 > cap) return 0;
    memcpy(buf + n, r->seq, (size_t)r->seq_len);
    n += r->seq_len;
    buf[n++] = '\n';
    return (size_t)n;
}

/* oracle_cli cns [options] wrk_dir candidates cns_out raw_out */
int run_cns(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: cns [options] wrk_dir candidates cns_out raw_out\n"); return 1; }
    OCnsOptions o;
    o_cns_options_default(&o);
    int c;
    optind = 1;
    while ((c = getopt(argc - 4, argv, "a:x:y:l:f:e:p:t:r:u:s:M:P:")) != -1) {
        switch (c) {
        case 'a': o.min_align_size = atoi(optarg); break;
        case 'x': o.min_cov = atoi(optarg); break;
        case 'y': o.max_cov = atoi(optarg); break;
        case 'l': o.min_size = atoi(optarg); break;
        case 'f': o.full_consensus = atoi(optarg); break;
        case 'e': o.error = atof(optarg); break;
        case 'p': o.mapping_ratio = atof(optarg); break;
        case 't': o.num_threads = atoi(optarg); break;
        case 'r': o.rescue_long_indels = atoi(optarg); break;
        case 'u': o.use_fixed_ident_cutoff = atoi(optarg); break;
        case 's': o.small_memory = atoi(optarg); break;
        case 'M': o_cns_tag_order_mode = atoi(optarg); break;     /* oracle only: 1 = klib introsort emulation */
        case 'P': o_cns_partition_size = atoi(optarg); break;     /* oracle only: what oc2pcan -p was given */
        default: return 1;
        }
    }
    const char* wrk = argv[argc - 4];
    const char* canp = argv[argc - 3];
    OVolumesInfo vi;
    if (o_volumes_info_load(wrk, &vi)) { fprintf(stderr, "cannot load volume info\n"); return 1; }
    OVolume reads;
    if (o_volume_load(vi.names[0], &reads)) { fprintf(stderr, "cannot load %s\n", vi.names[0]); return 1; }
    for (int v = 1; v < vi.num_volumes; ++v) {
        OVolume x;
        if (o_volume_load(vi.names[v], &x)) { fprintf(stderr, "cannot load %s\n", vi.names[v]); return 1; }
        o_volume_append(&reads, &x);
        o_volume_free(&x);
    }
    FILE* f = fopen(canp, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", canp); return 1; }
    fseek(f, 0, SEEK_END);
    size_t bytes = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    size_t n = bytes / sizeof(OPackedCandidate);
    OPackedCandidate* cans = (OPackedCandidate*)malloc(sizeof(OPackedCandidate) * (n ? n : 1));
    if (fread(cans, sizeof(OPackedCandidate), n, f) != n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    OCnsRecVec recs = {0, 0, 0};
    o_cns_job(&reads, cans, n, &o, &recs);
    FILE* fc = fopen(argv[argc - 2], "w");
    FILE* fr = fopen(argv[argc - 1], "w");
    if (!fc || !fr) { fprintf(stderr, "cannot open outputs\n"); return 1; }
    size_t cap = 1 << 20;
    char* buf = (char*)malloc(cap);
    for (size_t i = 0; i < recs.n; ++i) {
        if ((size_t)recs.a[i].seq_len + 4096 > cap) { cap = (size_t)recs.a[i].seq_len * 2 + 4096; buf = (char*)realloc(buf, cap); }
        size_t w = o_format_cns_record(&reads, &recs.a[i], buf, cap);
        fwrite(buf, 1, w, recs.a[i].is_raw ? fr : fc);
    }
    fclose(fc); fclose(fr);
    free(buf);
    o_cns_recvec_free(&recs);
    free(cans);
    o_volume_free(&reads);
    o_volumes_info_free(&vi);
    return 0;
}
