/*
 * oracle_pm.c -- CPU restatement of the oc2pmov half (volumes, k-mer index,
 * seed collection / block voting, chaining, candidate emission).
 * TEST INFRASTRUCTURE ONLY -- see necat_oracle.h.
 */
#include "necat_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================ volumes */

static const char kPacMagic[] = "ontcns_pac_header_hofuwhogfuewo"; /* packed_db.c:7 */

int o_volume_load(const char* path, OVolume* v)
{
    memset(v, 0, sizeof(*v));
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char magic[64];
    size_t ml = strlen(kPacMagic);
    if (fread(magic, 1, ml, f) != ml || memcmp(magic, kPacMagic, ml) != 0) { fclose(f); return -2; }
    if (fread(&v->nseq, 8, 1, f) != 1) goto bad;
    if (fread(&v->db_size, 8, 1, f) != 1) goto bad;
    v->seq = (OSeqInfo*)malloc(sizeof(OSeqInfo) * (v->nseq ? v->nseq : 1));
    if (fread(v->seq, sizeof(OSeqInfo), v->nseq, f) != v->nseq) goto bad;
    if (fread(&v->hdr_size, 8, 1, f) != 1) goto bad;
    v->hdr = (char*)malloc(v->hdr_size + 1);
    if (v->hdr_size && fread(v->hdr, 1, v->hdr_size, f) != v->hdr_size) goto bad;
    {
        uint64_t nb = (v->db_size + 3) >> 2;
        v->pac = (uint8_t*)calloc(nb + 16, 1);
        if (fread(v->pac, 1, nb, f) != nb) goto bad;
    }
    fclose(f);
    return 0;
bad:
    fclose(f);
    o_volume_free(v);
    return -3;
}

void o_volume_free(OVolume* v)
{
    free(v->pac); free(v->seq); free(v->hdr);
    memset(v, 0, sizeof(*v));
}

static void set_base(uint8_t* pac, uint64_t i, int c)
{
    pac[i >> 2] |= (uint8_t)(c << (((~i) & 3) << 1));
}

int o_volume_append(OVolume* dst, const OVolume* src)
{
    uint64_t base = dst->db_size, hbase = dst->hdr_size;
    dst->seq = (OSeqInfo*)realloc(dst->seq, sizeof(OSeqInfo) * (dst->nseq + src->nseq + 1));
    for (uint64_t i = 0; i < src->nseq; ++i) {
        OSeqInfo s = src->seq[i];
        s.offset += base;
        s.hdr_offset += hbase;
        dst->seq[dst->nseq + i] = s;
    }
    dst->hdr = (char*)realloc(dst->hdr, dst->hdr_size + src->hdr_size + 1);
    memcpy(dst->hdr + hbase, src->hdr, src->hdr_size);
    dst->hdr_size += src->hdr_size;
    uint64_t nb = (base + src->db_size + 3) >> 2;
    uint8_t* np = (uint8_t*)calloc(nb + 16, 1);
    memcpy(np, dst->pac, (base + 3) >> 2);
    for (uint64_t i = 0; i < src->db_size; ++i) set_base(np, base + i, o_base(src, i));
    free(dst->pac);
    dst->pac = np;
    dst->db_size = base + src->db_size;
    dst->nseq += src->nseq;
    return 0;
}

void o_extract(const OVolume* v, uint64_t id, int strand, uint8_t* out)
{
    uint64_t s = v->seq[id].offset, n = v->seq[id].size;
    if (strand == 0) {
        for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)o_base(v, s + i);
    } else {
        for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)(3 - o_base(v, s + n - 1 - i));
    }
}

uint64_t o_offset_to_id(const OVolume* v, uint64_t offset)
{
    /* last sequence whose start is <= offset */
    uint64_t lo = 0, hi = v->nseq;
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (v->seq[mid].offset <= offset) lo = mid; else hi = mid;
    }
    return lo;
}

int o_volumes_info_load(const char* wrk_dir, OVolumesInfo* vi)
{
    memset(vi, 0, sizeof(*vi));
    char path[4096];
    size_t wl = strlen(wrk_dir);
    const char* sep = (wl && wrk_dir[wl - 1] == '/') ? "" : "/";
    snprintf(path, sizeof path, "%s%sreads_info.txt", wrk_dir, sep);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    if (fscanf(f, "%d%d", &vi->num_volumes, &vi->num_reads) != 2) { fclose(f); return -2; }
    fclose(f);
    snprintf(path, sizeof path, "%s%svolume_names.txt", wrk_dir, sep);
    f = fopen(path, "r");
    if (!f) return -3;
    vi->names = (char**)calloc(vi->num_volumes, sizeof(char*));
    vi->start_id = (int*)calloc(vi->num_volumes, sizeof(int));
    vi->count = (int*)calloc(vi->num_volumes, sizeof(int));
    char line[4096];
    for (int i = 0; i < vi->num_volumes; ++i) {
        if (!fgets(line, sizeof line, f)) { fclose(f); return -4; }
        char name[4096];
        if (sscanf(line, "%4095s %d %d", name, &vi->start_id[i], &vi->count[i]) != 3) { fclose(f); return -5; }
        vi->names[i] = strdup(name);
    }
    fclose(f);
    return 0;
}

void o_volumes_info_free(OVolumesInfo* vi)
{
    for (int i = 0; i < vi->num_volumes; ++i) free(vi->names ? vi->names[i] : NULL);
    free(vi->names); free(vi->start_id); free(vi->count);
    memset(vi, 0, sizeof(*vi));
}

/* ================================================================== index */

/* Stable LSD radix sort of u64 by bits [lo, hi). */
static void radix_sort_u64(uint64_t* a, uint64_t n, int lo, int hi)
{
    uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
    uint64_t* src = a; uint64_t* dst = tmp;
    for (int sh = lo; sh < hi; sh += 16) {
        int bits = hi - sh < 16 ? hi - sh : 16;
        uint64_t mask = (1ull << bits) - 1;
        uint64_t* cnt = (uint64_t*)calloc((size_t)mask + 2, sizeof(uint64_t));
        for (uint64_t i = 0; i < n; ++i) ++cnt[((src[i] >> sh) & mask) + 1];
        for (uint64_t d = 0; d < mask + 1; ++d) cnt[d + 1] += cnt[d];
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & mask]++] = src[i];
        free(cnt);
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, sizeof(uint64_t) * n);
    free(tmp);
}

/*
 * lookup_table.c:15-124: every position of every read gets the hash of the
 * k-mer starting there (rolling hash restarted per read, :30-40); k-mers that
 * occur more than q times are dropped (:44-47); the kept (hash, position) pairs
 * end up grouped by hash with ascending positions inside each group (scan order
 * + stable sort, hash_list_bucket_sort.c:134).
 */
int o_index_build(const OVolume* v, int k, int q, OIndex* idx)
{
    memset(idx, 0, sizeof(*idx));
    idx->k = k;
    uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t cap = 0;
    for (uint64_t i = 0; i < v->nseq; ++i)
        if (v->seq[i].size >= (uint64_t)k) cap += v->seq[i].size - k + 1;
    uint64_t* pairs = (uint64_t*)malloc(sizeof(uint64_t) * (cap ? cap : 1));
    uint64_t n = 0;
    for (uint64_t i = 0; i < v->nseq; ++i) {
        uint64_t off = v->seq[i].offset, sz = v->seq[i].size;
        if (sz < (uint64_t)k) continue;
        uint64_t h = 0;
        for (uint64_t j = 0; j < sz; ++j) {
            h = ((h << 2) | (uint64_t)o_base(v, off + j)) & mask;
            if (j + 1 >= (uint64_t)k) pairs[n++] = (h << 34) | (off + j + 1 - k);
        }
    }
    radix_sort_u64(pairs, n, 34, 34 + 2 * k);
    /* run-length: count distinct kept hashes and kept positions */
    uint64_t nk = 0, no = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && (pairs[j] >> 34) == (pairs[i] >> 34)) ++j;
        if (j - i <= (uint64_t)q) { ++nk; no += j - i; }
        i = j;
    }
    idx->n_keys = nk; idx->n_off = no;
    idx->key = (uint32_t*)malloc(sizeof(uint32_t) * (nk ? nk : 1));
    idx->start = (uint64_t*)malloc(sizeof(uint64_t) * (nk + 1));
    idx->off = (uint64_t*)malloc(sizeof(uint64_t) * (no ? no : 1));
    nk = 0; no = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && (pairs[j] >> 34) == (pairs[i] >> 34)) ++j;
        if (j - i <= (uint64_t)q) {
            idx->key[nk] = (uint32_t)(pairs[i] >> 34);
            idx->start[nk] = no;
            for (uint64_t t = i; t < j; ++t) idx->off[no++] = pairs[t] & ((1ull << 34) - 1);
            ++nk;
        }
        i = j;
    }
    idx->start[nk] = no;
    free(pairs);
    return 0;
}

void o_index_free(OIndex* idx)
{
    free(idx->key); free(idx->start); free(idx->off);
    memset(idx, 0, sizeof(*idx));
}

const uint64_t* o_index_lookup(const OIndex* idx, uint64_t hash, uint64_t* n)
{
    uint64_t lo = 0, hi = idx->n_keys;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (idx->key[mid] < hash) lo = mid + 1; else hi = mid;
    }
    if (lo < idx->n_keys && idx->key[lo] == hash) {
        *n = idx->start[lo + 1] - idx->start[lo];
        return idx->off + idx->start[lo];
    }
    *n = 0;
    return NULL;
}

/* ============================================================= candidates */

void o_map_options_default(OMapOptions* o)
{
    o->kmer_size = 15; o->scan_window = 10; o->kmer_cnt_cutoff = 500; o->block_size = 2000;
    o->block_score_cutoff = 3; o->num_candidates = 500; o->align_size_cutoff = 500;
    o->ddfs_cutoff = 0.25; o->error = 0.5; o->num_output = 500; o->num_threads = 1;
    o->job = 1; o->binary_output = 0; o->use_hdr_as_id = 1;
}

void o_pack_candidate(const OCandidate* c, OPackedCandidate* p)
{
    uint32_t w = 0;
    if (c->sdir == 1) w |= 1u << 31;
    if (c->qdir == 1) w |= 1u << 30;
    if (c->qoff == c->qbeg) w |= 1u << 29;
    w |= (uint32_t)(c->score > 1000000 ? 1000000 : c->score);
    p->item[0] = w;
    p->item[1] = (uint32_t)c->sid;  p->item[2] = (uint32_t)c->sbeg; p->item[3] = (uint32_t)c->send;
    p->item[4] = (uint32_t)c->qid;  p->item[5] = (uint32_t)c->qbeg; p->item[6] = (uint32_t)c->qend;
}

void o_unpack_candidate(OCandidate* c, const OPackedCandidate* p)
{
    uint32_t w = p->item[0];
    c->sdir = (int)(w >> 31); c->qdir = (int)((w >> 30) & 1); c->score = (int)(w & ((1u << 29) - 1));
    c->sid = (int)p->item[1]; c->sbeg = p->item[2]; c->send = p->item[3];
    c->qid = (int)p->item[4]; c->qbeg = p->item[5]; c->qend = p->item[6];
    if (w & (1u << 29)) { c->qoff = c->qbeg; c->soff = c->sbeg; }
    else                { c->qoff = c->qend; c->soff = c->send; }
    c->qsize = c->ssize = 0;
}

void o_swap_roles(const OPackedCandidate* s, OPackedCandidate* d)
{
    uint32_t w = s->item[0], r = w & ((1u << 30) - 1);      /* keeps off flag + score */
    if (w >> 31) r |= 1u << 30;
    if ((w >> 30) & 1) r |= 1u << 31;
    OPackedCandidate t;
    t.item[0] = r;
    t.item[1] = s->item[4]; t.item[2] = s->item[5]; t.item[3] = s->item[6];
    t.item[4] = s->item[1]; t.item[5] = s->item[2]; t.item[6] = s->item[3];
    *d = t;
}

void o_normalise_sdir(OPackedCandidate* p, uint32_t qsize, uint32_t ssize)
{
    uint32_t w = p->item[0];
    if (!(w >> 31)) return;
    uint32_t r = w & ((1u << 29) - 1);
    if (!((w >> 30) & 1)) r |= 1u << 30;    /* query direction inverted */
    if (!(w & (1u << 29))) r |= 1u << 29;   /* off flag toggled */
    p->item[0] = r;
    uint32_t qb = qsize - p->item[6], qe = qsize - p->item[5];
    p->item[5] = qb; p->item[6] = qe;
    uint32_t sb = ssize - p->item[3], se = ssize - p->item[2];
    p->item[2] = sb; p->item[3] = se;
}

static void cand_push(OCandVec* v, const OCandidate* c)
{
    if (v->n == v->m) {
        v->m = v->m ? v->m * 2 : 64;
        v->a = (OCandidate*)realloc(v->a, sizeof(OCandidate) * v->m);
    }
    v->a[v->n++] = *c;
}

#define O_BLK_SEEDS 40      /* word_finder_aux.h:9 */

/* word_finder_aux.h:16-22 */
typedef struct {
    short score;
    short off[O_BLK_SEEDS];
    int   kid[O_BLK_SEEDS];     /* 1-based probe number */
    int   last_kid;
    int   slot;                 /* position in the first-touch list, -1 if untouched */
} OBlock;

typedef struct { int score; int blk; } OTouched;
typedef struct { int64_t qoff, soff; } OSeed;

struct OFinder {
    OBlock*   store;    /* store[0] is the sentinel in front of block 0 (word_finder.c:22-23) */
    OBlock*   blk;
    OTouched* touched;
    int       n_touched;
    int64_t   nblk;
    OSeed*    seeds; size_t n_seeds, m_seeds;
    int *f, *p, *t, *v; size_t m_dp;
    OCandidate* chains; size_t n_chains, m_chains;
    int       kmer_size, min_cnt;
};

OFinder* o_finder_new(const OVolume* ref, const OMapOptions* opt)
{
    OFinder* f = (OFinder*)calloc(1, sizeof(OFinder));
    f->nblk = (int64_t)(ref->db_size / (uint64_t)opt->block_size) + 5;
    f->store = (OBlock*)malloc(sizeof(OBlock) * (size_t)(f->nblk + 1));
    f->blk = f->store + 1;
    f->touched = (OTouched*)malloc(sizeof(OTouched) * (size_t)f->nblk);
    for (int64_t i = -1; i < f->nblk; ++i) { f->blk[i].score = 0; f->blk[i].slot = -1; f->blk[i].last_kid = -1; }
    f->kmer_size = opt->kmer_size;
    f->min_cnt = opt->block_score_cutoff;      /* chain_dp.c:177 */
    return f;
}

void o_finder_free(OFinder* f)
{
    if (!f) return;
    free(f->store); free(f->touched); free(f->seeds);
    free(f->f); free(f->p); free(f->t); free(f->v); free(f->chains);
    free(f);
}

static void finder_reset(OFinder* f)       /* word_finder.c:40-51 */
{
    for (int i = 0; i < f->n_touched; ++i) {
        OBlock* b = f->blk + f->touched[i].blk;
        b->slot = -1; b->score = 0; b->last_kid = -1;
    }
    f->n_touched = 0;
}

/* word_finder.c:85-104 */
static void add_hit(OFinder* f, int probe, int64_t blk, short off)
{
    OBlock* b = f->blk + blk;
    if (b->last_kid >= probe + 1) return;       /* one seed per probe per block */
    if (b->score >= O_BLK_SEEDS) return;
    int s = b->score++;
    b->off[s] = off;
    b->kid[s] = probe + 1;
    b->last_kid = probe + 1;
    if (b->slot < 0) {
        b->slot = f->n_touched++;
        f->touched[b->slot].blk = (int)blk;
    }
    f->touched[b->slot].score = b->score + (b - 1)->score;  /* as of this insertion */
}

/* word_finder.c:107-138 (+ :66-82) */
static void collect_seeds(OFinder* f, const uint8_t* read, int read_size, int read_id,
                          int read_start_id, int ref_start_id, const OVolume* ref,
                          const OIndex* idx, const OMapOptions* opt)
{
    uint64_t soff_max = ~0ull;
    int gid = read_id + read_start_id;
    if (gid >= ref_start_id && gid < ref_start_id + (int)ref->nseq) soff_max = ref->seq[read_id].offset;
    const int k = opt->kmer_size, z = opt->scan_window, bs = opt->block_size;
    int probe = 0;
    for (int i = 0; i + k <= read_size; i += z, ++probe) {
        uint64_t h = 0;
        for (int j = 0; j < k; ++j) h = (h << 2) | read[i + j];
        uint64_t n;
        const uint64_t* list = o_index_lookup(idx, h, &n);
        for (uint64_t t = 0; t < n; ++t) {
            if (list[t] >= soff_max) continue;
            add_hit(f, probe, (int64_t)(list[t] / (uint64_t)bs), (short)(list[t] % (uint64_t)bs));
        }
    }
}

/* The DDF test of word_finder.c:145: float division, double comparison. */
static int ddf_close(int dloc, int dkid, float window)
{
    float ratio = (float)dloc / ((float)dkid * window);
    return fabs((double)ratio - 1.0) < 0.25;
}

/*
 * word_finder.c:141-168.  Pairwise voting between the pooled seeds of two
 * neighbouring blocks; returns 0 if no seed collects >= 5 votes, else the
 * representative seed (location, 1-based probe number, index in the pool).
 */
static int vote_seeds(const int* loc, const int* kid, int* votes, int n, float window,
                      int read_size, int* rep_loc, int* rep_kid, int* rep_idx)
{
    for (int i = 0; i < n; ++i) votes[i] = 0;
    for (int i = 0; i + 1 < n; ++i) {
        int last = kid[i];
        for (int j = i + 1; j < n; ++j) {
            if (last == kid[j]) continue;
            int dk = kid[j] - kid[i], dl = loc[j] - loc[i];
            if (dk > 0 && dl > 0 && dl < read_size && ddf_close(dl, dk, window)) {
                ++votes[i]; ++votes[j];
                last = kid[j];
            }
        }
    }
    int best = 0, besti = 0, ties = 0, lasttie = 0;
    for (int i = 0; i < n; ++i) {
        if (votes[i] > best) { best = votes[i]; besti = i; ties = 0; }
        else if (votes[i] == best) { ++ties; lasttie = i; }
    }
    (void)lasttie;
    if (best < 5) return 0;
    if (ties == best) {         /* :152: number of ties compared with the vote count */
        *rep_loc = loc[besti]; *rep_kid = kid[besti]; *rep_idx = besti;
        return 1;
    }
    /* first seed compatible with the winner; location 0 counts as "unset" (:155,159,162) */
    int r_loc = 0, r_kid = 0, r_idx = -1;
    for (int j = 0; j < besti; ++j) {
        int dk = kid[besti] - kid[j], dl = loc[besti] - loc[j];
        if (dk > 0 && dl > 0 && dl < read_size && ddf_close(dl, dk, window)) {
            if (r_loc == 0) { r_loc = loc[j]; r_kid = kid[j]; r_idx = j; }
        }
    }
    if (r_loc == 0) { r_loc = loc[besti]; r_kid = kid[besti]; r_idx = besti; }
    for (int j = besti + 1; j < n; ++j) {
        int dk = kid[j] - kid[besti], dl = loc[j] - loc[besti];
        if (dk > 0 && dl > 0 && dl <= read_size && ddf_close(dl, dk, window)) {
            if (r_loc == 0) { r_loc = loc[j]; r_kid = kid[j]; r_idx = j; }
        }
    }
    *rep_loc = r_loc; *rep_kid = r_kid; *rep_idx = r_idx;
    return 1;
}

static int ilog2_u32(uint32_t v) { int r = -1; while (v) { ++r; v >>= 1; } return r; }  /* chain_dp.c:19-24 */

static int seed_lt(const void* a, const void* b)        /* word_finder_aux.h:16 */
{
    const OSeed* x = (const OSeed*)a; const OSeed* y = (const OSeed*)b;
    if (x->soff != y->soff) return x->soff < y->soff ? -1 : 1;
    if (x->qoff != y->qoff) return x->qoff < y->qoff ? -1 : 1;
    return 0;
}

typedef struct { int score, end; } OChainEnd;
static int chain_end_gt(const void* a, const void* b)   /* chain_dp.c:8 */
{
    const OChainEnd* x = (const OChainEnd*)a; const OChainEnd* y = (const OChainEnd*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return 0;
}

static int chain_gt(const void* a, const void* b)       /* chain_dp.c:26-33 */
{
    const OCandidate* x = (const OCandidate*)a; const OCandidate* y = (const OCandidate*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->qoff != y->qoff) return x->qoff < y->qoff ? -1 : 1;
    if (x->soff != y->soff) return x->soff < y->soff ? -1 : 1;
    return 0;
}

/*
 * chain_dp.c:37-159: minimap-style chaining over seeds sorted by (soff, qoff).
 * Fills f->chains (sorted best first).
 */
static void chain_seeds(OFinder* F, int qid, int qdir, int64_t qsize, int sid, int64_t ssize)
{
    const int n = (int)F->n_seeds;
    const OSeed* s = F->seeds;
    const int k = F->kmer_size, max_dist = 5000, bw = 500, max_skip = 25, min_sc = 30;
    const int min_cnt = F->min_cnt;
    F->n_chains = 0;
    if ((size_t)n > F->m_dp) {
        F->m_dp = (size_t)n * 2;
        F->f = (int*)realloc(F->f, sizeof(int) * F->m_dp); F->p = (int*)realloc(F->p, sizeof(int) * F->m_dp);
        F->t = (int*)realloc(F->t, sizeof(int) * F->m_dp); F->v = (int*)realloc(F->v, sizeof(int) * F->m_dp);
    }
    int *f = F->f, *p = F->p, *t = F->t, *v = F->v;
    for (int i = 0; i < n; ++i) { f[i] = 0; p[i] = -1; t[i] = 0; v[i] = 0; }
    int st = 0;
    for (int i = 0; i < n; ++i) {
        int64_t ri = s[i].soff, qi = s[i].qoff;
        int best_j = -1, best_f = k, n_skip = 0;
        while (st < i && ri - s[st].soff > max_dist) ++st;
        for (int j = i - 1; j >= st; --j) {
            if (ri <= s[j].soff || qi <= s[j].qoff || qi - s[j].qoff > max_dist) continue;
            int64_t dr = ri - s[j].soff, dq = qi - s[j].qoff;
            int64_t dd = dr > dq ? dr - dq : dq - dr;
            if (dd > bw) continue;
            int64_t md = dq < dr ? dq : dr;
            int sc = (int)(md < k ? md : k);
            int lg = dd ? ilog2_u32((uint32_t)dd) : 0;
            sc -= (int)((double)dd * 0.01 * k) + (lg >> 1);
            sc += f[j];
            if (sc > best_f) {
                best_f = sc; best_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = best_f; p[i] = best_j;
        v[i] = (best_j >= 0 && v[best_j] > best_f) ? v[best_j] : best_f;
    }
    /* chain ends: seeds nobody links to, with peak score >= min_sc */
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    int n_u = 0;
    for (int i = 0; i < n; ++i) if (!t[i] && v[i] >= min_sc) ++n_u;
    if (!n_u) return;
    OChainEnd* u = (OChainEnd*)malloc(sizeof(OChainEnd) * (size_t)n_u);
    n_u = 0;
    for (int i = 0; i < n; ++i) {
        if (t[i] || v[i] < min_sc) continue;
        int j = i;
        while (j >= 0 && f[j] < v[j]) j = p[j];     /* walk back to the peak */
        if (j < 0) j = i;
        u[n_u].score = f[j]; u[n_u].end = j; ++n_u;
    }
    qsort(u, (size_t)n_u, sizeof(OChainEnd), chain_end_gt);

    for (int i = 0; i < n; ++i) t[i] = 0;
    OCandidate c;
    memset(&c, 0, sizeof c);
    c.qid = qid; c.qdir = qdir; c.qsize = qsize; c.sid = sid; c.sdir = 0; c.ssize = ssize;
    int n_v = 0, kk = 0;
    for (int i = 0; i < n_u; ++i) {
        int n_v0 = n_v, k0 = kk;
        int j = u[i].end;
        c.qend = s[j].qoff + k; c.send = s[j].soff + k;
        c.qoff = c.qend; c.soff = c.send;
        int first;
        do { first = j; ++n_v; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
        int emit = 0, score = 0;
        if (j < 0) { emit = 1; score = u[i].score; }
        else if (u[i].score - f[j] >= min_sc) { emit = 1; score = u[i].score - f[j]; }
        if (emit && n_v - n_v0 >= min_cnt) {
            c.qbeg = s[first].qoff; c.sbeg = s[first].soff; c.score = score;
            if (F->n_chains == F->m_chains) {
                F->m_chains = F->m_chains ? F->m_chains * 2 : 16;
                F->chains = (OCandidate*)realloc(F->chains, sizeof(OCandidate) * F->m_chains);
            }
            F->chains[F->n_chains++] = c;
            ++kk;
        }
        if (k0 == kk) n_v = n_v0;
    }
    free(u);
    if (F->n_chains > 1) qsort(F->chains, F->n_chains, sizeof(OCandidate), chain_gt);
}

static void seed_push(OFinder* f, int64_t q, int64_t s)
{
    if (f->n_seeds == f->m_seeds) {
        f->m_seeds = f->m_seeds ? f->m_seeds * 2 : 256;
        f->seeds = (OSeed*)realloc(f->seeds, sizeof(OSeed) * f->m_seeds);
    }
    f->seeds[f->n_seeds].qoff = q; f->seeds[f->n_seeds].soff = s; ++f->n_seeds;
}

static void zero_span(OFinder* f, const OCandidate* c, int64_t sstart, int64_t bs)  /* word_finder.c:171-181 */
{
    for (int64_t b = (c->sbeg + sstart) / bs; b <= (c->send + sstart) / bs; ++b) f->blk[b].score = 0;
}

static int contains_seed(const OCandidate* c, int64_t q, int64_t s)
{
    return q >= c->qbeg && q < c->qend && s >= c->sbeg && s < c->send;
}

/* word_finder.c:184-359 */
static int try_block(OFinder* f, int block_id, const OVolume* ref, const OMapOptions* opt,
                     int qid, int qdir, int64_t qsize, OCandVec* out)
{
    const int64_t bs = opt->block_size;
    const int z = opt->scan_window;
    const int cutoff = 2 * opt->block_score_cutoff;
    int kid[2 * O_BLK_SEEDS], loc[2 * O_BLK_SEEDS], votes[2 * O_BLK_SEEDS];
    int n = 0, shift = 0;
    int64_t pool_start = bs * block_id;
    const OBlock* pb = f->blk + block_id - 1;
    if (pb->score) {
        for (int i = 0; i < pb->score; ++i) { kid[n] = pb->kid[i]; loc[n] = pb->off[i]; ++n; }
        shift = (int)bs;
        pool_start = bs * (block_id - 1);
    }
    const OBlock* cb = f->blk + block_id;
    for (int i = 0; i < cb->score; ++i) { kid[n] = cb->kid[i]; loc[n] = cb->off[i] + shift; ++n; }

    int r_loc, r_kid, r_idx;
    if (!vote_seeds(loc, kid, votes, n, (float)z, (int)qsize, &r_loc, &r_kid, &r_idx)) return 0;
    if (votes[r_idx] < cutoff) return 0;

    int64_t seed_t = r_loc + pool_start;
    int64_t seed_q = (int64_t)(r_kid - 1) * z;
    int seed_b = (int)(seed_t / bs);
    int64_t tid = (int64_t)o_offset_to_id(ref, (uint64_t)seed_t);
    int64_t tsize = (int64_t)ref->seq[tid].size, tstart = (int64_t)ref->seq[tid].offset, tend = tstart + tsize;
    seed_t -= tstart;
    int64_t L = seed_t < seed_q ? seed_t : seed_q;
    int b_lo = (int)(seed_b - L / bs - 1);
    if (b_lo < 0) b_lo = 0;
    int64_t tr = tsize - seed_t, qr = qsize - seed_q;
    L = tr < qr ? tr : qr;
    int b_hi = (int)(seed_b + (L + bs - 1) / bs);

    f->n_seeds = 0;
    int support = 0;
    for (int b = b_lo; b <= seed_b; ++b) {                 /* :249-276 */
        OBlock* B = f->blk + b;
        if (!B->score) continue;
        int rel = 0;
        for (int s = 0; s < B->score; ++s) {
            int64_t to = bs * b + B->off[s];
            int64_t qo = (int64_t)(B->kid[s] - 1) * z;
            if (to < tstart) continue;
            to -= tstart;
            if (to < seed_t && qo < seed_q) {
                double d = 1.0 * (double)(seed_t - to) / (double)(seed_q - qo) - 1.0;
                if (!(fabs(d) < 0.25)) continue;
                ++rel;
                seed_push(f, qo, to);
            }
        }
        if (b != seed_b && 1.0 * rel / B->score >= 0.4) B->score = 0;
        support += rel;
    }
    seed_push(f, seed_q, seed_t);
    for (int b = seed_b; b <= b_hi; ++b) {                 /* :282-307 */
        OBlock* B = f->blk + b;
        if (!B->score) continue;
        int rel = 0;
        for (int s = 0; s < B->score; ++s) {
            int64_t to = bs * b + B->off[s];
            int64_t qo = (int64_t)(B->kid[s] - 1) * z;
            if (to >= tend) continue;
            to -= tstart;       /* may be negative for a block straddling the read start: never relevant */
            if (to > seed_t && qo > seed_q) {
                double d = 1.0 * (double)(to - seed_t) / (double)(qo - seed_q) - 1.0;
                if (!(fabs(d) < 0.25)) continue;
                ++rel;
                seed_push(f, qo, to);
            }
        }
        if (b != seed_b && 1.0 * rel / B->score >= 0.4) B->score = 0;
        support += rel;
    }

    qsort(f->seeds, f->n_seeds, sizeof(OSeed), seed_lt);
    chain_seeds(f, qid, qdir, qsize, (int)tid, tsize);
    if (!f->n_chains) return 0;

    /* :318-357: best chain if it contains the seed, else the widest containing
     * chain, else the best chain if it spans >= 5 kb of the query. */
    int pick = -1;
    if (contains_seed(&f->chains[0], seed_q, seed_t)) pick = 0;
    else {
        int best_cov = 0;
        for (size_t i = 0; i < f->n_chains; ++i) {
            if (!contains_seed(&f->chains[i], seed_q, seed_t)) continue;
            int cov = (int)(f->chains[i].qend - f->chains[i].qbeg);
            if (cov > best_cov) { best_cov = cov; pick = (int)i; }
        }
        if (pick >= 0) {
            /* reference quirk (:337-346): the record emitted here is whatever the scan
             * variable last held, i.e. the LAST chain in the list, not chain `pick`. */
            pick = (int)f->n_chains - 1;
        } else if (f->chains[0].qend - f->chains[0].qbeg >= 5000) pick = 0;
    }
    if (pick < 0) return 0;
    OCandidate c = f->chains[pick];
    c.score = support; c.qoff = seed_q; c.soff = seed_t;
    zero_span(f, &c, tstart, bs);
    int keep = (c.send - c.sbeg >= opt->align_size_cutoff) || (c.qend - c.qbeg >= opt->align_size_cutoff);
    if (keep) cand_push(out, &c);
    return keep;
}

void o_find_candidates(OFinder* f, const uint8_t* read, int read_size, int qid, int qdir,
                       int read_start_id, int ref_start_id, const OVolume* ref,
                       const OIndex* idx, const OMapOptions* opt, OCandVec* out)
{
    finder_reset(f);
    collect_seeds(f, read, read_size, qid, read_start_id, ref_start_id, ref, idx, opt);
    const int s = opt->block_score_cutoff;
    for (int i = 0; i < f->n_touched; ++i) {            /* word_finder.c:392-411 */
        int b = f->touched[i].blk;
        if (f->blk[b].score >= s && f->touched[i].score >= 2 * s)
            try_block(f, b, ref, opt, qid, qdir, read_size, out);
    }
}

static int cand_pm_gt(const void* a, const void* b)     /* pm_worker.c:15-24 */
{
    const OCandidate* x = (const OCandidate*)a; const OCandidate* y = (const OCandidate*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->qdir != y->qdir) return x->qdir < y->qdir ? -1 : 1;
    if (x->sid != y->sid) return x->sid < y->sid ? -1 : 1;
    if (x->qoff != y->qoff) return x->qoff < y->qoff ? -1 : 1;
    if (x->soff != y->soff) return x->soff < y->soff ? -1 : 1;
    return 0;
}

void o_map_volume(const OVolume* reads, int read_start_id, const OVolume* ref,
                  int ref_start_id, const OIndex* idx, const OMapOptions* opt, OCandVec* out)
{
    OFinder* f = o_finder_new(ref, opt);
    uint64_t maxlen = 0;
    for (uint64_t i = 0; i < reads->nseq; ++i) if (reads->seq[i].size > maxlen) maxlen = reads->seq[i].size;
    uint8_t* buf = (uint8_t*)malloc(maxlen + 1);
    OCandVec one = {0, 0, 0};
    for (uint64_t i = 0; i < reads->nseq; ++i) {            /* pm_worker.c:108-136,162-171 */
        one.n = 0;
        int L = (int)reads->seq[i].size;
        for (int dir = 0; dir < 2; ++dir) {
            o_extract(reads, i, dir, buf);
            o_find_candidates(f, buf, L, (int)i, dir, read_start_id, ref_start_id, ref, idx, opt, &one);
        }
        for (size_t c = 0; c < one.n; ++c) { one.a[c].qid += read_start_id; one.a[c].sid += ref_start_id; }
        if (one.n > (size_t)opt->num_candidates) {
            qsort(one.a, one.n, sizeof(OCandidate), cand_pm_gt);
            one.n = (size_t)opt->num_candidates;
        }
        for (size_t c = 0; c < one.n; ++c) cand_push(out, &one.a[c]);
    }
    free(one.a);
    free(buf);
    o_finder_free(f);
}

/* ---------------------------------------------------------------- ALN mode */

void o_map_volume_aln(const OVolume* reads, int read_start_id, const OVolume* ref,
                      int ref_start_id, const OIndex* idx, const OMapOptions* opt, OM4Vec* out)
{
    OFinder* f = o_finder_new(ref, opt);
    OAligner* al = o_aligner_new(opt->error);
    uint64_t maxlen = 0, maxref = 0;
    for (uint64_t i = 0; i < reads->nseq; ++i) if (reads->seq[i].size > maxlen) maxlen = reads->seq[i].size;
    for (uint64_t i = 0; i < ref->nseq; ++i) if (ref->seq[i].size > maxref) maxref = ref->seq[i].size;
    uint8_t* fwd = (uint8_t*)malloc(maxlen + 1);
    uint8_t* rev = (uint8_t*)malloc(maxlen + 1);
    uint8_t* sub = (uint8_t*)malloc(maxref + 1);
    OCandVec one = {0, 0, 0};
    OM4Record* m4 = NULL; size_t nm4 = 0, mm4 = 0;
    for (uint64_t i = 0; i < reads->nseq; ++i) {
        one.n = 0;
        const int L = (int)reads->seq[i].size;
        o_extract(reads, i, 0, fwd);
        o_find_candidates(f, fwd, L, (int)i, 0, read_start_id, ref_start_id, ref, idx, opt, &one);
        o_extract(reads, i, 1, rev);
        o_find_candidates(f, rev, L, (int)i, 1, read_start_id, ref_start_id, ref, idx, opt, &one);
        qsort(one.a, one.n, sizeof(OCandidate), cand_pm_gt);                      /* pm_worker.c:139 */
        if (one.n > (size_t)opt->num_candidates) one.n = (size_t)opt->num_candidates;
        nm4 = 0;
        for (size_t c = 0; c < one.n; ++c) {                                       /* extend_candidates, :28-82 */
            const OCandidate* k = &one.a[c];
            int contained = 0;
            for (size_t x = 0; x < nm4 && !contained; ++x) {                       /* map_aux.c:4-18 */
                const OM4Record* m = &m4[x];
                contained = k->qdir == m->qdir && k->sid == m->sid &&
                            (uint64_t)k->qoff >= m->qoff && (uint64_t)k->qoff <= m->qend &&
                            (uint64_t)k->soff >= m->soff && (uint64_t)k->soff <= m->send;
            }
            if (contained) continue;
            o_extract(ref, (uint64_t)k->sid, 0, sub);
            OAlignment A;
            if (!o_onc_align(al, k->qdir == 0 ? fwd : rev, (int)k->qoff, (int)k->qsize, sub, (int)k->soff, (int)k->ssize,
                             512, opt->align_size_cutoff, 1, &A)) continue;
            if (nm4 == mm4) { mm4 = mm4 ? mm4 * 2 : 64; m4 = (OM4Record*)realloc(m4, sizeof(OM4Record) * mm4); }
            OM4Record* m = &m4[nm4++];
            memset(m, 0, sizeof *m);
            m->qid = k->qid; m->sid = k->sid; m->ident_perc = A.ident_perc; m->vscore = k->score; m->qdir = k->qdir;
            m->qoff = (uint64_t)A.qoff; m->qend = (uint64_t)A.qend; m->qext = (uint64_t)k->qoff; m->qsize = (uint64_t)k->qsize;
            m->sdir = 0; m->soff = (uint64_t)A.toff; m->send = (uint64_t)A.tend; m->sext = (uint64_t)k->soff; m->ssize = (uint64_t)k->ssize;
            if (m->qdir == 1) {                                                     /* :73-78 */
                uint64_t a = m->qsize - m->qend, b = m->qsize - m->qoff;
                m->qoff = a; m->qend = b;
            }
        }
        for (size_t x = 0; x < nm4; ++x) {
            m4[x].qid += read_start_id; m4[x].sid += ref_start_id;
            if (out->n == out->m) { out->m = out->m ? out->m * 2 : 256; out->a = (OM4Record*)realloc(out->a, sizeof(OM4Record) * out->m); }
            out->a[out->n++] = m4[x];
        }
    }
    free(m4); free(one.a); free(fwd); free(rev); free(sub);
    o_aligner_free(al);
    o_finder_free(f);
}
