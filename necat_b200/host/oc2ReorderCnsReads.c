/* oc2ReorderCnsReads -- sort the corrected and the uncorrected FASTA of a correction iteration by read id.  Drop-in for
 * the reference's reorder step (reorder_cns_reads/main.c:26-112, necat.pl:351):
 *
 *   oc2ReorderCnsReads cns_reads_in raw_reads_in cns_reads_out raw_reads_out
 *
 * The id of a record is the first run of digits in its name (main.c:26-38).  Records are written as ">name\nBASES\n"
 * with the bases as they come back from the 2-bit store (upper case ACGT; main.c:59-69).  Records with EQUAL ids
 * (the fragments of one read under oc2cns -f 0) keep their input order here; the reference leaves them in the order its
 * unstable introsort produces, which is one of the permutations of the same records.
 *
 * What is different from the reference: input that is already in id order -- what this repo's oc2cns writes -- is
 * streamed straight through without being held in memory; otherwise the records are held 2-bit packed and ordered with
 * a stable LSD radix sort; the two files are processed side by side; output is decoded four bases per byte into large
 * blocks instead of one fprintf per base.  No GPU involved.
 */
#include "nb_fastx.h"
#include "nb_host.h"

#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct { int32_t id; uint32_t order; } rec_key;
typedef struct { uint64_t offset, size, hdr_offset; } rec_info;

typedef struct {
    const char* in_path;
    const char* out_path;
    char        err[512];
    uint64_t    n_records, n_bases;
    int         streamed;
} job;

/* main.c:26-38.  A name without a digit has no id there (the scan runs off the string); here it is an error. */
static int header_id(const char* name, int32_t* id)
{
    const char* p = name;
    while (*p && !isdigit((unsigned char)*p)) ++p;
    if (!*p) return -1;
    uint32_t v = 0;
    while (isdigit((unsigned char)*p)) { v = v * 10u + (uint32_t)(*p - '0'); ++p; }
    *id = (int32_t)v;
    return 0;
}

typedef struct { FILE* f; char* buf; size_t len, cap; int failed; } writer;

static void w_flush(writer* w)
{
    if (w->len && fwrite(w->buf, 1, w->len, w->f) != w->len) w->failed = 1;
    w->len = 0;
}
static char* w_reserve(writer* w, size_t n)
{
    if (w->len + n > w->cap) {
        w_flush(w);
        if (n > w->cap) {
            char* q = (char*)realloc(w->buf, n);
            if (!q) { w->failed = 1; return NULL; }
            w->buf = q; w->cap = n;
        }
    }
    return w->buf + w->len;
}
static void w_record(writer* w, const char* name, size_t name_l, const nbh_pac* pac, uint64_t from, uint64_t n)
{
    char* o = w_reserve(w, name_l + n + 3);
    if (!o) return;
    *o++ = '>';
    memcpy(o, name, name_l); o += name_l;
    *o++ = '\n';
    nbh_pac_decode(pac, from, n, o); o += n;
    *o++ = '\n';
    w->len += name_l + n + 3;
}

static void* run_job(void* arg)
{
    job* j = (job*)arg;
    nbh_fastx* fx = nbh_fastx_open(j->in_path);
    if (!fx) { snprintf(j->err, sizeof j->err, "cannot open '%s'", j->in_path); return NULL; }
    FILE* out = fopen(j->out_path, "w");
    if (!out) { snprintf(j->err, sizeof j->err, "cannot open '%s' for writing: %s", j->out_path, strerror(errno)); nbh_fastx_close(fx); return NULL; }
    writer w = { out, (char*)malloc(8u << 20), 0, 8u << 20, 0 };
    nbh_pac all = { 0, 0, 0 }, one = { 0, 0, 0 };
    rec_key* keys = NULL; rec_info* info = NULL; size_t n = 0, m = 0;
    char* hdr = NULL; size_t hdr_l = 0, hdr_m = 0;
    int sorted = 1, rc;
    int32_t last_id = INT32_MIN;
    if (!w.buf) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }

    /* While the ids keep ascending, records go straight to the output.  The first record that breaks the order ends
     * that: the output is rewound and everything (re-read from the start) is held and sorted. */
    while ((rc = nbh_fastx_next(fx)) == 1) {
        size_t nl, sl;
        const char* name = nbh_fastx_name(fx, &nl);
        const uint8_t* seq = nbh_fastx_seq(fx, &sl);
        int32_t id;
        if (header_id(name, &id)) { snprintf(j->err, sizeof j->err, "record '%s' of '%s' has no numeric id", name, j->in_path); goto done; }
        if (id < last_id) { sorted = 0; break; }
        last_id = id;
        /* through the 2-bit store, so that the bases come out exactly as the reference's would (case, non-ACGT) */
        nbh_pac_clear(&one);
        if (nbh_pac_append(&one, seq, sl)) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }
        w_record(&w, name, nl, &one, 0, sl);
        j->n_records++; j->n_bases += sl;
    }
    if (sorted) {
        if (rc < 0) snprintf(j->err, sizeof j->err, "read error in '%s'", j->in_path);
        j->streamed = 1;
        goto done;
    }

    /* general case: start over, hold everything packed, stable sort by id */
    nbh_fastx_close(fx);
    fx = nbh_fastx_open(j->in_path);
    if (!fx) { snprintf(j->err, sizeof j->err, "cannot reopen '%s'", j->in_path); goto done; }
    w.len = 0;
    if (fflush(out) || fseek(out, 0, SEEK_SET) || ftruncate(fileno(out), 0)) { snprintf(j->err, sizeof j->err, "cannot rewind '%s'", j->out_path); goto done; }
    j->n_records = 0; j->n_bases = 0;
    while ((rc = nbh_fastx_next(fx)) == 1) {
        size_t nl, sl;
        const char* name = nbh_fastx_name(fx, &nl);
        const uint8_t* seq = nbh_fastx_seq(fx, &sl);
        if (n == m) {
            m = m ? m * 2 : 65536;
            keys = (rec_key*)realloc(keys, m * sizeof *keys);
            info = (rec_info*)realloc(info, m * sizeof *info);
            if (!keys || !info) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }
        }
        if (hdr_l + nl + 1 > hdr_m) {
            while (hdr_l + nl + 1 > hdr_m) hdr_m = hdr_m ? hdr_m * 2 : (1 << 20);
            hdr = (char*)realloc(hdr, hdr_m);
            if (!hdr) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }
        }
        if (header_id(name, &keys[n].id)) { snprintf(j->err, sizeof j->err, "record '%s' of '%s' has no numeric id", name, j->in_path); goto done; }
        keys[n].order = (uint32_t)n;
        info[n].offset = all.size; info[n].size = sl; info[n].hdr_offset = hdr_l;
        memcpy(hdr + hdr_l, name, nl + 1);
        hdr_l += nl + 1;
        if (nbh_pac_append(&all, seq, sl)) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }
        ++n;
        j->n_bases += sl;
    }
    if (rc < 0) { snprintf(j->err, sizeof j->err, "read error in '%s'", j->in_path); goto done; }
    j->n_records = n;
    {   /* stable LSD radix sort on the id (signed: flip the top bit), 4 passes of 8 bits */
        rec_key* tmp = (rec_key*)malloc((n ? n : 1) * sizeof *tmp);
        if (!tmp) { snprintf(j->err, sizeof j->err, "out of memory"); goto done; }
        rec_key *a = keys, *b = tmp;
        for (int pass = 0; pass < 4; ++pass) {
            size_t cnt[257];
            memset(cnt, 0, sizeof cnt);
            const int sh = pass * 8;
            for (size_t i = 0; i < n; ++i) cnt[((((uint32_t)a[i].id) ^ 0x80000000u) >> sh & 255u) + 1]++;
            for (int k = 0; k < 256; ++k) cnt[k + 1] += cnt[k];
            for (size_t i = 0; i < n; ++i) b[cnt[(((uint32_t)a[i].id) ^ 0x80000000u) >> sh & 255u]++] = a[i];
            rec_key* t = a; a = b; b = t;
        }
        /* 4 passes: the result is back in `keys` */
        free(tmp);
    }
    for (size_t i = 0; i < n && !w.failed; ++i) {
        const rec_info* r = &info[keys[i].order];
        const char* name = hdr + r->hdr_offset;
        w_record(&w, name, strlen(name), &all, r->offset, r->size);
    }
done:
    if (w.buf) w_flush(&w);
    if (w.failed && !j->err[0]) snprintf(j->err, sizeof j->err, "writing '%s' failed", j->out_path);
    if (fclose(out) && !j->err[0]) snprintf(j->err, sizeof j->err, "closing '%s' failed", j->out_path);
    if (fx) nbh_fastx_close(fx);
    nbh_pac_free(&all); nbh_pac_free(&one);
    free(keys); free(info); free(hdr); free(w.buf);
    return NULL;
}

int main(int argc, char** argv)
{
    if (argc != 5) {
        fprintf(stderr, "USAGE:\n%s cns_reads_in raw_reads_in cns_reads_out raw_reads_out\n", argv[0]);
        return 1;
    }
    job jobs[2];
    memset(jobs, 0, sizeof jobs);
    jobs[0].in_path = argv[1]; jobs[0].out_path = argv[3];
    jobs[1].in_path = argv[2]; jobs[1].out_path = argv[4];
    pthread_t th;
    const int threaded = pthread_create(&th, NULL, run_job, &jobs[1]) == 0;
    run_job(&jobs[0]);
    if (threaded) pthread_join(th, NULL); else run_job(&jobs[1]);
    for (int k = 0; k < 2; ++k) {
        if (jobs[k].err[0]) NBH_ERROR("%s", jobs[k].err);
        if (getenv("NB_TRACE"))
            fprintf(stderr, "[reorder] %s: %lu records, %lu bases, %s\n", jobs[k].in_path, (unsigned long)jobs[k].n_records,
                    (unsigned long)jobs[k].n_bases, jobs[k].streamed ? "already in id order: streamed" : "sorted");
    }
    return 0;
}
