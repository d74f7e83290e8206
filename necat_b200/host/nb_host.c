/* nb_host.c -- PackedDB / volume-list readers and logging for the drop-in hosts. */
#include "nb_host.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

static const char kMagic[] = "ontcns_pac_header_hofuwhogfuewo";   /* 31 bytes, no NUL on disk */

double nbh_phase_t0 = 0.0;

double nbh_now(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

void nbh_log(const char* level, const char* file, int line, const char* fmt, ...)
{
    FILE* out = strcmp(level, "INFO") ? stderr : stdout;
    time_t t = time(NULL);
    char ts[64];
    strncpy(ts, ctime(&t), sizeof ts - 1);
    ts[sizeof ts - 1] = 0;
    size_t n = strlen(ts);
    if (n && ts[n - 1] == '\n') ts[n - 1] = 0;
    fprintf(out, "[%s] %s: ", ts, level);
    va_list ap;
    va_start(ap, fmt);
    vfprintf(out, fmt, ap);
    va_end(ap);
    fprintf(out, " (%s, %d)\n", file, line);
    fflush(out);
}

int nbh_volume_load(const char* path, nbh_volume* v)
{
    memset(v, 0, sizeof *v);
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char magic[32];
    const size_t ml = sizeof kMagic - 1;
    int rc = -2;
    typedef struct { uint64_t offset, size, hdr_offset; int32_t platform, pad; } rec_t;
    rec_t* recs = NULL;
    if (fread(magic, 1, ml, f) != ml || memcmp(magic, kMagic, ml)) goto done;
    if (fread(&v->nseq, 8, 1, f) != 1 || fread(&v->db_size, 8, 1, f) != 1) goto done;
    recs = (rec_t*)malloc(sizeof(rec_t) * (v->nseq ? v->nseq : 1));
    v->offset = (uint64_t*)malloc(8 * (v->nseq + 1));
    v->size = (uint64_t*)malloc(8 * (v->nseq + 1));
    v->hdr_offset = (uint64_t*)malloc(8 * (v->nseq + 1));
    if (!recs || !v->offset || !v->size || !v->hdr_offset) goto done;
    if (fread(recs, sizeof(rec_t), v->nseq, f) != v->nseq) goto done;
    for (uint64_t i = 0; i < v->nseq; ++i) {
        v->offset[i] = recs[i].offset; v->size[i] = recs[i].size; v->hdr_offset[i] = recs[i].hdr_offset;
    }
    if (fread(&v->hdr_size, 8, 1, f) != 1) goto done;
    v->hdr = (char*)malloc(v->hdr_size + 1);
    if (!v->hdr) goto done;
    if (v->hdr_size && fread(v->hdr, 1, v->hdr_size, f) != v->hdr_size) goto done;
    v->hdr[v->hdr_size] = 0;
    {
        uint64_t nb = (v->db_size + 3) >> 2;
        v->pac = (uint8_t*)calloc(nb + 16, 1);
        if (!v->pac || fread(v->pac, 1, nb, f) != nb) goto done;
    }
    rc = 0;
done:
    free(recs);
    fclose(f);
    if (rc) nbh_volume_free(v);
    return rc;
}

int nbh_volume_append(nbh_volume* d, const nbh_volume* s)
{
    const uint64_t base = d->db_size, hbase = d->hdr_size, n0 = d->nseq;
    d->offset = (uint64_t*)realloc(d->offset, 8 * (n0 + s->nseq + 1));
    d->size = (uint64_t*)realloc(d->size, 8 * (n0 + s->nseq + 1));
    d->hdr_offset = (uint64_t*)realloc(d->hdr_offset, 8 * (n0 + s->nseq + 1));
    d->hdr = (char*)realloc(d->hdr, hbase + s->hdr_size + 1);
    if (!d->offset || !d->size || !d->hdr_offset || !d->hdr) return -1;
    for (uint64_t i = 0; i < s->nseq; ++i) {
        d->offset[n0 + i] = s->offset[i] + base;
        d->size[n0 + i] = s->size[i];
        d->hdr_offset[n0 + i] = s->hdr_offset[i] + hbase;
    }
    memcpy(d->hdr + hbase, s->hdr, s->hdr_size);
    d->hdr_size += s->hdr_size;
    d->hdr[d->hdr_size] = 0;
    /* the store grows in place; the bases of s follow those of d without a gap (pdb_merge), which puts them at a bit
     * offset of 0, 2, 4 or 6 inside the bytes: whole 64-bit words are shifted (a 2 Gbp volume in ~0.1 s; the
     * base-by-base loop this replaces took 2.5 s per volume on the 5.6 Gbp set of BASELINE config 4) */
    const uint64_t nb = (base + s->db_size + 3) >> 2, old_nb = (base + 3) >> 2;
    uint8_t* np = (uint8_t*)realloc(d->pac, nb + 16);
    if (!np) return -1;
    memset(np + old_nb, 0, nb + 16 - old_nb);
    {
        const unsigned r = (unsigned)(base & 3) * 2;
        const uint64_t sb = (s->db_size + 3) >> 2;
        uint8_t* o = np + (base >> 2);
        if (r == 0) {
            memcpy(o, s->pac, sb);
        } else {
            uint8_t carry = (uint8_t)(o[0] & (0xff << (8 - r)));     /* its top r bits are the last bases of d */
            uint64_t k = 0;
            for (; k + 8 <= sb; k += 8) {
                uint64_t w;
                memcpy(&w, s->pac + k, 8);
                w = __builtin_bswap64(w);               /* the first base of a byte sits in its top bits */
                uint64_t x = ((uint64_t)carry << 56) | (w >> r);
                carry = (uint8_t)(w << (8 - r));
                x = __builtin_bswap64(x);
                memcpy(o + k, &x, 8);
            }
            for (; k < sb; ++k) {
                const uint8_t b = s->pac[k];
                o[k] = (uint8_t)(carry | (b >> r));
                carry = (uint8_t)(b << (8 - r));
            }
            o[k] = carry;                               /* (inside the 16 bytes of slack when nothing is left) */
        }
        /* whatever the last byte of s held beyond its bases must not reach the store */
        const uint64_t last = base + s->db_size;
        if (last & 3) np[last >> 2] &= (uint8_t)(0xff << (8 - 2 * (unsigned)(last & 3)));
        memset(np + nb, 0, 16);
    }
    d->pac = np;
    d->db_size = base + s->db_size;
    d->nseq = n0 + s->nseq;
    return 0;
}

void nbh_volume_free(nbh_volume* v)
{
    free(v->pac); free(v->offset); free(v->size); free(v->hdr_offset); free(v->hdr);
    memset(v, 0, sizeof *v);
}

int nbh_volumes_info_load(const char* wrk_dir, nbh_volumes_info* vi)
{
    memset(vi, 0, sizeof *vi);
    char path[4096], line[4096], name[4096];
    const size_t wl = strlen(wrk_dir);
    const char* sep = (wl && wrk_dir[wl - 1] == '/') ? "" : "/";
    snprintf(path, sizeof path, "%s%sreads_info.txt", wrk_dir, sep);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    if (fscanf(f, "%d%d", &vi->num_volumes, &vi->num_reads) != 2) { fclose(f); return -2; }
    fclose(f);
    snprintf(path, sizeof path, "%s%svolume_names.txt", wrk_dir, sep);
    f = fopen(path, "r");
    if (!f) return -3;
    vi->names = (char**)calloc((size_t)vi->num_volumes, sizeof(char*));
    vi->start_id = (int*)calloc((size_t)vi->num_volumes, sizeof(int));
    vi->count = (int*)calloc((size_t)vi->num_volumes, sizeof(int));
    for (int i = 0; i < vi->num_volumes; ++i) {
        if (!fgets(line, sizeof line, f) || sscanf(line, "%4095s %d %d", name, &vi->start_id[i], &vi->count[i]) != 3) {
            fclose(f);
            return -4;
        }
        vi->names[i] = strdup(name);
    }
    fclose(f);
    return 0;
}

void nbh_volumes_info_free(nbh_volumes_info* vi)
{
    if (vi->names) for (int i = 0; i < vi->num_volumes; ++i) free(vi->names[i]);
    free(vi->names); free(vi->start_id); free(vi->count);
    memset(vi, 0, sizeof *vi);
}

int nbh_order_by_id(const nb_cns_record* recs, uint64_t n, uint64_t* order)
{
    uint64_t* tmp = (uint64_t*)malloc(8 * (n ? n : 1));
    if (!tmp) return -1;
    for (uint64_t i = 0; i < n; ++i) order[i] = i;
    uint64_t *a = order, *b = tmp;
    for (int pass = 0; pass < 4; ++pass) {                      /* stable LSD radix sort, 8 bits a pass; ids are >= 0 */
        uint64_t cnt[257];
        memset(cnt, 0, sizeof cnt);
        const int sh = pass * 8;
        for (uint64_t i = 0; i < n; ++i) cnt[(((uint32_t)recs[a[i]].id >> sh) & 255u) + 1]++;
        for (int k = 0; k < 256; ++k) cnt[k + 1] += cnt[k];
        for (uint64_t i = 0; i < n; ++i) b[cnt[((uint32_t)recs[a[i]].id >> sh) & 255u]++] = a[i];
        uint64_t* t = a; a = b; b = t;
    }
    free(tmp);                                                  /* an even number of passes: the result is in `order` */
    return 0;
}
