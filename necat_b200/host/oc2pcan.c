/*
 * oc2pcan -- drop-in replacement of the reference's oc2pcan
 * (partition_candidates/main.c:17-28, pcan.c:39-146): regroups the candidate
 * records of oc2pmov by template.  Every record goes to the partition of its
 * subject and, with roles swapped (gapped_candidate.c:54-69), to the partition of
 * its query; partitions hold `batch` (-p, default 100000) consecutive read ids.
 * Writes <candidates>.partitions and <candidates>.p<i>.  Host-only glue (the
 * exchange step between the two GPU halves); options as the
 * reference: -p batch size, -f number of files opened at once, -t threads (unused).
 */
#include <getopt.h>
#include <stdlib.h>
#include <string.h>

#include "nb_host.h"

int main(int argc, char* argv[])
{
    int batch = 100000, nfiles = 100, threads = 1;      /* pcan_options.c:10-14 */
    if (argc < 3) { fprintf(stderr, "USAGE:\n%s [options] wrk_dir candidates\n", argv[0]); return 1; }
    int c;
    while ((c = getopt(argc - 2, argv, "p:f:t:")) != -1) {
        switch (c) {
        case 'p': batch = atoi(optarg); break;
        case 'f': nfiles = atoi(optarg); break;
        case 't': threads = atoi(optarg); break;
        default: return 1;
        }
    }
    (void)threads;
    if (batch < 1) NBH_ERROR("invalid partition size -p %d", batch);
    if (nfiles < 1) nfiles = 1;
    const char* wrk_dir = argv[argc - 2];
    const char* can_path = argv[argc - 1];
    nbh_volumes_info vi;
    if (nbh_volumes_info_load(wrk_dir, &vi)) NBH_ERROR("cannot read volume info in '%s'", wrk_dir);
    const int nparts = (vi.num_reads + batch - 1) / batch;
    char path[4096];
    snprintf(path, sizeof path, "%s.partitions", can_path);
    FILE* pf = fopen(path, "w");
    if (!pf) NBH_ERROR("cannot open '%s'", path);
    fprintf(pf, "%d\n", nparts);
    fclose(pf);

    /* The candidate file is read in blocks of at most 256 MB (pcan.c:39-99 does the same): a first pass sizes nothing -- the
     * partition files are simply appended to -- so memory stays at three blocks whatever the file holds.  Inside a block
     * the records are put in partition order by a counting sort (input order kept inside a partition, as before). */
    FILE* in = fopen(can_path, "rb");
    if (!in) NBH_ERROR("cannot open '%s'", can_path);
    const uint64_t block = ((uint64_t)256 << 20) / sizeof(nb_packed_candidate);
    nb_packed_candidate* cans = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * block);
    nb_packed_candidate* out = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * 2 * block);
    uint64_t* cnt = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)nparts + 1));
    uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)nparts + 1));
    if (!cans || !out || !cnt || !cur) NBH_ERROR("out of memory");
    /* at most -f partition files stay open; with more partitions than that every block reopens the ones it touches */
    const int keep_open = nparts <= nfiles;
    FILE** pfile = (FILE**)calloc((size_t)nparts, sizeof(FILE*));
    if (!pfile) NBH_ERROR("out of memory");
    for (int p = 0; p < nparts; ++p) {                          /* every partition file exists, also an empty one */
        snprintf(path, sizeof path, "%s.p%d", can_path, p);
        FILE* f = fopen(path, "wb");
        if (!f) NBH_ERROR("cannot open '%s'", path);
        if (keep_open) pfile[p] = f; else fclose(f);
    }
    uint64_t seen = 0;
    for (;;) {
        const uint64_t n = fread(cans, sizeof(nb_packed_candidate), block, in);
        if (n == 0) { if (ferror(in)) NBH_ERROR("cannot read '%s'", can_path); break; }
        memset(cnt, 0, sizeof(uint64_t) * ((size_t)nparts + 1));
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t ps = cans[i].item[1] / (uint32_t)batch, pq = cans[i].item[4] / (uint32_t)batch;
            if ((int)ps >= nparts || (int)pq >= nparts) NBH_ERROR("candidate %lu refers to a read id beyond reads_info.txt", (unsigned long)(seen + i));
            ++cnt[ps + 1]; ++cnt[pq + 1];
        }
        for (int p = 0; p < nparts; ++p) cnt[p + 1] += cnt[p];
        memcpy(cur, cnt, sizeof(uint64_t) * ((size_t)nparts + 1));
        for (uint64_t i = 0; i < n; ++i) {
            const nb_packed_candidate* s = &cans[i];
            out[cur[s->item[1] / (uint32_t)batch]++] = *s;
            nb_packed_candidate d;                               /* change_pcan_roles */
            const uint32_t w = s->item[0];
            uint32_t r = w & ((1u << 30) - 1);
            if (w >> 31) r |= 1u << 30;
            if ((w >> 30) & 1) r |= 1u << 31;
            d.item[0] = r;
            d.item[1] = s->item[4]; d.item[2] = s->item[5]; d.item[3] = s->item[6];
            d.item[4] = s->item[1]; d.item[5] = s->item[2]; d.item[6] = s->item[3];
            out[cur[d.item[1] / (uint32_t)batch]++] = d;
        }
        for (int p = 0; p < nparts; ++p) {
            const uint64_t m = cnt[p + 1] - cnt[p];
            if (m == 0) continue;
            FILE* f = pfile[p];
            if (!f) {
                snprintf(path, sizeof path, "%s.p%d", can_path, p);
                f = fopen(path, "ab");
                if (!f) NBH_ERROR("cannot open '%s'", path);
            }
            if (fwrite(out + cnt[p], sizeof(nb_packed_candidate), m, f) != m) NBH_ERROR("write to partition %d failed", p);
            if (!pfile[p] && fclose(f)) NBH_ERROR("write to partition %d failed", p);
        }
        seen += n;
    }
    fclose(in);
    for (int p = 0; p < nparts; ++p) if (pfile[p] && fclose(pfile[p])) NBH_ERROR("write to partition %d failed", p);
    free(pfile); free(out); free(cur); free(cnt); free(cans);
    nbh_volumes_info_free(&vi);
    return 0;
}
