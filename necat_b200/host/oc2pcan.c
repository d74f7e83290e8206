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
    (void)threads; (void)nfiles;
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

    FILE* in = fopen(can_path, "rb");
    if (!in) NBH_ERROR("cannot open '%s'", can_path);
    fseek(in, 0, SEEK_END);
    const uint64_t n = (uint64_t)ftell(in) / sizeof(nb_packed_candidate);
    fseek(in, 0, SEEK_SET);
    nb_packed_candidate* cans = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * (n ? n : 1));
    if (!cans || fread(cans, sizeof(nb_packed_candidate), n, in) != n) NBH_ERROR("cannot read '%s'", can_path);
    fclose(in);
    /* counting sort by partition: sizes, then fill */
    uint64_t* cnt = (uint64_t*)calloc((size_t)nparts + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t ps = cans[i].item[1] / (uint32_t)batch, pq = cans[i].item[4] / (uint32_t)batch;
        if ((int)ps >= nparts || (int)pq >= nparts) NBH_ERROR("candidate %lu refers to a read id beyond reads_info.txt", (unsigned long)i);
        ++cnt[ps + 1]; ++cnt[pq + 1];
    }
    for (int p = 0; p < nparts; ++p) cnt[p + 1] += cnt[p];
    nb_packed_candidate* out = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * (2 * n + 1));
    uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)nparts + 1));
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
        snprintf(path, sizeof path, "%s.p%d", can_path, p);
        FILE* f = fopen(path, "wb");
        if (!f) NBH_ERROR("cannot open '%s'", path);
        const uint64_t m = cnt[p + 1] - cnt[p];
        if (fwrite(out + cnt[p], sizeof(nb_packed_candidate), m, f) != m) NBH_ERROR("write to '%s' failed", path);
        fclose(f);
    }
    free(out); free(cur); free(cnt); free(cans);
    nbh_volumes_info_free(&vi);
    return 0;
}
