/*
 * oc2cns -- drop-in replacement of the reference's oc2cns (consensus/main.c:24-78)
 * running on one B200 through libnecat_b200.  Same command line:
 *     oc2cns [options] wrk_dir candidates cns_out raw_out [-mn node_id num_nodes]
 * Reads reads_info.txt / volume_names.txt / vol<N> (all volumes merged,
 * main.c:66), <candidates>.partitions and <candidates>.p<i> as oc2pcan writes them,
 * and writes the corrected / uncorrected FASTA records of common/cns_seq.h:23-41.
 * Record order inside the files is unspecified in the reference as well
 * (SURVEY.md 0.4); here it is deterministic.
 */
#include <getopt.h>
#include <stdlib.h>
#include <string.h>

#include "nb_helper.h"

static void usage(const char* prog)
{
    fprintf(stderr, "USAGE:\n%s [options] wrk_dir candidates cns_out raw_out\n\n"
                    "If Multiple Nodes Are Used:\n%s [options] wrk_dir candidates cns_out raw_out -mn node_id num_nodes\n\n"
                    "OPTIONS AND DESCRIPTIONS:\n"
                    "-a <Integer>\talign length cutoff\n-x <Integer>\tminimal coverage\n-y <Integer>\tmaximal coverage\n"
                    "-l <Integer>\tminimal length of corrected reads.\n-f <0 or 1>\tfull consensus or not: 1 = yes, 0 = no\n"
                    "-e <Real>\tsequencing error\n-p <Real>\tminimal mapping ratio\n-t <Integer>\tnumber of cpu threads\n"
                    "-r <0 or 1>\trescue long indels or not: 1 = yes, 0 = no\n"
                    "-u <0 or 1>\tuse dynamic or fixed ident cutoff: 1 = fixed, 0 = dynamic\n-s <0 or 1>\tuse small memoty\n",
            prog, prog);
}

/* ">ontsa_id_<id>_(...)" or ">name_(...)" when the read is already a corrected one (cns_seq.h:25-29) */
static size_t format_header(char* buf, size_t cap, const char* name, const nb_cns_record* r)
{
    int n;
    if (strncmp(name, "ontsa_id", 8) == 0) n = snprintf(buf, cap, ">%s_", name);
    else n = snprintf(buf, cap, ">ontsa_id_%d_", r->id);
    n += snprintf(buf + n, cap - (size_t)n, "(%d_%d_%d_%d_%d_%d_%lf)\n", r->left, r->right, r->seq_len, r->org_size,
                  r->num_can, r->num_ovlps, r->ident_cutoff);
    return (size_t)n;
}

#ifdef NBH_AS_LIBRARY
int oc2cns_main(int argc, char* argv[])
#else
int main(int argc, char* argv[])
#endif
{
#ifndef NBH_AS_LIBRARY
    { int st = 0; if (nbh_try_helper(NBH_TOOL_CNS, argc, argv, &st)) return st; }      /* NECAT_B200_HELPER=1: the resident helper runs it */
#endif
    int spid = 0, nnode = 1;
    if (argc >= 4 && strcmp(argv[argc - 3], "-mn") == 0) {       /* consensus/main.c:37-41 */
        spid = atoi(argv[argc - 2]);
        nnode = atoi(argv[argc - 1]);
        argc -= 3;
    }
    if (argc < 5) { usage(argv[0]); return 1; }
    nb_cns_options o;
    nb_cns_options_default(&o);
    int c;
    while ((c = getopt(argc - 4, argv, "a:x:y:l:f:e:p:t:r:u:s:")) != -1) {
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
        default: usage(argv[0]); return 1;
        }
    }
    const char* wrk_dir = argv[argc - 4];
    const char* can_path = argv[argc - 3];
    const char* cns_path = argv[argc - 2];
    const char* raw_path = argv[argc - 1];
    if (nnode < 1) NBH_ERROR("invalid node count %d", nnode);

    nbh_phase_t0 = nbh_now();
    const int dev = nbh_pick_device();
    nbh_ctx_begin(dev);                                          /* the CUDA context comes up while the volumes are read */
    nbh_volumes_info vi;
    if (nbh_volumes_info_load(wrk_dir, &vi)) NBH_ERROR("cannot read volume info in '%s'", wrk_dir);
    nbh_volume reads;
    if (nbh_volume_load(vi.names[0], &reads)) NBH_ERROR("cannot load volume '%s'", vi.names[0]);
    for (int v = 1; v < vi.num_volumes; ++v) {                   /* merge_volumes, makedb_aux.c:138-154 */
        nbh_volume x;
        if (nbh_volume_load(vi.names[v], &x)) NBH_ERROR("cannot load volume '%s'", vi.names[v]);
        if (nbh_volume_append(&reads, &x)) NBH_ERROR("out of memory merging volumes");
        nbh_volume_free(&x);
    }
    char path[4096];
    snprintf(path, sizeof path, "%s.partitions", can_path);      /* pcan_aux.c:41 */
    FILE* pf = fopen(path, "r");
    if (!pf) NBH_ERROR("cannot open '%s'", path);
    int num_partitions = 0;
    if (fscanf(pf, "%d", &num_partitions) != 1) NBH_ERROR("cannot parse '%s'", path);
    fclose(pf);

    NBH_PHASE("volume load");
    nb_ctx* ctx = NULL;
    if (nbh_ctx_wait(&ctx)) NBH_ERROR("%s", nb_last_error(NULL));
    NBH_PHASE("context (rest)");
    nb_volume* dreads = NULL;
    if (nb_volume_upload(ctx, reads.pac, reads.db_size, reads.offset, reads.size, reads.nseq, &dreads)) NBH_ERROR("%s", nb_last_error(ctx));
    NBH_PHASE("volume upload");

    FILE* cns_out = fopen(cns_path, "w");
    FILE* raw_out = fopen(raw_path, "w");
    if (!cns_out || !raw_out) NBH_ERROR("cannot open the output files");
    static char obuf_c[1 << 22], obuf_r[1 << 22];
    setvbuf(cns_out, obuf_c, _IOFBF, sizeof obuf_c);
    setvbuf(raw_out, obuf_r, _IOFBF, sizeof obuf_r);
    for (int p = spid; p < num_partitions; p += nnode) {         /* consensus/main.c:71-73 */
        char job[256];
        snprintf(job, sizeof job, "consensus partition %d", p);
        NBH_LOG("'%s' BEGINS", job);
        const double t0 = nbh_now();
        snprintf(path, sizeof path, "%s.p%d", can_path, p);
        FILE* f = fopen(path, "rb");
        uint64_t n = 0;
        nb_packed_candidate* cans = NULL;
        if (f) {
            fseek(f, 0, SEEK_END);
            n = (uint64_t)ftell(f) / sizeof(nb_packed_candidate);
            fseek(f, 0, SEEK_SET);
            cans = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * (n ? n : 1));
            if (!cans || fread(cans, sizeof(nb_packed_candidate), n, f) != n) NBH_ERROR("cannot read '%s'", path);
            fclose(f);
        }
        if (n) {
            nb_cns_record* recs = NULL; uint64_t nrec = 0; char* seq = NULL; uint64_t nseq = 0;
            if (nb_consensus_job(ctx, dreads, cans, n, &o, &recs, &nrec, &seq, &nseq)) NBH_ERROR("%s", nb_last_error(ctx));
            NBH_PHASE("consensus job");
            char hdr[4096];
            uint64_t* order = (uint64_t*)malloc(8 * (nrec ? nrec : 1));
            if (!order || nbh_order_by_id(recs, nrec, order)) NBH_ERROR("out of memory");
            for (uint64_t k = 0; k < nrec; ++k) {               /* in read id order */
                const nb_cns_record* r = &recs[order[k]];
                FILE* out = r->is_raw ? raw_out : cns_out;
                const char* name = reads.hdr + reads.hdr_offset[r->id];
                const size_t hl = format_header(hdr, sizeof hdr, name, r);
                fwrite(hdr, 1, hl, out);
                fwrite(seq + r->seq_offset, 1, (size_t)r->seq_len, out);
                fputc('\n', out);
            }
            free(order);
            nb_free(recs);
            nb_free(seq);
            NBH_PHASE("fasta output");
        }
        free(cans);
        NBH_LOG("'%s' takes %.2lf secs.", job, nbh_now() - t0);
    }
    if (fclose(cns_out) || fclose(raw_out)) NBH_ERROR("closing the output files failed");
    if (nbh_fast_exit() && !nbh_in_helper) { fflush(NULL); _exit(0); }
    nb_volume_free(ctx, dreads);
    nbh_ctx_put(ctx);
    nbh_volume_free(&reads);
    nbh_volumes_info_free(&vi);
    NBH_PHASE("tear-down");
    return 0;
}
