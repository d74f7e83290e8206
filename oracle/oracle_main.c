/*
 * oracle_main.c -- command line front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *   oracle_cli pmov [oc2pmov options] wrk_dir volume_id output      (pm_one_volume/main.c:28-47)
 *   oracle_cli cns  [oc2cns options]  wrk_dir candidates cns_out raw_out
 * The cns form reads the concatenated oc2pmov output directly (it performs
 * oc2pcan's regrouping in memory).
 */
#include "necat_oracle.h"

#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int run_pmov(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: pmov [options] wrk_dir vid output\n"); return 1; }
    OMapOptions o;
    o_map_options_default(&o);
    int c;
    optind = 1;
    while ((c = getopt(argc - 3, argv, "k:z:q:b:s:n:a:d:e:m:t:j:u:i:")) != -1) {
        switch (c) {
        case 'k': o.kmer_size = atoi(optarg); break;
        case 'z': o.scan_window = atoi(optarg); break;
        case 'q': o.kmer_cnt_cutoff = atoi(optarg); break;
        case 'b': o.block_size = atoi(optarg); break;
        case 's': o.block_score_cutoff = atoi(optarg); break;
        case 'n': o.num_candidates = atoi(optarg); break;
        case 'a': o.align_size_cutoff = atoi(optarg); break;
        case 'd': o.ddfs_cutoff = atof(optarg); break;
        case 'e': o.error = atof(optarg); break;
        case 'm': o.num_output = atoi(optarg); break;
        case 't': o.num_threads = atoi(optarg); break;
        case 'j': o.job = atoi(optarg); break;
        case 'u': o.binary_output = atoi(optarg); break;
        case 'i': o.use_hdr_as_id = atoi(optarg); break;
        default: return 1;
        }
    }
    const char* wrk = argv[argc - 3];
    int vid = atoi(argv[argc - 2]);
    const char* outp = argv[argc - 1];
    OVolumesInfo vi;
    if (o_volumes_info_load(wrk, &vi)) { fprintf(stderr, "cannot load volume info from %s\n", wrk); return 1; }
    OVolume ref;
    if (o_volume_load(vi.names[vid], &ref)) { fprintf(stderr, "cannot load %s\n", vi.names[vid]); return 1; }
    OIndex idx;
    o_index_build(&ref, o.kmer_size, o.kmer_cnt_cutoff, &idx);
    FILE* out = fopen(outp, "wb");
    if (!out) { fprintf(stderr, "cannot open %s\n", outp); return 1; }
    for (int v = vid; v < vi.num_volumes; ++v) {
        OVolume reads;
        if (o_volume_load(vi.names[v], &reads)) { fprintf(stderr, "cannot load %s\n", vi.names[v]); return 1; }
        if (o.job == 1) {                                   /* ALN mode: M4 records (pm_worker.c:138-161) */
            OM4Vec mv = {0, 0, 0};
            o_map_volume_aln(&reads, vi.start_id[v], &ref, vi.start_id[vid], &idx, &o, &mv);
            for (size_t i = 0; i < mv.n; ++i) {
                const OM4Record* m = &mv.a[i];
                if (o.binary_output) { fwrite(m, sizeof *m, 1, out); continue; }
                if (o.use_hdr_as_id)                          /* m4_record.h:99-124 */
                    fprintf(out, "%s\t%s\t", reads.hdr + reads.seq[m->qid - vi.start_id[v]].hdr_offset,
                            ref.hdr + ref.seq[m->sid - vi.start_id[vid]].hdr_offset);
                else fprintf(out, "%d\t%d\t", m->qid, m->sid);
                fprintf(out, "%.2f\t%d\t%d\t%lu\t%lu\t%lu\t%d\t%lu\t%lu\t%lu\n", m->ident_perc, m->vscore, m->qdir,
                        (unsigned long)m->qoff, (unsigned long)m->qend, (unsigned long)m->qsize, m->sdir,
                        (unsigned long)m->soff, (unsigned long)m->send, (unsigned long)m->ssize);
            }
            free(mv.a);
            o_volume_free(&reads);
            continue;
        }
        OCandVec cv = {0, 0, 0};
        o_map_volume(&reads, vi.start_id[v], &ref, vi.start_id[vid], &idx, &o, &cv);
        for (size_t i = 0; i < cv.n; ++i) {
            if (o.binary_output) {
                OPackedCandidate p;
                o_pack_candidate(&cv.a[i], &p);
                fwrite(&p, sizeof p, 1, out);
            } else {
                const OCandidate* k = &cv.a[i];     /* gapped_candidate.h:26-41 */
                fprintf(out, "%d\t%d\t%d\t%d\t%lu\t%lu\t%lu\t%lu\t%d\t%lu\t%lu\t%lu\t%lu\n", k->qid, k->sid, k->score, k->qdir,
                        (unsigned long)k->qbeg, (unsigned long)k->qend, (unsigned long)k->qoff, (unsigned long)k->qsize, k->sdir,
                        (unsigned long)k->sbeg, (unsigned long)k->send, (unsigned long)k->soff, (unsigned long)k->ssize);
            }
        }
        free(cv.a);
        o_volume_free(&reads);
    }
    fclose(out);
    o_index_free(&idx);
    o_volume_free(&ref);
    o_volumes_info_free(&vi);
    return 0;
}

int run_cns(int argc, char** argv);

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s pmov|cns ...\n", argv[0]); return 1; }
    if (!strcmp(argv[1], "pmov")) return run_pmov(argc - 1, argv + 1);
    if (!strcmp(argv[1], "cns")) return run_cns(argc - 1, argv + 1);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 1;
}
