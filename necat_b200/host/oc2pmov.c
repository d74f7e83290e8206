/*
 * oc2pmov -- drop-in replacement of the reference's oc2pmov
 * (pm_one_volume/main.c:28-47, pm_worker.c:337-400) running on one B200 through
 * libnecat_b200.  Same command line:
 *     oc2pmov [options] wrk-dir volume-id output
 * same files read (reads_info.txt, volume_names.txt, vol<N>) and written
 * (28-byte candidate records with -u 1, candidate text with -u 0).  Options may
 * come in any order and default to sDefaultPairwiseMapingOptions
 * (map_options.c:12-28) -- the reference leaves them uninitialised.
 * CAN mode (-j 0) is what necat.pl uses (necat.pl:31-32,197); ALN mode (-j 1)
 * writes M4 records: text (m4_record.h:72-124, read names instead of ids with
 * -i 1) or the raw 96-byte struct with -u 1.
 */
#include <getopt.h>
#include <stdlib.h>
#include <string.h>

#include "nb_helper.h"

static void usage(const char* prog)
{
    fprintf(stderr, "USAGE:\n%s [options] wrk-dir volume-id output\n\nOPTIONS AND DESCRIPTIONS:\n", prog);
    fprintf(stderr, "-k <Integer>\tkmer size\n-z <Integer>\tscan window size\n-q <Integer>\tkmer occurs > q times will be ignored\n"
                    "-b <Integer>\tblock size\n-n <Integer>\tnumber of candidates\n-a <Integer>\tmin align length\n"
                    "-d <Real>\tddf score cutoff\n-e <Real>\tsequencing error\n-m <Integer>\tnumber of output\n"
                    "-t <Integer>\tnumber of cpu threads\n-j <0 or 1>\tjob: 0 = find candidates only, 1 = perform alignemnt\n"
                    "-u <0 or 1>\toutput binary results: 0 = no, 1 = yes\n-i <0 or 1>\tuse header as sequence id: 0 = no, 1 = yes\n");
}

#ifdef NBH_AS_LIBRARY
int oc2pmov_main(int argc, char* argv[])
#else
int main(int argc, char* argv[])
#endif
{
#ifndef NBH_AS_LIBRARY
    { int st = 0; if (nbh_try_helper(NBH_TOOL_PMOV, argc, argv, &st)) return st; }      /* NECAT_B200_HELPER=1: the resident helper runs it */
#endif
    if (argc < 4) { usage(argv[0]); return 1; }
    nb_map_options o;
    nb_map_options_default(&o);
    int c;
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
        default: usage(argv[0]); return 1;
        }
    }
    const char* wrk_dir = argv[argc - 3];
    const int vid = atoi(argv[argc - 2]);
    const char* output = argv[argc - 1];
    if (o.job != 0 && o.job != 1) NBH_ERROR("invalid job -j %d", o.job);

    nbh_phase_t0 = nbh_now();
    nbh_volumes_info vi;
    if (nbh_volumes_info_load(wrk_dir, &vi)) NBH_ERROR("cannot read volume info in '%s'", wrk_dir);
    if (vid < 0 || vid >= vi.num_volumes) NBH_ERROR("volume id %d out of range", vid);
    const int dev = nbh_pick_device();
    nbh_ctx_begin(dev);                                          /* the CUDA context comes up while the volume is read */
    nbh_volume ref;
    if (nbh_volume_load(vi.names[vid], &ref)) NBH_ERROR("cannot load volume '%s'", vi.names[vid]);
    NBH_PHASE("volume load");
    nb_ctx* ctx = NULL;
    if (nbh_ctx_wait(&ctx)) NBH_ERROR("%s", nb_last_error(NULL));
    NBH_PHASE("context (rest)");
    nb_volume* dref = NULL;
    if (nb_volume_upload(ctx, ref.pac, ref.db_size, ref.offset, ref.size, ref.nseq, &dref)) NBH_ERROR("%s", nb_last_error(ctx));
    NBH_PHASE("volume upload");
    double t0 = nbh_now();
    NBH_LOG("'build_lookup_table' BEGINS");
    nb_index* idx = NULL;
    if (nb_index_build(ctx, dref, o.kmer_size, o.kmer_cnt_cutoff, &idx)) NBH_ERROR("%s", nb_last_error(ctx));
    NBH_LOG("'build_lookup_table' takes %.2lf secs.", nbh_now() - t0);
    NBH_PHASE("index");

    FILE* out = fopen(output, "w");
    if (!out) NBH_ERROR("cannot open '%s' for writing", output);
    for (int v = vid; v < vi.num_volumes; ++v) {          /* pm_worker.c:372-390 */
        char job[256];
        snprintf(job, sizeof job, "pairwise mapping v%d vs v%d", v, vid);
        NBH_LOG("'%s' BEGINS", job);
        t0 = nbh_now();
        nbh_volume rd;
        nb_volume* drd = dref;
        if (v != vid) {
            if (nbh_volume_load(vi.names[v], &rd)) NBH_ERROR("cannot load volume '%s'", vi.names[v]);
            if (nb_volume_upload(ctx, rd.pac, rd.db_size, rd.offset, rd.size, rd.nseq, &drd)) NBH_ERROR("%s", nb_last_error(ctx));
        }
        const uint64_t nreads = v == vid ? ref.nseq : rd.nseq;
        if (o.job == 1) {                                   /* ALN mode, pm_worker.c:138-161 */
            const nbh_volume* qv = (v == vid) ? &ref : &rd;
            nb_m4_record* m4 = NULL;
            uint64_t nm4 = 0;
            if (nb_map_volume_aln(ctx, idx, dref, vi.start_id[vid], drd, vi.start_id[v], 0, nreads, &o, &m4, &nm4))
                NBH_ERROR("%s", nb_last_error(ctx));
            if (o.binary_output) {
                if (fwrite(m4, sizeof(nb_m4_record), nm4, out) != nm4) NBH_ERROR("write to '%s' failed", output);
            } else {
                for (uint64_t i = 0; i < nm4; ++i) {
                    const nb_m4_record* m = &m4[i];
                    if (o.use_hdr_as_id)
                        fprintf(out, "%s\t%s\t", qv->hdr + qv->hdr_offset[m->qid - vi.start_id[v]],
                                ref.hdr + ref.hdr_offset[m->sid - vi.start_id[vid]]);
                    else fprintf(out, "%d\t%d\t", m->qid, m->sid);
                    fprintf(out, "%.2f\t%d\t%d\t%lu\t%lu\t%lu\t%d\t%lu\t%lu\t%lu\n", m->ident_perc, m->vscore, m->qdir,
                            (unsigned long)m->qoff, (unsigned long)m->qend, (unsigned long)m->qsize, m->sdir,
                            (unsigned long)m->soff, (unsigned long)m->send, (unsigned long)m->ssize);
                }
            }
            nb_free(m4);
        } else {
        nb_candidate* cans = NULL;
        uint64_t ncan = 0;
        if (nb_map_volume(ctx, idx, dref, vi.start_id[vid], drd, vi.start_id[v], 0, nreads, &o, &cans, &ncan))
            NBH_ERROR("%s", nb_last_error(ctx));
        if (o.binary_output) {
            nb_packed_candidate* pk = (nb_packed_candidate*)malloc(sizeof(nb_packed_candidate) * (ncan ? ncan : 1));
            for (uint64_t i = 0; i < ncan; ++i) nb_pack_candidate(&cans[i], &pk[i]);
            if (fwrite(pk, sizeof(nb_packed_candidate), ncan, out) != ncan) NBH_ERROR("write to '%s' failed", output);
            free(pk);
        } else {
            for (uint64_t i = 0; i < ncan; ++i) {           /* gapped_candidate.h:26-41 */
                const nb_candidate* k = &cans[i];
                fprintf(out, "%d\t%d\t%d\t%d\t%lu\t%lu\t%lu\t%lu\t%d\t%lu\t%lu\t%lu\t%lu\n", k->qid, k->sid, k->score, k->qdir,
                        (unsigned long)k->qbeg, (unsigned long)k->qend, (unsigned long)k->qoff, (unsigned long)k->qsize, k->sdir,
                        (unsigned long)k->sbeg, (unsigned long)k->send, (unsigned long)k->soff, (unsigned long)k->ssize);
            }
        }
        nb_free(cans);
        }
        if (v != vid) { nb_volume_free(ctx, drd); nbh_volume_free(&rd); }
        NBH_LOG("'%s' takes %.2lf secs.", job, nbh_now() - t0);
    }
    if (fclose(out)) NBH_ERROR("closing '%s' failed", output);
    NBH_PHASE("mapping + output");
    if (nbh_fast_exit() && !nbh_in_helper) { fflush(NULL); _exit(0); }
    nb_index_free(ctx, idx);
    nb_volume_free(ctx, dref);
    nbh_volume_free(&ref);
    nbh_volumes_info_free(&vi);
    nbh_ctx_put(ctx);
    NBH_PHASE("tear-down");
    return 0;
}
