/*
 * necat_b200.h -- C-ABI of libnecat_b200.so: the B200 (sm_100a) implementation of
 * NECAT's correct-stage hot path (oc2pmov -> oc2pcan -> oc2cns).
 *
 * The reference has no FFI/plugin layer: its boundary is "process + files"
 * (SURVEY.md 8b).  The drop-in executables necat_b200/bin/{oc2pmov,oc2cns} keep
 * that contract; they are thin C hosts over the entry points below, which are
 * what a maintainer binding the GPU path into the reference's own mains would
 * call (INTEGRATION.md shows the stubs).  Plain pointers and sizes only; all
 * buffers passed in are HOST memory and are not retained; every function
 * returns 0 on success or a negative NB_E* code, with nb_last_error() giving a
 * message.  One nb_ctx per GPU, used from one host thread at a time.
 *
 * Reference interfaces replaced (paths relative to /root/reference/src):
 *   nb_volume_upload      <- pdb_load                  common/packed_db.c:386
 *   nb_index_build        <- build_lookup_table        lookup_table/lookup_table.c:149
 *   nb_map_volume         <- pm_search_one_volume      pm_one_volume/pm_worker.c:84 (CAN mode)
 *                            = find_candidates         word_finder/word_finder.c:364 per (read, strand)
 *   nb_extend_candidates  <- cns_extension / onc_align consensus/consensus_aux.c:124, gapped_align/oc_aligner.c:303
 *   nb_consensus_job      <- consensus_one_partition   consensus/consensus_one_partition.c:110
 */
#ifndef NECAT_B200_H
#define NECAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_OK            0
#define NB_EINVAL       -1
#define NB_ECUDA        -2
#define NB_ENOMEM       -3
#define NB_ENODEVICE    -4
#define NB_ELIMIT       -5

typedef struct nb_ctx    nb_ctx;
typedef struct nb_volume nb_volume;
typedef struct nb_index  nb_index;

/* ------------------------------------------------------------------ context */
int         nb_ctx_create(int device, nb_ctx** out);
void        nb_ctx_destroy(nb_ctx* ctx);
const char* nb_last_error(const nb_ctx* ctx);      /* ctx may be NULL: last global error */
const char* nb_version(void);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t    nb_launch_count(const nb_ctx* ctx);
/* GPU time of the last call of each stage, milliseconds (CUDA events on the ctx stream). */
typedef struct {
    float index_ms;         /* nb_index_build */
    float map_ms;           /* nb_map_volume: all kernels */
    float map_probe_ms;     /*   of which: seed probe + vote kernels */
    float extend_ms;        /* nb_extend_candidates / extension inside nb_consensus_job */
    float consensus_ms;     /* consensus kernels inside nb_consensus_job */
    float h2d_ms, d2h_ms;
    uint64_t h2d_bytes, d2h_bytes;
} nb_timings;
void        nb_get_timings(const nb_ctx* ctx, nb_timings* out);

/* ------------------------------------------------------------------ volumes */
/*
 * A PackedDB volume (common/packed_db.h:12-29): `pac` holds ceil(db_size/4)
 * bytes, base l in byte l>>2 at bit shift ((~l)&3)<<1; seq_offset/seq_size give
 * each read's start and length in bases.
 */
int  nb_volume_upload(nb_ctx* ctx, const uint8_t* pac, uint64_t db_size,
                      const uint64_t* seq_offset, const uint64_t* seq_size, uint64_t nseq,
                      nb_volume** out);
void nb_volume_free(nb_ctx* ctx, nb_volume* vol);

/* -------------------------------------------------------------------- index */
/*
 * k-mer index of one volume: all k-mers occurring <= q times, positions ascending
 * per k-mer (lookup_table.c:15-124).  k <= 15, volume < 2^32 bases.
 */
int  nb_index_build(nb_ctx* ctx, const nb_volume* ref, int kmer_size, int kmer_cnt_cutoff,
                    nb_index** out);
void nb_index_free(nb_ctx* ctx, nb_index* idx);
/* For parity tests: number of indexed positions, and a copy of the CSR arrays
 * (starts has 4^k + 1 entries; offsets has n_offsets entries). */
int  nb_index_stats(nb_ctx* ctx, const nb_index* idx, uint64_t* n_offsets, uint64_t* n_distinct);
int  nb_index_download(nb_ctx* ctx, const nb_index* idx, uint32_t* starts, uint32_t* offsets);

/*
 * The index of one volume built by several GPUs together (multi-GPU form of build_lookup_table,
 * lookup_table/lookup_table.c:149-164): GPU number `shard` of `nshards` builds the CSR of the shard-th
 * slice of ceil(4^k / nshards) hash values; nb_index_shard_export writes its starts (plus `base`, the number of
 * positions of the slices before it) and positions into the device arrays the GPUs assemble -- the caller all-gathers
 * those (NCCL) and wraps the result with nb_index_adopt, which does not take ownership of the arrays.
 */
typedef struct nb_index_shard nb_index_shard;
int  nb_index_build_shard(nb_ctx* ctx, const nb_volume* ref, int k, int q, int shard, int nshards,
                          nb_index_shard** out, uint64_t* n_offsets, uint64_t* n_distinct);
int  nb_index_shard_export(nb_ctx* ctx, nb_index_shard* shard, uint64_t base, uint32_t* d_tbl_dst, uint32_t* d_off_dst);
int  nb_index_adopt(nb_ctx* ctx, int k, int q, uint32_t* d_tbl, uint32_t* d_off, uint64_t n_offsets, uint64_t n_distinct,
                    nb_index** out);

/* --------------------------------------------------------------- candidates */
/* common/map_options.h:10-25 (same meaning, same defaults via nb_map_options_default) */
typedef struct {
    int    kmer_size, scan_window, kmer_cnt_cutoff, block_size, block_score_cutoff;
    int    num_candidates, align_size_cutoff;
    double ddfs_cutoff, error;
    int    num_output, num_threads, job, binary_output, use_hdr_as_id;
} nb_map_options;
void nb_map_options_default(nb_map_options* o);     /* map_options.c:12-28 */

/* 28-byte PackedGappedCandidate, common/gapped_candidate.h:59-80 */
typedef struct { uint32_t item[7]; } nb_packed_candidate;

/* Unpacked record (gapped_candidate.h:9-19) for -u 0 text output and tests */
typedef struct {
    int32_t qid, sid, qdir, sdir, score;
    int64_t qbeg, qend, qsize, sbeg, send, ssize, qoff, soff;
} nb_candidate;

/*
 * All candidates of reads [read_lo, read_hi) of `reads` (both strands) against
 * (ref, idx), CAN mode of pm_worker.c:84-198: ids made global with the start
 * ids, per-read truncation to opt->num_candidates.  On return *out is a
 * malloc()ed array the caller frees with nb_free(); records of one read are
 * contiguous, reads ascending, FWD strand before REV, emission order inside.
 */
int  nb_map_volume(nb_ctx* ctx, const nb_index* idx, const nb_volume* ref, int ref_start_id,
                   const nb_volume* reads, int read_start_id, uint64_t read_lo, uint64_t read_hi,
                   const nb_map_options* opt, nb_candidate** out, uint64_t* n_out);
void nb_pack_candidate(const nb_candidate* c, nb_packed_candidate* p);  /* gapped_candidate.c:13-28 */
/* the same over n records: what the CAN-mode output loop does to every candidate before it is written (pm_worker.c:168-171) */
void nb_pack_candidates(const nb_candidate* c, uint64_t n, nb_packed_candidate* out);
/* change_pcan_roles (gapped_candidate.c:54-69) over n records: query and subject swapped, strands with them.
 * What oc2pcan does to every record before it files it under its query read as well (pcan.c:86-91). */
void nb_swap_candidate_roles(const nb_packed_candidate* in, uint64_t n, nb_packed_candidate* out);
void nb_free(void* p);

/* 96-byte M4Record, common/m4_record.h:10-25 (binary form of oc2pmov -j 1 -u 1) */
typedef struct {
    int32_t  qid, qdir;
    uint64_t qoff, qend, qext, qsize;
    int32_t  sid, sdir;
    uint64_t soff, send, sext, ssize;
    double   ident_perc;
    int32_t  vscore;
    int32_t  pad_;
} nb_m4_record;

/*
 * ALN mode of pm_worker.c:84-161 (oc2pmov -j 1): the candidates of every read are sorted by
 * (score desc, qdir, sid, qoff, soff), cut to opt->num_candidates, and extended from their seed with
 * onc_align (error opt->error, tail match length 1, minimum size opt->align_size_cutoff) unless the seed
 * lies inside an alignment already reported for the same strand and subject (common/map_aux.c:4-18).
 * Query coordinates of reverse-strand records are flipped to the forward strand (pm_worker.c:73-78).
 */
int  nb_map_volume_aln(nb_ctx* ctx, const nb_index* idx, const nb_volume* ref, int ref_start_id,
                       const nb_volume* reads, int read_start_id, uint64_t read_lo, uint64_t read_hi,
                       const nb_map_options* opt, nb_m4_record** out, uint64_t* n_out);

/* Work counters of the last nb_map_volume call (roofline arithmetic). */
typedef struct {
    uint64_t n_units;       /* (read, strand) pairs */
    uint64_t n_probes;      /* k-mer index probes */
    uint64_t n_hits;        /* index positions returned below soff_max */
    uint64_t n_seeds;       /* hits accepted into scoring blocks */
    uint64_t n_blocks;      /* touched scoring blocks */
    uint64_t n_eligible;    /* blocks passing the static score gates */
    uint64_t n_candidates;
    uint64_t query_bases;   /* sum of read lengths x 2 strands */
} nb_map_counters;
void nb_get_map_counters(const nb_ctx* ctx, nb_map_counters* out);

/* ---------------------------------------------------------------- extension */
/* consensus/cns_options.h (same meaning; defaults cns_options.c:10-22) */
typedef struct {
    int    min_align_size, min_cov, max_cov, min_size, full_consensus;
    double error, mapping_ratio;
    int    num_threads, rescue_long_indels, use_fixed_ident_cutoff, small_memory;
} nb_cns_options;
void nb_cns_options_default(nb_cns_options* o);

/* Result of extending one candidate (oc_aligner.h:7-27 minus the strings). */
typedef struct {
    int32_t ok;                 /* onc_align's return value */
    int32_t qoff, qend, toff, tend;
    int32_t aln_size;           /* alignment columns */
    int32_t n_match;            /* identical columns: ident_perc = 100.0 * n_match / aln_size */
    int32_t long_ins;           /* 1 = the alignment holds a run of >= 255 query-only columns: get_cns_tags drops it
                                   whole (tasc/align_tags.c:40-44) */
    uint64_t ops_offset;        /* into the ops buffer returned alongside */
} nb_alignment;

/*
 * Extend n candidates (already normalised so the subject is forward, as
 * consensus_one_partition.c:55-95 leaves them) with the block-wise Myers
 * extension of oc_aligner.c:303-451 (block 512, error opt->error,
 * tail_match_len given).  `ops` receives one byte per alignment column:
 * 0 = both bases (match or mismatch), 1 = query base vs gap, 2 = gap vs target
 * base, in left-to-right order; the gapped strings follow from the sequences.
 * Both outputs are malloc()ed; free with nb_free().
 */
int  nb_extend_candidates(nb_ctx* ctx, const nb_volume* reads,
                          const nb_packed_candidate* cans, uint64_t n,
                          double error, int min_align_size, int tail_match_len,
                          nb_alignment** out_aln, uint8_t** out_ops, uint64_t* out_ops_bytes);

/* ------------------------------------------------------------------- rescue */
/*
 * First half of the oc2cns -r 1 rescue (consensus/consensus_aux.c:169-178): ocda_go (gapped_align/oc_daligner.c:34-79),
 * Myers' O(nd) local alignment with path-tip trimming (gapped_align/align.c:382-1945) through the point
 * (qstart, tstart) of query read qid (strand qdir) and template read sid (forward).  Returns the two ranges and the
 * number of differences; ok = both ranges span at least min_align_size.  The second half is nb_rescue_global below;
 * nb_consensus_job runs both when its options ask for -r 1.
 */
typedef struct { uint32_t qid, sid; int32_t qdir, qstart, tstart, pad_; } nb_rescue_task;
typedef struct { int32_t ok, qbeg, qend, tbeg, tend, diffs, overflow, pad_; } nb_rescue_result;
int  nb_rescue_local(nb_ctx* ctx, const nb_volume* reads, const nb_rescue_task* tasks, uint64_t n, double error,
                     int min_align_size, nb_rescue_result* results);

/*
 * Second half of the rescue: edlib_go with find_path (edlib/edlib_wrapper.c:111-255) -- the global alignment of query
 * range [qfrom, qto) (strand qdir) with template range [tfrom, tto) if their edit distance is <= tolerance, by edlib's
 * rules for choosing among co-optimal alignments (Hirschberg's first split row, up / left / diagonal traceback below
 * 1 MiB of alignment data), trimmed to its first and last run of k_match matches.  Results as nb_extend_candidates.
 * Checked against the oracle operation by operation on a B200 (tests/test_gpu_rescue.py) and in host emulation.
 */
typedef struct { uint32_t qid, sid; int32_t qdir, qfrom, qto, tfrom, tto, tolerance; } nb_rescue_range;
int  nb_rescue_global(nb_ctx* ctx, const nb_volume* reads, const nb_rescue_range* ranges, uint64_t n, double error,
                      int min_align_size, int k_match, nb_alignment** out_aln, uint8_t** out_ops, uint64_t* out_ops_bytes);

/* ---------------------------------------------------------------- consensus */
/* One FASTA record of oc2cns (common/cns_seq.h:7-41). */
typedef struct {
    int32_t  id, left, right, org_size, num_can, num_ovlps;
    double   ident_cutoff;
    uint64_t seq_offset;    /* into the sequence buffer (ASCII ACGT) */
    int32_t  seq_len;
    int32_t  is_raw;        /* 0 -> cns_out, 1 -> raw_out */
} nb_cns_record;

/*
 * Whole oc2cns job for the templates of one candidate partition
 * (consensus_one_partition.c:110-201): cans[0..n) are the partition's records
 * (as oc2pcan wrote them: grouped by template in item[1], any order inside).
 * Emits the corrected fragments and the raw leftovers of corrected templates,
 * plus the whole reads of [min_sid, max_sid) that were never corrected (:173-194).
 */
int  nb_consensus_job(nb_ctx* ctx, const nb_volume* reads,
                      const nb_packed_candidate* cans, uint64_t n,
                      const nb_cns_options* opt,
                      nb_cns_record** out_recs, uint64_t* n_recs,
                      char** out_seq, uint64_t* out_seq_bytes);

/*
 * Multi-GPU form: the same partition is handed to every rank, rank `shard_index`
 * of `shard_count` corrects the templates with id % shard_count == shard_index
 * (and dumps only those uncorrected reads); the union of the ranks' records is
 * exactly nb_consensus_job's output.  Templates are independent
 * (consensus/main.c:71-73), so no collective is needed inside.
 */
int  nb_consensus_job_sharded(nb_ctx* ctx, const nb_volume* reads,
                              const nb_packed_candidate* cans, uint64_t n,
                              const nb_cns_options* opt, int shard_count, int shard_index,
                              nb_cns_record** out_recs, uint64_t* n_recs,
                              char** out_seq, uint64_t* out_seq_bytes);

typedef struct {
    uint64_t n_templates, n_extensions, n_block_aligns, n_dp_columns, n_dp_word_updates;
    uint64_t n_accepted, n_tags, n_cns_records, cns_bases;
    /* the largest k_extend launch of the job (the roofline of bench.py describes this launch) */
    uint64_t big_ext_tasks, big_ext_block_aligns, big_ext_word_updates, big_ext_usec, big_ext_dp_columns;
} nb_cns_counters;
void nb_get_cns_counters(const nb_ctx* ctx, nb_cns_counters* out);

#ifdef __cplusplus
}
#endif
#endif
