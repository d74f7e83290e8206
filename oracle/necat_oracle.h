/*
 * necat_oracle.h -- CPU restatement ("oracle") of NECAT's oc2pmov -> oc2cns hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (necat_b200/, the C-ABI
 * library, the oc2pmov/oc2cns executables) may include, link or call this.  It is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker.  Parity is pinned: every function here is validated against the
 * unmodified reference binaries built by oracle/Makefile.ref into oracle/_ref/
 * (tests/test_oracle_vs_ref.py, fixtures under tests/golden/).
 *
 * Each function cites the reference file:line (relative to /root/reference/src)
 * whose behaviour it restates.  The code is written from the algorithm, not
 * transcribed: different data structures (sorted sparse index instead of the
 * dense 4^k table, merge sorts instead of klib introsort, ...).
 */
#ifndef NECAT_ORACLE_H
#define NECAT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- volumes */

/* common/packed_db.h:12-18 (32-byte on-disk record) */
typedef struct {
    uint64_t offset;
    uint64_t size;
    uint64_t hdr_offset;
    int32_t  platform;
    int32_t  pad_;
} OSeqInfo;

typedef struct {
    uint8_t*  pac;      /* 2-bit bases, first base in the top bits of byte 0 */
    uint64_t  db_size;  /* bases */
    uint64_t  nseq;
    OSeqInfo* seq;
    char*     hdr;      /* NUL separated names */
    uint64_t  hdr_size;
} OVolume;

int  o_volume_load(const char* path, OVolume* v);           /* packed_db.c:317-348 */
void o_volume_free(OVolume* v);
int  o_volume_append(OVolume* dst, const OVolume* src);     /* packed_db.c:419-448 (pdb_merge) */
static inline int o_base(const OVolume* v, uint64_t i)      /* ontcns_aux.h:118-119 */
{
    return (v->pac[i >> 2] >> (((~i) & 3) << 1)) & 3;
}
/* packed_db.c:255-275: one byte per base, strand 1 = reverse complement */
void o_extract(const OVolume* v, uint64_t id, int strand, uint8_t* out);
uint64_t o_offset_to_id(const OVolume* v, uint64_t offset); /* packed_db.c:197-213 */

typedef struct {
    int       num_volumes;
    int       num_reads;
    char**    names;
    int*      start_id;
    int*      count;
} OVolumesInfo;
int  o_volumes_info_load(const char* wrk_dir, OVolumesInfo* vi); /* makedb_aux.c:78-118 */
void o_volumes_info_free(OVolumesInfo* vi);

/* ------------------------------------------------------------------ index */

/* lookup_table.c:149-190.  Same content as the reference's (kmer_stats,
 * offset_list) pair, stored sparsely: distinct kept hashes ascending, each with
 * its run of ascending positions. */
typedef struct {
    int       k;
    uint64_t  n_keys;
    uint32_t* key;      /* hash value, ascending */
    uint64_t* start;    /* n_keys + 1 */
    uint64_t  n_off;
    uint64_t* off;      /* volume positions */
} OIndex;

int  o_index_build(const OVolume* v, int k, int q, OIndex* idx);
void o_index_free(OIndex* idx);
/* lookup_table.c:176-190 */
const uint64_t* o_index_lookup(const OIndex* idx, uint64_t hash, uint64_t* n);

/* -------------------------------------------------------------- candidates */

/* common/map_options.h:10-25 */
typedef struct {
    int    kmer_size, scan_window, kmer_cnt_cutoff, block_size, block_score_cutoff;
    int    num_candidates, align_size_cutoff;
    double ddfs_cutoff, error;
    int    num_output, num_threads, job, binary_output, use_hdr_as_id;
} OMapOptions;
void o_map_options_default(OMapOptions* o);    /* map_options.c:12-28 */

/* common/gapped_candidate.h:9-19 */
typedef struct {
    int     qid, sid, qdir, sdir, score;
    int64_t qbeg, qend, qsize;
    int64_t sbeg, send, ssize;
    int64_t qoff, soff;
} OCandidate;

typedef struct { uint32_t item[7]; } OPackedCandidate;    /* gapped_candidate.h:59-61 */
void o_pack_candidate(const OCandidate* c, OPackedCandidate* p);     /* gapped_candidate.c:13-28 */
void o_unpack_candidate(OCandidate* c, const OPackedCandidate* p);   /* gapped_candidate.c:31-51 */
void o_swap_roles(const OPackedCandidate* s, OPackedCandidate* d);   /* gapped_candidate.c:54-69 */
void o_normalise_sdir(OPackedCandidate* p, uint32_t qsize, uint32_t ssize); /* :72-93 */

typedef struct OFinder OFinder;     /* per-thread scratch, word_finder.c:15-38 */
OFinder* o_finder_new(const OVolume* ref, const OMapOptions* opt);
void     o_finder_free(OFinder* f);

typedef struct { OCandidate* a; size_t n, m; } OCandVec;

/* word_finder.c:364-412: candidates of one (read, strand) against the index */
void o_find_candidates(OFinder* f, const uint8_t* read, int read_size, int qid, int qdir,
                       int read_start_id, int ref_start_id, const OVolume* ref,
                       const OIndex* idx, const OMapOptions* opt, OCandVec* out);

/* pm_worker.c:84-198 (CAN mode) for every read of `reads` against (ref, idx):
 * appends the per-read candidate lists (global ids, truncated to -n) to out. */
void o_map_volume(const OVolume* reads, int read_start_id, const OVolume* ref,
                  int ref_start_id, const OIndex* idx, const OMapOptions* opt,
                  OCandVec* out);

/* common/m4_record.h:10-25 (96-byte struct, 4 bytes of tail padding) */
typedef struct {
    int32_t  qid, qdir;
    uint64_t qoff, qend, qext, qsize;
    int32_t  sid, sdir;
    uint64_t soff, send, sext, ssize;
    double   ident_perc;
    int32_t  vscore;
    int32_t  pad_;
} OM4Record;
typedef struct { OM4Record* a; size_t n, m; } OM4Vec;

/* pm_worker.c:84-161 (ALN mode, -j 1) for every read of `reads`: candidates sorted, truncated to -n,
 * extended with onc_align (tail match 1) unless their seed lies inside an alignment already reported
 * for the same (strand, subject) (common/map_aux.c:4-18); ids made global. */
void o_map_volume_aln(const OVolume* reads, int read_start_id, const OVolume* ref,
                      int ref_start_id, const OIndex* idx, const OMapOptions* opt, OM4Vec* out);

/* -------------------------------------------------------------- extension */

typedef struct {
    int    ok;             /* onc_align return value (align_size >= min_align_size) */
    int    qoff, qend, toff, tend;
    double ident_perc;
    char*  qaln;           /* gapped strings, NUL terminated, owned by OAligner */
    char*  taln;
    int    aln_size;
} OAlignment;

typedef struct OAligner OAligner;
OAligner* o_aligner_new(double error);      /* oc_aligner.c:11-31 */
void      o_aligner_free(OAligner* a);

/* edlib_ex.c:733-800.  Returns 1 and fills qaln/taln (NUL terminated) + ends. */
int o_edlib_align(OAligner* a, const uint8_t* q, int qn, const uint8_t* t, int tn,
                  char* qaln, char* taln, int* qend, int* tend);

/* oc_aligner.c:303-451 */
int o_onc_align(OAligner* a, const uint8_t* query, int qstart, int qsize,
                const uint8_t* target, int tstart, int tsize,
                int block_size, int min_align_size, int tail_match_len, OAlignment* out);

/* --------------------------------------------------------------- consensus */

/* consensus/cns_options.h */
typedef struct {
    int    min_align_size, min_cov, max_cov, min_size, full_consensus;
    double error, mapping_ratio;
    int    num_threads, rescue_long_indels, use_fixed_ident_cutoff, small_memory;
} OCnsOptions;
void o_cns_options_default(OCnsOptions* o);   /* cns_options.c:10-22 */

/* One output record of oc2cns (common/cns_seq.h:7-41) */
typedef struct {
    int    id, left, right, org_size, num_can, num_ovlps;
    double ident_cutoff;
    char*  seq;     /* ASCII ACGT */
    int    seq_len;
    int    is_raw;  /* 0 -> cns_out, 1 -> raw_out */
} OCnsRecord;

typedef struct { OCnsRecord* a; size_t n, m; } OCnsRecVec;
void o_cns_recvec_free(OCnsRecVec* v);

/* weight summation order inside equal-key tag groups (tasc/cns_aux.c:48):
 * 0 = canonical (order alignments were added), 1 = emulate klib introsort
 * exactly (what the reference binary does). */
extern int o_cns_tag_order_mode;
extern int o_cns_partition_size;      /* reads per partition: 100000 as oc2pcan's default (pcan_options.c:10-14), -P in oracle_cli */

/* consensus_one_read.c:216-392 for one template: cans[0..n) are the template's
 * candidates already normalised so that the template is forward.
 * Returns 1 if the template counts as corrected (consensus_one_read.c:391). */
int o_consensus_one_read(const OVolume* reads, const OPackedCandidate* cans, size_t n,
                         const OCnsOptions* opt, OCnsRecVec* out);

/* partition_candidates/pcan.c:39-103 + consensus_one_partition.c:10-95,173-194:
 * whole oc2cns job over an in-memory candidate list (as written by oc2pmov). */
int o_cns_job(const OVolume* reads, const OPackedCandidate* cans, size_t n,
              const OCnsOptions* opt, OCnsRecVec* out);

/* FASTA formatting of one record, common/cns_seq.h:23-41; returns bytes written */
size_t o_format_cns_record(const OVolume* reads, const OCnsRecord* r, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
