/* oracle_rescue.h -- the rescue branch of cns_extension (oc2cns -r 1); see oracle_rescue.c.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_RESCUE_H
#define ORACLE_RESCUE_H
#include <stdint.h>

typedef struct ORescue ORescue;
ORescue* o_rescue_new(double error);        /* new_OcDalignData + new_FullEdlibAlignData, consensus_aux.c:34-35 */
void     o_rescue_free(ORescue* r);

/* ocda_go, gapped_align/oc_daligner.c:34-79: local alignment through (qstart, tstart); 1 = both spans >= min_align_size */
int o_rescue_local(ORescue* r, const uint8_t* query, int qstart, int qsize, const uint8_t* target, int tstart, int tsize,
                   int min_align_size, int* qb, int* qe, int* tb, int* te, int* diffs);

/* edlib_go (find_path), edlib/edlib_wrapper.c:111-255; qaln / taln are malloc()ed */
int o_rescue_global(ORescue* r, const uint8_t* query, int qfrom, int qto, const uint8_t* target, int tfrom, int tto,
                    int tolerance, int min_align_size, int k_match, int* qoff, int* qend, int* toff, int* tend,
                    double* ident_perc, char** qaln, char** taln, int* aln_size);
/* last column of the global DP (col[i] = D[i][tn], i = 0 .. qn); strides of -1 walk backwards from q / t */
void o_nw_last_column(const uint8_t* q, int qn, const uint8_t* t, int tn, int q_step, int t_step, int* col);
/* a leaf of the global alignment: full-matrix traceback (up, left, diagonal); edlib op codes 0 / 1 / 2 / 3 */
int o_nw_traceback_ops(const uint8_t* q, int qn, const uint8_t* t, int tn, uint8_t* ops_out);
#endif
