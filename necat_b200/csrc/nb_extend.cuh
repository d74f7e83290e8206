// nb_extend.cuh -- device-side interface of the block-wise Myers extension kernel.
#pragma once
#include "nb_common.cuh"

// One extension request: a candidate normalised so that the template (subject) is forward.
struct ExtTask {
    uint32_t qid, sid;      // read ids in the (merged) read volume
    int32_t  qoff, soff;    // extension start (gapped_candidate.c:44-50)
    int32_t  qdir;          // query strand
    int32_t  pad_;
    uint64_t ops_base;      // start of this task's region in the ops pool
    uint32_t ops_cap;       // region size
    uint32_t lcap;          // columns reserved for the left extension (= qoff + soff)
};

// op codes, one byte per alignment column, left to right along the template
#define OP_MATCH 0      // query base == template base
#define OP_QONLY 1      // query base against a gap   (insertion w.r.t. the template)
#define OP_TONLY 2      // gap against a template base (deletion)
#define OP_MISMATCH 3   // query base != template base

#define EXT_MAX_BLK 832         // >= 1.3 * 611 + slack (oc_aligner.c:127-137), multiple of 32
#define EXT_WARPS 8             // warps per CTA

struct ExtParams {
    const uint8_t*  pac;            // template (subject) volume
    const uint64_t* seq_off;
    const uint32_t* seq_len;
    const uint8_t*  q_pac;          // query volume (the same one in oc2cns, possibly another in oc2pmov -j 1)
    const uint64_t* q_off;
    const uint32_t* q_len;
    const ExtTask*  tasks;
    uint32_t        n_tasks;
    nb_alignment*   results;
    uint8_t*        ops;            // pool
    uint2*          mat;            // per candidate in flight: EXT_MAT_STRIDE entries, the (Pv, Ph) words of the band, [column][4]
    uint8_t*        rev;            // per candidate in flight: EXT_REV_STRIDE bytes, the columns of the current block, last to first
    uint32_t*       next_task;      // work queue cursors: [0] over the tasks [n_long, n_tasks), [1] over the longest ones [0, n_long)
    uint32_t        n_long;         // the first n_long tasks (the caller lists the longest first) go to warps that keep only two in flight
    uint32_t        n_long_warps;   // ... the first n_long_warps warps of the grid, until that queue is empty
    unsigned long long* counters;   // [0] block alignments, [1] DP columns, [2] word updates
    double          error;
    int             min_align_size;
    int             tail_match_len;
    int             block_size;     // 512
    int             slots;          // candidates in flight per warp (2 .. EXT_SLOTS): few when the launch is small, so that a candidate's chain of blocks
                                    // does not wait for the forward passes of fifteen others
    int             band_test;      // 1: the walk-back only trusts the two middle words of a column's band (tests: exercises the moving band)
};

#define EXT_MAT_STRIDE ((size_t)(EXT_MAX_BLK + 32) * 8)
#define EXT_REV_STRIDE ((size_t)2 * EXT_MAX_BLK + 32)

int nb_run_extension(nb_ctx* ctx, const nb_volume* queries, const nb_volume* templates, const ExtTask* d_tasks, uint32_t n_tasks,
                     nb_alignment* d_results, uint8_t* d_ops, double error, int min_align_size, int tail_match_len);
