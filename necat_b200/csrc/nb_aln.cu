// nb_aln.cu -- nb_map_volume_aln: oc2pmov's ALN mode (-j 1, M4 records).
//
// Replaces the ALN branch of pm_search_one_volume / extend_candidates
// (pm_one_volume/pm_worker.c:28-82,138-161).  Candidates come from nb_map_volume; every candidate that
// survives the per-read cut is extended speculatively on the GPU (k_extend, tail match length 1), then the
// reference's sequential "skip candidates whose seed lies inside an alignment already reported" rule
// (common/map_aux.c:4-18) is replayed on the host over the results.
#include <algorithm>
#include <vector>

#include "nb_extend.cuh"

extern "C" int nb_map_volume_aln(nb_ctx* ctx, const nb_index* idx, const nb_volume* ref, int ref_start_id,
                                 const nb_volume* reads, int read_start_id, uint64_t read_lo, uint64_t read_hi,
                                 const nb_map_options* opt, nb_m4_record** out, uint64_t* n_out)
{
    if (!out || !n_out) return nb_fail(ctx, NB_EINVAL, "nb_map_volume_aln: bad arguments");
    *out = NULL; *n_out = 0;
    nb_candidate* cans = NULL;
    uint64_t n = 0;
    int rc = nb_map_volume(ctx, idx, ref, ref_start_id, reads, read_start_id, read_lo, read_hi, opt, &cans, &n);
    if (rc) return rc;
    // per read: order of pm_worker.c:15-24 (the CAN-mode cut already kept the same top -n set)
    std::vector<nb_candidate> c(cans, cans + n);
    nb_free(cans);
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && c[j].qid == c[i].qid) ++j;
        std::stable_sort(c.begin() + (long)i, c.begin() + (long)j, [](const nb_candidate& a, const nb_candidate& b) {
            if (a.score != b.score) return a.score > b.score;
            if (a.qdir != b.qdir) return a.qdir < b.qdir;
            if (a.sid != b.sid) return a.sid < b.sid;
            if (a.qoff != b.qoff) return a.qoff < b.qoff;
            return a.soff < b.soff;
        });
        i = j;
    }
    std::vector<nb_m4_record> res;
    if (n) {
        if (n >= (1ull << 31)) return nb_fail(ctx, NB_ELIMIT, "too many candidates in one call");
        NB_CUDA(ctx, cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        std::vector<ExtTask> tasks(n);
        uint64_t pool = 0;
        for (uint64_t i = 0; i < n; ++i) {
            ExtTask& T = tasks[i];
            T.qid = (uint32_t)(c[i].qid - read_start_id); T.sid = (uint32_t)(c[i].sid - ref_start_id);
            T.qoff = (int32_t)c[i].qoff; T.soff = (int32_t)c[i].soff; T.qdir = c[i].qdir; T.pad_ = 0;
            const uint64_t mn = (uint64_t)std::min(c[i].qoff, c[i].soff);
            T.lcap = (uint32_t)(c[i].qoff + c[i].soff);
            T.ops_cap = (uint32_t)((uint64_t)c[i].qsize + (uint64_t)c[i].ssize + mn + 64);
            T.ops_base = pool;
            pool += (T.ops_cap + 15) & ~15ull;
        }
        NbDevBuf D(ctx);
        ExtTask* d_tasks; nb_alignment* d_res; uint8_t* d_ops;
        if ((rc = D.get(&d_tasks, n))) return rc;
        if ((rc = D.get(&d_res, n))) return rc;
        if ((rc = D.get(&d_ops, pool + 64))) return rc;
        NB_CUDA(ctx, cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(ExtTask) * n, cudaMemcpyHostToDevice, st));
        cudaEventRecord(ctx->ev0, st);
        rc = nb_run_extension(ctx, reads, ref, d_tasks, (uint32_t)n, d_res, d_ops, opt->error, opt->align_size_cutoff, 1 /* ONC_TAIL_MATCH_LEN_SHORT */);
        if (rc) return rc;
        cudaEventRecord(ctx->ev1, st);
        std::vector<nb_alignment> al(n);
        NB_CUDA(ctx, cudaMemcpyAsync(al.data(), d_res, sizeof(nb_alignment) * n, cudaMemcpyDeviceToHost, st));
        NB_CUDA(ctx, cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ctx->tm.extend_ms, ctx->ev0, ctx->ev1);
        // replay of extend_candidates (pm_worker.c:39-80)
        for (uint64_t i = 0; i < n;) {
            uint64_t j = i + 1;
            while (j < n && c[j].qid == c[i].qid) ++j;
            const size_t first = res.size();
            for (uint64_t k = i; k < j; ++k) {
                const nb_candidate& k_ = c[k];
                bool contained = false;
                for (size_t x = first; x < res.size() && !contained; ++x) {          // map_aux.c:4-18
                    const nb_m4_record& m = res[x];
                    contained = k_.qdir == m.qdir && k_.sid == m.sid && (uint64_t)k_.qoff >= m.qoff && (uint64_t)k_.qoff <= m.qend &&
                                (uint64_t)k_.soff >= m.soff && (uint64_t)k_.soff <= m.send;
                }
                if (contained || !al[k].ok) continue;
                nb_m4_record m;
                memset(&m, 0, sizeof m);
                m.qid = k_.qid; m.sid = k_.sid; m.qdir = k_.qdir; m.sdir = 0; m.vscore = k_.score;
                m.ident_perc = al[k].aln_size ? 100.0 * al[k].n_match / al[k].aln_size : 0.0;
                m.qoff = (uint64_t)al[k].qoff; m.qend = (uint64_t)al[k].qend; m.qext = (uint64_t)k_.qoff; m.qsize = (uint64_t)k_.qsize;
                m.soff = (uint64_t)al[k].toff; m.send = (uint64_t)al[k].tend; m.sext = (uint64_t)k_.soff; m.ssize = (uint64_t)k_.ssize;
                if (m.qdir == 1) { const uint64_t a = m.qsize - m.qend, b = m.qsize - m.qoff; m.qoff = a; m.qend = b; }
                res.push_back(m);
            }
            i = j;
        }
    }
    nb_m4_record* r = (nb_m4_record*)malloc(sizeof(nb_m4_record) * (res.size() ? res.size() : 1));
    if (!r) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    if (!res.empty()) memcpy(r, res.data(), sizeof(nb_m4_record) * res.size());
    *out = r; *n_out = res.size();
    return NB_OK;
}
