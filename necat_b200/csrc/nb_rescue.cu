// nb_rescue.cu -- the oc2cns -r 1 rescue (consensus/consensus_aux.c:169-201).  First half: the O(nd) local alignment
// of ocda_go (gapped_align/oc_daligner.c:34-79) for a batch of candidates, one warp per candidate.
//
// The wave recurrence itself lives in nb_rescue_core.cuh (lane = diagonal, two generations of the wave vectors in a
// per-warp band in HBM / L2); this file is the persistent kernel around it and the C-ABI entry.  The second half of
// the rescue (the global alignment of the range found here, edlib_go) follows further down; nb_consensus_job calls
// both for -r 1, and both are exported so that they can be checked against the oracle (tests/test_gpu_rescue.py).
#include <vector>

#include "nb_common.cuh"
#include "nb_rescue_core.cuh"

#define RS_BAND_CAP 8192            // diagonals a warp can keep around the start diagonal (overflow is reported)
#define RS_WARPS 4

struct GpuWarp {
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
};

struct RescueParams {
    const uint8_t*  pac;
    const uint64_t* seq_off;
    const uint32_t* seq_len;
    const nb_rescue_task* tasks;
    nb_rescue_result*     results;
    uint32_t        n_tasks;
    uint32_t*       next_task;
    int             min_align_size;
    int             path_ave;
    const int16_t*  tip_table;
    const int16_t*  tip_score;
    int*            band_i;         // per warp: V[2], M[2] (4 x RS_BAND_CAP ints)
    unsigned long long* band_t;     // per warp: T[2]
};

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescue_local(RescueParams P)
{
    const int lane = threadIdx.x & 31;
    const size_t slot = (size_t)blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    GpuWarp w;
    RsBand bd;
    int* bi = P.band_i + slot * 4 * RS_BAND_CAP;
    bd.V[0] = bi; bd.V[1] = bi + RS_BAND_CAP; bd.M[0] = bi + 2 * RS_BAND_CAP; bd.M[1] = bi + 3 * RS_BAND_CAP;
    bd.T[0] = reinterpret_cast<uint64_t*>(P.band_t + slot * 2 * RS_BAND_CAP); bd.T[1] = bd.T[0] + RS_BAND_CAP;
    bd.cap = RS_BAND_CAP; bd.kbase = 0;
    RsTables tb; tb.path_ave = P.path_ave; tb.tip_table = P.tip_table; tb.tip_score = P.tip_score;
    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(P.next_task, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= P.n_tasks) break;
        const nb_rescue_task T = P.tasks[ti];
        RsSeq A; A.pac = P.pac; A.off = P.seq_off[T.qid]; A.len = (int)P.seq_len[T.qid]; A.strand = T.qdir;
        RsSeq B; B.pac = P.pac; B.off = P.seq_off[T.sid]; B.len = (int)P.seq_len[T.sid]; B.strand = 0;
        const int k0 = T.qstart - T.tstart, anti = T.qstart + T.tstart;
        const RsTip f = rs_wave(w, A, B, k0, anti, +1, tb, bd);
        __syncwarp();
        const RsTip r = rs_wave(w, A, B, k0, anti, -1, tb, bd);
        __syncwarp();
        if (lane == 0) {
            nb_rescue_result R;
            R.qbeg = r.apos; R.tbeg = r.bpos; R.qend = f.apos; R.tend = f.bpos; R.diffs = f.diffs + r.diffs;
            R.overflow = f.overflow | r.overflow;
            R.ok = !R.overflow && (R.qend - R.qbeg >= P.min_align_size) && (R.tend - R.tbeg >= P.min_align_size);   // oc_daligner.c:71-74
            R.pad_ = 0;
            P.results[ti] = R;
        }
    }
}

extern "C" int nb_rescue_local(nb_ctx* ctx, const nb_volume* reads, const nb_rescue_task* tasks, uint64_t n, double error,
                               int min_align_size, nb_rescue_result* results)
{
    if (!ctx || !reads || (!tasks && n) || (!results && n)) return nb_fail(ctx, NB_EINVAL, "nb_rescue_local: bad arguments");
    if (n == 0) return NB_OK;
    if (n >= (1ull << 31)) return nb_fail(ctx, NB_ELIMIT, "too many rescue tasks in one call");
    for (uint64_t i = 0; i < n; ++i) {
        const nb_rescue_task& t = tasks[i];
        if (t.qid >= reads->nseq || t.sid >= reads->nseq || t.qstart < 0 || t.tstart < 0 || (uint32_t)t.qstart > reads->h_len[t.qid] ||
            (uint32_t)t.tstart > reads->h_len[t.sid] || (t.qdir != 0 && t.qdir != 1))
            return nb_fail(ctx, NB_EINVAL, "rescue task %llu out of range", (unsigned long long)i);
    }
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<int16_t> tt(1 << 15), ts(1 << 15);
    int path_ave = 0;
    rs_fill_tables(error, &path_ave, tt.data(), ts.data());
    NbDevBuf D(ctx);
    RescueParams P;
    memset(&P, 0, sizeof P);
    int rc;
    nb_rescue_task* d_tasks; nb_rescue_result* d_res; int16_t *d_tt, *d_ts;
    const unsigned grid = NB_SMS * 2;
    const size_t slots = (size_t)grid * RS_WARPS;
    if ((rc = D.get(&d_tasks, n))) return rc;
    if ((rc = D.get(&d_res, n))) return rc;
    if ((rc = D.get(&d_tt, tt.size()))) return rc;
    if ((rc = D.get(&d_ts, ts.size()))) return rc;
    if ((rc = D.get(&P.next_task, 1))) return rc;
    if ((rc = D.get(&P.band_i, slots * 4 * RS_BAND_CAP))) return rc;
    if ((rc = D.get(&P.band_t, slots * 2 * RS_BAND_CAP))) return rc;
    NB_CUDA(ctx, cudaMemcpyAsync(d_tasks, tasks, sizeof(nb_rescue_task) * n, cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemcpyAsync(d_tt, tt.data(), sizeof(int16_t) * tt.size(), cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemcpyAsync(d_ts, ts.data(), sizeof(int16_t) * ts.size(), cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemsetAsync(P.next_task, 0, sizeof(uint32_t), st));
    P.pac = reads->d_pac; P.seq_off = reads->d_off; P.seq_len = reads->d_len;
    P.tasks = d_tasks; P.results = d_res; P.n_tasks = (uint32_t)n; P.min_align_size = min_align_size;
    P.path_ave = path_ave; P.tip_table = d_tt; P.tip_score = d_ts;
    k_rescue_local<<<grid, RS_WARPS * 32, 0, st>>>(P);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(ctx, cudaMemcpyAsync(results, d_res, sizeof(nb_rescue_result) * n, cudaMemcpyDeviceToHost, st));
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Second half of the rescue: edlib_go (edlib/edlib_wrapper.c:111-255) for a batch of ranges.  The recursion is the
// host's (nb_rescue_driver.h, checked against the oracle in emulation); the two kernels below are the persistent-warp
// shells around its two warp-level cores.  The cores and the driver are also run on the host, one thread per lane
// (tests/test_emu_rescue.py); on the GPU, tests/test_gpu_rescue.py::test_rescue_global_matches_oracle compares
// nb_rescue_global with the oracle operation by operation.
#include "nb_rescue_driver.h"

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescue_columns(const RgColJobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t* next_job, int* cols, int8_t* hbuf)
{
    const int lane = threadIdx.x & 31;
    GpuWarp w;
    for (;;) {
        uint32_t ji = 0;
        if (lane == 0) ji = atomicAdd(next_job, 1u);
        ji = __shfl_sync(0xffffffffu, ji, 0);
        if (ji >= n_jobs) break;
        const RgColJobDev J = jobs[ji];
        rs_nw_column(w, J.Q, J.q0, J.qn, J.qstep, J.T, J.t0, J.tn, J.tstep, cols + J.col_off, hbuf + J.h_off, hbuf + J.h_off + J.tn + 1);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescue_leaves(const RgLeafJobDev* __restrict__ jobs, uint32_t n_jobs, uint32_t* next_job, RsWord2* mat, int8_t* hbuf, uint8_t* ops, int* n_ops)
{
    const int lane = threadIdx.x & 31;
    GpuWarp w;
    for (;;) {
        uint32_t ji = 0;
        if (lane == 0) ji = atomicAdd(next_job, 1u);
        ji = __shfl_sync(0xffffffffu, ji, 0);
        if (ji >= n_jobs) break;
        const RgLeafJobDev J = jobs[ji];
        const int n = rs_nw_leaf(w, J.Q, J.q0, J.qn, J.T, J.t0, J.tn, mat + J.mat_off, hbuf + J.h_off, hbuf + J.h_off + J.tn + 1, ops + J.ops_off, J.cap);
        if (lane == 0) n_ops[ji] = n;
        __syncwarp();
    }
}

namespace {

// what actually runs a slice of pooled jobs on the GPU; the pooling itself (RgPooledBackend, nb_rescue_driver.h) is shared
// with the host emulation
struct GpuRescueExec {
    nb_ctx* ctx;
    int rc = NB_OK;

    bool run_columns(const std::vector<RgColJobDev>& dj, unsigned long long col_total, unsigned long long h_total, std::vector<int>& h_cols)
    {
        cudaStream_t st = ctx->stream;
        NbDevBuf D(ctx);
        RgColJobDev* d_jobs; int* d_cols; int8_t* d_h; uint32_t* d_next;
        if ((rc = D.get(&d_jobs, dj.size())) || (rc = D.get(&d_cols, col_total)) || (rc = D.get(&d_h, h_total + 16)) || (rc = D.get(&d_next, 1))) return false;
        h_cols.resize((size_t)col_total);
        cudaMemcpyAsync(d_jobs, dj.data(), sizeof(RgColJobDev) * dj.size(), cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(d_next, 0, sizeof(uint32_t), st);
        k_rescue_columns<<<NB_SMS * 4, RS_WARPS * 32, 0, st>>>(d_jobs, (uint32_t)dj.size(), d_next, d_cols, d_h);
        ctx->launches++;
        cudaMemcpyAsync(h_cols.data(), d_cols, sizeof(int) * (size_t)col_total, cudaMemcpyDeviceToHost, st);
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = nb_fail(ctx, NB_ECUDA, "k_rescue_columns -> %s", cudaGetErrorString(e)); return false; }
        return true;
    }

    bool run_leaves(const std::vector<RgLeafJobDev>& dj, unsigned long long mat_total, unsigned long long h_total, unsigned long long ops_total,
                    std::vector<uint8_t>& h_ops, std::vector<int>& h_n)
    {
        cudaStream_t st = ctx->stream;
        NbDevBuf D(ctx);
        RgLeafJobDev* d_jobs; RsWord2* d_mat; int8_t* d_h; uint8_t* d_ops; int* d_n; uint32_t* d_next;
        if ((rc = D.get(&d_jobs, dj.size())) || (rc = D.get(&d_mat, mat_total + 1)) || (rc = D.get(&d_h, h_total + 16)) ||
            (rc = D.get(&d_ops, ops_total + 16)) || (rc = D.get(&d_n, dj.size())) || (rc = D.get(&d_next, 1))) return false;
        h_ops.resize((size_t)ops_total);
        h_n.resize(dj.size());
        cudaMemcpyAsync(d_jobs, dj.data(), sizeof(RgLeafJobDev) * dj.size(), cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(d_next, 0, sizeof(uint32_t), st);
        k_rescue_leaves<<<NB_SMS * 4, RS_WARPS * 32, 0, st>>>(d_jobs, (uint32_t)dj.size(), d_next, d_mat, d_h, d_ops, d_n);
        ctx->launches++;
        cudaMemcpyAsync(h_ops.data(), d_ops, (size_t)ops_total, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(h_n.data(), d_n, sizeof(int) * dj.size(), cudaMemcpyDeviceToHost, st);
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = nb_fail(ctx, NB_ECUDA, "k_rescue_leaves -> %s", cudaGetErrorString(e)); return false; }
        return true;
    }
};

}  // namespace

extern "C" int nb_rescue_global(nb_ctx* ctx, const nb_volume* reads, const nb_rescue_range* ranges, uint64_t n, double error,
                                int min_align_size, int k_match, nb_alignment** out_aln, uint8_t** out_ops, uint64_t* out_ops_bytes)
{
    if (!ctx || !reads || (!ranges && n) || !out_aln || !out_ops || !out_ops_bytes) return nb_fail(ctx, NB_EINVAL, "nb_rescue_global: bad arguments");
    *out_aln = NULL; *out_ops = NULL; *out_ops_bytes = 0;
    if (n == 0) return NB_OK;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<RgTask> tasks((size_t)n);
    for (uint64_t i = 0; i < n; ++i) {
        const nb_rescue_range& r = ranges[i];
        if (r.qid >= reads->nseq || r.sid >= reads->nseq || (r.qdir != 0 && r.qdir != 1) || r.qfrom < 0 || r.tfrom < 0 || r.qto < r.qfrom || r.tto < r.tfrom ||
            (uint32_t)r.qto > reads->h_len[r.qid] || (uint32_t)r.tto > reads->h_len[r.sid])
            return nb_fail(ctx, NB_EINVAL, "rescue range %llu out of bounds", (unsigned long long)i);
        RgTask& t = tasks[(size_t)i];
        t.Q.pac = reads->d_pac; t.Q.off = reads->h_off[r.qid]; t.Q.len = (int)reads->h_len[r.qid]; t.Q.strand = r.qdir;
        t.T.pac = reads->d_pac; t.T.off = reads->h_off[r.sid]; t.T.len = (int)reads->h_len[r.sid]; t.T.strand = 0;
        t.qfrom = r.qfrom; t.qto = r.qto; t.tfrom = r.tfrom; t.tto = r.tto; t.tolerance = r.tolerance;
    }
    GpuRescueExec ex;
    ex.ctx = ctx;
    RgPooledBackend<GpuRescueExec> be(ex);
    std::vector<RgResult> res;
    rg_run(be, tasks, error, min_align_size, k_match, res);
    if (ex.rc) return ex.rc;
    uint64_t total = 0;
    for (const RgResult& r : res) total += r.ops.size();
    nb_alignment* aln = (nb_alignment*)calloc((size_t)n, sizeof(nb_alignment));
    uint8_t* ops = (uint8_t*)malloc(total ? total : 1);
    if (!aln || !ops) { free(aln); free(ops); return nb_fail(ctx, NB_ENOMEM, "out of host memory"); }
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const RgResult& r = res[(size_t)i];
        nb_alignment& a = aln[i];
        a.ok = r.ok; a.qoff = r.qoff; a.qend = r.qend; a.toff = r.toff; a.tend = r.tend; a.aln_size = r.aln_size; a.n_match = r.n_match;
        a.long_ins = 0; a.ops_offset = off;
        if (!r.ops.empty()) memcpy(ops + off, r.ops.data(), r.ops.size());
        off += r.ops.size();
    }
    *out_aln = aln; *out_ops = ops; *out_ops_bytes = total;
    return NB_OK;
}
