// nb_rescue.cu -- first half of the oc2cns -r 1 rescue (consensus/consensus_aux.c:169-178): the O(nd) local alignment
// of ocda_go (gapped_align/oc_daligner.c:34-79) for a batch of candidates, one warp per candidate.
//
// The wave recurrence itself lives in nb_rescue_core.cuh (lane = diagonal, two generations of the wave vectors in a
// per-warp band in HBM / L2); this file is the persistent kernel around it and the C-ABI entry.  The second half of
// the rescue (the global alignment of the range found here) is not on the GPU yet, so nb_consensus_job still refuses
// -r 1; this entry point exists so that the kernel can be checked against the oracle (tests/test_gpu_rescue.py).
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
