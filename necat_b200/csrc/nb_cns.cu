// nb_cns.cu -- nb_consensus_job: one oc2cns partition on the GPU.
//
// Replaces consensus_one_partition / consensus_one_read
// (consensus/consensus_one_partition.c:110-201, consensus/consensus_one_read.c:216-392).
// Division of labour:
//   * GPU: every candidate extension (k_extend, nb_extend.cu) and the whole
//     consensus (k_consensus, nb_cns_kernel.cuh);
//   * host (C++, this file): the order-dependent accept/skip logic of
//     consensus_one_read.  An extension's result does not depend on that state,
//     only whether it is used does -- so extensions are computed speculatively in
//     rounds (first 50 candidates of every template, then the remaining windows of
//     50 minus the candidates that are already skippable: both skip tests are
//     monotone) and the reference's loop is replayed over the results.
#include <algorithm>
#include <cmath>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "nb_cns_kernel.cuh"

#include <chrono>

namespace {

struct HostTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap() { auto t1 = std::chrono::steady_clock::now(); double d = std::chrono::duration<double, std::milli>(t1 - t0).count(); t0 = t1; return d; }
};

struct Can {
    int qid, qdir, qoff, soff, sbeg, send, qsize, score, qbeg, qend;
    int res;            // index into the template chunk's result arrays, -1 = not extended yet
};

struct Accepted { int can; double weight; };

struct Tpl {
    int tid, T;
    size_t first, n;            // candidates (sorted, truncated) in the chunk's `cans`
    // replay state
    int stage;                  // 0 = phase 1 pending, 1 = main loop, 2 = finished
    size_t next, win_to, win_i;
    bool in_window;
    double cutoff;
    int num_can, num_ovlps;
    std::vector<uint16_t> cov;
    int n_below;                // positions whose coverage is still below max_cov
    std::vector<int> used;
    std::vector<Accepted> acc;
    std::vector<size_t> request; // candidate indices to extend in the next round
};

inline bool good_overlap(int qoff, int qend, int qsize, int soff, int send, int ssize)   // error_estimate.c:8-30
{
    const int M = 200;
    const int ql = qoff, qr = qsize - qend, sl = soff, sr = ssize - send;
    return (ql <= M && qr <= M) || (sl <= M && sr <= M) || (qr <= M && sl <= M) || (sr <= M && ql <= M);
}
inline bool full_cov_ovlp(int ql, int qr, int qs, int tl, int tr, int ts, int L, int M)  // consensus_aux.c:93-112
{
    if (ql <= M && qs - qr <= M) return true;
    if (tl <= M && ts - tr <= M) return true;
    if (qs - qr <= M) { if (tl > M) return false; if (qr - ql >= L) return true; }
    if (ts - tr <= M) { if (ql > M) return false; if (qr - ql >= L) return true; }
    return false;
}
inline bool mapping_range_ok(int ql, int qr, int qs, int tl, int tr, int ts, int min_size, double ratio)  // :115-121
{
    if (qr - ql >= min_size || tr - tl >= min_size) return true;
    return (qr - ql >= qs * ratio) || (tr - tl >= ts * ratio);
}
inline double cns_weight(double ident)      // consensus_one_read.c:11-16
{
    double e = (100.0 - ident) / 100.0 / 2.0;
    double w = (1.0 - e) * (1.0 - e) + e * e / 3.0;
    if (100.0 - ident <= 1.0e-6) w = 1.0;
    return w;
}
inline double ident_of(const nb_alignment& a) { return a.aln_size ? 100.0 * a.n_match / a.aln_size : 0.0; }

// "was this read already extended against the template" (ReadIdPool, consensus_aux.h): the template keeps the list, the
// thread that replays it a stamp per read id (set up from the list when the replay starts) so that a look-up is one load.
struct UsedStamps {
    std::vector<uint32_t> stamp;
    uint32_t epoch = 0;
    void begin(const Tpl& t, size_t nseq)
    {
        if (stamp.size() < nseq) stamp.assign(nseq, 0);
        if (++epoch == 0) { std::fill(stamp.begin(), stamp.end(), 0u); epoch = 1; }
        for (int u : t.used) stamp[(size_t)u] = epoch;
    }
};
static thread_local UsedStamps tl_used;
inline bool is_used(const Tpl&, int qid) { return tl_used.stamp[(size_t)qid] == tl_used.epoch; }
inline void set_used(Tpl& t, int qid) { if (!is_used(t, qid)) { t.used.push_back(qid); tl_used.stamp[(size_t)qid] = tl_used.epoch; } }
// The two per-base loops of the replay (2 G coverage increments on the E. coli-size set): plain loops the compiler turns
// into 16-lane AVX2 code where the host has it (chosen once, at load time), scalar otherwise.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__CUDA_ARCH__)
#define NB_SIMD_TARGET __attribute__((target("avx2"), optimize("O3")))
#else
#define NB_SIMD_TARGET
#endif
static int cov_add_plain(uint16_t* c, int n, int max_cov) { int r = 0; for (int i = 0; i < n; ++i) r += (++c[i] == max_cov); return r; }
static int cov_below_plain(const uint16_t* c, int n, int max_cov) { for (int i = 0; i < n; ++i) if (c[i] < max_cov) return 1; return 0; }
NB_SIMD_TARGET static int cov_add_simd(uint16_t* __restrict__ c, int n, int max_cov)
{
    const uint16_t m = (uint16_t)max_cov;
    int r = 0;
    for (int i = 0; i < n; ++i) { const uint16_t v = (uint16_t)(c[i] + 1); c[i] = v; r += v == m; }
    return r;
}
NB_SIMD_TARGET static int cov_below_simd(const uint16_t* __restrict__ c, int n, int max_cov)
{
    int i = 0;
    for (; i + 64 <= n; i += 64) {                  // whole 64-base pieces without an early exit inside: they vectorise
        int any = 0;
        for (int k = 0; k < 64; ++k) any |= c[i + k] < max_cov;
        if (any) return 1;
    }
    for (; i < n; ++i) if (c[i] < max_cov) return 1;
    return 0;
}
static bool host_has_avx2()
{
#if defined(__x86_64__) && defined(__GNUC__)
    return __builtin_cpu_supports("avx2");
#else
    return false;
#endif
}
static int (*const cov_add)(uint16_t*, int, int) = host_has_avx2() ? cov_add_simd : cov_add_plain;
static int (*const cov_below)(const uint16_t*, int, int) = host_has_avx2() ? cov_below_simd : cov_below_plain;

inline bool cov_full(const Tpl& t, int from, int to, int max_cov)   // consensus_one_read.c:139-145
{
    return to <= from || !cov_below(t.cov.data() + from, to - from, max_cov);
}
inline void accept(Tpl& t, size_t ci, const nb_alignment& a, int max_cov)
{
    Accepted ac; ac.can = (int)ci; ac.weight = cns_weight(ident_of(a));
    t.acc.push_back(ac);
    if (a.tend > a.toff) t.n_below -= cov_add(t.cov.data() + a.toff, a.tend - a.toff, max_cov);
}

// Replays consensus_one_read for one template as far as the available extension results allow.
void replay(Tpl& t, std::vector<Can>& cans, const std::vector<nb_alignment>& res, const nb_cns_options& o, size_t nseq)
{
    tl_used.begin(t, nseq);
    t.request.clear();
    Can* c = cans.data() + t.first;
    if (t.stage == 0) {
        if (o.use_fixed_ident_cutoff) {                     // :276-281
            t.cutoff = 100.0 * (1.0 - o.error); t.next = 0; t.num_can = 0; t.num_ovlps = 0;
        } else {
            // get_good_overlaps, error_estimate.c:96-184
            std::vector<size_t> pool;
            double ident[15]; int n_ident = 0;
            size_t i;
            for (i = 0; i < t.n && i < 50; ++i) {
                if (is_used(t, c[i].qid)) continue;
                const nb_alignment& a = res[(size_t)c[i].res];
                if (!a.ok) continue;
                pool.push_back(i);
                set_used(t, c[i].qid);
                if (good_overlap(a.qoff, a.qend, c[i].qsize, a.toff, a.tend, t.T)) {
                    ident[n_ident++] = ident_of(a);
                    if (n_ident == 15) break;
                }
            }
            t.next = i;
            if (n_ident < 15) {                             // get_idents, :66-93
                int m = 0;
                for (size_t p : pool) {
                    const nb_alignment& a = res[(size_t)c[p].res];
                    if (good_overlap(a.qoff, a.qend, c[p].qsize, a.toff, a.tend, t.T)) { ident[m++] = ident_of(a); if (m == 15) break; }
                }
                if (m < 15) {
                    m = 0;
                    for (size_t p : pool) {
                        const nb_alignment& a = res[(size_t)c[p].res];
                        if ((a.qend - a.qoff >= c[p].qsize * 0.6) || (a.tend - a.toff >= t.T * 0.6)) { ident[m++] = ident_of(a); if (m == 15) break; }
                    }
                }
                n_ident = m;
            }
            std::sort(ident, ident + n_ident, [](double x, double y) { return x > y; });
            if (n_ident >= 8) n_ident = (int)(n_ident * 0.7);
            if (n_ident < 5) t.cutoff = 0.0;                 // :38-43
            else {
                double sum = 0.0;
                for (int x = 0; x < n_ident; ++x) sum += ident[x];
                const double avg = sum / n_ident;
                double se = 0.0;
                for (int x = 0; x < n_ident; ++x) se += (avg - ident[x]) * (avg - ident[x]);
                se /= n_ident;
                se = std::sqrt(se);
                t.cutoff = avg - se * 5;
            }
            t.num_ovlps = 0;
            for (size_t p : pool) {                         // add_extended_overlaps, consensus_one_read.c:148-183
                const nb_alignment& a = res[(size_t)c[p].res];
                if (ident_of(a) < t.cutoff) continue;
                if (!mapping_range_ok(a.qoff, a.qend, c[p].qsize, a.toff, a.tend, t.T, o.min_align_size, o.mapping_ratio)) continue;
                ++t.num_ovlps;
                accept(t, p, a, o.max_cov);
                set_used(t, c[p].qid);
            }
            t.num_can = (int)t.next;
        }
        t.stage = 1;
        t.in_window = false;
    }
    // main loop, consensus_one_read.c:316-372
    for (;;) {
        if (!t.in_window) {
            if (t.next >= t.n || t.n_below == 0) { t.stage = 2; return; }     // region_coverage_is_full(0, T)
            t.win_i = t.next;
            t.win_to = std::min(t.next + 50, t.n);
            t.next = t.win_to;
            t.in_window = true;
        }
        for (; t.win_i < t.win_to; ++t.win_i) {
            const size_t i = t.win_i;
            if (is_used(t, c[i].qid)) continue;
            if (cov_full(t, c[i].sbeg, c[i].send, o.max_cov)) continue;
            if (c[i].res < 0) {
                // result missing: ask for everything in the window that is not skippable right now
                for (size_t k = i; k < t.win_to; ++k) {
                    if (c[k].res >= 0 || is_used(t, c[k].qid)) continue;
                    if (cov_full(t, c[k].sbeg, c[k].send, o.max_cov)) continue;
                    t.request.push_back(k);
                }
                return;
            }
            const nb_alignment& a = res[(size_t)c[i].res];
            ++t.num_can;
            if (!a.ok) continue;
            const double id = ident_of(a);
            if (id < t.cutoff && !full_cov_ovlp(a.qoff, a.qend, c[i].qsize, a.toff, a.tend, t.T, 5000, 100)) continue;
            if (!mapping_range_ok(a.qoff, a.qend, c[i].qsize, a.toff, a.tend, t.T, o.min_align_size, o.mapping_ratio)) continue;
            ++t.num_ovlps;
            accept(t, i, a, o.max_cov);
            set_used(t, c[i].qid);
        }
        t.in_window = false;
    }
}

inline int host_base(const nb_volume* v, uint64_t p) { return (v->h_pac[p >> 2] >> (((~p) & 3) << 1)) & 3; }

// Output sequences are written in bulk: records are planned first (offsets into one growing
// buffer), then the bytes are filled by several host threads.
struct Fill {
    uint64_t dst;       // offset in the output buffer
    uint32_t len;
    uint32_t kind;      // 0 = copy from the chunk's ASCII consensus buffer at src, 1 = decode read bases from volume position src
    uint64_t src;
};

struct SeqOut {
    char*    buf = NULL;
    uint64_t cap = 0, used = 0;
    bool reserve(uint64_t need)
    {
        if (need <= cap) return true;
        uint64_t nc = cap ? cap : (1u << 20);
        while (nc < need) nc += nc / 2 + (1u << 20);
        char* nb = (char*)realloc(buf, nc);
        if (!nb) return false;
        buf = nb; cap = nc;
        return true;
    }
};

void run_fills(const std::vector<Fill>& fills, char* out, const char* ascii, const nb_volume* v, unsigned nthreads)
{
    if (fills.empty()) return;
    auto work = [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            const Fill& f = fills[i];
            char* d = out + f.dst;
            if (f.kind == 0) memcpy(d, ascii + f.src, f.len);
            else {
                uint64_t p = f.src;
                for (uint32_t k = 0; k < f.len; ++k, ++p) d[k] = "ACGT"[(v->h_pac[p >> 2] >> (((~p) & 3) << 1)) & 3];
            }
        }
    };
    uint64_t total = 0;
    for (const Fill& f : fills) total += f.len;
    std::vector<std::thread> th;
    const uint64_t per = total / nthreads + 1;
    size_t a = 0;
    while (a < fills.size()) {
        uint64_t acc = 0;
        size_t b = a;
        while (b < fills.size() && acc < per) acc += fills[b++].len;
        th.emplace_back(work, a, b);
        a = b;
    }
    for (auto& t : th) t.join();
}

// codes (0..3) of the consensus stretches -> ASCII, gathered densely: one CTA per stretch
__global__ void k_gather_ascii(const uint8_t* __restrict__ pool, const uint64_t* __restrict__ src, const uint64_t* __restrict__ dst,
                               const uint32_t* __restrict__ len, uint32_t n, char* __restrict__ out)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint8_t* s = pool + src[i];
        char* d = out + dst[i];
        for (uint32_t k = threadIdx.x; k < len[i]; k += blockDim.x) d[k] = "ACGT"[s[k] & 3];
    }
}

}  // namespace

// Shared between the pipeline workers of one job.
struct JobShared {
    char*    buf;                       // output sequences: one allocation sized by an upper bound, dense prefix in use
    uint64_t cap;
    std::atomic<uint64_t> used;
    std::vector<uint8_t>* corrected;
    size_t   n_long;                    // the first n_long templates of the job are its longest (0 = not arranged)
    std::mutex gpu;                     // one chunk at a time submits to the GPU and waits for it: the kernels are sized
                                        // to fill the machine, running two side by side only makes both slower;
                                        // what overlaps is the host work of one chunk with the kernels of another
};

struct ChunkTimes { double replay = 0, gather = 0, ext = 0, prep = 0, kern = 0, rec = 0, wait = 0, ext_wall = 0, init = 0; float ext_ms = 0, cns_ms = 0; };

// Extension rounds + consensus + records of the templates [c0, c1): everything the GPU does for them runs on the
// stream of `ctx` (a pipeline worker: its own stream, events, staging buffer and walk-back scratch).
static int process_chunk(nb_ctx* ctx, const nb_volume* reads, const nb_cns_options& o, std::vector<Tpl>& tpls, std::vector<Can>& cans,
                         size_t c0, size_t c1, JobShared& sh, std::vector<nb_cns_record>& recs, ChunkTimes& tms, unsigned nthreads)
{
    HostTimer ht;
    double t_replay = 0, t_gather = 0, t_ext = 0, t_prep = 0, t_kern = 0, t_rec = 0, t_wait = 0, t_init = 0;
    float ext_ms = 0, cns_ms = 0;
    std::vector<uint8_t>& corrected = *sh.corrected;
    // ---- extension rounds.  The templates of the chunk form up to NB_MAX_WORKERS groups, each driven by its own host
    //      thread on its own stream; one group at a time holds the GPU (sh.gpu), so the kernels run alone and at full
    //      width while the other groups replay the accept/skip logic and gather their next requests on the host.
    struct ExtGroup {
        size_t t0 = 0, t1 = 0;
        nb_ctx* wc = NULL;
        NbDevBuf* D = NULL;                     // the groups' ops pools stay alive until the consensus has read them
        std::vector<nb_alignment> res;          // results of all rounds of this group
        std::vector<const uint8_t*> res_ops;    // device pointer of each result's ops
        ChunkTimes tm;
        int rc = NB_OK;
        std::atomic<int>* round0_done = NULL;    // set once the group's first (by far largest) extension launch has finished
        std::atomic<int>* submitted = NULL;      // pipelined groups: set when the group's first launch is being queued
        std::atomic<int>* prev_submitted = NULL; // ... that of the group before
    };
    const char* env_groups = getenv("NB_CNS_GROUPS");
    // default 1: with the replay's per-base loops vectorised, what several groups hide of the host phases is less than what their
    // smaller extension launches lose (measured, E. coli-size set: 355 ms per step with one group, 378 with three)
    int n_groups = env_groups ? atoi(env_groups) : 1;
    const char* env_gmin = getenv("NB_CNS_GROUPS_MIN");       // (tests lower it to run small sets through several groups)
    if (ctx->is_worker || c1 - c0 < (size_t)(env_gmin ? atoi(env_gmin) : 1024)) n_groups = 1;   // workers of the chunk pipeline have no workers of their own
    if (n_groups > NB_MAX_WORKERS) n_groups = NB_MAX_WORKERS;
    if (n_groups < 1) n_groups = 1;
    std::vector<ExtGroup> groups((size_t)n_groups);
    // the job put its longest templates first (sh.n_long of them): they form group 0 of the first chunk
    const bool long_mode = n_groups >= 2 && c0 == 0 && sh.n_long > 0 && sh.n_long < c1;
    // Pipelined groups (several groups, not the long-template mode): group k takes every n-th template of the
    // length-sorted list, so the groups are alike; each has its own stream, walk-back scratch and host thread, nothing is
    // locked.  The first extension launches are queued in group order: the persistent CTAs of group k + 1 move in as those
    // of group k drain (no idle tail), and while group k + 1 extends, the host replays group k and prepares its consensus,
    // whose kernels in turn run behind group k + 1's extension.
    static const int kOverlap = getenv("NB_CNS_OVERLAP") ? atoi(getenv("NB_CNS_OVERLAP")) : 1;
    const bool overlap = kOverlap && n_groups > 1 && !long_mode;
    if (overlap) {
        const size_t n = c1 - c0, ng = (size_t)n_groups;
        std::vector<Tpl> re;
        re.reserve(n);
        size_t t = c0;
        for (size_t k = 0; k < ng; ++k) {
            groups[k].t0 = t;
            for (size_t i = k; i < n; i += ng) { re.push_back(std::move(tpls[c0 + i])); ++t; }
            groups[k].t1 = t;
        }
        for (size_t i = 0; i < n; ++i) tpls[c0 + i] = std::move(re[i]);
    } else {
        const size_t r0 = long_mode ? sh.n_long : c0;
        const int first = long_mode ? 1 : 0;
        if (long_mode) { groups[0].t0 = c0; groups[0].t1 = r0; }
        uint64_t total = 0, acc = 0;                        // equal numbers of first-round extensions x template length
        auto weight = [&](size_t t) { return (uint64_t)std::min<size_t>(tpls[t].n, 50) * (uint64_t)tpls[t].T; };
        for (size_t t = r0; t < c1; ++t) total += weight(t);
        size_t t = r0;
        for (int k = first; k < n_groups; ++k) {
            groups[(size_t)k].t0 = t;
            const uint64_t want = total * (uint64_t)(k - first + 1) / (uint64_t)(n_groups - first);
            while (t < c1 && (acc < want || k == n_groups - 1)) { acc += weight(t); ++t; }
            groups[(size_t)k].t1 = t;
        }
    }
    std::vector<NbDevBuf> group_bufs;
    group_bufs.reserve((size_t)n_groups);
    for (int k = 0; k < n_groups; ++k) {
        ExtGroup& g = groups[(size_t)k];
        g.wc = ctx;
        if (k > 0) { const int rc = nb_worker_ctx(ctx, k, &g.wc); if (rc) return rc; }
        // with the long group's consensus running beside them, the other groups' extension kernels leave it room on the SMs
        g.wc->ext_ctas = (long_mode && k > 0) ? 2 : 0;
        g.wc->ext_private = (overlap && k > 0) ? 1 : 0;
        group_bufs.emplace_back(g.wc);
        g.D = &group_bufs.back();
    }
    auto ext_phase = [&](ExtGroup& g) -> int {
        nb_ctx* wc = g.wc;
        if (cudaSetDevice(wc->device) != cudaSuccess) return nb_fail(wc, NB_ECUDA, "cudaSetDevice failed");
        cudaStream_t gst = wc->stream;
        HostTimer gt;
        // The round-0 requests: the first 50 candidates (phase 1), or nothing with -u 1 (the replay then asks for the first
        // window).  The coverage arrays (2 bytes per template base: hundreds of MB to clear) are only needed by the first
        // replay: they are set up by side threads while the first extension launch runs.
        std::vector<std::thread> cov_threads;
        {
            auto init = [&tpls, &o](size_t a, size_t b) {
                for (size_t t = a; t < b; ++t) { Tpl& T = tpls[t]; T.cov.assign((size_t)T.T + 1, 0); T.n_below = o.max_cov > 0 ? T.T : 0; }
            };
            const unsigned nside = std::max(1u, nthreads / 2);
            const size_t nt = g.t1 - g.t0, per = (nt + nside - 1) / nside;
            for (unsigned w = 0; w < nside && per; ++w) {
                const size_t a = g.t0 + w * per, b = std::min(g.t1, a + per);
                if (a >= b) break;
                cov_threads.emplace_back(init, a, b);
            }
        }
        struct Joiner { std::vector<std::thread>& th; ~Joiner() { for (auto& x : th) if (x.joinable()) x.join(); } } cov_joiner{cov_threads};
        g.tm.prep += gt.lap();
        for (int round = 0;; ++round) {
            if (round == 0 && o.use_fixed_ident_cutoff) continue;         // -u 1: nothing to extend before the first replay
            if (round > 0) {
                for (auto& x : cov_threads) if (x.joinable()) x.join();
                // replay in parallel over templates
                std::vector<std::thread> th;
                std::vector<size_t> todo;
                for (size_t t = g.t0; t < g.t1; ++t) if (tpls[t].stage != 2) todo.push_back(t);
                // the templates are sorted by length: handed out a few at a time, not in equal ranges
                std::atomic<size_t> next_todo(0);
                auto work = [&]() {
                    for (;;) {
                        const size_t a = next_todo.fetch_add(8), b = std::min(todo.size(), a + 8);
                        if (a >= todo.size()) break;
                        for (size_t x = a; x < b; ++x) replay(tpls[todo[x]], cans, g.res, o, (size_t)reads->nseq);
                    }
                };
                const unsigned nw = (unsigned)std::min<size_t>(nthreads, (todo.size() + 7) / 8);
                for (unsigned w = 1; w < nw; ++w) th.emplace_back(work);
                work();
                for (auto& x : th) x.join();
            }
            g.tm.replay += gt.lap();
            // gather requests: sizes per template, prefix, fill in parallel
            std::vector<ExtTask> tasks;
            std::vector<size_t> task_can;
            uint64_t pool = 0;
            {
                const size_t nt = g.t1 - g.t0;
                std::vector<uint64_t> t_cnt(nt + 1, 0), t_bytes(nt + 1, 0);
                std::atomic<int> bad(-1);
                auto each = [&](size_t a, size_t b, bool fill) {
                    for (size_t x = a; x < b; ++x) {
                        Tpl& T = tpls[g.t0 + x];
                        if (T.stage == 2) continue;
                        uint64_t n = 0, bytes = 0;
                        // round 0 asks for the first 50 candidates of every template (phase 1); later rounds for what the replay listed
                        const size_t nreq = round == 0 ? std::min<size_t>(T.n, 50) : T.request.size();
                        for (size_t ri = 0; ri < nreq; ++ri) {
                            const size_t k = round == 0 ? ri : T.request[ri];
                            const Can& c = cans[T.first + k];
                            const uint64_t mn = (uint64_t)std::min(c.qoff, c.soff);
                            const uint32_t cap = (uint32_t)((uint64_t)c.qsize + (uint64_t)T.T + mn + 64);
                            if (fill) {
                                ExtTask e;
                                e.qid = (uint32_t)c.qid; e.sid = (uint32_t)T.tid; e.qoff = c.qoff; e.soff = c.soff; e.qdir = c.qdir; e.pad_ = 0;
                                e.lcap = (uint32_t)(c.qoff + c.soff);
                                e.ops_cap = cap;
                                e.ops_base = t_bytes[x] + bytes;
                                tasks[t_cnt[x] + n] = e;
                                task_can[t_cnt[x] + n] = T.first + k;
                            } else if (c.qoff < 0 || c.soff < 0 || c.qoff > c.qsize || c.soff > T.T) bad.store(T.tid);
                            ++n; bytes += (cap + 15) & ~15ull;
                        }
                        if (!fill) { t_cnt[x + 1] = n; t_bytes[x + 1] = bytes; }
                    }
                };
                auto spread = [&](bool fill) {
                    const size_t per = (nt + nthreads - 1) / nthreads;
                    std::vector<std::thread> th;
                    for (unsigned w = 1; w < nthreads && per; ++w) {
                        const size_t a = w * per, b = std::min(nt, a + per);
                        if (a >= b) break;
                        th.emplace_back(each, a, b, fill);
                    }
                    if (per) each(0, std::min(nt, per), fill);
                    for (auto& x : th) x.join();
                };
                spread(false);
                if (bad.load() >= 0) return nb_fail(wc, NB_EINVAL, "candidate offsets out of range (template %d)", bad.load());
                for (size_t x = 0; x < nt; ++x) { t_cnt[x + 1] += t_cnt[x]; t_bytes[x + 1] += t_bytes[x]; }
                tasks.resize(t_cnt[nt]); task_can.resize(t_cnt[nt]);
                pool = t_bytes[nt];
                if (!tasks.empty()) spread(true);
            }
            g.tm.gather += gt.lap();
            if (tasks.empty()) break;
            if (tasks.size() >= (1ull << 31)) return nb_fail(wc, NB_ELIMIT, "too many extension tasks in one round");
            ExtTask* d_tasks; nb_alignment* d_res; uint8_t* d_ops;
            int rc;
            if ((rc = g.D->get(&d_tasks, tasks.size()))) return rc;
            if ((rc = g.D->get(&d_res, tasks.size()))) return rc;
            if ((rc = g.D->get(&d_ops, pool + 64))) return rc;
            std::unique_lock<std::mutex> gpu(sh.gpu, std::defer_lock);
            if (!overlap) gpu.lock();
            else if (round == 0 && g.prev_submitted) while (!g.prev_submitted->load()) std::this_thread::yield();   // launches queue in group order
            g.tm.wait += gt.lap();
            NB_CUDA(wc, cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(ExtTask) * tasks.size(), cudaMemcpyHostToDevice, gst));
            cudaEventRecord(wc->ev0, gst);
            if (round == 0 && g.submitted) g.submitted->store(1);
            rc = nb_run_extension(wc, reads, reads, d_tasks, (uint32_t)tasks.size(), d_res, d_ops, o.error, o.min_align_size, 4 /* ONC_TAIL_MATCH_LEN_LONG */);
            if (rc) return rc;
            cudaEventRecord(wc->ev1, gst);
            const size_t r0 = g.res.size();
            g.res.resize(r0 + tasks.size());
            NB_CUDA(wc, cudaMemcpyAsync(g.res.data() + r0, d_res, sizeof(nb_alignment) * tasks.size(), cudaMemcpyDeviceToHost, gst));
            NB_CUDA(wc, cudaStreamSynchronize(gst));
            std::vector<const uint8_t*> rescued_ops(tasks.size(), (const uint8_t*)NULL);
            if (o.rescue_long_indels) {
                // -r 1: the rescue branch of cns_extension, consensus_aux.c:150-201.  Candidates of this
                // round whose block-wise extension failed, or left more than 200 bases of their chain unaligned, go through
                // the O(nd) local alignment and, if that spans enough, the global alignment of its range; where both succeed
                // the result replaces the block-wise one.
                std::vector<nb_rescue_task> rt;
                std::vector<size_t> rt_k;
                for (size_t k = 0; k < tasks.size(); ++k) {
                    const nb_alignment& a = g.res[r0 + k];
                    const Can& c = cans[task_can[k]];
                    bool fine = a.ok != 0;
                    if (fine) {
                        const int lhang = a.qoff > c.qbeg ? a.qoff - c.qbeg : 0, rhang = a.qend < c.qend ? c.qend - a.qend : 0;
                        if (lhang + rhang > 200) fine = false;
                    }
                    if (fine) continue;
                    nb_rescue_task t;
                    t.qid = tasks[k].qid; t.sid = tasks[k].sid; t.qdir = tasks[k].qdir; t.qstart = tasks[k].qoff; t.tstart = tasks[k].soff; t.pad_ = 0;
                    rt.push_back(t); rt_k.push_back(k);
                }
                if (!rt.empty()) {
                    std::vector<nb_rescue_result> rr(rt.size());
                    if ((rc = nb_rescue_local(wc, reads, rt.data(), rt.size(), o.error, o.min_align_size, rr.data()))) return rc;
                    std::vector<nb_rescue_range> rg;
                    std::vector<size_t> rg_k;
                    for (size_t x = 0; x < rt.size(); ++x) {
                        if (rr[x].overflow) return nb_fail(wc, NB_ELIMIT, "rescue: the wave band of a candidate exceeds its capacity");
                        if (!rr[x].ok) continue;
                        nb_rescue_range q;
                        q.qid = rt[x].qid; q.sid = rt[x].sid; q.qdir = rt[x].qdir; q.qfrom = rr[x].qbeg; q.qto = rr[x].qend;
                        q.tfrom = rr[x].tbeg; q.tto = rr[x].tend; q.tolerance = rr[x].diffs;
                        rg.push_back(q); rg_k.push_back(rt_k[x]);
                    }
                    if (!rg.empty()) {
                        nb_alignment* ga = NULL; uint8_t* gops = NULL; uint64_t gbytes = 0;
                        if ((rc = nb_rescue_global(wc, reads, rg.data(), rg.size(), o.error, o.min_align_size, 4, &ga, &gops, &gbytes))) return rc;
                        uint8_t* d_rops = NULL;
                        if (gbytes) {
                            if ((rc = g.D->get(&d_rops, gbytes))) { free(ga); free(gops); return rc; }
                            const cudaError_t e1 = cudaMemcpyAsync(d_rops, gops, gbytes, cudaMemcpyHostToDevice, gst);
                            const cudaError_t e2 = cudaStreamSynchronize(gst);
                            if (e1 != cudaSuccess || e2 != cudaSuccess) { free(ga); free(gops); return nb_fail(wc, NB_ECUDA, "rescue: upload of the operations failed"); }
                        }
                        for (size_t x = 0; x < rg.size(); ++x) {
                            if (!ga[x].ok) continue;
                            nb_alignment a = ga[x];
                            // a run of >= 255 query-only columns drops the alignment from the tags (align_tags.c:40-44)
                            int run = 0; a.long_ins = 0;
                            for (int y = 0; y < a.aln_size; ++y) { run = gops[a.ops_offset + (uint64_t)y] == 1 ? run + 1 : 0; if (run >= 255) { a.long_ins = 1; break; } }
                            rescued_ops[rg_k[x]] = d_rops + a.ops_offset;
                            g.res[r0 + rg_k[x]] = a;
                        }
                        free(ga); free(gops);
                    }
                }
            }
            if (gpu.owns_lock()) gpu.unlock();
            float ms = 0; cudaEventElapsedTime(&ms, wc->ev0, wc->ev1); g.tm.ext_ms += ms;
            wc->tm.h2d_bytes += sizeof(ExtTask) * tasks.size();
            wc->tm.d2h_bytes += sizeof(nb_alignment) * tasks.size();
            g.res_ops.resize(g.res.size());
            for (size_t k = 0; k < tasks.size(); ++k) {
                cans[task_can[k]].res = (int)(r0 + k);
                g.res_ops[r0 + k] = rescued_ops[k] ? rescued_ops[k] : d_ops + g.res[r0 + k].ops_offset;
            }
            g.tm.ext += gt.lap();
            if (g.round0_done) g.round0_done->store(1);
            if (round > 64) return nb_fail(wc, NB_ECUDA, "extension rounds do not converge");
        }
        return NB_OK;
    };
    std::vector<uint8_t> group_of(c1 - c0, 0);
    for (int k = 0; k < n_groups; ++k)
        for (size_t t = groups[(size_t)k].t0; t < groups[(size_t)k].t1; ++t) group_of[t - c0] = (uint8_t)k;

    // ---- consensus of templates [a0, b0) on the stream of `cx`: link building, best path, walk-back, ASCII gather,
    //      record planning.  take_gpu: hold the GPU like an extension round does; the long-template group runs it
    //      WITHOUT the lock, side by side with the other groups' extension kernels (see below).
    struct CnsOut {
        std::vector<nb_cns_record> recs;
        std::vector<Fill> fills;
        const char* ascii = NULL;       // consensus stretches, dense (the context's pinned staging buffer)
        uint64_t used = 0;              // bytes the records take in the output buffer
        ChunkTimes tm;
    };
    auto cns_phase = [&](nb_ctx* cx, size_t a0, size_t b0, CnsOut& out, bool take_gpu) -> int {
        if (cudaSetDevice(cx->device) != cudaSuccess) return nb_fail(cx, NB_ECUDA, "cudaSetDevice failed");
        cudaStream_t cst = cx->stream;
        HostTimer pt;
        NbDevBuf D(cx);                         // everything this consensus allocates

        // ---- consensus of the chunk
        std::vector<CnsAln> alns;
        std::vector<CnsTemplate> dt;
        std::vector<size_t> dt_tpl;
        uint64_t out_pool = 0, pos_total = 0, link_total = 0, level_total = 0; uint32_t seg_total = 0;
        std::vector<size_t> order;
        for (size_t t = a0; t < b0; ++t) order.push_back(t);
        static const int kSort = getenv("NB_CNS_SORT") ? atoi(getenv("NB_CNS_SORT")) : 1;
        static const int kDpTest = getenv("NB_CNS_DP_TEST") ? atoi(getenv("NB_CNS_DP_TEST")) : 0;   // 1 = test build of k_cns_dp: score ring of 8 (most templates go to k_cns_dp_serial), 2 = all of them
        int span = getenv("NB_CNS_SPAN") ? atoi(getenv("NB_CNS_SPAN")) : CNS_SPAN_DEFAULT;           // (tests lower it: many span boundaries on short reads)
        if (span < CNS_TILE) span = CNS_TILE;
        span -= span % CNS_TILE;
        if (kSort) std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tpls[a].T != tpls[b].T ? tpls[a].T > tpls[b].T : a < b; });
        dt.resize(order.size()); dt_tpl.resize(order.size());
        std::vector<CnsUnit> units;
        uint64_t n_ckpt = 0;
        {
            uint32_t aln_total = 0;
            for (size_t x = 0; x < order.size(); ++x) {
                const Tpl& T = tpls[order[x]];
                CnsTemplate& ct = dt[x];
                memset(&ct, 0, sizeof ct);
                ct.aln_first = aln_total; ct.n_aln = (uint32_t)T.acc.size(); ct.T = T.T;
                aln_total += ct.n_aln;
                ct.seg_cap = (uint32_t)(T.T / std::max(1, (int)(o.min_size * 0.85)) + 2);
                ct.seg_first = seg_total; seg_total += ct.seg_cap;
                ct.out_cap = (uint32_t)(T.T * 2 + 1024);
                ct.out_base = out_pool; out_pool += ct.out_cap;
                ct.level_cap = (uint32_t)(T.T * 3 + 2048);
                ct.unit_first = (uint32_t)units.size();
                for (int s = 0; s * span < T.T; ++s) { CnsUnit u; u.tpl = (uint32_t)x; u.span = (uint32_t)s; units.push_back(u); }
                cx->cc.n_accepted += T.acc.size();
                dt_tpl[x] = order[x];
            }
            if (units.size() >= 0xfffffff0ull) return nb_fail(cx, NB_ELIMIT, "too many consensus units in one chunk");
            alns.resize(aln_total);
            auto fill = [&](size_t x0, size_t x1) {             // the alignment descriptions, in parallel over templates
                for (size_t x = x0; x < x1; ++x) {
                    const size_t t = order[x];
                    const Tpl& T = tpls[t];
                    const ExtGroup& G = groups[group_of[t - c0]];
                    CnsAln* dst = alns.data() + dt[x].aln_first;
                    uint64_t cols = 0;
                    for (const Accepted& a : T.acc) {
                        const Can& c = cans[T.first + (size_t)a.can];
                        const nb_alignment& r = G.res[(size_t)c.res];
                        CnsAln ca;
                        ca.ops = G.res_ops[(size_t)c.res]; ca.qseq_off = reads->h_off[c.qid];
                        ca.aln_size = r.aln_size; ca.toff = r.toff; ca.tend = r.tend; ca.qstart = r.qoff;
                        ca.qlen = c.qsize; ca.qdir = c.qdir; ca.weight = a.weight; ca.drop = r.long_ins;
                        ca.ckpt_first = 0; ca.unit_first = dt[x].unit_first; ca.pad_ = 0;
                        *dst++ = ca;
                        cols += (uint64_t)r.aln_size;
                    }
                    dt[x].link_cap = (uint32_t)std::min<uint64_t>(cols + 32, 0xfffffff0ull);      // links <= columns
                }
            };
            const size_t per = (order.size() + nthreads - 1) / nthreads;
            std::vector<std::thread> th;
            for (unsigned w = 1; w < nthreads && per; ++w) {
                const size_t x0 = w * per, x1 = std::min(order.size(), x0 + per);
                if (x0 >= x1) break;
                th.emplace_back(fill, x0, x1);
            }
            if (per) fill(0, std::min(order.size(), per));
            for (auto& x : th) x.join();
            // span-boundary checkpoints: one per boundary an alignment crosses
            for (CnsAln& ca : alns) {
                ca.ckpt_first = (uint32_t)n_ckpt;
                if (!ca.drop && ca.tend > ca.toff) n_ckpt += (uint64_t)((ca.tend - 1) / span - ca.toff / span);
            }
            if (n_ckpt >= 0xfffffff0ull) return nb_fail(cx, NB_ELIMIT, "too many alignment checkpoints in one chunk");
        }
        for (auto& ct : dt) {
            ct.pos_base = pos_total; pos_total += (uint64_t)ct.T + 2;
            ct.link_base = link_total; link_total += ct.link_cap;
            ct.level_base = level_total; level_total += ct.level_cap;
        }
        std::vector<CnsSegment> segs(seg_total);
        std::vector<uint32_t> nseg(dt.size(), 0);
            std::vector<uint64_t> seg_dense(seg_total, 0);
        uint8_t* d_seq_pool = NULL;
        if (!dt.empty()) {
            int rc;
            CnsParams P;
            memset(&P, 0, sizeof P);
            CnsTemplate* d_tpl; CnsAln* d_aln; CnsUnit* d_unit;
            if ((rc = D.get(&d_tpl, dt.size()))) return rc;
            if ((rc = D.get(&d_aln, alns.size()))) return rc;
            if ((rc = D.get(&d_unit, units.size()))) return rc;
            if ((rc = D.get(&P.seg, seg_total))) return rc;
            if ((rc = D.get(&P.tpl_nseg, dt.size()))) return rc;
            if ((rc = D.get(&P.seq_pool, out_pool))) return rc;
            d_seq_pool = P.seq_pool;
            if ((rc = D.get(&P.span_cols, units.size()))) return rc;
            if ((rc = D.get(&P.span_links, units.size()))) return rc;
            if ((rc = D.get(&P.ckpt, n_ckpt))) return rc;
            if ((rc = D.get(&P.pos, pos_total))) return rc;
            if ((rc = D.get(&P.link, link_total + 64))) return rc;
            if ((rc = D.get(&P.bp, level_total))) return rc;
            if ((rc = D.get(&P.rev, level_total))) return rc;
            static const int kLinksCtas = getenv("NB_CNS_LINKS_CTAS") ? atoi(getenv("NB_CNS_LINKS_CTAS")) : 8;
            const unsigned links_grid = NB_SMS * (unsigned)(kLinksCtas < 1 ? 1 : kLinksCtas);
            if ((rc = D.get(&P.cursor, (size_t)links_grid * CNS_WARPS * CNS_MAX_ALN))) return rc;
            if ((rc = D.get(&P.ovpool, (size_t)links_grid * CNS_WARPS * CNS_OV))) return rc;
            if ((rc = D.get(&P.error_flag, 1))) return rc;
            if ((rc = D.get(&P.counters, 4))) return rc;
            if ((rc = D.get(&P.runs, seg_total))) return rc;
            if ((rc = D.get(&P.tpl_nrun, dt.size()))) return rc;
            if ((rc = D.get(&P.redo, dt.size()))) return rc;
            NB_CUDA(cx, cudaMemcpyAsync(d_tpl, dt.data(), sizeof(CnsTemplate) * dt.size(), cudaMemcpyHostToDevice, cst));
            NB_CUDA(cx, cudaMemcpyAsync(d_aln, alns.data(), sizeof(CnsAln) * alns.size(), cudaMemcpyHostToDevice, cst));
            NB_CUDA(cx, cudaMemcpyAsync(d_unit, units.data(), sizeof(CnsUnit) * units.size(), cudaMemcpyHostToDevice, cst));
            NB_CUDA(cx, cudaMemsetAsync(P.span_cols, 0, sizeof(uint32_t) * units.size(), cst));
            NB_CUDA(cx, cudaMemsetAsync(P.error_flag, 0, sizeof(uint32_t), cst));
            NB_CUDA(cx, cudaMemsetAsync(P.counters, 0, sizeof(unsigned long long) * 4, cst));
            P.pac = reads->d_pac; P.tpl = d_tpl; P.aln = d_aln; P.n_tpl = (uint32_t)dt.size(); P.n_aln = (uint32_t)alns.size();
            P.unit = d_unit; P.n_units = (uint32_t)units.size(); P.span = span;
            P.min_cov = o.min_cov; P.min_size = o.min_size;
            P.flush_slow = getenv("NB_CNS_FLUSH_SLOW") ? atoi(getenv("NB_CNS_FLUSH_SLOW")) : 0;
            P.dp_serial = (o.min_cov < 1 || kDpTest == 2) ? 1 : 0;
            const size_t smem = sizeof(CnsTileShared) * CNS_WARPS;
            if (smem > 48 * 1024) NB_CUDA(cx, cudaFuncSetAttribute(k_cns_links, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            out.tm.prep += pt.lap();
            std::unique_lock<std::mutex> gpu(sh.gpu, std::defer_lock);
            if (take_gpu) gpu.lock();
            out.tm.wait += pt.lap();
            cudaEventRecord(cx->ev2, cst);
            {
                // The templates are sorted by length.  The best-path kernel is a dependent chain over a template's levels: it
                // lasts as long as its longest template has levels and keeps few warps busy, while the link kernel is
                // throughput work that fills the machine.  So the templates are cut into slices, every slice on a stream
                // of its own: the link kernels run one after the other (events), and the chain of a slice runs beside the
                // link kernels of the slices behind it.  Slice 0 holds the longest templates (kLongPct of the bases) and
                // gets a warp per template; the rest is cut into kRest slices of equal size.
                static const int kLongPct = getenv("NB_CNS_LONG_PCT") ? atoi(getenv("NB_CNS_LONG_PCT")) : 8;
                static const int kRest = getenv("NB_CNS_SLICES") ? atoi(getenv("NB_CNS_SLICES")) : 3;
                std::vector<size_t> cut(1, 0);
                if (kLongPct > 0 && dt.size() >= (getenv("NB_CNS_LONG_PCT") ? 4u : 64u)) {
                    uint64_t total_T = 0, acc = 0;
                    for (const auto& ct : dt) total_T += (uint64_t)ct.T;
                    size_t cutx = 0;
                    while (cutx < dt.size() && acc * 100 < total_T * (uint64_t)kLongPct) acc += (uint64_t)dt[cutx++].T;
                    if (cutx < dt.size()) {
                        cut.push_back(cutx);
                        const int nrest = std::max(1, std::min(kRest, NB_AUX - 1));
                        const uint64_t rest_T = total_T - acc;
                        uint64_t racc = 0;
                        for (size_t x = cutx; x < dt.size(); ++x) {
                            racc += (uint64_t)dt[x].T;
                            if ((int)cut.size() < nrest + 1 && racc * (uint64_t)nrest >= rest_T * (uint64_t)(cut.size() - 1) && x + 1 < dt.size()) cut.push_back(x + 1);
                        }
                    }
                }
                cut.push_back(dt.size());
                const int nsl = (int)cut.size() - 1;
                uint32_t* d_ctr;                                    // per slice: unit queue, redo count
                if ((rc = D.get(&d_ctr, 2 * NB_AUX))) return rc;
                NB_CUDA(cx, cudaMemsetAsync(d_ctr, 0, sizeof(uint32_t) * 2 * NB_AUX, cst));
                if (nsl > 1) {
                    NB_CUDA(cx, cudaEventRecord(cx->aux_ev[NB_AUX], cst));
                    for (int k = 0; k < nsl; ++k) NB_CUDA(cx, cudaStreamWaitEvent(cx->aux[k], cx->aux_ev[NB_AUX], 0));
                }
                for (int sl = 0; sl < nsl; ++sl) {
                    const size_t a = cut[(size_t)sl], b = cut[(size_t)sl + 1];
                    cudaStream_t ss = nsl == 1 ? cst : cx->aux[sl];
                    CnsParams Q = P;
                    Q.tpl0 = d_tpl; Q.tpl = d_tpl + a; Q.n_tpl = (uint32_t)(b - a);
                    Q.tpl_nseg = P.tpl_nseg + a; Q.tpl_nrun = P.tpl_nrun + a; Q.redo = P.redo + a;
                    Q.next_unit = d_ctr + sl; Q.n_redo = d_ctr + NB_AUX + sl;
                    const uint32_t u0 = dt[a].unit_first, u1 = b < dt.size() ? dt[b].unit_first : (uint32_t)units.size();
                    Q.unit = d_unit + u0; Q.n_units = u1 - u0;
                    const uint32_t al0 = dt[a].aln_first, al1 = b < dt.size() ? dt[b].aln_first : (uint32_t)alns.size();
                    Q.aln = d_aln;                                  // (CnsTemplate::aln_first indexes all alignments)
                    Q.cursor = P.cursor; Q.ovpool = P.ovpool;      // (the link kernels run one at a time)
                    const unsigned n_tpl = Q.n_tpl;
                    unsigned g = (n_tpl + 3) / 4;
                    if (g > NB_SMS * 16) g = NB_SMS * 16;
                    k_cns_prep<<<g, 128, 0, ss>>>(Q);
                    NB_LAUNCH_CHECK(cx);
                    if (al1 > al0) {
                        CnsParams R = Q;
                        R.aln = d_aln + al0; R.n_aln = al1 - al0;
                        g = (unsigned)((al1 - al0 + 3) / 4);
                        if (g > NB_SMS * 16) g = NB_SMS * 16;
                        k_cns_ckpt<<<g, 128, 0, ss>>>(R);
                        NB_LAUNCH_CHECK(cx);
                    }
                    // with the long slice's chain running beside them, the links of the rest leave it room on the SMs
                    unsigned lmax = links_grid;
                    if (nsl > 1 && sl >= 1 && links_grid > NB_SMS) lmax = links_grid - NB_SMS * 2 > NB_SMS ? links_grid - NB_SMS * 2 : NB_SMS;
                    unsigned lg = (unsigned)((Q.n_units + CNS_WARPS - 1) / CNS_WARPS);
                    if (lg > lmax) lg = lmax;
                    if (nsl > 1 && sl > 0) NB_CUDA(cx, cudaStreamWaitEvent(ss, cx->link_ev[sl - 1], 0));
                    if (lg) {
                        k_cns_links<<<lg, CNS_WARPS * 32, smem, ss>>>(Q);
                        NB_LAUNCH_CHECK(cx);
                    }
                    if (nsl > 1) NB_CUDA(cx, cudaEventRecord(cx->link_ev[sl], ss));
                    if (lg) {
                        k_cns_fix<<<(unsigned)((Q.n_units + 127) / 128), 128, 0, ss>>>(Q);
                        NB_LAUNCH_CHECK(cx);
                    }
                    // the longest templates (slice 0 of two) get a warp each: their chain of levels is what lasts; the rest
                    // four to a warp
                    {
                        const bool wide = nsl > 1 && sl == 0;           // (slice 0 of several = the longest templates)
                        const unsigned per_cta = wide ? 4u : 16u;
                        const int sr = kDpTest == 1 ? 8 : 128;
                        const size_t dsm = (size_t)per_cta * sr * 8;
                        const unsigned dg = (n_tpl + per_cta - 1) / per_cta;
                        if (wide) { if (sr == 8) k_cns_dp<32, 8><<<dg, 128, dsm, ss>>>(Q); else k_cns_dp<32, 128><<<dg, 128, dsm, ss>>>(Q); }
                        else      { if (sr == 8) k_cns_dp<8, 8><<<dg, 128, dsm, ss>>>(Q);  else k_cns_dp<8, 128><<<dg, 128, dsm, ss>>>(Q); }
                    }
                    NB_LAUNCH_CHECK(cx);
                    k_cns_dp_serial<<<(n_tpl + 127) / 128, 128, 0, ss>>>(Q);
                    NB_LAUNCH_CHECK(cx);
                    k_cns_walk<<<(n_tpl + 3) / 4, 128, 0, ss>>>(Q);
                    NB_LAUNCH_CHECK(cx);
                }
                if (nsl > 1)
                    for (int k = 0; k < nsl; ++k) {
                        NB_CUDA(cx, cudaEventRecord(cx->aux_ev[k], cx->aux[k]));
                        NB_CUDA(cx, cudaStreamWaitEvent(cst, cx->aux_ev[k], 0));
                    }
            }
            cudaEventRecord(cx->ev3, cst);
            uint32_t eflag = 0; unsigned long long cnt[4];
            NB_CUDA(cx, cudaMemcpyAsync(&eflag, P.error_flag, sizeof eflag, cudaMemcpyDeviceToHost, cst));
            NB_CUDA(cx, cudaMemcpyAsync(cnt, P.counters, sizeof cnt, cudaMemcpyDeviceToHost, cst));
            NB_CUDA(cx, cudaMemcpyAsync(segs.data(), P.seg, sizeof(CnsSegment) * seg_total, cudaMemcpyDeviceToHost, cst));
            NB_CUDA(cx, cudaMemcpyAsync(nseg.data(), P.tpl_nseg, sizeof(uint32_t) * dt.size(), cudaMemcpyDeviceToHost, cst));
            NB_CUDA(cx, cudaStreamSynchronize(cst));
            float ms = 0; cudaEventElapsedTime(&ms, cx->ev2, cx->ev3); out.tm.cns_ms += ms;
            cx->tm.d2h_bytes += sizeof(CnsSegment) * seg_total;
            cx->cc.n_tags += cnt[1];
            if (eflag) return nb_fail(cx, NB_ELIMIT, "consensus kernel capacity exceeded (flags %u)", eflag);
            {
                // gather the stretches densely as ASCII, one exact-size D2H copy
                std::vector<uint64_t> gsrc, gdst; std::vector<uint32_t> glen;
                uint64_t dense = 0;
                for (size_t x = 0; x < dt.size(); ++x)
                    for (uint32_t k = 0; k < nseg[x]; ++k) {
                        const CnsSegment& g = segs[dt[x].seg_first + k];
                        gsrc.push_back(g.seq_off); gdst.push_back(dense); glen.push_back(g.len);
                        seg_dense[dt[x].seg_first + k] = dense;
                        dense += g.len;
                    }
                if ((rc = nb_pinned_reserve(cx, dense + 16))) return rc;
                out.ascii = (const char*)cx->pinned;
                if (!gsrc.empty()) {
                    uint64_t *d_src, *d_dst; uint32_t* d_len; char* d_out;
                    if ((rc = D.get(&d_src, gsrc.size()))) return rc;
                    if ((rc = D.get(&d_dst, gsrc.size()))) return rc;
                    if ((rc = D.get(&d_len, gsrc.size()))) return rc;
                    if ((rc = D.get(&d_out, dense))) return rc;
                    NB_CUDA(cx, cudaMemcpyAsync(d_src, gsrc.data(), 8 * gsrc.size(), cudaMemcpyHostToDevice, cst));
                    NB_CUDA(cx, cudaMemcpyAsync(d_dst, gdst.data(), 8 * gsrc.size(), cudaMemcpyHostToDevice, cst));
                    NB_CUDA(cx, cudaMemcpyAsync(d_len, glen.data(), 4 * gsrc.size(), cudaMemcpyHostToDevice, cst));
                    k_gather_ascii<<<NB_SMS * 8, 256, 0, cst>>>(d_seq_pool, d_src, d_dst, d_len, (uint32_t)gsrc.size(), d_out);
                    NB_LAUNCH_CHECK(cx);
                    NB_CUDA(cx, cudaMemcpyAsync(cx->pinned, d_out, dense, cudaMemcpyDeviceToHost, cst));
                    NB_CUDA(cx, cudaStreamSynchronize(cst));
                    cx->tm.d2h_bytes += dense;
                }
            }
            if (take_gpu) gpu.unlock();
            out.tm.kern += pt.lap();
        }
        // ---- records (tasc/cbcns.c:108-264, consensus_one_read.c:20-136): plan, then fill in parallel
            auto plan = [&](int id, int left, int right, int org, int ncan, int novl, double cutoff, int is_raw, uint32_t len) {
            nb_cns_record r;
            r.id = id; r.left = left; r.right = right; r.org_size = org; r.num_can = ncan; r.num_ovlps = novl;
            r.ident_cutoff = cutoff; r.seq_offset = out.used; r.seq_len = (int)len; r.is_raw = is_raw;
            out.recs.push_back(r);
        };
        auto add_raw = [&](int id, int from, int to) {      // read bases [from, to)
            Fill f; f.dst = out.used; f.len = (uint32_t)(to - from); f.kind = 1; f.src = reads->h_off[id] + (uint64_t)from;
            out.fills.push_back(f); out.used += f.len;
        };
        auto add_cns = [&](const CnsSegment& g, uint64_t dense_off) {
            Fill f; f.dst = out.used; f.len = g.len; f.kind = 0; f.src = dense_off;
            out.fills.push_back(f); out.used += f.len;
        };
        for (size_t x = 0; x < dt.size(); ++x) {
            Tpl& T = tpls[dt_tpl[x]];
            const CnsSegment* sg = segs.data() + dt[x].seg_first;
            const uint64_t* sd = seg_dense.data() + dt[x].seg_first;
            const uint32_t ns = nseg[x];
            if (!o.full_consensus) {
                for (uint32_t k = 0; k < ns; ++k) {
                    plan(T.tid, sg[k].left, sg[k].right, T.T, T.num_can, T.num_ovlps, T.cutoff, 0, sg[k].len);
                    add_cns(sg[k], sd[k]);
                    cx->cc.cns_bases += sg[k].len;
                }
                // raw stretches between the corrected runs (get_raw_intvs, consensus_one_read.c:20-65)
                bool first = true, open = true;
                int left = 0;
                const int end_off = T.T - 1;
                for (uint32_t k = 0; k < ns; ++k) {
                    const int fs = sg[k].left, fe = sg[k].right;
                    if (first) {
                        open = false; first = false;
                        if (fs > 0) left = 0;
                        else { left = fe + 1; continue; }
                    }
                    const int right = fs - 1;
                    if (right - left + 1 >= 1000) {
                        plan(T.tid, left, right + 1, T.T, T.num_can, T.num_ovlps, T.cutoff, 1, (uint32_t)(right + 1 - left));
                        add_raw(T.tid, left, right + 1);
                    }
                    if (fe >= end_off) { open = false; break; }
                    left = fe + 1;
                }
                if (open) {
                    const int right = end_off;
                    if (right - left + 1 >= 1000) {
                        plan(T.tid, left, right + 1, T.T, T.num_can, T.num_ovlps, T.cutoff, 1, (uint32_t)(right + 1 - left));
                        add_raw(T.tid, left, right + 1);
                    }
                }
                corrected[(size_t)T.tid] = 1;                       // consensus_one_read.c:391 (r stays 1)
            } else if (ns) {
                // consensus_unbroken (cbcns.c:218-257): corrected runs spliced into the raw read
                const uint64_t start = out.used;
                const size_t ri = out.recs.size();
                plan(T.tid, 0, T.T, T.T, T.num_can, T.num_ovlps, T.cutoff, 0, 0);
                int last_raw = 0;
                for (uint32_t k = 0; k < ns; ++k) {
                    const int rf = sg[k].cns_from, rt = sg[k].cns_to;
                    if (rf > last_raw) add_raw(T.tid, last_raw, rf);
                    add_cns(sg[k], sd[k]);
                    last_raw = rt;
                }
                if (last_raw < T.T) add_raw(T.tid, last_raw, T.T);
                out.recs[ri].seq_len = (int)(out.used - start);
                cx->cc.cns_bases += out.used - start;
                corrected[(size_t)T.tid] = 1;
            }
            std::vector<uint16_t>().swap(T.cov);
        }
        out.tm.rec += pt.lap();
        return NB_OK;
    };
    // records of one consensus into the job's buffer: place, rebase, fill in parallel
    auto emit = [&](CnsOut& out) -> int {
        HostTimer et;
        const uint64_t base = sh.used.fetch_add(out.used);
        if (base + out.used > sh.cap) return nb_fail(ctx, NB_ELIMIT, "output exceeds its upper bound");
        for (auto& r : out.recs) r.seq_offset += base;
        for (auto& f : out.fills) f.dst += base;
        run_fills(out.fills, sh.buf, out.ascii, reads, nthreads);
        recs.insert(recs.end(), out.recs.begin(), out.recs.end());
        t_prep += out.tm.prep; t_kern += out.tm.kern; t_wait += out.tm.wait; t_rec += out.tm.rec + et.lap(); cns_ms += out.tm.cns_ms;
        return NB_OK;
    };
    auto fold = [&](nb_ctx* wc) {                       // a worker's counters into this context
        if (wc == ctx) return;
        ctx->launches += wc->launches; wc->launches = 0;
        ctx->tm.h2d_bytes += wc->tm.h2d_bytes; ctx->tm.d2h_bytes += wc->tm.d2h_bytes;
        ctx->cc.n_extensions += wc->cc.n_extensions; ctx->cc.n_block_aligns += wc->cc.n_block_aligns;
        ctx->cc.n_dp_columns += wc->cc.n_dp_columns; ctx->cc.n_dp_word_updates += wc->cc.n_dp_word_updates;
        ctx->cc.n_accepted += wc->cc.n_accepted; ctx->cc.n_tags += wc->cc.n_tags; ctx->cc.cns_bases += wc->cc.cns_bases;
        if (wc->cc.big_ext_word_updates > ctx->cc.big_ext_word_updates) {
            ctx->cc.big_ext_tasks = wc->cc.big_ext_tasks; ctx->cc.big_ext_block_aligns = wc->cc.big_ext_block_aligns;
            ctx->cc.big_ext_word_updates = wc->cc.big_ext_word_updates; ctx->cc.big_ext_usec = wc->cc.big_ext_usec; ctx->cc.big_ext_dp_columns = wc->cc.big_ext_dp_columns;
        }
        memset(&wc->cc, 0, sizeof wc->cc); memset(&wc->tm, 0, sizeof wc->tm);
    };
    std::mutex emit_mx;
    std::vector<CnsOut> outs((size_t)n_groups);
    std::vector<int> cns_rc((size_t)n_groups, NB_OK);
    CnsOut out_long, out_main;
    int rc_long = NB_OK;
    {
        std::vector<std::atomic<int>> r0done((size_t)n_groups);
        for (auto& x : r0done) x.store(0);
        std::atomic<int> long_ext_done(long_mode ? 0 : 1);
        auto group_work = [&](int k) {
            ExtGroup& g = groups[(size_t)k];
            if (overlap) {
                g.submitted = &r0done[(size_t)k];
                if (k > 0) g.prev_submitted = &r0done[(size_t)k - 1];
            }
            g.rc = ext_phase(g);
            if (overlap) {
                r0done[(size_t)k].store(1);                 // (also when the group had nothing to launch)
                if (g.rc == NB_OK && g.t0 < g.t1) {
                    cns_rc[(size_t)k] = cns_phase(g.wc, g.t0, g.t1, outs[(size_t)k], false);
                    if (cns_rc[(size_t)k] == NB_OK) { std::lock_guard<std::mutex> lk(emit_mx); cns_rc[(size_t)k] = emit(outs[(size_t)k]); }
                }
            }
        };
        std::vector<std::thread> th;
        for (int k = 1; k < n_groups; ++k)
            th.emplace_back([&, k]() {
                while (!long_ext_done.load()) std::this_thread::yield();
                group_work(k);
            });
        group_work(0);
        long_ext_done.store(1);
        if (long_mode && groups[0].rc == NB_OK) {
            // Long-template mode (off by default): group 0 holds the longest templates; its consensus runs on this thread
            // without the GPU lock while the other groups extend.
            rc_long = cns_phase(ctx, groups[0].t0, groups[0].t1, out_long, false);
            if (rc_long == NB_OK) { std::lock_guard<std::mutex> lk(emit_mx); rc_long = emit(out_long); }
        }
        for (auto& x : th) x.join();
    }
    for (int k = 0; k < n_groups; ++k) {
        ExtGroup& g = groups[(size_t)k];
        if (g.rc) { if (g.wc != ctx) memcpy(ctx->err, g.wc->err, sizeof ctx->err); return g.rc; }
        if (cns_rc[(size_t)k]) { if (g.wc != ctx) memcpy(ctx->err, g.wc->err, sizeof ctx->err); return cns_rc[(size_t)k]; }
        t_replay += g.tm.replay; t_gather += g.tm.gather; t_wait += g.tm.wait; t_ext += g.tm.ext; ext_ms += g.tm.ext_ms; t_init += g.tm.prep;
        fold(g.wc);
    }
    if (rc_long) return rc_long;
    const double t_ext_wall = ht.lap();
    if (!overlap) {
        const size_t a0 = long_mode ? groups[0].t1 : c0;
        if (a0 < c1) {
            int rc = cns_phase(ctx, a0, c1, out_main, true);
            if (rc) return rc;
            if ((rc = emit(out_main))) return rc;
        }
    }
    tms.replay = t_replay; tms.gather = t_gather; tms.ext = t_ext; tms.prep = t_prep; tms.kern = t_kern; tms.rec = t_rec; tms.wait = t_wait; tms.ext_wall = t_ext_wall; tms.init = t_init;
    tms.ext_ms = ext_ms; tms.cns_ms = cns_ms;
    return NB_OK;
}

extern "C" int nb_consensus_job_sharded(nb_ctx* ctx, const nb_volume* reads, const nb_packed_candidate* cans_in, uint64_t n_in,
                                        const nb_cns_options* opt, int shard_count, int shard_index,
                                        nb_cns_record** out_recs, uint64_t* n_recs, char** out_seq, uint64_t* out_seq_bytes);

extern "C" int nb_consensus_job(nb_ctx* ctx, const nb_volume* reads, const nb_packed_candidate* cans_in, uint64_t n_in,
                                const nb_cns_options* opt, nb_cns_record** out_recs, uint64_t* n_recs, char** out_seq,
                                uint64_t* out_seq_bytes)
{
    return nb_consensus_job_sharded(ctx, reads, cans_in, n_in, opt, 1, 0, out_recs, n_recs, out_seq, out_seq_bytes);
}

extern "C" int nb_consensus_job_sharded(nb_ctx* ctx, const nb_volume* reads, const nb_packed_candidate* cans_in, uint64_t n_in,
                                        const nb_cns_options* opt, int shard_count, int shard_index,
                                        nb_cns_record** out_recs, uint64_t* n_recs, char** out_seq, uint64_t* out_seq_bytes)
{
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
        return nb_fail(ctx, NB_EINVAL, "nb_consensus_job_sharded: bad shard %d of %d", shard_index, shard_count);
    if (!ctx || !reads || (!cans_in && n_in) || !opt || !out_recs || !n_recs || !out_seq || !out_seq_bytes)
        return nb_fail(ctx, NB_EINVAL, "nb_consensus_job: bad arguments");
    *out_recs = NULL; *n_recs = 0; *out_seq = NULL; *out_seq_bytes = 0;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->cc, 0, sizeof ctx->cc);
    ctx->tm.extend_ms = 0; ctx->tm.consensus_ms = 0;
    const nb_cns_options o = *opt;
    const bool trace = getenv("NB_TRACE") != NULL;
    HostTimer ht;
    double t_setup = 0, t_tail = 0;
    std::vector<nb_cns_record> recs;
    std::vector<Fill> fills;
    if (n_in == 0) return NB_OK;

    // ---- load order + normalisation (consensus_one_partition.c:10-95): records grouped by template, input order
    //      kept inside a group (a stable counting sort on the template id)
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 4;
    // NB_HOST_THREADS: the host threads of this job (several ranks on one host share its cores: necat_b200/pipeline.py
    // gives every rank cores / local world size)
    if (getenv("NB_HOST_THREADS") && atoi(getenv("NB_HOST_THREADS")) > 0) nthreads = (unsigned)atoi(getenv("NB_HOST_THREADS"));
    if (nthreads > 64) nthreads = 64;
    std::vector<uint32_t> grp_start(reads->nseq + 1, 0);
    for (uint64_t i = 0; i < n_in; ++i) {
        const nb_packed_candidate& p = cans_in[i];
        if (p.item[1] >= reads->nseq || p.item[4] >= reads->nseq) return nb_fail(ctx, NB_EINVAL, "candidate read id out of range");
        ++grp_start[p.item[1] + 1];
    }
    for (uint64_t i = 0; i < reads->nseq; ++i) grp_start[i + 1] += grp_start[i];
    std::vector<nb_packed_candidate> pc(n_in);
    {
        std::vector<uint32_t> cur(grp_start.begin(), grp_start.end() - 1);
        for (uint64_t i = 0; i < n_in; ++i) pc[cur[cans_in[i].item[1]]++] = cans_in[i];
    }
    const int min_id = (int)pc.front().item[1], max_id = (int)pc.back().item[1];
    std::vector<uint8_t> corrected(reads->nseq, 0);

    // ---- templates: at least -x records (consensus_one_read.c:218), this shard's share; their candidates sorted and
    //      cut to MAX_EXAMINED_CAN, in parallel over templates
    std::vector<Tpl> tpls;
    std::vector<Can> cans;
    {
        size_t ncan = 0;
        for (uint64_t id = 0; id < reads->nseq; ++id) {
            const size_t cnt = grp_start[id + 1] - grp_start[id];
            if (cnt == 0 || (size_t)o.min_cov > cnt || (int)(id % (uint64_t)shard_count) != shard_index) continue;
            Tpl t;
            t.tid = (int)id; t.T = (int)reads->h_len[id];
            t.first = ncan; t.n = cnt > 300 ? 300 : cnt;               // MAX_EXAMINED_CAN
            t.stage = 0; t.next = 0; t.win_to = 0; t.win_i = 0; t.in_window = false; t.cutoff = 0; t.num_can = 0; t.num_ovlps = 0;
            t.n_below = 0;
            ncan += t.n;
            tpls.push_back(std::move(t));
        }
        cans.resize(ncan);
        auto build = [&](size_t t0, size_t t1) {
            std::vector<nb_packed_candidate> one;
            for (size_t x = t0; x < t1; ++x) {
                const Tpl& t = tpls[x];
                one.assign(pc.begin() + grp_start[(size_t)t.tid], pc.begin() + grp_start[(size_t)t.tid + 1]);
                for (auto& p : one) {                                   // normalise_pcan_sdir, gapped_candidate.c:72-93
                    const uint32_t w = p.item[0];
                    if (!(w >> 31)) continue;
                    uint32_t r = w & ((1u << 29) - 1);
                    if (!((w >> 30) & 1)) r |= 1u << 30;
                    if (!(w & (1u << 29))) r |= 1u << 29;
                    p.item[0] = r;
                    const uint32_t qs = reads->h_len[p.item[4]], ss = reads->h_len[p.item[1]];
                    const uint32_t qb = qs - p.item[6], qe = qs - p.item[5];
                    p.item[5] = qb; p.item[6] = qe;
                    const uint32_t sb = ss - p.item[3], se = ss - p.item[2];
                    p.item[2] = sb; p.item[3] = se;
                }
                // -s 1 (consensus_one_read.c:249-255): more than MAX_EXAMINED_CAN records are cut in load order and stay unsorted
                if (!(o.small_memory && one.size() > 300))
                std::sort(one.begin(), one.end(), [](const nb_packed_candidate& a, const nb_packed_candidate& b) {
                    // PackedGappedCandidate_CnsScoreGT, gapped_candidate.c:96-126
                    const int as = (int)(a.item[0] & ((1u << 29) - 1)), bs = (int)(b.item[0] & ((1u << 29) - 1));
                    if (as != bs) return as > bs;
                    if (a.item[4] != b.item[4]) return (int)a.item[4] < (int)b.item[4];
                    const int ad = (a.item[0] >> 30) & 1, bd = (b.item[0] >> 30) & 1;
                    if (ad != bd) return ad < bd;
                    if (a.item[5] != b.item[5]) return (int)a.item[5] < (int)b.item[5];
                    return (int)a.item[2] < (int)b.item[2];
                });
                for (size_t k = 0; k < t.n; ++k) {
                    const nb_packed_candidate& p = one[k];
                    Can c;
                    const uint32_t w = p.item[0];
                    c.qid = (int)p.item[4]; c.qdir = (int)((w >> 30) & 1); c.score = (int)(w & ((1u << 29) - 1));
                    c.sbeg = (int)p.item[2]; c.send = (int)p.item[3]; c.qbeg = (int)p.item[5]; c.qend = (int)p.item[6];
                    if (w & (1u << 29)) { c.qoff = (int)p.item[5]; c.soff = (int)p.item[2]; }
                    else                { c.qoff = (int)p.item[6]; c.soff = (int)p.item[3]; }
                    c.qsize = (int)reads->h_len[c.qid];
                    c.res = -1;
                    cans[t.first + k] = c;
                }
            }
        };
        const size_t per = (tpls.size() + nthreads - 1) / nthreads;
        std::vector<std::thread> th;
        for (unsigned w = 1; w < nthreads && per; ++w) {
            const size_t t0 = w * per, t1 = std::min(tpls.size(), t0 + per);
            if (t0 >= t1) break;
            th.emplace_back(build, t0, t1);
        }
        if (per) build(0, std::min(tpls.size(), per));
        for (auto& x : th) x.join();
    }
    // Longest templates first: their candidates are the longest chains of block alignments of an extension launch (k_extend
    // takes the tasks in order), started first they do not end up as its tail.  (Records are put in id order by the caller.)
    {
        // (an index is sorted and the templates are moved once: a Tpl is four vectors and a dozen fields)
        std::vector<uint32_t> ix(tpls.size());
        for (size_t i = 0; i < ix.size(); ++i) ix[i] = (uint32_t)i;
        std::stable_sort(ix.begin(), ix.end(), [&](uint32_t a, uint32_t b) { return tpls[a].T > tpls[b].T; });
        std::vector<Tpl> re;
        re.reserve(tpls.size());
        for (uint32_t i : ix) re.push_back(std::move(tpls[i]));
        tpls.swap(re);
    }
    // The consensus kernels of a template are a dependent chain over its positions, so the longest templates decide
    // how long the consensus of a batch lasts.  They are moved to the front and get their own group, whose consensus
    // runs beside the extension of everything else (process_chunk).
    size_t n_long = 0;
    {
        const char* env_long = getenv("NB_CNS_LONG");
        const char* env_lmin = getenv("NB_CNS_LONG_MIN");   // (tests lower it)
        const size_t lmin = env_lmin ? (size_t)atoi(env_lmin) : 2048;
        // off by default: measured on the E. coli workload, the extension kernels lose more (3 CTAs per SM, shared L2 / HBM)
        // than the overlap gains (537 ms per step against 390), see DESIGN.md
        if (env_long && atoi(env_long) != 0 && tpls.size() >= lmin && tpls.size() >= 32) {
            n_long = std::min<size_t>(512, std::max<size_t>(8, tpls.size() / 32));
            std::vector<uint32_t> idx(tpls.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
            std::nth_element(idx.begin(), idx.begin() + (long)n_long, idx.end(), [&](uint32_t a, uint32_t b) {
                return tpls[a].T != tpls[b].T ? tpls[a].T > tpls[b].T : a < b; });
            std::vector<uint8_t> is_long(tpls.size(), 0);
            for (size_t i = 0; i < n_long; ++i) is_long[idx[i]] = 1;
            std::vector<Tpl> re;
            re.reserve(tpls.size());
            for (size_t i = 0; i < tpls.size(); ++i) if (is_long[i]) re.push_back(std::move(tpls[i]));
            for (size_t i = 0; i < tpls.size(); ++i) if (!is_long[i]) re.push_back(std::move(tpls[i]));
            tpls.swap(re);
        }
    }
    ctx->cc.n_templates = tpls.size();
    t_setup += ht.lap();

    // ---- the templates are cut into chunks of equal total length that go through a small pipeline: every
    //      worker owns a stream, events, a pinned staging buffer and walk-back scratch, so the host phases of one
    //      chunk (replay of the accept/skip logic, request gathering, record planning and filling) overlap the
    //      kernels of the others.  Chunks also bound the device memory of a 100 000-template partition.
    const char* env_chunks = getenv("NB_CNS_CHUNKS");
    const char* env_pipes = getenv("NB_CNS_PIPES");
    size_t n_chunks = env_chunks ? (size_t)atoi(env_chunks) : 1;
    const size_t kChunkMax = 32768;
    if (n_chunks < (tpls.size() + kChunkMax - 1) / kChunkMax) n_chunks = (tpls.size() + kChunkMax - 1) / kChunkMax;
    // ... and the device memory of a chunk: the op bytes of its extension rounds (every speculative extension reserves
    // query + template + slack bytes; round 0 is counted exactly, the later rounds as 60 % more) and the link, level and
    // position records of its consensus (16 B per alignment column, ~1.3 columns per template base and alignment)
    std::vector<uint64_t> need(tpls.size());
    uint64_t need_total = 0;
    for (size_t t = 0; t < tpls.size(); ++t) {
        const Tpl& T = tpls[t];
        uint64_t r0 = 0;
        for (size_t k = 0; k < T.n && k < 50; ++k) {
            const Can& c = cans[T.first + k];
            r0 += ((uint64_t)c.qsize + (uint64_t)T.T + (uint64_t)std::min(c.qoff, c.soff) + 64 + 15) & ~15ull;
        }
        const uint64_t alns = std::min<uint64_t>(T.n, (uint64_t)o.max_cov + 4);
        need[t] = r0 + r0 * 6 / 10 + (uint64_t)T.T * (21 * alns + 120);
        need_total += need[t];
    }
    {
        size_t fr = 0, tot = 0;
        NB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
        uint64_t budget = (uint64_t)fr / 2;
        if (getenv("NB_CNS_MEM")) budget = strtoull(getenv("NB_CNS_MEM"), NULL, 10);       // (tests)
        if (budget < (64ull << 20)) budget = 64ull << 20;
        const size_t by_mem = (size_t)((need_total + budget - 1) / budget);
        if (n_chunks < by_mem) n_chunks = by_mem;
    }
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > tpls.size()) n_chunks = tpls.size() ? tpls.size() : 1;
    int n_workers = env_pipes ? atoi(env_pipes) : 1;
    if (n_workers > NB_MAX_WORKERS) n_workers = NB_MAX_WORKERS;
    if ((size_t)n_workers > n_chunks) n_workers = (int)n_chunks;
    if (n_workers < 1) n_workers = 1;
    std::vector<size_t> cut(n_chunks + 1, tpls.size());
    {
        uint64_t acc = 0;
        size_t k = 1;
        cut[0] = 0;
        for (size_t t = 0; t < tpls.size() && k < n_chunks; ++t) {
            acc += need[t];
            if (acc * n_chunks >= need_total * k || t + 1 - cut[k - 1] >= kChunkMax) cut[k++] = t + 1;
        }
    }
    JobShared sh;
    sh.used = 0; sh.corrected = &corrected; sh.n_long = n_long;
    {
        // upper bound of the output: a template emits at most its consensus pool (2 T + 1024, enforced by the kernel)
        // plus its raw leftovers; every other read of the id range may be dumped whole
        uint64_t bound = 64;
        for (const Tpl& t : tpls) bound += (uint64_t)t.T * 3 + 1024;
        for (int id = min_id; id < max_id; ++id) bound += reads->h_len[id];
        sh.cap = bound;
        sh.buf = (char*)malloc(bound);      // untouched pages cost nothing; shrunk to the used prefix at the end
        if (!sh.buf) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    }
    std::vector<std::vector<nb_cns_record>> chunk_recs(n_chunks);
    std::vector<ChunkTimes> chunk_tm(n_chunks);
    std::vector<int> worker_rc((size_t)n_workers, NB_OK);
    std::vector<nb_ctx*> wctx((size_t)n_workers, ctx);
    for (int w = 1; w < n_workers; ++w) {
        int rc = nb_worker_ctx(ctx, w, &wctx[(size_t)w]);
        if (rc) { free(sh.buf); return rc; }
    }
    std::atomic<size_t> next_chunk(0);
    auto work = [&](int w) {
        nb_ctx* wc = wctx[(size_t)w];
        if (cudaSetDevice(wc->device) != cudaSuccess) { worker_rc[(size_t)w] = nb_fail(wc, NB_ECUDA, "cudaSetDevice failed"); return; }
        for (;;) {
            const size_t k = next_chunk.fetch_add(1);
            if (k >= n_chunks) break;
            if (cut[k] >= cut[k + 1]) continue;
            const int rc = process_chunk(wc, reads, o, tpls, cans, cut[k], cut[k + 1], sh, chunk_recs[k], chunk_tm[k],
                                         n_workers > 1 ? std::max(2u, nthreads / 2) : nthreads);
            if (rc) { worker_rc[(size_t)w] = rc; next_chunk.store(n_chunks); break; }
        }
    };
    {
        std::vector<std::thread> th;
        for (int w = 1; w < n_workers; ++w) th.emplace_back(work, w);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int w = 1; w < n_workers; ++w) {               // fold the workers' counters into the job's context
        nb_ctx* wc = wctx[(size_t)w];
        ctx->launches += wc->launches; wc->launches = 0;
        ctx->tm.h2d_bytes += wc->tm.h2d_bytes; ctx->tm.d2h_bytes += wc->tm.d2h_bytes;
        ctx->cc.n_extensions += wc->cc.n_extensions; ctx->cc.n_block_aligns += wc->cc.n_block_aligns;
        ctx->cc.n_dp_columns += wc->cc.n_dp_columns; ctx->cc.n_dp_word_updates += wc->cc.n_dp_word_updates;
        ctx->cc.n_accepted += wc->cc.n_accepted; ctx->cc.n_tags += wc->cc.n_tags; ctx->cc.cns_bases += wc->cc.cns_bases;
        memset(&wc->cc, 0, sizeof wc->cc); memset(&wc->tm, 0, sizeof wc->tm);
    }
    for (int w = 0; w < n_workers; ++w)
        if (worker_rc[(size_t)w]) {
            if (w) memcpy(ctx->err, wctx[(size_t)w]->err, sizeof ctx->err);
            free(sh.buf);
            return worker_rc[(size_t)w];
        }
    float ext_ms = 0, cns_ms = 0;
    for (size_t k = 0; k < n_chunks; ++k) {
        recs.insert(recs.end(), chunk_recs[k].begin(), chunk_recs[k].end());
        ext_ms += chunk_tm[k].ext_ms; cns_ms += chunk_tm[k].cns_ms;
    }
    const double t_chunks = ht.lap();
    ctx->tm.extend_ms = ext_ms;
    ctx->tm.consensus_ms = cns_ms;

    // ---- reads of [min_id, max_id) that were never corrected go to raw_out whole (consensus_one_partition.c:173-194)
    fills.clear();
    uint64_t used = sh.used.load();
    // (-s 1: the reference then has no merged read store and skips this dump, consensus_one_partition.c:173)
    for (int id = min_id; id < max_id && !o.small_memory; ++id) {
        if (corrected[(size_t)id] || id % shard_count != shard_index) continue;
        const int L = (int)reads->h_len[id];
        nb_cns_record r;
        r.id = id; r.left = 0; r.right = L; r.org_size = L; r.num_can = 0; r.num_ovlps = 0; r.ident_cutoff = 0.0;
        r.seq_offset = used; r.seq_len = L; r.is_raw = 1;
        recs.push_back(r);
        Fill f; f.dst = used; f.len = (uint32_t)L; f.kind = 1; f.src = reads->h_off[id];
        fills.push_back(f); used += f.len;
    }
    if (used > sh.cap) { free(sh.buf); return nb_fail(ctx, NB_ELIMIT, "output exceeds its upper bound"); }
    run_fills(fills, sh.buf, NULL, reads, nthreads);
    ctx->cc.n_cns_records = recs.size();

    nb_cns_record* r_out = (nb_cns_record*)malloc(sizeof(nb_cns_record) * (recs.size() ? recs.size() : 1));
    char* s_out = (char*)realloc(sh.buf, used ? used : 1);
    if (!s_out) s_out = sh.buf;
    if (!r_out || !s_out) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    if (!recs.empty()) memcpy(r_out, recs.data(), sizeof(nb_cns_record) * recs.size());
    t_tail += ht.lap();
    if (trace) {
        ChunkTimes s;
        for (const ChunkTimes& c : chunk_tm) { s.replay += c.replay; s.gather += c.gather; s.ext += c.ext; s.prep += c.prep; s.kern += c.kern; s.rec += c.rec; s.wait += c.wait; s.ext_wall += c.ext_wall; s.init += c.init; }
        fprintf(stderr, "[nb_consensus_job] %zu chunks on %d workers; wall ms: setup %.1f chunks %.1f tail %.1f | summed over chunks: init %.1f replay %.1f gather %.1f "
                        "gpu-wait %.1f ext(call) %.1f [extension phase wall %.1f] cns_prep %.1f cns(call) %.1f records %.1f | gpu ms: extend %.1f consensus %.1f\n",
                n_chunks, n_workers, t_setup, t_chunks, t_tail, s.init, s.replay, s.gather, s.wait, s.ext, s.ext_wall, s.prep, s.kern, s.rec, ext_ms, cns_ms);
    }
    *out_recs = r_out; *n_recs = recs.size(); *out_seq = s_out; *out_seq_bytes = used;
    return NB_OK;
}
