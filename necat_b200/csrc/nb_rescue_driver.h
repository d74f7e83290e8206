// nb_rescue_driver.h -- host side of the global alignment of the oc2cns -r 1 rescue (edlib_go with find_path,
// edlib/edlib_wrapper.c:111-255, over edlib's NW + Hirschberg + traceback, edlib/edlib.cpp:887-1361).
//
// The recursion is driven from the host, level by level over ALL tasks of a batch: every level asks its backend for
//   columns(jobs)  the last column of a global edit-distance matrix per job (rs_nw_column: the two halves of every
//                  node that is still too big for the traceback), and
//   leaves(jobs)   the operations of every node that is small enough (rs_nw_leaf);
// the host only looks for the split row (the FIRST row where left + right == best, edlib.cpp:1262-1296), cuts the
// nodes, and at the end strings the leaves together and trims the alignment to its first and last run of kMatchSize
// matches.  The backend is a template parameter: the CUDA kernels of nb_rescue.cu on the GPU, or the host emulation of
// the very same warp-level cores in tests/emu (how this file is checked without a GPU).
#pragma once
#include <stdint.h>

#include <vector>

#include "nb_rescue_core.cuh"

struct RgTask {                 // align Q[qfrom, qto) (a read on its strand) with T[tfrom, tto) if their distance is <= tolerance
    RsSeq Q, T;
    int   qfrom, qto, tfrom, tto, tolerance;
};

struct RgResult {
    int ok = 0, qoff = 0, qend = 0, toff = 0, tend = 0, aln_size = 0, n_match = 0;
    std::vector<uint8_t> ops;   // 0 match, 1 query base against a gap, 2 gap against a template base, 3 mismatch (edlib's codes = ours)
};

struct RgColJob  { int task, q0, qn, qstep, t0, tn, tstep; };       // q(i) = Q.at(q0 + i * qstep), likewise t
struct RgLeafJob { int task, q0, qn, t0, tn; };

// A slice of jobs laid out in pools (what the kernels k_rescue_columns / k_rescue_leaves consume): per job the offsets of
// its column scores, its two generations of bottom-row deltas, its bit-vector matrix and its operations.
struct RgColJobDev  { RsSeq Q, T; int q0, qn, qstep, t0, tn, tstep; unsigned long long col_off, h_off; };
struct RgLeafJobDev { RsSeq Q, T; int q0, qn, t0, tn, cap, pad_; unsigned long long mat_off, h_off, ops_off; };

// Backend of rg_run over an executor that runs one pooled slice: Exec::run_columns(jobs, col_total, h_total, h_cols) and
// Exec::run_leaves(jobs, mat_total, h_total, ops_total, h_ops, h_n) (false = failed).  Slices keep the scratch of one
// launch bounded.  The executor is the GPU (nb_rescue.cu) or the host emulation (tests/emu).
template <class Exec>
struct RgPooledBackend {
    Exec& ex;
    bool failed = false;
    explicit RgPooledBackend(Exec& e) : ex(e) {}

    void columns(const std::vector<RgTask>& tasks, const std::vector<RgColJob>& jobs, std::vector<std::vector<int>>& cols)
    {
        cols.assign(jobs.size(), std::vector<int>());
        size_t a = 0;
        while (a < jobs.size()) {
            std::vector<RgColJobDev> dj;
            unsigned long long col_total = 0, h_total = 0;
            size_t b = a;
            while (b < jobs.size() && (b == a || (col_total + h_total) < (1ull << 29))) {
                const RgColJob& j = jobs[b];
                const RgTask& t = tasks[(size_t)j.task];
                RgColJobDev d;
                d.Q = t.Q; d.T = t.T; d.q0 = j.q0; d.qn = j.qn; d.qstep = j.qstep; d.t0 = j.t0; d.tn = j.tn; d.tstep = j.tstep;
                d.col_off = col_total; d.h_off = h_total;
                col_total += (unsigned long long)j.qn + 1;
                h_total += 2ull * ((unsigned long long)j.tn + 1);
                dj.push_back(d);
                ++b;
            }
            std::vector<int> h_cols;
            if (failed || !ex.run_columns(dj, col_total, h_total, h_cols)) { failed = true; h_cols.assign((size_t)col_total, 0); }
            for (size_t x = a; x < b; ++x)
                cols[x].assign(h_cols.begin() + (long)dj[x - a].col_off, h_cols.begin() + (long)dj[x - a].col_off + jobs[x].qn + 1);
            a = b;
        }
    }

    void leaves(const std::vector<RgTask>& tasks, const std::vector<RgLeafJob>& jobs, std::vector<std::vector<uint8_t>>& ops)
    {
        ops.assign(jobs.size(), std::vector<uint8_t>());
        size_t a = 0;
        while (a < jobs.size()) {
            std::vector<RgLeafJobDev> dj;
            unsigned long long mat_total = 0, h_total = 0, ops_total = 0;
            size_t b = a;
            while (b < jobs.size() && (b == a || mat_total * sizeof(RsWord2) < (1ull << 31))) {
                const RgLeafJob& j = jobs[b];
                const RgTask& t = tasks[(size_t)j.task];
                RgLeafJobDev d;
                d.Q = t.Q; d.T = t.T; d.q0 = j.q0; d.qn = j.qn; d.t0 = j.t0; d.tn = j.tn; d.cap = j.qn + j.tn; d.pad_ = 0;
                d.mat_off = mat_total; d.h_off = h_total; d.ops_off = ops_total;
                mat_total += (unsigned long long)((j.qn + 31) / 32) * (unsigned long long)j.tn;
                h_total += 2ull * ((unsigned long long)j.tn + 1);
                ops_total += (unsigned long long)d.cap;
                dj.push_back(d);
                ++b;
            }
            std::vector<uint8_t> h_ops;
            std::vector<int> h_n;
            if (failed || !ex.run_leaves(dj, mat_total, h_total, ops_total, h_ops, h_n)) { failed = true; h_ops.assign((size_t)ops_total, 0); h_n.assign(dj.size(), 0); }
            for (size_t x = a; x < b; ++x) {
                const RgLeafJobDev& d = dj[x - a];
                const int n = h_n[x - a];
                ops[x].assign(h_ops.begin() + (long)(d.ops_off + (unsigned long long)(d.cap - n)), h_ops.begin() + (long)(d.ops_off + (unsigned long long)d.cap));
            }
            a = b;
        }
    }
};

// edlib.cpp:1139-1141: bytes of alignment data the traceback would need; below 1 MiB it is used, else Hirschberg
static inline long long rg_traceback_bytes(int qn, int tn)
{
    const long long blocks = (qn + 63) / 64;
    return (long long)(2 * 8 + 4) * blocks * tn + (long long)2 * 4 * tn;
}

template <class Backend>
void rg_run(Backend& be, const std::vector<RgTask>& tasks, double error, int min_align_size, int k_match, std::vector<RgResult>& out)
{
    const size_t n = tasks.size();
    out.assign(n, RgResult());
    // ---- edit distance of every task (edlibAlign NW with k = tolerance: none above it)
    std::vector<RgColJob> cj;
    std::vector<std::vector<int>> cols;
    for (size_t i = 0; i < n; ++i) {
        const RgTask& t = tasks[i];
        const int qn0 = t.qto > t.qfrom ? t.qto - t.qfrom : 0, tn0 = t.tto > t.tfrom ? t.tto - t.tfrom : 0;
        cj.push_back({(int)i, t.qfrom, qn0, 1, t.tfrom, tn0, 1});
    }
    be.columns(tasks, cj, cols);
    struct Node { int task, q0, qn, t0, tn, best, left, right; std::vector<uint8_t> ops; bool done; };
    std::vector<Node> nodes;
    std::vector<int> root(n, -1), frontier;
    std::vector<int> dist(n, 0);
    for (size_t i = 0; i < n; ++i) {
        const RgTask& t = tasks[i];
        const int qn = t.qto - t.qfrom, tn = t.tto - t.tfrom;
        if (qn <= 0 || tn <= 0) continue;
        dist[i] = cols[i][(size_t)qn];
        if (t.tolerance >= 0 && dist[i] > t.tolerance) continue;            // numLocations == 0
        const int align_len = tn - 1;                                       // endLocations[0] - startLocations[0], :165
        if (align_len < min_align_size) continue;
        if (dist[i] / (double)align_len > error) continue;
        root[i] = (int)nodes.size();
        frontier.push_back((int)nodes.size());
        nodes.push_back({(int)i, t.qfrom, qn, t.tfrom, tn, dist[i], -1, -1, {}, false});
    }
    // ---- the recursion, one level of all tasks at a time
    std::vector<char> failed(n, 0);
    while (!frontier.empty()) {
        std::vector<RgLeafJob> lj;
        std::vector<int> lj_node, split_node;
        cj.clear();
        for (int ni : frontier) {
            Node& nd = nodes[(size_t)ni];
            if (nd.qn == 0 || nd.tn == 0) {                                 // edlib.cpp:1113-1120
                nd.ops.assign((size_t)(nd.qn + nd.tn), nd.qn == 0 ? 2 : 1);
                nd.done = true;
            } else if (rg_traceback_bytes(nd.qn, nd.tn) < 1024 * 1024) {
                lj.push_back({nd.task, nd.q0, nd.qn, nd.t0, nd.tn});
                lj_node.push_back(ni);
            } else {
                const int lw = nd.tn / 2, rw = nd.tn - lw;
                cj.push_back({nd.task, nd.q0, nd.qn, 1, nd.t0, lw, 1});                                         // left half
                cj.push_back({nd.task, nd.q0 + nd.qn - 1, nd.qn, -1, nd.t0 + nd.tn - 1, rw, -1});             // right half, reversed
                split_node.push_back(ni);
            }
        }
        std::vector<std::vector<uint8_t>> lops;
        if (!lj.empty()) be.leaves(tasks, lj, lops);
        for (size_t x = 0; x < lj.size(); ++x) { nodes[(size_t)lj_node[x]].ops.swap(lops[x]); nodes[(size_t)lj_node[x]].done = true; }
        if (!cj.empty()) be.columns(tasks, cj, cols);
        std::vector<int> next;
        for (size_t x = 0; x < split_node.size(); ++x) {
            const int ni = split_node[x];
            const int qn = nodes[(size_t)ni].qn, tn = nodes[(size_t)ni].tn, best = nodes[(size_t)ni].best;
            const int lw = tn / 2, rw = tn - lw;
            const std::vector<int>& left = cols[2 * x];         // left[i] = D(q[0..i), t[0..lw))
            const std::vector<int>& right = cols[2 * x + 1];    // right[r] = D(last r query bases, last rw template bases)
            int split = -2, ls = 0, rs = 0;
            for (int i = 0; i <= qn - 2; ++i) {
                const int l = left[(size_t)i + 1], r = right[(size_t)(qn - 1 - i)];
                if (l + r == best) { split = i; ls = l; rs = r; break; }
            }
            if (split == -2 && lw + right[(size_t)qn] == best) { split = -1; ls = lw; rs = right[(size_t)qn]; }
            if (split == -2 && left[(size_t)qn] + rw == best) { split = qn - 1; ls = left[(size_t)qn]; rs = rw; }
            if (split == -2) { failed[(size_t)nodes[(size_t)ni].task] = 1; continue; }
            const int ul = split + 1;
            const Node p = nodes[(size_t)ni];
            const int a = (int)nodes.size();
            nodes.push_back({p.task, p.q0, ul, p.t0, lw, ls, -1, -1, {}, false});
            nodes.push_back({p.task, p.q0 + ul, p.qn - ul, p.t0 + lw, rw, rs, -1, -1, {}, false});
            nodes[(size_t)ni].left = a; nodes[(size_t)ni].right = a + 1;
            next.push_back(a); next.push_back(a + 1);
        }
        frontier.swap(next);
    }
    // ---- string the leaves together, trim (edlib_wrapper.c:197-247)
    for (size_t i = 0; i < n; ++i) {
        if (root[i] < 0 || failed[i]) continue;
        std::vector<uint8_t> ops;
        std::vector<int> stack(1, root[i]);
        bool whole = true;
        while (!stack.empty()) {
            const int ni = stack.back();
            stack.pop_back();
            const Node& nd = nodes[(size_t)ni];
            if (nd.left >= 0) { stack.push_back(nd.right); stack.push_back(nd.left); }
            else if (nd.done) ops.insert(ops.end(), nd.ops.begin(), nd.ops.end());
            else whole = false;
        }
        if (!whole) continue;
        const RgTask& t = tasks[i];
        const int qn = t.qto - t.qfrom, tn = t.tto - t.tfrom, len = (int)ops.size();
        int from = 0, pq = 0, pt = 0, m = 0;
        while (from < len) {
            const int op = ops[(size_t)from];
            m = op == 0 ? m + 1 : 0;
            if (op != 2) ++pq;
            if (op != 1) ++pt;
            ++from;
            if (m == k_match) break;
        }
        if (m != k_match) continue;
        from -= k_match; pq -= k_match; pt -= k_match;
        int to = len, tq = 0, tt = 0;
        m = 0;
        while (to) {
            const int op = ops[(size_t)to - 1];
            m = op == 0 ? m + 1 : 0;
            if (op != 2) ++tq;
            if (op != 1) ++tt;
            --to;
            if (m == k_match) break;
        }
        to += k_match; tq -= k_match; tt -= k_match;
        RgResult& r = out[i];
        r.ok = 1;
        r.qoff = t.qfrom + pq; r.qend = t.qfrom + (qn - tq); r.toff = t.tfrom + pt; r.tend = t.tfrom + (tn - tt);
        r.ops.assign(ops.begin() + from, ops.begin() + to);
        r.aln_size = to - from;
        r.n_match = 0;
        for (uint8_t op : r.ops) r.n_match += op == 0;
    }
}
