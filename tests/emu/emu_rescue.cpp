// emu_rescue.cpp -- TEST INFRASTRUCTURE: runs the warp-level core of the -r 1 local alignment kernel
// (necat_b200/csrc/nb_rescue_core.cuh, the code the CUDA kernel instantiates) on the host, one std::thread per lane
// with barriers for the warp collectives, and compares its end points with the oracle's (oracle/oracle_rescue.c) on
// the candidates of a work directory.  No GPU needed: this is how the kernel logic is checked on the CPU-only CI.
//
//   emu_rescue <wrk_dir> <candidates.bin> <wave_tasks> [column_tasks]
#include <algorithm>
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../necat_b200/csrc/nb_rescue_core.cuh"
#include "../../necat_b200/csrc/nb_rescue_driver.h"
extern "C" {
#include "../../oracle/necat_oracle.h"
#include "../../oracle/oracle_rescue.h"
}

struct EmuShared {
    std::barrier<> bar{32};
    int ival[32];
    bool pred[32];
};

struct EmuWarp {
    int l;
    EmuShared* s;
    int lane() const { return l; }
    void sync() { s->bar.arrive_and_wait(); }
    uint32_t ballot(bool p)
    {
        s->pred[l] = p;
        s->bar.arrive_and_wait();
        uint32_t r = 0;
        for (int i = 0; i < 32; ++i) r |= (uint32_t)s->pred[i] << i;
        s->bar.arrive_and_wait();
        return r;
    }
    int shfl(int v, int src)
    {
        s->ival[l] = v;
        s->bar.arrive_and_wait();
        const int r = s->ival[src & 31];
        s->bar.arrive_and_wait();
        return r;
    }
};

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: emu_rescue wrk_dir candidates max_tasks [error]\n"); return 2; }
    const int max_tasks = atoi(argv[3]);
    const int max_cols = argc > 4 ? atoi(argv[4]) : max_tasks;
    const double error = 0.5;
    OVolumesInfo vi;
    if (o_volumes_info_load(argv[1], &vi)) { fprintf(stderr, "cannot load %s\n", argv[1]); return 2; }
    OVolume vol;
    if (o_volume_load(vi.names[0], &vol)) { fprintf(stderr, "cannot load volume\n"); return 2; }
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    std::vector<OPackedCandidate> pc;
    OPackedCandidate p;
    while (fread(&p, sizeof p, 1, f) == 1) pc.push_back(p);
    fclose(f);

    int path_ave;
    std::vector<int16_t> tt(1 << 15), ts(1 << 15);
    rs_fill_tables(error, &path_ave, tt.data(), ts.data());
    RsTables tb{path_ave, tt.data(), ts.data()};
    const int cap = 8192;
    std::vector<int> V0(cap), V1(cap), M0(cap), M1(cap);
    std::vector<uint64_t> T0(cap), T1(cap);
    ORescue* rs = o_rescue_new(error);

    int n_done = 0, n_bad = 0;
    const size_t stride = pc.size() / (size_t)max_tasks + 1;
    for (size_t ci = 0; ci < pc.size() && n_done < max_tasks; ci += stride) {
        OPackedCandidate q = pc[ci];
        o_normalise_sdir(&q, (uint32_t)vol.seq[q.item[4]].size, (uint32_t)vol.seq[q.item[1]].size);
        OCandidate c;
        o_unpack_candidate(&c, &q);
        const int qsize = (int)vol.seq[c.qid].size, tsize = (int)vol.seq[c.sid].size;
        std::vector<uint8_t> qs((size_t)qsize), tsq((size_t)tsize);
        o_extract(&vol, (uint64_t)c.qid, c.qdir, qs.data());
        o_extract(&vol, (uint64_t)c.sid, 0, tsq.data());
        int qb, qe, tb_, te, diffs;
        const int ok = o_rescue_local(rs, qs.data(), (int)c.qoff, qsize, tsq.data(), (int)c.soff, tsize, 1000, &qb, &qe, &tb_, &te, &diffs);
        (void)ok;
        RsSeq A{vol.pac, vol.seq[c.qid].offset, qsize, c.qdir};
        RsSeq B{vol.pac, vol.seq[c.sid].offset, tsize, 0};
        const int k0 = (int)c.qoff - (int)c.soff, anti = (int)c.qoff + (int)c.soff;
        RsTip fwd[32], rev[32];
        EmuShared sh;
        std::vector<std::thread> th;
        for (int l = 0; l < 32; ++l)
            th.emplace_back([&, l]() {
                EmuWarp w{l, &sh};
                RsBand bd{{V0.data(), V1.data()}, {M0.data(), M1.data()}, {T0.data(), T1.data()}, cap, 0};
                fwd[l] = rs_wave(w, A, B, k0, anti, +1, tb, bd);
                w.sync();
                rev[l] = rs_wave(w, A, B, k0, anti, -1, tb, bd);
            });
        for (auto& t : th) t.join();
        const bool same = fwd[0].apos == qe && fwd[0].bpos == te && rev[0].apos == qb && rev[0].bpos == tb_ &&
                          fwd[0].diffs + rev[0].diffs == diffs && !fwd[0].overflow && !rev[0].overflow;
        bool uniform = true;
        for (int l = 1; l < 32; ++l)
            uniform = uniform && fwd[l].apos == fwd[0].apos && fwd[l].bpos == fwd[0].bpos && fwd[l].diffs == fwd[0].diffs &&
                      rev[l].apos == rev[0].apos && rev[l].bpos == rev[0].bpos && rev[l].diffs == rev[0].diffs;
        if (!uniform)
            for (int l = 1; l < 32; ++l)
                if (fwd[l].apos != fwd[0].apos || fwd[l].bpos != fwd[0].bpos || fwd[l].diffs != fwd[0].diffs || rev[l].apos != rev[0].apos || rev[l].bpos != rev[0].bpos || rev[l].diffs != rev[0].diffs)
                    fprintf(stderr, "  lane %d: fwd (%d,%d,%d) rev (%d,%d,%d) vs lane 0 fwd (%d,%d,%d) rev (%d,%d,%d)\n", l, fwd[l].apos, fwd[l].bpos, fwd[l].diffs,
                            rev[l].apos, rev[l].bpos, rev[l].diffs, fwd[0].apos, fwd[0].bpos, fwd[0].diffs, rev[0].apos, rev[0].bpos, rev[0].diffs);
        if (!same || !uniform) {
            ++n_bad;
            fprintf(stderr, "MISMATCH candidate %zu: oracle (%d,%d)-(%d,%d) d=%d | warp (%d,%d)-(%d,%d) d=%d uniform=%d\n", ci, qb, tb_, qe, te,
                    diffs, rev[0].apos, rev[0].bpos, fwd[0].apos, fwd[0].bpos, fwd[0].diffs + rev[0].diffs, (int)uniform);
        }
        ++n_done;
    }
    // ---- the strip-mined last-column routine (Hirschberg's building block) against the oracle's, both directions
    int n_col = 0, n_col_bad = 0;
    const size_t cstride = pc.size() / (size_t)(max_cols > 0 ? max_cols : 1) + 1;
    for (size_t ci = 0; ci < pc.size() && n_col < max_cols; ci += cstride) {
        OPackedCandidate q = pc[ci];
        o_normalise_sdir(&q, (uint32_t)vol.seq[q.item[4]].size, (uint32_t)vol.seq[q.item[1]].size);
        OCandidate c;
        o_unpack_candidate(&c, &q);
        const int qsize = (int)vol.seq[c.qid].size, tsize = (int)vol.seq[c.sid].size;
        std::vector<uint8_t> qs((size_t)qsize), tsq((size_t)tsize);
        o_extract(&vol, (uint64_t)c.qid, c.qdir, qs.data());
        o_extract(&vol, (uint64_t)c.sid, 0, tsq.data());
        RsSeq A{vol.pac, vol.seq[c.qid].offset, qsize, c.qdir};
        RsSeq B{vol.pac, vol.seq[c.sid].offset, tsize, 0};
        // a query stretch of 1 to 3 strips against a shorter template stretch; even tasks forwards, odd tasks backwards
        const int qn = std::min(qsize - 1, 700 + (int)(ci * 37 % 1900)), tn = std::min(tsize - 1, 150 + (int)(ci * 53 % 500));
        const int back = n_col & 1;
        const int q0 = back ? qn - 1 : 0, t0 = back ? tn - 1 : 0, step = back ? -1 : 1;
        std::vector<int> want((size_t)qn + 1), got((size_t)qn + 1, -777);
        o_nw_last_column(qs.data() + q0, qn, tsq.data() + t0, tn, step, step, want.data());
        std::vector<int8_t> h0((size_t)tn + 1), h1((size_t)tn + 1);
        EmuShared sh;
        std::vector<std::thread> th;
        for (int l = 0; l < 32; ++l)
            th.emplace_back([&, l]() { EmuWarp w{l, &sh}; rs_nw_column(w, A, q0, qn, step, B, t0, tn, step, got.data(), h0.data(), h1.data()); });
        for (auto& t : th) t.join();
        if (want != got) {
            ++n_col_bad;
            size_t x = 0;
            while (x < want.size() && want[x] == got[x]) ++x;
            fprintf(stderr, "COLUMN MISMATCH candidate %zu (qn %d tn %d back %d) at row %zu: %d vs %d\n", ci, qn, tn, back, x, want[x], got[x]);
        }
        ++n_col;
    }
    // ---- the leaf (forward pass with stored bit-vectors + walk-back) against the oracle's full-matrix traceback
    int n_leaf = 0, n_leaf_bad = 0;
    for (size_t ci = 0; ci < pc.size() && n_leaf < max_cols; ci += cstride) {
        OPackedCandidate q = pc[ci];
        o_normalise_sdir(&q, (uint32_t)vol.seq[q.item[4]].size, (uint32_t)vol.seq[q.item[1]].size);
        OCandidate c;
        o_unpack_candidate(&c, &q);
        const int qsize = (int)vol.seq[c.qid].size, tsize = (int)vol.seq[c.sid].size;
        std::vector<uint8_t> qs((size_t)qsize), tsq((size_t)tsize);
        o_extract(&vol, (uint64_t)c.qid, c.qdir, qs.data());
        o_extract(&vol, (uint64_t)c.sid, 0, tsq.data());
        RsSeq A{vol.pac, vol.seq[c.qid].offset, qsize, c.qdir};
        RsSeq B{vol.pac, vol.seq[c.sid].offset, tsize, 0};
        // aligned stretches around the candidate's start so that the alignment is not degenerate; 1 or 2 strips
        const int qn = std::min(qsize - (int)c.qoff, 300 + (int)(ci * 29 % 1100)), tn = std::min(tsize - (int)c.soff, 280 + (int)(ci * 31 % 900));
        if (qn < 50 || tn < 50) continue;
        const int q0 = (int)c.qoff, t0 = (int)c.soff;
        std::vector<uint8_t> want((size_t)qn + tn), got((size_t)qn + tn);
        const int nwant = o_nw_traceback_ops(qs.data() + q0, qn, tsq.data() + t0, tn, want.data());
        const int nwords = (qn + 31) / 32;
        std::vector<RsWord2> mat((size_t)nwords * tn);
        std::vector<int8_t> h0((size_t)tn + 1), h1((size_t)tn + 1);
        int ngot[32];
        EmuShared sh;
        std::vector<std::thread> th;
        for (int l = 0; l < 32; ++l)
            th.emplace_back([&, l]() { EmuWarp w{l, &sh}; ngot[l] = rs_nw_leaf(w, A, q0, qn, B, t0, tn, mat.data(), h0.data(), h1.data(), got.data(), qn + tn); });
        for (auto& t : th) t.join();
        bool okk = ngot[0] == nwant;
        for (int x = 0; okk && x < nwant; ++x) okk = got[(size_t)(qn + tn - nwant + x)] == want[(size_t)x];
        if (!okk) { ++n_leaf_bad; fprintf(stderr, "LEAF MISMATCH candidate %zu (qn %d tn %d): %d vs %d ops\n", ci, qn, tn, ngot[0], nwant); }
        ++n_leaf;
    }
    printf("emu_rescue: leaves %d, %d mismatches\n", n_leaf, n_leaf_bad);
    // ---- the whole global alignment (host-driven Hirschberg of nb_rescue_driver.h over the emulated cores) against
    //      the oracle's o_rescue_global, on the ranges the local alignment finds
    // the executor the kernels k_rescue_columns / k_rescue_leaves are on the GPU: same pooled jobs, same per-job calls
    struct EmuExec {
        bool run_columns(const std::vector<RgColJobDev>& dj, unsigned long long col_total, unsigned long long h_total, std::vector<int>& h_cols)
        {
            h_cols.assign((size_t)col_total, -1);
            std::vector<int8_t> hbuf((size_t)h_total + 16);
            for (const RgColJobDev& J : dj) {
                EmuShared sh;
                std::vector<std::thread> th;
                for (int l = 0; l < 32; ++l)
                    th.emplace_back([&, l]() { EmuWarp w{l, &sh}; rs_nw_column(w, J.Q, J.q0, J.qn, J.qstep, J.T, J.t0, J.tn, J.tstep, h_cols.data() + J.col_off, hbuf.data() + J.h_off, hbuf.data() + J.h_off + J.tn + 1); });
                for (auto& q : th) q.join();
            }
            return true;
        }
        bool run_leaves(const std::vector<RgLeafJobDev>& dj, unsigned long long mat_total, unsigned long long h_total, unsigned long long ops_total,
                        std::vector<uint8_t>& h_ops, std::vector<int>& h_n)
        {
            h_ops.assign((size_t)ops_total, 9);
            h_n.assign(dj.size(), 0);
            std::vector<RsWord2> mat((size_t)mat_total + 1);
            std::vector<int8_t> hbuf((size_t)h_total + 16);
            for (size_t x = 0; x < dj.size(); ++x) {
                const RgLeafJobDev& J = dj[x];
                int nops[32];
                EmuShared sh;
                std::vector<std::thread> th;
                for (int l = 0; l < 32; ++l)
                    th.emplace_back([&, l]() { EmuWarp w{l, &sh}; nops[l] = rs_nw_leaf(w, J.Q, J.q0, J.qn, J.T, J.t0, J.tn, mat.data() + J.mat_off, hbuf.data() + J.h_off, hbuf.data() + J.h_off + J.tn + 1, h_ops.data() + J.ops_off, J.cap); });
                for (auto& q : th) q.join();
                h_n[x] = nops[0];
            }
            return true;
        }
    } ex;
    RgPooledBackend<EmuExec> be(ex);

    int n_glob = 0, n_glob_bad = 0, n_glob_ok = 0;
    const int max_glob = argc > 5 ? atoi(argv[5]) : 0;
    for (size_t ci = 0; ci < pc.size() && n_glob < max_glob; ++ci) {
        OPackedCandidate q = pc[ci];
        o_normalise_sdir(&q, (uint32_t)vol.seq[q.item[4]].size, (uint32_t)vol.seq[q.item[1]].size);
        OCandidate c;
        o_unpack_candidate(&c, &q);
        const int qsize = (int)vol.seq[c.qid].size, tsize = (int)vol.seq[c.sid].size;
        std::vector<uint8_t> qs((size_t)qsize), tsq((size_t)tsize);
        o_extract(&vol, (uint64_t)c.qid, c.qdir, qs.data());
        o_extract(&vol, (uint64_t)c.sid, 0, tsq.data());
        int qb, qe, tb_, te, diffs;
        if (!o_rescue_local(rs, qs.data(), (int)c.qoff, qsize, tsq.data(), (int)c.soff, tsize, 1000, &qb, &qe, &tb_, &te, &diffs)) continue;
        if (qe - qb > 3200 || qe - qb < 2100) continue;                     // big enough for one or two Hirschberg levels, small enough to emulate
        int oqoff, oqend, otoff, otend, olen; double oid; char *oqa = NULL, *ota = NULL;
        const int ook = o_rescue_global(rs, qs.data(), qb, qe, tsq.data(), tb_, te, diffs, 1000, 4, &oqoff, &oqend, &otoff, &otend, &oid, &oqa, &ota, &olen);
        std::vector<RgTask> tasks(1);
        tasks[0].Q = RsSeq{vol.pac, vol.seq[c.qid].offset, qsize, c.qdir};
        tasks[0].T = RsSeq{vol.pac, vol.seq[c.sid].offset, tsize, 0};
        tasks[0].qfrom = qb; tasks[0].qto = qe; tasks[0].tfrom = tb_; tasks[0].tto = te; tasks[0].tolerance = diffs;
        std::vector<RgResult> res;
        rg_run(be, tasks, 0.5, 1000, 4, res);
        bool same = res[0].ok == ook;
        if (same && ook) {
            same = res[0].qoff == oqoff && res[0].qend == oqend && res[0].toff == otoff && res[0].tend == otend && res[0].aln_size == olen;
            for (int x = 0; same && x < olen; ++x) {
                const int op = res[0].ops[(size_t)x];
                const int want = oqa[x] == '-' ? 2 : (ota[x] == '-' ? 1 : (oqa[x] == ota[x] ? 0 : 3));
                same = op == want;
            }
        }
        free(oqa); free(ota);
        if (!same) { ++n_glob_bad; fprintf(stderr, "GLOBAL MISMATCH candidate %zu: ok %d vs %d\n", ci, res[0].ok, ook); }
        n_glob_ok += ook;
        ++n_glob;
    }
    n_bad += n_glob_bad;
    printf("emu_rescue: global alignments %d (%d ok), %d mismatches\n", n_glob, n_glob_ok, n_glob_bad);
    n_bad += n_col_bad + n_leaf_bad;
    printf("emu_rescue: %d tasks, %d mismatches; %d columns, %d mismatches\n", n_done, n_bad - n_col_bad - n_leaf_bad, n_col, n_col_bad);
    o_rescue_free(rs);
    return n_bad ? 1 : 0;
}
