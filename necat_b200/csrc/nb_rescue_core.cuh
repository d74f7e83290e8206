// nb_rescue_core.cuh -- one direction of the O(nd) local alignment of oc2cns -r 1, written against a small "warp"
// interface so that the very same code runs in the CUDA kernel (nb_rescue.cu: W = GpuWarp, CUDA intrinsics) and in
// a host emulation with 32 threads and barriers (tests/emu_rescue.cpp: W = EmuWarp) that checks it against the
// oracle without a GPU.
//
// Replaces forward_wave / reverse_wave (gapped_align/align.c:382-1041, :1043-1750) as Local_Alignment uses them from
// ocda_go (gapped_align/oc_daligner.c:34-79): one start diagonal, no trace points (nothing downstream reads them), end
// points and difference count only.
//
// A wave (one more difference) is computed by the whole warp, lane = diagonal, 32 diagonals at a time in the order the
// reference sweeps them (forward: highest diagonal first; reverse: lowest first).  Every new furthest point depends
// on the OLD points of its diagonal and its two neighbours only, so the two generations live in two buffers and the
// diagonals of a wave are independent; what IS order dependent -- the running best point, the "last good" point
// and the trim point, all updated only when a point beats the best so far -- is reproduced with an exclusive prefix
// maximum over the sweep order and "highest set bit" picks on ballots.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

#define RS_TRIM_LEN 15               // align.c:200
#define RS_PATH_LEN 60               // align.c:204
#define RS_PATH_TOP (1ull << RS_PATH_LEN)
#define RS_PATH_INT (RS_PATH_TOP - 1)
#define RS_TRIM_MASK ((1 << RS_TRIM_LEN) - 1)
#define RS_TRIM_MLAG 200             // align.c:211
#define RS_WAVE_LAG 30               // align.c:212
#define RS_INT_MAX 0x7fffffff

struct RsSeq {                       // a read of a PackedDB volume, possibly on its reverse strand
    const uint8_t* pac;
    uint64_t off;
    int      len;
    int      strand;
    NB_HD int at(int i) const        // base i of the strand; 4 outside (the sentinels of oc_daligner.c:42-50)
    {
        if (i < 0 || i >= len) return 4;
        const uint64_t p = strand ? off + (uint64_t)(len - 1 - i) : off + (uint64_t)i;
        const int b = (pac[p >> 2] >> (((~p) & 3) << 1)) & 3;
        return strand ? 3 - b : b;
    }
};

struct RsTables {                    // New_Align_Spec(1 - error, 100, uniform base frequencies), align.c:257-301
    int            path_ave;
    const int16_t* tip_table;        // [1 << 15]
    const int16_t* tip_score;        // [1 << 15]
};

struct RsBand {                      // two generations of the wave vectors of one warp; index = diagonal - kbase
    int*      V[2];
    int*      M[2];
    uint64_t* T[2];
    int       cap, kbase;
};

struct RsTip { int apos, bpos, diffs, overflow; };

// host side of the tables (set_table, align.c:242-253): first edit of the window in the top bit
static inline void rs_fill_tables(double error, int* path_ave, int16_t* tip_table, int16_t* tip_score)
{
    static const double bias_factor[10] = { .690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000 };
    const double ave_corr = 1.0 - error;
    const int bias = (int)((0.5 + .025) * 20. - 1.);
    *path_ave = (int)(RS_PATH_LEN * (1. - bias_factor[bias] * (1. - ave_corr)));
    const int mscore = (int)(1000 * bias_factor[bias] * (1. - ave_corr)), dscore = 1000 - mscore;
    for (int p = 0; p <= RS_TRIM_MASK; ++p) {
        int s = 0, mx = 0;
        for (int i = RS_TRIM_LEN - 1; i >= 0; --i) {
            if (s > mx) mx = s;
            s += ((p >> i) & 1) ? mscore : -dscore;
        }
        tip_table[p] = (int16_t)(s - mx);
        tip_score[p] = (int16_t)s;
    }
}

// W: lane(), ballot(bool) -> uint32_t, shfl(int, src) -> int, sync()
template <class W>
NB_HD RsTip rs_wave(W& w, const RsSeq& A, const RsSeq& B, const int k0, const int anti, const int dir, const RsTables& tb, RsBand& bd)
{
    const int lane = w.lane();
    const int sh = dir > 0 ? 0 : 1;                      // reverse: b position y means "b[y - 1] is compared next"
    const int INV = dir > 0 ? -1 : RS_INT_MAX;
#define RS_BETTER(x, y) (dir > 0 ? (x) > (y) : (x) < (y))
    int low = k0, hgh = k0, dif = 0, more = 1, cur = 0, overflow = 0;
    int aclip = dir > 0 ? RS_INT_MAX : -RS_INT_MAX, bclip = dir > 0 ? -RS_INT_MAX : RS_INT_MAX;
    int besta = anti, trima = anti, lasta = anti, morea = anti;
    int besty = (anti - k0) >> 1, trimy = besty, morey = besty;
    int trimd = 0, mored = 0, morem = -1;
    bd.kbase = k0 - bd.cap / 2;
#define RS_IX(k) ((k) - bd.kbase)
    // ---- wave 0: the start diagonal, by lane 0 (history full of matches, align.c:577-583)
    {
        int y = (anti - k0) >> 1, hit_a = 0, hit_b = 0;
        if (lane == 0) {
            for (;;) {
                const int cb = B.at(y - sh);
                if (cb == 4) { hit_b = 1; break; }
                const int ca = A.at(y + k0 - sh);
                if (cb != ca) { hit_a = ca == 4; break; }
                y += dir;
            }
            bd.V[0][RS_IX(k0)] = (y << 1) + k0; bd.M[0][RS_IX(k0)] = RS_PATH_LEN; bd.T[0][RS_IX(k0)] = RS_PATH_INT;
        }
        y = w.shfl(y, 0); hit_a = w.shfl(hit_a, 0); hit_b = w.shfl(hit_b, 0);
        const int c = (y << 1) + k0;
        if (RS_BETTER(c, besta)) { besta = trima = lasta = c; besty = trimy = y; }
        if (hit_b) { more = 0; bclip = k0; }
        if (hit_a) { more = 0; aclip = k0; }
        w.sync();
    }
    // a diagonal that ran into the end of a (aclip) or of b (bclip) leaves the band; the best such point is remembered
#define RS_CLIPS(BUF)                                                                                              \
    if (more == 0) {                                                                                               \
        if (B.at(besty - sh) != 4 && A.at(besta - besty - sh) != 4) more = 1;                                      \
        const int c1__ = dir > 0 ? (hgh >= aclip) : (low <= aclip);                                                \
        if (c1__) {                                                                                                \
            if (dir > 0) hgh = aclip - 1; else low = aclip + 1;                                                    \
            const int mm__ = bd.M[BUF][RS_IX(aclip)];                                                              \
            if (morem <= mm__) { morem = mm__; morea = bd.V[BUF][RS_IX(aclip)]; morey = (morea - aclip) / 2; mored = dif; } \
        }                                                                                                          \
        const int c2__ = dir > 0 ? (low <= bclip) : (hgh >= bclip);                                                \
        if (c2__) {                                                                                                \
            if (dir > 0) low = bclip + 1; else hgh = bclip - 1;                                                    \
            const int mm__ = bd.M[BUF][RS_IX(bclip)];                                                              \
            if (morem <= mm__) { morem = mm__; morea = bd.V[BUF][RS_IX(bclip)]; morey = (morea - bclip) / 2; mored = dif; } \
        }                                                                                                          \
        aclip = dir > 0 ? RS_INT_MAX : -RS_INT_MAX; bclip = dir > 0 ? -RS_INT_MAX : RS_INT_MAX;                    \
    }
    RS_CLIPS(0)
    // ---- waves 1, 2, ...
    while (more && (dir > 0 ? lasta >= besta - RS_TRIM_MLAG : lasta <= besta + RS_TRIM_MLAG)) {
        low -= 1; hgh += 1;
        if (RS_IX(low - 1) < 0 || RS_IX(hgh + 1) >= bd.cap) { overflow = 1; break; }
        const int nxt = cur ^ 1;
        w.sync();                                   // every lane is done reading the previous wave (clips, pruning)
        if (lane == 0) {
            bd.V[cur][RS_IX(low)] = INV; bd.V[cur][RS_IX(hgh)] = INV;
            bd.V[cur][RS_IX(low - 1)] = INV; bd.V[cur][RS_IX(hgh + 1)] = INV;
        }
        w.sync();
        dif += 1;
        const int span = hgh - low + 1;
        for (int base = 0; base < span; base += 32) {
            const int s_idx = base + lane;                          // place in the sweep order
            const bool on = s_idx < span;
            const int k = dir > 0 ? hgh - s_idx : low + s_idx;
            int c = INV, y = 0, m = 0, hit_a = 0, hit_b = 0;
            uint64_t bv = 0;
            if (on) {
                // forward: side = k + 1 (swept before k), far = k - 1; reverse mirrored.  Old generation throughout.
                const int v_side = bd.V[cur][RS_IX(k + dir)], v_self = bd.V[cur][RS_IX(k)], v_far = bd.V[cur][RS_IX(k - dir)];
                int src;                                            // tie order of align.c:737-768 / :1391-1422
                if (RS_BETTER(v_far, v_self)) {
                    if (RS_BETTER(v_side, v_far)) { c = v_side + dir; src = k + dir; } else { c = v_far + dir; src = k - dir; }
                } else {
                    if (RS_BETTER(v_side, v_self)) { c = v_side + dir; src = k + dir; } else { c = v_self + 2 * dir; src = k; }
                }
                m = bd.M[cur][RS_IX(src)]; bv = bd.T[cur][RS_IX(src)];
                // a neighbour outside the old band contributes the "empty" history (t = PATH_INT, n = PATH_LEN, align.c:726-727)
                if (bd.V[cur][RS_IX(src)] == INV) { m = RS_PATH_LEN; bv = RS_PATH_INT; }
                if (bv & RS_PATH_TOP) m -= 1;                       // the difference enters the history
                bv <<= 1;
                y = (c - k) >> 1;
                for (;;) {
                    const int cb = B.at(y - sh);
                    if (cb == 4) { hit_b = 1; break; }
                    const int ca = A.at(y + k - sh);
                    if (cb != ca) { hit_a = ca == 4; break; }
                    y += dir;
                    if ((bv & RS_PATH_TOP) == 0) m += 1;
                    bv = (bv << 1) | 1;
                }
                c = (y << 1) + k;
                bd.V[nxt][RS_IX(k)] = c; bd.M[nxt][RS_IX(k)] = m; bd.T[nxt][RS_IX(k)] = bv;
            }
            // ---- ends of the sequences (align.c:772-790): bclip = extreme diagonal, aclip = the last one swept
            const uint32_t hb = w.ballot(on && hit_b), ha = w.ballot(on && hit_a);
            if (hb | ha) more = 0;
            if (hb) {
                // forward sweeps downwards: the first lane holds the largest k; reverse: the first lane holds the smallest
                const int first = 0;
                int fl = first; { uint32_t x = hb; fl = 0; while (!(x & 1u)) { x >>= 1; ++fl; } }
                const int kk = dir > 0 ? hgh - (base + fl) : low + (base + fl);
                if (dir > 0 ? bclip < kk : bclip > kk) bclip = kk;
            }
            if (ha) {
                int ll = 31; { uint32_t x = ha; while (!(x & 0x80000000u)) { x <<= 1; --ll; } }
                aclip = dir > 0 ? hgh - (base + ll) : low + (base + ll);
            }
            // ---- running best, last good point, trim point (align.c:842-856): only records update them
            int pre = besta;                                        // best before my place in the sweep
            {
                int run = on ? c : INV;
                // inclusive prefix "best" over lanes, then shift by one lane
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = w.shfl(run, lane >= d ? lane - d : lane);
                    if (lane >= d && RS_BETTER(o, run)) run = o;
                }
                const int ex = w.shfl(run, lane > 0 ? lane - 1 : 0);
                if (lane > 0 && RS_BETTER(ex, pre)) pre = ex;
            }
            const bool rec = on && RS_BETTER(c, pre);
            const bool good = rec && m >= tb.path_ave;
            const bool trim = good && tb.tip_table[bv & RS_TRIM_MASK] >= 0 &&
                              tb.tip_table[(bv >> RS_TRIM_LEN) & RS_TRIM_MASK] + tb.tip_score[bv & RS_TRIM_MASK] >= 0;
            const uint32_t br = w.ballot(rec), bg = w.ballot(good), bt = w.ballot(trim);
            if (br) {
                int l = 31; { uint32_t x = br; while (!(x & 0x80000000u)) { x <<= 1; --l; } }
                besta = w.shfl(c, l); besty = w.shfl(y, l);
            }
            if (bg) {
                int l = 31; { uint32_t x = bg; while (!(x & 0x80000000u)) { x <<= 1; --l; } }
                lasta = w.shfl(c, l);
            }
            if (bt) {
                int l = 31; { uint32_t x = bt; while (!(x & 0x80000000u)) { x <<= 1; --l; } }
                trima = w.shfl(c, l); trimy = w.shfl(y, l); trimd = dif;
            }
        }
        w.sync();
        cur = nxt;
        RS_CLIPS(cur)
        // drop the diagonals that lag more than WAVE_LAG behind the best point (align.c:906-914)
        const int lim = dir > 0 ? besta - RS_WAVE_LAG : besta + RS_WAVE_LAG;
        while (hgh >= low) {
            if (RS_BETTER(lim, bd.V[cur][RS_IX(hgh)])) hgh -= 1;
            else { while (RS_BETTER(lim, bd.V[cur][RS_IX(low)])) low += 1; break; }
        }
    }
    RsTip tip;
    if (morem >= 0) { tip.apos = morea - morey; tip.bpos = morey; tip.diffs = mored; }
    else { tip.apos = trima - trimy; tip.bpos = trimy; tip.diffs = trimd; }
    tip.overflow = overflow;
    return tip;
#undef RS_BETTER
#undef RS_IX
#undef RS_CLIPS
}

// ------------------------------------------------------------------------------------------------------------------
// Second half of the rescue, building block: the last column of a GLOBAL edit-distance matrix -- what edlib's
// Hirschberg step needs from each half (edlib/edlib.cpp:1177-1260: scores of the left half's last column, and of the
// reversed right half's).  col[i] = D[i][tn] for i = 0 .. qn, D[0][j] = j, D[i][0] = i.
//
// Myers' bit-vector recurrence on 32-bit words, strip-mined: a strip is 32 words (1024 query rows), lane = word, the
// lanes run along the anti-diagonal of the strip exactly like the forward pass of k_extend (the horizontal delta of
// the row above and the target symbol travel lane to lane in one shuffle).  The horizontal deltas of a strip's bottom
// row are kept per column (hbuf, one byte each, two generations) and enter the first word of the next strip.
// q(i) = Q.at(q0 + i * qstep), t(j) = T.at(t0 + j * tstep): qstep = tstep = -1 walks the reversed sequences.
template <class W>
NB_HD void rs_nw_column(W& w, const RsSeq& Q, const int q0, const int qn, const int qstep, const RsSeq& T, const int t0, const int tn,
                        const int tstep, int* col, int8_t* hbuf0, int8_t* hbuf1)
{
    const int lane = w.lane();
    if (lane == 0) col[0] = tn;
    int base = tn;                                           // D[r0][tn] of the current strip
    int8_t* hin_buf = hbuf0;
    int8_t* hout_buf = hbuf1;
    for (int r0 = 0, strip = 0; r0 < qn; r0 += 1024, ++strip) {
        const int rows = qn - r0 < 1024 ? qn - r0 : 1024;
        const int nw = (rows + 31) >> 5;
        const bool mine = lane < nw;
        uint32_t pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0;
        if (mine) {
            const int rb = r0 + (lane << 5);
            for (int r = 0; r < 32 && rb + r < qn; ++r) {
                const int c = Q.at(q0 + (rb + r) * qstep);
                const uint32_t bit = 1u << r;
                pe0 |= c == 0 ? bit : 0; pe1 |= c == 1 ? bit : 0; pe2 |= c == 2 ? bit : 0; pe3 |= c == 3 ? bit : 0;
            }
        }
        uint32_t Pv = 0xffffffffu, Mv = 0;
        int carry = 0;                                       // (hout + 1) | symbol << 2 of the column this lane just finished
        const int nsteps = tn + nw - 1;
        for (int s = 0; s < nsteps; ++s) {
            int cin = w.shfl(carry, lane > 0 ? lane - 1 : 0);
            const int col_j = s - lane;
            if (lane == 0 && col_j < tn) {
                const int h = strip == 0 ? 1 : (int)hin_buf[col_j];
                cin = (h + 1) | (T.at(t0 + col_j * tstep) << 2);
            }
            if (mine && col_j >= 0 && col_j < tn) {
                const int hin = (cin & 3) - 1, tc = cin >> 2;
                uint32_t Eq = tc == 0 ? pe0 : (tc == 1 ? pe1 : (tc == 2 ? pe2 : pe3));
                const uint32_t Xv = Eq | Mv;
                if (hin < 0) Eq |= 1u;
                const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint32_t Ph = Mv | ~(Xh | Pv);
                uint32_t Mh = Pv & Xh;
                const int hout = (int)(Ph >> 31) - (int)(Mh >> 31);
                Ph <<= 1; Mh <<= 1;
                if (hin < 0) Mh |= 1u; else if (hin > 0) Ph |= 1u;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
                carry = (hout + 1) | (tc << 2);
                if (lane == nw - 1 && rows == 1024) hout_buf[col_j] = (int8_t)hout;     // bottom row of a full strip
            }
        }
        // scores of this strip's rows in the last column: prefix sums of the vertical deltas
        int live = 0;                                        // rows of my word
        if (mine) { live = rows - (lane << 5); if (live > 32) live = 32; }
        const uint32_t lm = live >= 32 ? 0xffffffffu : ((1u << live) - 1u);
        int sum = 0;
        { uint32_t p = Pv & lm, m = Mv & lm; while (p) { p &= p - 1; ++sum; } while (m) { m &= m - 1; --sum; } }
        int incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = w.shfl(incl, lane >= d ? lane - d : lane);
            if (lane >= d) incl += o;
        }
        int v = base + incl - sum;                           // D[r0 + 32 lane][tn]
        for (int b = 0; b < live; ++b) {
            v += (int)((Pv >> b) & 1u) - (int)((Mv >> b) & 1u);
            col[r0 + (lane << 5) + b + 1] = v;
        }
        base += w.shfl(incl, 31);
        w.sync();                                            // hout_buf complete before it is read as hin_buf
        int8_t* x = hin_buf; hin_buf = hout_buf; hout_buf = x;
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------------------------------
// Second half of the rescue, leaf of edlib's recursion: a GLOBAL alignment small enough for the traceback
// (obtainAlignmentTraceback, edlib/edlib.cpp:887-1090: below 1 MiB of alignment data).  Forward pass as in
// rs_nw_column, but every (word, column) keeps two bit-vectors -- Pv (cell = upper + 1) and Ph (cell = left + 1) -- in
// mat[word * tn + column]; the walk back from (qn, tn) prefers up (a query base against a gap, EDLIB_EDOP_INSERT = 1),
// then left (EDLIB_EDOP_DELETE = 2), then the diagonal (match 0 / mismatch 3), like the reference's traceback.
// The operations are written backwards into ops[cap - 1], ops[cap - 2], ...; returns their number (cap >= qn + tn).
struct RsWord2 { uint32_t pv, ph; };

template <class W>
NB_HD int rs_nw_leaf(W& w, const RsSeq& Q, const int q0, const int qn, const RsSeq& T, const int t0, const int tn,
                     RsWord2* mat, int8_t* hbuf0, int8_t* hbuf1, uint8_t* ops, const int cap)
{
    const int lane = w.lane();
    // ---- forward, strip by strip
    int8_t* hin_buf = hbuf0;
    int8_t* hout_buf = hbuf1;
    for (int r0 = 0, strip = 0; r0 < qn; r0 += 1024, ++strip) {
        const int rows = qn - r0 < 1024 ? qn - r0 : 1024;
        const int nw = (rows + 31) >> 5;
        const bool mine = lane < nw;
        uint32_t pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0;
        if (mine) {
            const int rb = r0 + (lane << 5);
            for (int r = 0; r < 32 && rb + r < qn; ++r) {
                const int c = Q.at(q0 + rb + r);
                const uint32_t bit = 1u << r;
                pe0 |= c == 0 ? bit : 0; pe1 |= c == 1 ? bit : 0; pe2 |= c == 2 ? bit : 0; pe3 |= c == 3 ? bit : 0;
            }
        }
        RsWord2* row = mat + (size_t)((r0 >> 5) + lane) * (size_t)tn;
        uint32_t Pv = 0xffffffffu, Mv = 0;
        int carry = 0;
        const int nsteps = tn + nw - 1;
        for (int s = 0; s < nsteps; ++s) {
            int cin = w.shfl(carry, lane > 0 ? lane - 1 : 0);
            const int col_j = s - lane;
            if (lane == 0 && col_j < tn) {
                const int h = strip == 0 ? 1 : (int)hin_buf[col_j];
                cin = (h + 1) | (T.at(t0 + col_j) << 2);
            }
            if (mine && col_j >= 0 && col_j < tn) {
                const int hin = (cin & 3) - 1, tc = cin >> 2;
                uint32_t Eq = tc == 0 ? pe0 : (tc == 1 ? pe1 : (tc == 2 ? pe2 : pe3));
                const uint32_t Xv = Eq | Mv;
                if (hin < 0) Eq |= 1u;
                const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                const uint32_t Ph = Mv | ~(Xh | Pv);
                const uint32_t Mh = Pv & Xh;
                const int hout = (int)(Ph >> 31) - (int)(Mh >> 31);
                uint32_t Phs = Ph << 1, Mhs = Mh << 1;
                if (hin < 0) Mhs |= 1u; else if (hin > 0) Phs |= 1u;
                Pv = Mhs | ~(Xv | Phs);
                Mv = Phs & Xv;
                carry = (hout + 1) | (tc << 2);
                RsWord2 x; x.pv = Pv; x.ph = Ph;
                row[col_j] = x;
                if (lane == nw - 1 && rows == 1024) hout_buf[col_j] = (int8_t)hout;
            }
        }
        w.sync();
        int8_t* x = hin_buf; hin_buf = hout_buf; hout_buf = x;
    }
    w.sync();
    // ---- walk back (lane 0; the leaf is small)
    int n = 0;
    if (lane == 0) {
        int i = qn, j = tn;
        while (i > 0 || j > 0) {
            int op;
            if (j == 0) { op = 1; --i; }                                    // first column: only up
            else if (i == 0) { op = 2; --j; }                               // first row: only left
            else {
                const RsWord2 x = mat[(size_t)((i - 1) >> 5) * (size_t)tn + (size_t)(j - 1)];
                const int b = (i - 1) & 31;
                if ((x.pv >> b) & 1u) { op = 1; --i; }
                else if ((x.ph >> b) & 1u) { op = 2; --j; }
                else { op = Q.at(q0 + i - 1) == T.at(t0 + j - 1) ? 0 : 3; --i; --j; }
            }
            ops[cap - 1 - n] = (uint8_t)op;
            ++n;
        }
    }
    n = w.shfl(n, 0);
    w.sync();
    return n;
}
