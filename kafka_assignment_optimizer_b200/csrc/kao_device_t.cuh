// kao_device_t.cuh — column-major ("transposed") full evaluator of one candidate by one warp.
//
// Same result as eval_candidate (kao_device.cuh, docs/MODEL.md §3) for the layout class of the
// headline configuration: rows of up to 64 slots, 8-slot rack fields, "at most one replica of a
// partition per rack" (C7 bounds 0..1), three weighted mask planes.  The base is ALSO kept transposed
// in shared memory: for every slot s a bitmap over the partitions (32 per word).  Five planes:
//   q = 0  replicas            T0[s] bit p  <=>  partition p has a replica on slot s
//   q = 1  leader one-hot      T1[s] bit p  <=>  ... and is led from s
//   q = 2,3,4  objective masks Mc[s] bit p  <=>  row-major mask plane c of partition p has bit s
// The evaluation walks this matrix twice, each time with all data of a constraint inside one lane:
//   rows     (C1, C7)  a lane owns 32 PARTITIONS (one word of every slot): a carry-save counter
//            network over the slot words gives the bit-sliced replica count n_p of its partitions and
//            the bit-sliced count z_p of non-empty rack fields; rows with n = z = RF cost nothing more
//            (no POPC per row), the others are charged |n - RF| + (n - z) one by one;
//   columns  (C2-C6, objective)  a lane owns one SLOT per row word: replica and leader totals of its
//            columns are plain popcount sums — no bit-sliced counters, no cross-lane reduce-scatter —
//            and the objective is popc(column & mask column).
// The candidate's <= 3 patched rows are substituted while loading (columns) / skipped and scored
// from the patch itself (rows), so every row of the candidate is evaluated, none is taken from a
// previous evaluation.  Model: /root/reference/README.md:144-185.
#pragma once
#include "kao_device.cuh"

namespace kao {

// kNW: partition words per slot fixed at compile time (32 = 1024 padded partitions, the headline
// shape: every shared-memory offset of the evaluator is then an immediate), 0 = read at run time.
// The other parameters are SCHEDULES of the same arithmetic (kao_set_schedule; results identical):
//   kSync      how the warps of a CTA meet before an evaluation: 0 block barrier (all warps walk the
//              evaluator together: instruction cache), 1 warp only, 2 one named barrier per scheduler,
//              3 two groups that each hold half of every scheduler's warps (one group can generate
//              while the other evaluates), 4 the same two groups started in anti-phase
//   kCompress  carry-save compression of popcount streams: 0 none, 1 column / leader / bonus totals,
//              2 also the two objective streams (pooled over the lane's slots)
//   kThreads   threads per CTA (0 = threads_for<W>()); fewer threads = more registers per thread
//   kUnroll    unroll factor of the column chunk loop
//   kRoll      1: the row pass loops over pairs of rack fields instead of being unrolled (a quarter of
//              the code: matters when the warps of a scheduler are not in step and share the instruction cache)
//   kFuse      1 (two-word rows, 32 partition words): the row network is not a pass of its own — column
//              chunk j also folds rack field j into it, so that the ALU work of the rows and the
//              popcounts of the columns overlap inside one warp even when all warps are in step
template <int W_, int kNW_ = 0, int kSync_ = 0, int kCompress_ = 1, int kThreads_ = 0, int kUnroll_ = 1, int kRoll_ = 0, int kFuse_ = 0>
struct EvalCfgT {
    static constexpr int W = W_, NPH = 3, kRack = 3, kObj = 3, kNW = kNW_;
    static constexpr int kSync = kSync_, kCompress = kCompress_, kThreads = kThreads_, kUnroll = kUnroll_, kRoll = kRoll_, kFuse = kFuse_;
    static_assert(!kFuse_ || (W_ == 2 && kNW_ == 32), "fused passes: one rack field per column chunk");
    static constexpr bool kTrans = true;
};
constexpr int kTPlanes = 5;

// physical word of (plane q, slot s, partition word w).  Rows are rotated by 4 * (s & 7) words so
// that the 128-bit column loads of a quarter warp (8 consecutive slots) hit 8 different bank groups;
// the 32-bit row loads of a warp (32 consecutive words of one slot) stay conflict-free.
__host__ __device__ __forceinline__ int t_word(int q, int s, int w, int nW, int NSL)
{
    int t = w + (nW >= 32 ? 4 * (s & 7) : 0);
    t -= (t >= nW) ? nW : 0;
    return (q * NSL + s) * nW + t;
}

// ------------------------------------------------------------------------------------------
// rows: C1 + C7 of every partition that is not patched, bit-sliced over 32 partitions per lane
// ------------------------------------------------------------------------------------------
template <int W, bool kShared, int kNW, int kRoll>
__device__ __forceinline__ int rows_vertical(const MemRef<kShared> &T, int nW_rt, int P, int RF, int lane, const PatchSet &ps)
{
    const int nW = kNW ? kNW : nW_rt;
    constexpr int NB = 4 * W;                       // 8-slot blocks = rack fields
    int viol = 0;
    const uint32_t rf0 = (RF & 1) ? ~0u : 0u, rf1 = (RF & 2) ? ~0u : 0u, rf2 = (RF & 4) ? ~0u : 0u, rf3 = (RF & 8) ? ~0u : 0u;
#pragma unroll 1
    for (int w = lane; w < nW; w += 32) {
        const int left = P - 32 * w;                // partitions of this word that exist
        uint32_t valid = left >= 32 ? ~0u : (left <= 0 ? 0u : ((1u << left) - 1u));
#pragma unroll
        for (int i = 0; i < kMaxOps; ++i)           // unused patches hold -1: (-1 >> 5) never equals w
            valid &= ((ps.p[i] >> 5) == w) ? ~(1u << (ps.p[i] & 31)) : ~0u;
        int tk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int t = w + (nW >= 32 ? 4 * k : 0);
            t -= (t >= nW) ? nW : 0;
            tk[k] = t;
        }
        uint32_t ones = 0, twos = 0, fours = 0;     // n_p, weights 1 2 4
        uint32_t e1 = 0, e2 = 0, e4 = 0, e8 = 0;    // n_p, weights 8 16 32 64
        uint32_t z1 = 0, z2 = 0, z4 = 0, z8 = 0;    // non-empty fields of p, 0..8
#pragma unroll (kRoll ? 1 : NB / 2)
        for (int b = 0; b < NB; b += 2) {
            uint32_t o8[2], ne[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = T.ld32((uint32_t)(((b + h) * 8 + k) * nW + tk[k]) * 4u);
                uint32_t a2, b2, qa, qb;
                csa(a2, ones, ones, x[0], x[1]);
                csa(b2, ones, ones, x[2], x[3]);
                csa(qa, twos, twos, a2, b2);
                csa(a2, ones, ones, x[4], x[5]);
                csa(b2, ones, ones, x[6], x[7]);
                csa(qb, twos, twos, a2, b2);
                csa(o8[h], fours, fours, qa, qb);
                ne[h] = (x[0] | x[1] | x[2]) | (x[3] | x[4] | x[5]) | (x[6] | x[7]);
            }
            uint32_t c16, c;
            csa(c16, e1, e1, o8[0], o8[1]);
            c = e2 & c16; e2 ^= c16;
            c16 = e4 & c; e4 ^= c;
            e8 ^= c16;
            uint32_t y2;
            csa(y2, z1, z1, ne[0], ne[1]);
            c = z2 & y2; z2 ^= y2;
            y2 = z4 & c; z4 ^= c;
            z8 ^= y2;
        }
        // partitions whose replica count or non-empty field count is not RF (RF <= 8)
        uint32_t bad = (ones ^ rf0) | (twos ^ rf1) | (fours ^ rf2) | (e1 ^ rf3) | e2 | e4 | e8;
        bad |= (z1 ^ rf0) | (z2 ^ rf1) | (z4 ^ rf2) | (z8 ^ rf3);
        bad &= valid;
        for (uint32_t m = bad; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            const int n = (int)(((ones >> s) & 1u) | (((twos >> s) & 1u) << 1) | (((fours >> s) & 1u) << 2) |
                                (((e1 >> s) & 1u) << 3) | (((e2 >> s) & 1u) << 4) | (((e4 >> s) & 1u) << 5) |
                                (((e8 >> s) & 1u) << 6));
            const int z = (int)(((z1 >> s) & 1u) | (((z2 >> s) & 1u) << 1) | (((z4 >> s) & 1u) << 2) | (((z8 >> s) & 1u) << 3));
            viol += abs(n - RF) + (n - z);
        }
    }
    return viol;
}

// C1 + C7 of one row held row-major (a patched row of the candidate): same terms as row_rack_terms<W, 3>
template <int W> __device__ __forceinline__ int row_terms_hi1_s8(const uint32_t (&x)[W], int RF)
{
    int n = 0, nz = 0;
#pragma unroll
    for (int t = 0; t < W; ++t) {
        n += __popc(x[t]);
        nz += __popc((((x[t] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x[t]) & 0x80808080u);
    }
    return abs(n - RF) + (n - nz);
}

__device__ __forceinline__ void set_comp(uint4 &v, int k, uint32_t clear, uint32_t set)
{
    if (k == 0) v.x = (v.x & ~clear) | set;
    else if (k == 1) v.y = (v.y & ~clear) | set;
    else if (k == 2) v.z = (v.z & ~clear) | set;
    else v.w = (v.w & ~clear) | set;
}

// Row network held across the column loop (kFuse): the planes of rows_vertical, one rack field at a time.
struct RowNet {
    uint32_t ones = 0, twos = 0, fours = 0, e1 = 0, e2 = 0, e4 = 0, e8 = 0, z1 = 0, z2 = 0, z4 = 0, z8 = 0;
    __device__ __forceinline__ void fold(const uint32_t (&x)[8])
    {
        uint32_t a2, b2, qa, qb, o8;
        csa(a2, ones, ones, x[0], x[1]);
        csa(b2, ones, ones, x[2], x[3]);
        csa(qa, twos, twos, a2, b2);
        csa(a2, ones, ones, x[4], x[5]);
        csa(b2, ones, ones, x[6], x[7]);
        csa(qb, twos, twos, a2, b2);
        csa(o8, fours, fours, qa, qb);
        const uint32_t ne = (x[0] | x[1] | x[2]) | (x[3] | x[4] | x[5]) | (x[6] | x[7]);
        uint32_t c = e1 & o8; e1 ^= o8;             // + o8 at weight 8
        uint32_t c2 = e2 & c; e2 ^= c;
        c = e4 & c2; e4 ^= c2;
        e8 ^= c;
        c = z1 & ne; z1 ^= ne;                      // + 1 non-empty field
        c2 = z2 & c; z2 ^= c;
        c = z4 & c2; z4 ^= c2;
        z8 ^= c;
    }
    __device__ __forceinline__ int charge(uint32_t valid, int RF) const
    {
        const uint32_t rf0 = (RF & 1) ? ~0u : 0u, rf1 = (RF & 2) ? ~0u : 0u, rf2 = (RF & 4) ? ~0u : 0u, rf3 = (RF & 8) ? ~0u : 0u;
        uint32_t bad = (ones ^ rf0) | (twos ^ rf1) | (fours ^ rf2) | (e1 ^ rf3) | e2 | e4 | e8;
        bad |= (z1 ^ rf0) | (z2 ^ rf1) | (z4 ^ rf2) | (z8 ^ rf3);
        bad &= valid;
        int viol = 0;
        for (uint32_t m = bad; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            const int n = (int)(((ones >> s) & 1u) | (((twos >> s) & 1u) << 1) | (((fours >> s) & 1u) << 2) |
                                (((e1 >> s) & 1u) << 3) | (((e2 >> s) & 1u) << 4) | (((e4 >> s) & 1u) << 5) |
                                (((e8 >> s) & 1u) << 6));
            const int z = (int)(((z1 >> s) & 1u) | (((z2 >> s) & 1u) << 1) | (((z4 >> s) & 1u) << 2) | (((z8 >> s) & 1u) << 3));
            viol += abs(n - RF) + (n - z);
        }
        return viol;
    }
};

// ------------------------------------------------------------------------------------------
// the whole candidate.  T: the five transposed planes; prow: this warp's patched rows [kMaxOps * W]
// ------------------------------------------------------------------------------------------
template <class Cfg, bool kShared>
__device__ void eval_candidate_t(const Params &d, const uint32_t *Tp, int nW_rt, const Consts *cs, const PatchSet &ps,
                                 const uint32_t *prow, int lane, int &viol_out, int &obj_out)
{
    constexpr int W = Cfg::W, kNW = Cfg::kNW;
    constexpr int NSL = 32 * W;
    const int nW = kNW ? kNW : nW_rt;
    const MemRef<kShared> T(Tp);
    // ---- rows: unpatched partitions from the transposed bit-plane, patched ones from the patch
    int viol = 0;
    RowNet net;                                     // kFuse: filled inside the column loop
    uint32_t net_valid = 0;
    int net_tk[8];
    if constexpr (Cfg::kFuse) {
        const int left = d.P - 32 * lane;           // 32 partition words: word `lane` is this lane's
        net_valid = left >= 32 ? ~0u : (left <= 0 ? 0u : ((1u << left) - 1u));
#pragma unroll
        for (int i = 0; i < kMaxOps; ++i) net_valid &= ((ps.p[i] >> 5) == lane) ? ~(1u << (ps.p[i] & 31)) : ~0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) net_tk[k] = ((lane + 4 * k) & 31) * 4;
    } else {
        viol = rows_vertical<W, kShared, kNW, Cfg::kRoll>(T, nW, d.P, d.RF, lane, ps);
    }
    if (lane < kMaxOps) {
        const int i = lane;
        const int pp = i == 0 ? ps.p[0] : (i == 1 ? ps.p[1] : ps.p[2]);
        if (pp >= 0) {
            uint32_t x[W];
#pragma unroll
            for (int t = 0; t < W; ++t) x[t] = prow[i * W + t];
            viol += row_terms_hi1_s8<W>(x, d.RF);
        }
    }
    // ---- columns: this lane owns slot `lane` of every row word
    int cnt[W], lcnt[W], cnt2[W], lcnt2[W], o0 = 0, o1 = 0, o2 = 0, o2b = 0, o0b = 0, o1b = 0;
#pragma unroll
    for (int t = 0; t < W; ++t) cnt[t] = lcnt[t] = cnt2[t] = lcnt2[t] = 0;
    const int rot = nW >= 32 ? 4 * (lane & 7) : 0;
    const int nch = nW >> 2;
#pragma unroll (Cfg::kUnroll)
    for (int j = 0; j < nch; ++j) {
        int tw = 4 * j + rot;
        tw -= (tw >= nW) ? nW : 0;
        uint4 col[W], oh[W], m0[W], m1[W], m2[W];
#pragma unroll
        for (int t = 0; t < W; ++t) {
            const int s = lane + 32 * t;
            col[t] = T.ld128((uint32_t)((0 * NSL + s) * nW + tw) * 4u);
            oh[t] = T.ld128((uint32_t)((1 * NSL + s) * nW + tw) * 4u);
            m0[t] = T.ld128((uint32_t)((2 * NSL + s) * nW + tw) * 4u);
            m1[t] = T.ld128((uint32_t)((3 * NSL + s) * nW + tw) * 4u);
            m2[t] = T.ld128((uint32_t)((4 * NSL + s) * nW + tw) * 4u);
        }
        if constexpr (Cfg::kFuse) {                 // rack field j of the row network (8 slots x this lane's word)
            uint32_t x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = T.ld32((uint32_t)((j * 8 + k) * (nW * 4) + net_tk[k]));
            net.fold(x);
        }
        // the candidate's patched rows replace their partition's bit in this lane's columns
        if (((ps.p[0] >> 7) == j) | ((ps.p[1] >> 7) == j) | ((ps.p[2] >> 7) == j)) {
#pragma unroll
            for (int i = 0; i < kMaxOps; ++i) {
                const int pp = ps.p[i];
                if ((pp >> 7) == j) {
                    const int k = (pp >> 5) & 3;
                    const uint32_t bit = 1u << (pp & 31);
#pragma unroll
                    for (int t = 0; t < W; ++t) {
                        const bool has = (prow[i * W + t] >> lane) & 1u;
                        const bool led = has && ((int)ps.ld[i] == lane + 32 * t);
                        set_comp(col[t], k, bit, has ? bit : 0u);
                        set_comp(oh[t], k, bit, led ? bit : 0u);
                    }
                }
            }
        }
        // The XU pipe (POPC) is the scarce one here: the four words of a chunk that feed the same total
        // go through one carry-save adder first, a + b + c = 2 * maj + xor, so 3 popcounts replace 4
        // (totals of a column, valid leaders of a column, leader bonus); the doubled parts are summed
        // apart and weighted once per candidate.
        uint32_t hit[4];                            // leader bonus: the one-hot columns of a lane are disjoint
        uint32_t y0[4 * W], y1[4 * W];              // objective streams (kCompress 2)
#pragma unroll
        for (int i = 0; i < 4; ++i) hit[i] = 0;
#pragma unroll
        for (int t = 0; t < W; ++t) {
            if constexpr (Cfg::kCompress >= 1) {
                uint32_t h, l;
                csa(h, l, col[t].x, col[t].y, col[t].z);
                cnt[t] += __popc(l) + __popc(col[t].w);
                cnt2[t] += __popc(h);
                csa(h, l, oh[t].x, oh[t].y, oh[t].z);
                lcnt[t] += __popc(l) + __popc(oh[t].w);
                lcnt2[t] += __popc(h);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t c = comp(col[t], i);
                if constexpr (Cfg::kCompress == 0) {
                    cnt[t] += __popc(c);
                    lcnt[t] += __popc(comp(oh[t], i));
                }
                if constexpr (Cfg::kCompress >= 2) {
                    y0[4 * t + i] = c & comp(m0[t], i);
                    y1[4 * t + i] = c & comp(m1[t], i);
                } else {
                    o0 += __popc(c & comp(m0[t], i));
                    o1 += __popc(c & comp(m1[t], i));
                }
                hit[i] |= comp(oh[t], i) & comp(m2[t], i);
            }
        }
        if constexpr (Cfg::kCompress >= 2) {
            // 4 * W words per stream: every full group of three goes through one carry-save adder
#pragma unroll
            for (int g = 0; g + 3 <= 4 * W; g += 3) {
                uint32_t h, l;
                csa(h, l, y0[g], y0[g + 1], y0[g + 2]);
                o0 += __popc(l);
                o0b += __popc(h);
                csa(h, l, y1[g], y1[g + 1], y1[g + 2]);
                o1 += __popc(l);
                o1b += __popc(h);
            }
#pragma unroll
            for (int g = (4 * W) / 3 * 3; g < 4 * W; ++g) { o0 += __popc(y0[g]); o1 += __popc(y1[g]); }
        }
        if constexpr (Cfg::kCompress >= 1) {
            uint32_t h, l;
            csa(h, l, hit[0], hit[1], hit[2]);
            o2 += __popc(l) + __popc(hit[3]);
            o2b += __popc(h);
        } else {
            o2 += __popc(hit[0]) + __popc(hit[1]) + __popc(hit[2]) + __popc(hit[3]);
        }
    }
    if constexpr (Cfg::kFuse) viol += net.charge(net_valid, d.RF);
#pragma unroll
    for (int t = 0; t < W; ++t) { cnt[t] += 2 * cnt2[t]; lcnt[t] += 2 * lcnt2[t]; }
    o2 += 2 * o2b;
    o0 += 2 * o0b;
    o1 += 2 * o1b;
    // ---- C3 / C4 on this lane's columns, C2/C5 as P - sum of valid leaders, C6 per 8-lane rack group
#pragma unroll
    for (int t = 0; t < W; ++t) {
        const int s = lane + 32 * t;
        viol += band_violation(cnt[t], cs->bnd_rep[s]) + band_violation(lcnt[t], cs->bnd_ldr[s]) - lcnt[t];
        const int tot = __reduce_add_sync(0xFFu << (lane & 24), cnt[t]);
        const int rk = s >> 3;
        if ((lane & 7) == 0 && rk < d.R) viol += max(tot - cs->rack_hi[rk], 0) + max(cs->rack_lo[rk] - tot, 0);
    }
    const int obj = o0 * d.plane_value[0] + o1 * d.plane_value[1] + o2 * d.plane_value[2];
    viol_out = __reduce_add_sync(0xFFFFFFFFu, viol) + d.P;
    obj_out = __reduce_add_sync(0xFFFFFFFFu, obj);
}

// ------------------------------------------------------------------------------------------
// keeping the transposed planes in step with the row-major base (search kernels; tests/emu)
// ------------------------------------------------------------------------------------------
// one word (32 partitions) of plane q, slot s, gathered from the row-major tables
template <int W>
__device__ __forceinline__ uint32_t t_gather(int q, int s, int w, const uint32_t *bitsT, const uint8_t *leader,
                                             const uint32_t *planesT, int Ppad)
{
    uint32_t out = 0;
    const int sw = s >> 5, sb = s & 31;
    for (int b = 0; b < 32; ++b) {
        const int p = 32 * w + b;
        uint32_t bit;
        if (q == 0) bit = (bitsT[(size_t)sw * Ppad + p] >> sb) & 1u;
        else if (q == 1) bit = ((bitsT[(size_t)sw * Ppad + p] >> sb) & 1u) & ((int)leader[p] == s ? 1u : 0u);
        else bit = (planesT[((size_t)(q - 2) * W + sw) * Ppad + p] >> sb) & 1u;
        out |= bit << b;
    }
    return out;
}
// row p of the base changes from (oldrow, oldld) to (newrow, newld): planes 0 and 1 follow
template <int W>
__device__ __forceinline__ void t_patch_row(uint32_t *T, int nW, int p, const uint32_t (&oldrow)[W], uint32_t oldld,
                                            const uint32_t (&newrow)[W], uint32_t newld)
{
    constexpr int NSL = 32 * W;
    const int w = p >> 5;
    const uint32_t bit = 1u << (p & 31);
#pragma unroll
    for (int t = 0; t < W; ++t)
        for (uint32_t m = oldrow[t] ^ newrow[t]; m; m &= m - 1) T[t_word(0, 32 * t + __ffs(m) - 1, w, nW, NSL)] ^= bit;
    if ((int)oldld < NSL && row_has<W>(oldrow, (int)oldld)) T[t_word(1, (int)oldld, w, nW, NSL)] &= ~bit;
    if ((int)newld < NSL && row_has<W>(newrow, (int)newld)) T[t_word(1, (int)newld, w, nW, NSL)] |= bit;
}

}  // namespace kao
