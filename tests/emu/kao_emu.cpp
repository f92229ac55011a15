// kao_emu.cpp — TEST INFRASTRUCTURE.  The engine's own device code (csrc/kao_device.cuh: candidate
// generator, full evaluator, delta evaluator) compiled for the host on top of warp_emu.hpp, with the
// engine's own host model (csrc/kao_host.hpp) and shared-memory plan (csrc/kao_plan.hpp) deciding
// layout and evaluator configuration exactly as kao_create / dispatch do.  tests/test_device_emulation.py
// checks it against the committed golden streams and the oracle restatement, so the arithmetic of the
// CUDA path is exercised by the CPU suite as well; the kernels themselves (staging, grid barrier,
// cross-GPU exchange) are only tested on the GPU.  Never linked into libkao.so.
#define KAO_HOST_EMU 1
#include "warp_emu.hpp"

#include "../../kafka_assignment_optimizer_b200/csrc/kao_device.cuh"
#include "../../kafka_assignment_optimizer_b200/csrc/kao_device_t.cuh"
#include "../../kafka_assignment_optimizer_b200/csrc/kao_plan.hpp"
#include "../../kafka_assignment_optimizer_b200/csrc/kao_host.hpp"

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace kao;

namespace {

struct Emu {
    HostModel hm;
    Params prm{};
    Consts cs{};
    bool oh = false;                     // leader one-hot plane kept behind the bit-plane (has_oh_plane)
    int nph = 3, rack = 0, obj = 0;      // evaluator configuration (EvalCfg<W, NPH, kRack, kObj>)
    std::vector<uint32_t> bits;          // [W or 2W][Ppad]
    std::vector<uint8_t> leader;         // [Ppad]
    std::vector<uint16_t> D, DL;
    int nD = 0, nL = 0;
    std::vector<uint32_t> prow;          // [kMaxOps * W]
    // column-major evaluator (kao_device_t.cuh): supported shape, in use, the five transposed planes
    bool trans_ok = false;
    int trans = 0;                       // 0 row-major evaluator, 1.. forms (schedules) of the column-major one
    int nW = 0;
    std::vector<uint32_t> T, Z;          // transposed planes, term planes of the objective
    int n_invalid = 0;                   // partitions led from a slot that holds none of their replicas (s_counts[2] of the kernel)
    std::string err;
};

template <int W> void build_T(Emu &e)
{
    const HostModel &m = e.hm;
    e.T.assign((size_t)kTPlanes * 32 * W * e.nW, 0);
    for (int q = 0; q < kTPlanes; ++q)
        for (int s = 0; s < 32 * W; ++s)
            for (int w = 0; w < e.nW; ++w)
                e.T[t_word(q, s, w, e.nW, 32 * W)] = t_gather<W>(q, s, w, e.bits.data(), e.leader.data(), m.Ppad);
    e.Z.assign((size_t)(kZPlanes + kAPlanes<W>()) * e.nW, 0);
    for (int j = 0; j < kZPlanes; ++j)
        for (int w = 0; w < e.nW; ++w) e.Z[(size_t)j * e.nW + w] = z_gather<W>(e.prm, j, w, e.bits.data(), e.leader.data());
    for (int b = 0; b < kAPlanes<W>(); ++b)
        for (int w = 0; w < e.nW; ++w) e.Z[(size_t)(kZPlanes + b) * e.nW + w] = a_gather<W>(b, w, e.bits.data(), m.Ppad);
}

uint32_t oh_word(uint32_t x, uint32_t ld, int w) { return ((int)(ld >> 5) == w) ? (x & (1u << (ld & 31u))) : 0u; }

// rebuild_lists of kao_kernels.cuh: ascending partitions that miss a home slot (D) / hold home slot
// 0 without being led from it (DL)
void rebuild_lists(Emu &e)
{
    const HostModel &m = e.hm;
    e.nD = e.nL = 0;
    e.n_invalid = 0;
    for (int p = 0; p < m.P; ++p) {
        const int ld = e.leader[p];
        if (!(ld < 32 * m.W && ((e.bits[(size_t)(ld >> 5) * m.Ppad + p] >> (ld & 31)) & 1u))) ++e.n_invalid;
    }
    for (int p = 0; p < m.P; ++p) {
        const uint32_t h4 = m.homeT[p];
        bool miss = false, ldis = false;
        for (int i = 0; i < 4; ++i) {
            const int hs = (h4 >> (8 * i)) & 0xFF;
            if (hs == 0xFF) continue;
            const bool has = (e.bits[(size_t)(hs >> 5) * m.Ppad + p] >> (hs & 31)) & 1u;
            miss |= !has;
            if (i == 0) ldis = has && ((int)e.leader[p] != hs);
        }
        if (miss) e.D[e.nD++] = (uint16_t)p;
        if (ldis) e.DL[e.nL++] = (uint16_t)p;
    }
}

void set_base(Emu &e, const std::vector<uint32_t> &bitsT, const std::vector<uint8_t> &leader)
{
    const HostModel &m = e.hm;
    e.bits.assign((size_t)(e.oh ? 2 : 1) * m.W * m.Ppad, 0);
    std::copy(bitsT.begin(), bitsT.end(), e.bits.begin());
    e.leader = leader;
    if (e.oh)
        for (int p = 0; p < m.Ppad; ++p)
            for (int w = 0; w < m.W; ++w)
                e.bits[(size_t)(m.W + w) * m.Ppad + p] = oh_word(e.bits[(size_t)w * m.Ppad + p], e.leader[p], w);
    e.prm.bitsT = e.bits.data();
    e.prm.leader = e.leader.data();
    rebuild_lists(e);
    if (e.trans_ok) { if (m.W == 1) build_T<1>(e); else build_T<2>(e); }
}

template <class Cfg> struct Run {
    static constexpr int W = Cfg::W;
    template <bool kSmall> static Gen<W, false, kSmall> make_gen(Emu &e, int lane)
    {
        Gen<W, false, kSmall> g;
        g.bitsT = e.bits.data(); g.leader = e.leader.data(); g.cs = &e.cs; g.d = &e.prm;
        g.prow = e.prow.data(); g.lane = lane;
        g.D = e.D.data(); g.DL = e.DL.data(); g.nD = e.nD; g.nL = e.nL;
        return g;
    }
    // the per-thread generator of the column-major kernels: every lane generates the candidate for itself ("first holder
    // of slot s" = scan of the transposed planes), lane 0 parks the patched rows where the evaluator reads them
    static void gen_thread(Emu &e, int lane, uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size, PatchSet &ps)
    {
        Gen<W, true, true> tg;
        tg.bitsT = e.bits.data(); tg.leader = e.leader.data(); tg.cs = &e.cs; tg.d = &e.prm; tg.prow = nullptr; tg.lane = 0;
        tg.D = e.D.data(); tg.DL = e.DL.data(); tg.nD = e.nD; tg.nL = e.nL;
        tg.T = e.T.data(); tg.tnW = e.nW; tg.t_leaders_valid = e.n_invalid == 0;
        uint32_t rows[kMaxOps][W];
        for (int i = 0; i < kMaxOps; ++i)
            for (int t = 0; t < W; ++t) rows[i][t] = 0;
        tg.run(seed, round, idx, round_size, ps, rows);
        if (lane == 0)
            for (int i = 0; i < kMaxOps; ++i)
                for (int t = 0; t < W; ++t) e.prow[i * W + t] = rows[i][t];
    }
    // generate + evaluate in full, as one warp of the search kernels does
    static unsigned long long key(Emu &e, uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size)
    {
        unsigned long long out = 0;
        const uint32_t *objT = Cfg::kObj > 0 ? e.prm.planesT : e.prm.swT;
        emu::run_warp([&](int lane) {
            uint32_t no_rows[kMaxOps][W];
            PatchSet ps;
            if constexpr (W <= 2) {
                if (e.trans) gen_thread(e, lane, seed, round, idx, round_size, ps);               // as the column-major kernels
                else make_gen<false>(e, lane).run(seed, round, idx, round_size, ps, no_rows);
            } else {
                make_gen<false>(e, lane).run(seed, round, idx, round_size, ps, no_rows);
            }
            __syncwarp();                                     // __syncthreads() of the kernels
            int viol, obj;
            if constexpr (W <= 2) {
                // forms of the column-major evaluator: 1 as the engine picks it (32-word specialisation where it
                // applies, default popcount streams), 2 run-time word count, 3 a POPC per word, 4 three POPC per four
                // words, 5 Harley-Seal on the column totals, 6 Harley-Seal on both streams
                const bool fixed = e.nW == 32;
                uint32_t prows[kMaxOps][W];
                for (int i = 0; i < kMaxOps; ++i)
                    for (int t = 0; t < W; ++t) prows[i][t] = ps.p[i] >= 0 ? e.prow[i * W + t] : 0u;
                int pviol = 0, pobj = 0, pcount = 0;
                LaneBounds<W> lb;
                lb.load(&e.cs, lane, e.prm.R);
                alignas(16) uint8_t pdelta[32 * W];                                      // every fiber keeps its own copy (the kernel: the warp's batch)
                if (e.trans) {                                                           // per candidate, as the per-thread generator does
                    patch_terms<W>(e.prm, ps, prows, pviol, pobj, pcount);
                    patch_column_deltas<W>(e.prm, ps, prows, e.bits.data(), e.leader.data(), pdelta);
                }
#define KAO_EMU_T(NW_, POP_) eval_candidate_t<EvalCfgT<W, NW_, 1, POP_>, true>(e.prm, e.T.data(), e.nW, e.bits.data(), e.Z.data(), lb, ps, pdelta, pviol, pobj, pcount, lane, viol, obj)
                if (e.trans == 1 && fixed) KAO_EMU_T(32, 0x22);
                else if (e.trans == 1 || e.trans == 2) KAO_EMU_T(0, 0x22);
                else if (e.trans == 3 && fixed) KAO_EMU_T(32, 0x00);
                else if (e.trans == 3) KAO_EMU_T(0, 0x00);
                else if (e.trans == 4 && fixed) KAO_EMU_T(32, 0x11);
                else if (e.trans == 4) KAO_EMU_T(0, 0x11);
                else if (e.trans == 5 && fixed) KAO_EMU_T(32, 0x23);
                else if (e.trans == 5) KAO_EMU_T(0, 0x23);
                else if (e.trans == 6 && fixed) KAO_EMU_T(32, 0x33);
                else if (e.trans == 6) KAO_EMU_T(0, 0x33);
#undef KAO_EMU_T
                else eval_candidate<Cfg, true>(e.prm, e.bits.data(), e.leader.data(), objT, &e.cs, ps, e.prow.data(), lane, viol, obj);
            } else {
                eval_candidate<Cfg, true>(e.prm, e.bits.data(), e.leader.data(), objT, &e.cs, ps, e.prow.data(), lane, viol, obj);
            }
            if (lane == 0) out = pack_key(viol, obj, idx, e.prm.key_obj_bits);
        });
        return out;
    }
    // the winner becomes the base (search_persistent_kernel, after the grid barrier)
    static void apply(Emu &e, uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size)
    {
        PatchSet win;
        emu::run_warp([&](int lane) {
            uint32_t no_rows[kMaxOps][W];
            PatchSet ps;
            if constexpr (W <= 2) {
                if (e.trans) gen_thread(e, lane, seed, round, idx, round_size, ps);
                else make_gen<false>(e, lane).run(seed, round, idx, round_size, ps, no_rows);
            } else {
                make_gen<false>(e, lane).run(seed, round, idx, round_size, ps, no_rows);
            }
            if (lane == 0) win = ps;
        });
        const int Ppad = e.hm.Ppad;
        for (int i = 0; i < win.n; ++i) {
            if constexpr (W <= 2) {
                if (e.trans_ok) {                   // as the kernel does: every lane rewrites its own slots' words
                    uint32_t newrow[W];
                    for (int w = 0; w < W; ++w) newrow[w] = e.prow[i * W + w];
                    for (int lane = 0; lane < 32; ++lane)
                        t_patch_row<W>(e.prm, e.T.data(), e.Z.data(), e.nW, win.p[i], newrow, win.ld[i], lane);
                }
            }
            for (int w = 0; w < W; ++w) {
                const uint32_t v = e.prow[i * W + w];
                e.bits[(size_t)w * Ppad + win.p[i]] = v;
                if (e.oh) e.bits[(size_t)(W + w) * Ppad + win.p[i]] = oh_word(v, win.ld[i], w);
            }
            e.leader[win.p[i]] = (uint8_t)win.ld[i];
        }
        rebuild_lists(e);
    }
    // delta mode (search_persistent_kernel<..., kDelta>): per-thread generator + delta evaluator on
    // the base's totals; no inverted lists here (the generator's linear-scan form)
    static unsigned long long key_delta(Emu &e, uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size,
                                        const int *cnt, const int *lcnt, const int *rc, int base_viol, int base_obj)
    {
        Gen<W, true> tg;
        tg.bitsT = e.bits.data(); tg.leader = e.leader.data(); tg.cs = &e.cs; tg.d = &e.prm; tg.prow = nullptr; tg.lane = 0;
        tg.D = e.D.data(); tg.DL = e.DL.data(); tg.nD = e.nD; tg.nL = e.nL;
        const uint32_t *objT = Cfg::kObj > 0 ? e.prm.planesT : e.prm.swT;
        const MemRef<true> m_obj(objT);
        PatchSet ps;
        uint32_t rows[kMaxOps][W];
        tg.run(seed, round, idx, round_size, ps, rows);
        int viol, obj;
        delta_eval<Cfg>(e.prm, e.bits.data(), e.leader.data(), m_obj, &e.cs, ps, rows, cnt, lcnt, rc, base_viol, base_obj, viol, obj);
        return pack_key(viol, obj, idx, e.prm.key_obj_bits);
    }
};

// explicit candidate from HBM-style buffers (eval_batch_kernel): general rack form, packed entries
template <int W> void eval_explicit(Emu &e, const uint32_t *cb, const uint8_t *cl, long long &viol, long long &obj)
{
    int v = 0, o = 0;
    emu::run_warp([&](int lane) {
        PatchSet ps;
        ps.n = 0;
        for (int i = 0; i < kMaxOps; ++i) { ps.p[i] = -1; ps.ld[i] = 0xFF; }
        int lv, lo;
        eval_candidate<EvalCfg<W, 5, 0, kObjEntries>, false>(e.prm, cb, cl, e.prm.swT, &e.cs, ps, nullptr, lane, lv, lo);
        if (lane == 0) { v = lv; o = lo; }
    });
    viol = v; obj = o;
}

// kao_engine.cu: dispatch / dispatch_w / dispatch_obj
template <class F> auto dispatch(Emu &e, F f)
{
#define KAO_EMU_CASE(W_, NPH_, R_, O_) \
    if (e.hm.W == W_ && e.nph == NPH_ && e.rack == R_ && e.obj == O_) return f(Run<EvalCfg<W_, NPH_, R_, O_>>{});
#define KAO_EMU_RACKS(W_, NPH_, O_) KAO_EMU_CASE(W_, NPH_, 0, O_) KAO_EMU_CASE(W_, NPH_, 3, O_) KAO_EMU_CASE(W_, NPH_, 4, O_) KAO_EMU_CASE(W_, NPH_, 5, O_)
#define KAO_EMU_NARROW(W_, NPH_) KAO_EMU_RACKS(W_, NPH_, 0) KAO_EMU_RACKS(W_, NPH_, 3)
    KAO_EMU_NARROW(1, 5) KAO_EMU_NARROW(2, 5) KAO_EMU_RACKS(4, 5, 0) KAO_EMU_RACKS(8, 5, 0)
    fprintf(stderr, "kao_emu: no evaluator configuration W=%d NPH=%d rack=%d obj=%d\n", e.hm.W, e.nph, e.rack, e.obj);
    abort();
    return f(Run<EvalCfg<1, 5, 0, 0>>{});
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char *kao_emu_last_error(void) { return g_err.c_str(); }

void *kao_emu_create(const kao_problem *pb)
{
    auto e = std::make_unique<Emu>();
    if (!build_host_model(*pb, e->hm, g_err)) return nullptr;
    HostModel &m = e->hm;
    // kao_create: do mask planes + one-hot plane fit next to the base, else packed entries
    const int threads = m.W <= 2 ? 768 : 256;
    SmemPlan plan = make_plan(m.W, m.Ppad, threads / 32, m.nplanes > 0 ? m.nplanes * m.W : 4, m.P, m.RF, m.nplanes > 0);
    if (plan.total > 227u * 1024u && m.nplanes > 0) {
        m.nplanes = 0;
        plan = make_plan(m.W, m.Ppad, threads / 32, 4, m.P, m.RF, false);
    }
    if (plan.total > 227u * 1024u) { g_err = "problem too large for the shared-memory resident search kernel"; return nullptr; }
    e->nph = 5;                                       // one counter depth (kao_engine.cu: dispatch)
    e->rack = !m.hi1 ? 0 : (m.log2S == 3 ? 3 : (m.log2S == 4 ? 4 : 5));
    e->obj = (m.W <= 2 && m.nplanes == 3) ? 3 : 0;
    e->oh = m.W <= 2 && e->obj > 0;
    // kao_set_evaluator: the column-major evaluator covers 8-slot rack fields with C7 = "at most one
    // replica per rack" and an objective of up to eight term planes
    e->trans_ok = m.W <= 2 && m.hi1 && m.log2S == 3 && m.z_ok && column_major_fits(m.W, m.Ppad, 1024, m.P, m.RF);
    e->nW = t_words(m.Ppad);
    fill_consts(m, e->cs);
    Params &p = e->prm;
    p.P = m.P; p.Ppad = m.Ppad; p.B = m.B; p.R = m.R; p.RF = m.RF; p.NS = m.NS; p.log2S = m.log2S;
    set_rf_masks(p);
    p.key_obj_bits = m.key_obj_bits;
    p.ppr_lo = m.ppr_lo; p.ppr_hi = m.ppr_hi; p.dense = m.dense ? 1 : 0;
    p.nentries = m.nentries; p.nplanes = m.nplanes; p.plane_on_leader = m.plane_on_leader;
    for (int c = 0; c < 6; ++c) p.plane_value[c] = m.plane_value[c];
    p.swT = m.swT.data();
    p.planesT = m.nplanes > 0 ? m.planesT.data() : nullptr;
    p.dense_w = m.dense ? m.dense_w.data() : nullptr;
    p.homeT = m.homeT.data();
    p.nz = m.z_ok ? m.nz : 0; p.z_on_leader = m.z_on_leader; p.zslot = m.z_ok ? m.zslot.data() : nullptr;
    for (int j = 0; j < 8; ++j) p.z_value[j] = m.z_value[j];
    p.consts = &e->cs;
    e->D.assign(m.Ppad, 0);
    e->DL.assign(m.Ppad, 0);
    e->prow.assign((size_t)kMaxOps * m.W, 0);
    std::vector<uint32_t> bitsT;
    std::vector<uint8_t> leader;
    initial_base(m, bitsT, leader);
    set_base(*e, bitsT, leader);
    return e.release();
}

int kao_emu_set_evaluator(void *h, int32_t mode)
{
    Emu &e = *static_cast<Emu *>(h);
    if (mode != 0 && !e.trans_ok) { g_err = "column-major evaluator: unsupported layout"; return -1; }
    if (mode < 0 || mode > 6) { g_err = "unknown evaluator form"; return -1; }
    e.trans = mode;
    return 0;
}

void kao_emu_destroy(void *h) { delete static_cast<Emu *>(h); }

// rows the column-major evaluator charged one by one (its slow, exact path) since the last call
long long kao_emu_rows_charged_one_by_one(void)
{
    const long long n = emu_rows_charged_one_by_one;
    emu_rows_charged_one_by_one = 0;
    return n;
}

// 4 ints: words per row, counter planes above the fours (NPH), rack form, objective planes
void kao_emu_config(void *h, int32_t *out)
{
    Emu &e = *static_cast<Emu *>(h);
    out[0] = e.hm.W; out[1] = e.nph; out[2] = e.rack; out[3] = e.obj;
}

void kao_emu_set_base(void *h, const int32_t *replicas)
{
    Emu &e = *static_cast<Emu *>(h);
    std::vector<uint32_t> bitsT;
    std::vector<uint8_t> leader;
    encode_replicas(e.hm, replicas, bitsT, leader);
    set_base(e, bitsT, leader);
}

// replica lists of the base and its evaluation as the identity candidate of the search evaluator
void kao_emu_get_base(void *h, int32_t *replicas, int64_t *violation, int64_t *objective, int32_t *moves)
{
    Emu &e = *static_cast<Emu *>(h);
    std::vector<uint32_t> bitsT(e.bits.begin(), e.bits.begin() + (size_t)e.hm.W * e.hm.Ppad);
    decode_replicas(e.hm, bitsT, e.leader, replicas);
    if (moves) *moves = count_moves(e.hm, replicas);
    // idx + 1 == round_size is the identity candidate (docs/MODEL.md §5)
    const unsigned long long k = dispatch(e, [&](auto r) { return decltype(r)::key(e, 0, 0, 1, 2); });
    if (violation) *violation = (int64_t)key_violation(k, e.prm.key_obj_bits);
    if (objective) *objective = (int64_t)key_objective(k, e.prm.key_obj_bits);
}

void kao_emu_candidate_keys(void *h, uint64_t seed, uint32_t round, uint32_t round_size, uint32_t idx_begin,
                            uint32_t count, uint64_t *keys)
{
    Emu &e = *static_cast<Emu *>(h);
    dispatch(e, [&](auto r) {
        for (uint32_t i = 0; i < count; ++i) keys[i] = decltype(r)::key(e, seed, round, idx_begin + i, round_size);
        return 0;
    });
}

// whole rounds: argmin of the keys, the winner becomes the base (kao_search, one GPU, no early stop)
void kao_emu_search(void *h, uint64_t seed, uint32_t first_round, uint32_t rounds, uint32_t round_size, uint64_t *round_keys)
{
    Emu &e = *static_cast<Emu *>(h);
    dispatch(e, [&](auto r) {
        for (uint32_t t = 0; t < rounds; ++t) {
            unsigned long long best = kKeyNone;
            for (uint32_t idx = 0; idx < round_size; ++idx) best = std::min(best, decltype(r)::key(e, seed, first_round + t, idx, round_size));
            if (round_keys) round_keys[t] = best;
            if (best != kKeyNone) decltype(r)::apply(e, seed, first_round + t, (uint32_t)(best & kIdxMask), round_size);
        }
        return 0;
    });
}

// delta mode keys (rows of up to 64 slots): totals of the base as the kernel builds them per round
int kao_emu_candidate_keys_delta(void *h, uint64_t seed, uint32_t round, uint32_t round_size, uint32_t idx_begin,
                                 uint32_t count, uint64_t *keys)
{
    Emu &e = *static_cast<Emu *>(h);
    const HostModel &m = e.hm;
    std::vector<int> cnt(256, 0), lcnt(256, 0), rc(32 + 4, 0);
    for (int p = 0; p < m.P; ++p) {
        const int ld = e.leader[p];
        bool ok = false;
        for (int w = 0; w < m.W; ++w) {
            const uint32_t xw = e.bits[(size_t)w * m.Ppad + p];
            for (uint32_t b = xw; b; b &= b - 1) {
                const int sl = w * 32 + __ffs(b) - 1;
                ++cnt[sl];
                ++rc[sl >> m.log2S];
            }
            if ((ld >> 5) == w) ok = (xw >> (ld & 31)) & 1u;
        }
        if (ok) ++lcnt[ld];
    }
    int64_t bv, bo;
    std::vector<int32_t> reps((size_t)m.P * m.RF);
    kao_emu_get_base(h, reps.data(), &bv, &bo, nullptr);
    dispatch(e, [&](auto r) {
        for (uint32_t i = 0; i < count; ++i)
            keys[i] = decltype(r)::key_delta(e, seed, round, idx_begin + i, round_size, cnt.data(), lcnt.data(), rc.data(), (int)bv, (int)bo);
        return 0;
    });
    return 0;
}

// explicit candidates (kao_eval / eval_batch_kernel): replicas [n][P*RF]
void kao_emu_eval(void *h, const int32_t *replicas, int32_t n, int64_t *violation, int64_t *objective)
{
    Emu &e = *static_cast<Emu *>(h);
    const HostModel &m = e.hm;
    for (int i = 0; i < n; ++i) {
        std::vector<uint32_t> cb;
        std::vector<uint8_t> cl;
        encode_replicas(m, replicas + (size_t)i * m.P * m.RF, cb, cl);
        long long v = 0, o = 0;
        switch (m.W) {
        case 1: eval_explicit<1>(e, cb.data(), cl.data(), v, o); break;
        case 2: eval_explicit<2>(e, cb.data(), cl.data(), v, o); break;
        case 4: eval_explicit<4>(e, cb.data(), cl.data(), v, o); break;
        default: eval_explicit<8>(e, cb.data(), cl.data(), v, o); break;
        }
        violation[i] = v;
        objective[i] = o;
    }
}

}  // extern "C"
