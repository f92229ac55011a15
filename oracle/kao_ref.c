/*
 * TEST INFRASTRUCTURE — plain-C restatement of the candidate-search hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (kafka_assignment_optimizer_b200/csrc) never links
 * it and has no CPU fallback.
 *
 * PARITY STATUS: parity unpinned against lp_solve (see oracle/model.py header).  This file
 * restates, with scalar loops and no bit tricks, the SAME deterministic search the CUDA engine
 * runs (docs/MODEL.md §3-§5): counter-based candidate stream (Philox4x32-10 keyed by seed,
 * counter = (index, round)), full evaluation of C1..C7 + objective per candidate
 * (/root/reference/README.md:144-185), packed (violation, cost, index) argmin per round, winner
 * becomes the next base.  Same seed => bit-identical winner, per-candidate keys and trajectory.
 * The model semantics are validated against oracle/model.py (HiGHS), which is pinned on the
 * README known-answer vector (README.md:83-91).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KAO_MAX_B 256
#define KAO_MAX_R 16
#define KAO_MAX_W 8
#define KAO_MAX_OPS 3

typedef struct {
    int32_t P, B, R, RF, RFcur;
    const uint8_t *rack_of;          /* [B] */
    const uint16_t *wF, *wL;         /* [P*B] README.md:145-146 */
    const int32_t *rep_lo, *rep_hi;  /* [B] C3 README.md:158-161 */
    const int32_t *ldr_lo, *ldr_hi;  /* [B] C4 README.md:163-166 */
    const int32_t *rack_lo, *rack_hi;/* [R] C6 README.md:173-176 */
    int32_t ppr_lo, ppr_hi;          /*     C7 README.md:178-180 */
    const int32_t *cur;              /* [P*RFcur] dense broker index, -1 = absent */
} ref_problem;

#define OBJ_CAP 0xFFFFFFu
#define VIOL_CAP 0xFFFFu
#define IDX_BITS 24

static int W_of(const ref_problem *pb) { return (pb->B + 31) / 32; }

/* ---------------------------------------------------------------- Philox4x32-10 (Random123) */
void kao_ref_philox(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---------------------------------------------------------------- bit helpers (scalar) */
static int row_has(const uint32_t *row, int b) { return (row[b >> 5] >> (b & 31)) & 1u; }
static void row_set(uint32_t *row, int b) { row[b >> 5] |= 1u << (b & 31); }
static void row_clr(uint32_t *row, int b) { row[b >> 5] &= ~(1u << (b & 31)); }
static int row_count(const uint32_t *row, int W)
{
    int n = 0;
    for (int j = 0; j < W; ++j) n += __builtin_popcount(row[j]);
    return n;
}
/* k-th (0-based) set bit in ascending broker order, -1 if fewer */
static int row_kth(const uint32_t *row, int B, int k)
{
    for (int b = 0; b < B; ++b)
        if (row_has(row, b)) { if (k == 0) return b; --k; }
    return -1;
}

/* ---------------------------------------------------------------- full evaluation (docs/MODEL.md §3) */
/* violation = sum over every row of C1..C7 of the amount by which it is missed; objective =
 * README.md:145-146.  A leader that is not one of the partition's replicas violates C2 by 1 and
 * earns no leader weight. */
void kao_ref_eval(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                  int64_t *viol_out, int64_t *obj_out)
{
    const int P = pb->P, B = pb->B, R = pb->R, W = W_of(pb);
    int64_t viol = 0, obj = 0;
    int32_t cnt[KAO_MAX_B] = {0}, lcnt[KAO_MAX_B] = {0}, rc[KAO_MAX_R] = {0};
    for (int p = 0; p < P; ++p) {
        const uint32_t *row = bits + (size_t)p * W;
        int pr[KAO_MAX_R] = {0};
        int n = 0, ld = leader[p];
        int ld_ok = (ld < B) && row_has(row, ld);
        for (int b = 0; b < B; ++b) {
            if (!row_has(row, b)) continue;
            ++n; ++cnt[b]; ++pr[pb->rack_of[b]];
            if (!(ld_ok && b == ld)) obj += pb->wF[(size_t)p * B + b];
        }
        /* bits at positions >= B (padding) are replicas on non-existent brokers: count for C1 */
        for (int b = B; b < W * 32; ++b) if (row_has(row, b)) ++n;
        viol += abs(n - pb->RF);                                   /* C1 */
        if (ld_ok) { ++lcnt[ld]; obj += pb->wL[(size_t)p * B + ld]; }
        else viol += 1;                                            /* C2 (+C5) */
        for (int r = 0; r < R; ++r) {                              /* C7 */
            if (pr[r] > pb->ppr_hi) viol += pr[r] - pb->ppr_hi;
            if (pr[r] < pb->ppr_lo) viol += pb->ppr_lo - pr[r];
        }
    }
    for (int b = 0; b < B; ++b) {
        if (cnt[b] > pb->rep_hi[b]) viol += cnt[b] - pb->rep_hi[b];  /* C3 */
        if (cnt[b] < pb->rep_lo[b]) viol += pb->rep_lo[b] - cnt[b];
        if (lcnt[b] > pb->ldr_hi[b]) viol += lcnt[b] - pb->ldr_hi[b];/* C4 */
        if (lcnt[b] < pb->ldr_lo[b]) viol += pb->ldr_lo[b] - lcnt[b];
        rc[pb->rack_of[b]] += cnt[b];
    }
    for (int r = 0; r < R; ++r) {                                  /* C6 */
        if (rc[r] > pb->rack_hi[r]) viol += rc[r] - pb->rack_hi[r];
        if (rc[r] < pb->rack_lo[r]) viol += pb->rack_lo[r] - rc[r];
    }
    *viol_out = viol; *obj_out = obj;
}

uint64_t kao_ref_pack(int64_t viol, int64_t obj, uint32_t idx)
{
    uint64_t v = viol > VIOL_CAP ? VIOL_CAP : (uint64_t)viol;
    uint64_t c = obj > OBJ_CAP ? 0 : (uint64_t)(OBJ_CAP - obj);
    return (v << 48) | (c << IDX_BITS) | (idx & ((1u << IDX_BITS) - 1));
}

/* ---------------------------------------------------------------- initial base (docs/MODEL.md §4) */
/* cur restricted to the target brokers, order kept (leader = first survivor); surplus replicas
 * (RF lowered) dropped from the tail; missing replicas (broker removed / RF raised) added one at
 * a time on the broker minimising (replicas of p already in that rack, current load, index). */
void kao_ref_init_base(const ref_problem *pb, uint32_t *bits, uint8_t *leader)
{
    const int P = pb->P, B = pb->B, W = W_of(pb);
    int32_t load[KAO_MAX_B] = {0};
    memset(bits, 0, (size_t)P * W * 4);
    for (int p = 0; p < P; ++p) {
        uint32_t *row = bits + (size_t)p * W;
        int n = 0, ld = -1;
        for (int i = 0; i < pb->RFcur && n < pb->RF; ++i) {
            int b = pb->cur[(size_t)p * pb->RFcur + i];
            if (b < 0 || b >= B || row_has(row, b)) continue;
            row_set(row, b); ++n; ++load[b];
            if (ld < 0) ld = b;
        }
        leader[p] = (uint8_t)(ld < 0 ? 0 : ld);
    }
    for (int p = 0; p < P; ++p) {
        uint32_t *row = bits + (size_t)p * W;
        int n = row_count(row, W);
        int had_leader = n > 0;
        while (n < pb->RF && n < B) {
            int pr[KAO_MAX_R] = {0};
            for (int b = 0; b < B; ++b) if (row_has(row, b)) ++pr[pb->rack_of[b]];
            int best = -1;
            for (int b = 0; b < B; ++b) {
                if (row_has(row, b)) continue;
                if (best < 0) { best = b; continue; }
                int ra = pr[pb->rack_of[b]], rb = pr[pb->rack_of[best]];
                if (ra < rb || (ra == rb && load[b] < load[best])) best = b;
            }
            row_set(row, best); ++load[best]; ++n;
            if (!had_leader) { leader[p] = (uint8_t)best; had_leader = 1; }
        }
    }
}

/* ---------------------------------------------------------------- candidate generator (docs/MODEL.md §5) */
static uint32_t mulhi32(uint32_t a, uint32_t n) { return (uint32_t)(((uint64_t)a * n) >> 32); }

typedef struct { int32_t p; uint32_t row[KAO_MAX_W]; uint8_t leader; } ref_patch;
typedef struct { int n; ref_patch e[KAO_MAX_OPS]; } ref_patchset;

/* current view of row p under the patches made so far */
static void view_row(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                     const ref_patchset *ps, int p, uint32_t *row, int *ld)
{
    const int W = W_of(pb);
    memcpy(row, bits + (size_t)p * W, (size_t)W * 4);
    *ld = leader[p];
    for (int i = 0; i < ps->n; ++i)
        if (ps->e[i].p == p) { memcpy(row, ps->e[i].row, (size_t)W * 4); *ld = ps->e[i].leader; }
}
static int touched(const ref_patchset *ps, int p)
{
    for (int i = 0; i < ps->n; ++i) if (ps->e[i].p == p) return 1;
    return 0;
}
static void push_patch(const ref_problem *pb, ref_patchset *ps, int p, const uint32_t *row, int ld)
{
    const int W = W_of(pb);
    for (int i = 0; i < ps->n; ++i)
        if (ps->e[i].p == p) { memcpy(ps->e[i].row, row, (size_t)W * 4); ps->e[i].leader = (uint8_t)ld; return; }
    ref_patch *e = &ps->e[ps->n++];
    e->p = p; memset(e->row, 0, sizeof e->row); memcpy(e->row, row, (size_t)W * 4); e->leader = (uint8_t)ld;
}

/* REPLACE: in partition p, the replica on broker a moves to the first broker >= bt (cyclic) that
 * p does not already use; if a was the leader, the new broker inherits leadership. */
static int op_replace(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                      ref_patchset *ps, int p, int a, int bt, int *b_out)
{
    uint32_t row[KAO_MAX_W]; int ld;
    view_row(pb, bits, leader, ps, p, row, &ld);
    if (!row_has(row, a)) return 0;
    int b = bt, tries = 0;
    while (row_has(row, b)) { b = (b + 1 == pb->B) ? 0 : b + 1; if (++tries > pb->B) return 0; }
    row_clr(row, a); row_set(row, b);
    if (ld == a) ld = b;
    push_patch(pb, ps, p, row, ld);
    *b_out = b;
    return 1;
}
/* first partition q >= p0 (cyclic), not yet patched, that has a replica on broker src */
static int find_holder(const ref_problem *pb, const uint32_t *bits, const ref_patchset *ps,
                       int p0, int src)
{
    const int W = W_of(pb);
    for (int k = 0; k < pb->P; ++k) {
        int q = p0 + k; if (q >= pb->P) q -= pb->P;
        if (touched(ps, q)) continue;
        if (row_has(bits + (size_t)q * W, src)) return q;
    }
    return -1;
}

/* Fills `ps` with the row patches that turn the base into candidate (round, idx).  idx ==
 * round_size-1 is the identity.  Pure function of (base, seed, round, idx). */
void kao_ref_gen_patches(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                         uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size,
                         ref_patchset *ps)
{
    const int P = pb->P, B = pb->B, W = W_of(pb);
    ps->n = 0;
    if (idx + 1 == round_size) return;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c0[4] = {idx, round, 0, 0x4B414F21u}, c1[4] = {idx, round, 1, 0x4B414F21u};
    uint32_t r[4], s[4];
    kao_ref_philox(c0, key, r);
    kao_ref_philox(c1, key, s);
    const uint32_t sel = r[0] & 15u;
    const int p = (int)mulhi32(r[1], (uint32_t)P);
    uint32_t row[KAO_MAX_W]; int ld;
    view_row(pb, bits, leader, ps, p, row, &ld);
    const int n = row_count(row, W);
    if (n == 0) return;
    if (sel == 5 || sel == 6) {
        /* LEADER: the k-th non-leader replica of p becomes its leader (no data moves) */
        if (n < 2) return;
        int k = (int)mulhi32(r[2], (uint32_t)(n - 1));
        for (int b = 0; b < B; ++b) {
            if (!row_has(row, b) || b == ld) continue;
            if (k-- == 0) { push_patch(pb, ps, p, row, b); return; }
        }
        return;
    }
    const int a = row_kth(row, W * 32, (int)mulhi32(r[2], (uint32_t)n));
    int b = -1;
    if (!op_replace(pb, bits, leader, ps, p, a, (int)mulhi32(r[3], (uint32_t)B), &b)) return;
    if (sel <= 4) return;                                   /* single REPLACE */
    if (sel == 7) {                                         /* REPLACE + LEADER on the same p */
        view_row(pb, bits, leader, ps, p, row, &ld);
        if (n < 2) return;
        int k = (int)mulhi32(s[0], (uint32_t)(n - 1));
        for (int c = 0; c < B; ++c) {
            if (!row_has(row, c) || c == ld) continue;
            if (k-- == 0) { push_patch(pb, ps, p, row, c); return; }
        }
        return;
    }
    /* chains: the broker that just gained a replica (b) gives one up from another partition */
    int q = find_holder(pb, bits, ps, (int)mulhi32(s[0], (uint32_t)P), b);
    if (q < 0) return;
    int c = -1;
    const int closed2 = (sel <= 11);                        /* 8..11: swap  a<->b */
    if (!op_replace(pb, bits, leader, ps, q, b, closed2 ? a : (int)mulhi32(s[1], (uint32_t)B), &c))
        return;
    if (sel <= 13) return;                                  /* 12,13: open 2-chain */
    int q2 = find_holder(pb, bits, ps, (int)mulhi32(s[2], (uint32_t)P), c);   /* 14,15: 3-cycle */
    if (q2 < 0) return;
    int d;
    op_replace(pb, bits, leader, ps, q2, c, a, &d);
}

void kao_ref_gen(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                 uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size,
                 uint32_t *out_bits, uint8_t *out_leader)
{
    const int W = W_of(pb);
    ref_patchset ps;
    kao_ref_gen_patches(pb, bits, leader, seed, round, idx, round_size, &ps);
    if (out_bits != bits) memcpy(out_bits, bits, (size_t)pb->P * W * 4);
    if (out_leader != leader) memcpy(out_leader, leader, (size_t)pb->P);
    for (int i = 0; i < ps.n; ++i) {
        memcpy(out_bits + (size_t)ps.e[i].p * W, ps.e[i].row, (size_t)W * 4);
        out_leader[ps.e[i].p] = ps.e[i].leader;
    }
}

/* key of candidate (round, idx) by materialising it and evaluating it in full */
uint64_t kao_ref_candidate_key(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                               uint64_t seed, uint32_t round, uint32_t idx, uint32_t round_size,
                               uint32_t *scratch_bits, uint8_t *scratch_leader)
{
    int64_t v, o;
    kao_ref_gen(pb, bits, leader, seed, round, idx, round_size, scratch_bits, scratch_leader);
    kao_ref_eval(pb, scratch_bits, scratch_leader, &v, &o);
    return kao_ref_pack(v, o, idx);
}

/* ---------------------------------------------------------------- search (docs/MODEL.md §6) */
/* rounds x round_size candidates; per round the minimum key wins and becomes the next base.
 * bits/leader: in = base, out = final base.  round_keys (optional) receives each round's key.
 * Returns the key of the final base's winning candidate (idx field = winner's index). */
uint64_t kao_ref_search(const ref_problem *pb, uint32_t *bits, uint8_t *leader, uint64_t seed,
                        uint32_t first_round, uint32_t rounds, uint32_t round_size,
                        uint64_t *round_keys, int nthreads)
{
    const int W = W_of(pb);
    const size_t nb = (size_t)pb->P * W;
    uint64_t last = ~0ull;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (uint32_t t = first_round; t < first_round + rounds; ++t) {
        uint64_t best = ~0ull;
#pragma omp parallel
        {
            uint32_t *sb = (uint32_t *)malloc(nb * 4);
            uint8_t *sl = (uint8_t *)malloc((size_t)pb->P);
            uint64_t mine = ~0ull;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < (int64_t)round_size; ++i) {
                uint64_t k = kao_ref_candidate_key(pb, bits, leader, seed, t, (uint32_t)i,
                                                   round_size, sb, sl);
                if (k < mine) mine = k;
            }
#pragma omp critical
            if (mine < best) best = mine;
            free(sb); free(sl);
        }
        uint32_t widx = (uint32_t)(best & ((1u << IDX_BITS) - 1));
        kao_ref_gen(pb, bits, leader, seed, t, widx, round_size, bits, leader);
        if (round_keys) round_keys[t - first_round] = best;
        last = best;
    }
    return last;
}

/* bit-plane + leader -> replica lists, leader first then followers ascending (README.md:65-78,:88) */
void kao_ref_decode(const ref_problem *pb, const uint32_t *bits, const uint8_t *leader,
                    int32_t *replicas /* [P*RF], -1 padded */)
{
    const int W = W_of(pb);
    for (int p = 0; p < pb->P; ++p) {
        const uint32_t *row = bits + (size_t)p * W;
        int32_t *out = replicas + (size_t)p * pb->RF;
        int n = 0, ld = leader[p];
        for (int i = 0; i < pb->RF; ++i) out[i] = -1;
        if (ld < pb->B && row_has(row, ld)) out[n++] = ld;
        for (int b = 0; b < pb->B && n < pb->RF; ++b)
            if (row_has(row, b) && b != ld) out[n++] = b;
    }
}

int kao_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
