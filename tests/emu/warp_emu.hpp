// warp_emu.hpp — TEST INFRASTRUCTURE.  Runs the engine's __device__ functions (kao_device.cuh) on
// the host: one warp = 32 fibers that a round-robin scheduler advances from one warp
// collective to the next, so __ballot_sync / __shfl_xor_sync / __match_any_sync / __reduce_add_sync
// see the values of all 32 lanes exactly as the hardware would for converged code.  The integer
// intrinsics are restated with their PTX semantics.  Nothing here is linked into libkao.so; the
// product has no CPU path (tests/test_host.py::test_no_gpu_means_loud_failure).
#pragma once
#include <cuda_runtime.h>   // vector types; under g++ the __device__ / __forceinline__ qualifiers are harmless

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>

namespace emu {

constexpr int kLanes = 32;
constexpr size_t kStackBytes = 256 * 1024;

#if defined(__x86_64__) && !defined(KAO_EMU_UCONTEXT)
// Fiber switch without the signal-mask system call of swapcontext: callee-saved registers + stack pointer.
extern "C" void kao_emu_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".hidden kao_emu_switch\n"
    ".globl kao_emu_switch\n"
    ".type kao_emu_switch,@function\n"
    "kao_emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size kao_emu_switch, .-kao_emu_switch\n");
struct Ctx {
    void *sp = nullptr;
};
inline void ctx_switch(Ctx &from, Ctx &to) { kao_emu_switch(&from.sp, to.sp); }
inline void ctx_make(Ctx &c, char *stack, size_t bytes, void (*entry)(), Ctx &)
{
    void **sp = reinterpret_cast<void **>((reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15));
    *--sp = nullptr;                                  // fake return address: entry() never returns
    *--sp = reinterpret_cast<void *>(entry);
    for (int i = 0; i < 6; ++i) *--sp = nullptr;      // rbp rbx r12 r13 r14 r15
    c.sp = sp;
}
#else
struct Ctx {
    ucontext_t uc;
};
inline void ctx_switch(Ctx &from, Ctx &to) { swapcontext(&from.uc, &to.uc); }
inline void ctx_make(Ctx &c, char *stack, size_t bytes, void (*entry)(), Ctx &back)
{
    getcontext(&c.uc);
    c.uc.uc_stack.ss_sp = stack;
    c.uc.uc_stack.ss_size = bytes;
    c.uc.uc_link = &back.uc;
    makecontext(&c.uc, entry, 0);
}
#endif

struct Warp {
    Ctx main_ctx, ctx[kLanes];
    char *stack[kLanes] = {};
    bool done[kLanes];
    int cur = 0;
    uint64_t val[kLanes];       // what a lane contributes to the collective it waits at
    int op[kLanes];             // which collective that is (divergence check)
    const std::function<void(int)> *body = nullptr;
};
inline Warp &warp()
{
    static Warp w;
    return w;
}

inline void yield_lane(int op)
{
    Warp &w = warp();
    w.op[w.cur] = op;
    ctx_switch(w.ctx[w.cur], w.main_ctx);
}

inline void trampoline()
{
    Warp &w = warp();
    const int lane = w.cur;
    (*w.body)(lane);
    w.done[lane] = true;
    for (;;) ctx_switch(w.ctx[lane], w.main_ctx);     // a finished lane is never resumed
}

// Runs body(lane) for the 32 lanes of one warp to completion.  Every sweep of the scheduler moves
// each live lane from one collective to the next; lanes that wait at a collective must all wait at
// the same kind of collective (a mismatch means the emulated code diverged around a collective).
inline void run_warp(const std::function<void(int)> &body)
{
    Warp &w = warp();
    w.body = &body;
    for (int l = 0; l < kLanes; ++l) {
        if (!w.stack[l]) w.stack[l] = static_cast<char *>(malloc(kStackBytes));
        ctx_make(w.ctx[l], w.stack[l], kStackBytes, trampoline, w.main_ctx);
        w.done[l] = false;
        w.op[l] = 0;
    }
    for (;;) {
        bool any = false;
        int seen = -1;
        for (int l = 0; l < kLanes; ++l) {
            if (w.done[l]) continue;
            w.cur = l;
            ctx_switch(w.main_ctx, w.ctx[l]);
            any = true;
            if (!w.done[l]) {
                if (seen < 0) seen = w.op[l];
                if (seen != w.op[l]) { fprintf(stderr, "warp_emu: lanes wait at different collectives\n"); abort(); }
            }
        }
        if (!any) break;
    }
}

enum { OP_SYNC = 1, OP_BALLOT, OP_SHFL, OP_MATCH, OP_REDUCE };

// publish -> (all lanes arrive) -> combine -> (all lanes have read) -> continue
template <class F> inline uint64_t collective(int op, uint64_t mine, F combine)
{
    Warp &w = warp();
    w.val[w.cur] = mine;
    yield_lane(op);
    const uint64_t r = combine(w.val, w.cur);
    yield_lane(-op);
    return r;
}

}  // namespace emu

// ------------------------------------------------------------------------------------------
// CUDA intrinsics used by kao_device.cuh, restated for the host
// ------------------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xFu;
        if (sel & 8u) { fprintf(stderr, "warp_emu: __byte_perm sign mode not restated\n"); abort(); }
        r |= (unsigned)((src >> (8 * sel)) & 0xFFu) << (8 * i);
    }
    return r;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31u)) >> 32);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31u));
}
template <class T> inline T __ldg(const T *p) { return *p; }

inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::yield_lane(emu::OP_SYNC); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    return (unsigned)emu::collective(emu::OP_BALLOT, pred ? 1 : 0, [](const uint64_t *v, int) {
        uint64_t m = 0;
        for (int l = 0; l < emu::kLanes; ++l) m |= (v[l] & 1u) << l;
        return m;
    });
}
inline unsigned __shfl_xor_sync(unsigned, unsigned x, int lane_mask)
{
    return (unsigned)emu::collective(emu::OP_SHFL, x, [lane_mask](const uint64_t *v, int me) { return v[(me ^ lane_mask) & 31]; });
}
inline unsigned __match_any_sync(unsigned, int x)
{
    return (unsigned)emu::collective(emu::OP_MATCH, (uint32_t)x, [](const uint64_t *v, int me) {
        uint64_t m = 0;
        for (int l = 0; l < emu::kLanes; ++l) m |= (uint64_t)(v[l] == v[me]) << l;
        return m;
    });
}
inline int __reduce_add_sync(unsigned mask, int x)
{
    return (int)(uint32_t)emu::collective(emu::OP_REDUCE, (uint32_t)x, [mask](const uint64_t *v, int) {
        uint32_t s = 0;
        for (int l = 0; l < emu::kLanes; ++l)
            if ((mask >> l) & 1u) s += (uint32_t)v[l];
        return (uint64_t)s;
    });
}

namespace kao {
using std::abs;
using std::max;
using std::min;
}  // namespace kao
