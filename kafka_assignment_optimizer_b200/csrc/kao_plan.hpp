// kao_plan.hpp — shared-memory plan of the search kernels: which table sits where in the dynamic
// shared memory of a CTA.  Host-only arithmetic, shared by the kernels (kao_kernels.cuh), the engine
// (kao_engine.cu: does the mask-plane layout fit, else packed entries) and the host-side emulation
// harness of the device functions (tests/emu).
#pragma once
#include "kao_device.cuh"

#include <cstdint>

using namespace kao;

// ------------------------------------------------------------------------------------------
// shared-memory plan of the search kernel
// ------------------------------------------------------------------------------------------
struct SmemPlan {
    uint32_t off_bits, off_sw, off_leader, off_consts, off_prow, off_red, off_bar, off_lists, off_totals, off_inv, total;
    uint32_t cap_hold, cap_led;      // delta mode: capacity of the inverted lists (0 = not staged)
};
inline SmemPlan make_plan(int W, int Ppad, int warps, int obj_words_per_row, int P, int RF, bool oh_plane)
{
    SmemPlan s;
    uint32_t o = 0;
    s.off_bits = o;   o += (uint32_t)W * Ppad * 4;
    if (oh_plane) o += (uint32_t)W * Ppad * 4;            // leader one-hot plane, directly behind the bit-plane
    s.off_sw = o;     o += (uint32_t)obj_words_per_row * Ppad * 4;
    s.off_leader = o; o += (uint32_t)Ppad;
    s.off_consts = o; o += (uint32_t)sizeof(Consts);
    s.off_prow = o;   o += (uint32_t)warps * kMaxOps * W * 4;
    o = (o + 15u) & ~15u;
    s.off_red = o;    o += (uint32_t)(warps + 4) * 8;      // + early-stop state behind the per-warp minima
    s.off_bar = o;    o += 16;
    s.off_lists = o;  o += (uint32_t)Ppad * 4 + 16 + 2 * 36 * 4;   // D, DL (u16 each), counts, scan scratch
    s.off_totals = o; o += (256 + 256 + 32 + 4 + 256 + 2 * 66 + 4) * 4;   // delta mode: cnt, lcnt, rc, base (viol, obj),
                                                                           // led counts, list offsets, flag
    s.off_inv = o;
    s.cap_hold = s.cap_led = 0;
    if (W <= 2) {   // inverted lists (u16 partitions) if they fit next to everything else
        const uint32_t need = ((uint32_t)P * RF + 66 + (uint32_t)P + 2 + 8) * 2 + 2 * 512 * 4 + 16;   // + segment counts
        if (o + need <= 227u * 1024u) {
            s.cap_hold = ((uint32_t)P * RF + 64 + 1) & ~1u;     // even counts keep the int scratch behind them aligned
            s.cap_led = ((uint32_t)P + 1) & ~1u;
            o += (need + 15u) & ~15u;
        }
    }
    s.total = o;
    return s;
}
