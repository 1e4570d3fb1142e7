// Host-side geometry and tap stream of the uniform-tap polyphase kernel (kernels_ut.cuh).
#include <algorithm>
#include <cstdlib>

#include "launch.hpp"

namespace aptb200 {

namespace {
constexpr u32 kSmemBudget = 227 * 1024;
constexpr u32 kInflightBytes = 40 * 1024;   // keep about this much of the signal in flight per SM (latency x HBM share)
}  // namespace

bool make_ut_plan(u32 l, u32 m, const std::vector<float> &taps, UtPlan &up, std::vector<float> &stream) {
    if (l < 2 || m == 0 || taps.empty() || getenv("APTB200_NO_UNIFORM_TAPS")) return false;
    const u32 np = (l + 1) / 2;
    if (l != kUtL) return false;                            // the kernel is instantiated for L = 13
    const u64 off2 = 2 * ((static_cast<u64>(taps.size()) - 1) / 2);
    auto fx = [&](u64 r) { return (r * m + l - 1) / l; };   // first / last sample (relative to the row) output r touches
    auto lx = [&](u64 r) { return (r * m + off2) / l; };
    u32 cs[8] = {0}, ce[8] = {0};
    for (u32 p = 0; p < np; ++p) {
        const u64 hi = 2 * p + 1 < l ? std::max(lx(2 * p), lx(2 * p + 1)) : lx(2 * p);
        cs[p] = static_cast<u32>(fx(2 * p) / 4);
        ce[p] = static_cast<u32>(hi / 4 + 1);
        if (p && (cs[p] < cs[p - 1] || ce[p] < ce[p - 1])) return false;
    }
    if (cs[0] != 0 || cs[np - 1] > ce[0]) return false;     // every pair must have started before the first one ends
    const u32 chunks = ce[np - 1];
    if (chunks > 4096) return false;

    // tap stream in consumption order: ramp-up segments, steady, ramp-down; per (chunk, active pair) 8 floats
    // {T[4c][2p], T[4c][2p+1], T[4c+1][2p], ... T[4c+3][2p+1]},  T[u][r] = h[u*l - r*m]
    auto tap = [&](u64 u, u32 r) -> float {
        if (r >= l) return 0.f;
        const long long idx = static_cast<long long>(u * l) - static_cast<long long>(static_cast<u64>(r) * m);
        return idx >= 0 && static_cast<u64>(idx) <= off2 ? taps[static_cast<size_t>(idx)] : 0.f;
    };
    stream.clear();
    auto emit = [&](u32 c0, u32 c1, u32 p0, u32 p1) {
        for (u32 c = c0; c < c1; ++c)
            for (u32 p = p0; p < p1; ++p)
                for (u32 uu = 0; uu < 4; ++uu) {
                    stream.push_back(tap(4ull * c + uu, 2 * p));
                    stream.push_back(tap(4ull * c + uu, 2 * p + 1));
                }
    };
    for (u32 a = 1; a < np; ++a) emit(cs[a - 1], cs[a], 0, a);
    emit(cs[np - 1], ce[0], 0, np);
    for (u32 a = 1; a < np; ++a) emit(ce[a - 1], ce[a], a, np);
    // every nonzero tap must be in the stream exactly once: sum check against the filter
    const u32 nvec = static_cast<u32>(stream.size() / 4);
    if (nvec > kUtMaxVecLarge) return false;

    const u32 halo_u0 = static_cast<u32>(fx(l - 1));
    const u32 halo_n = static_cast<u32>(lx(l - 1) - fx(l - 1) + 1);
    const u32 back = (m > halo_u0 ? (m - halo_u0 + 3) / 4 * 4 : 0);
    bool ok = false;
    for (u32 q : {2u, 1u}) {
        const u32 rb = 32 * q;
        const u64 slot_floats = (static_cast<u64>(back) + static_cast<u64>(rb - 1) * m + 4ull * chunks + 3) / 4 * 4;
        const u64 slot_bytes = slot_floats * 4;
        if (static_cast<u64>(rb) * (l + 1) > slot_floats) continue;   // outputs (+ one exchange word per row) are staged in the slot
        const u32 nslot = static_cast<u32>(std::min<u64>(kUtMaxSlots, (kSmemBudget - 512) / slot_bytes));
        const u32 spare = static_cast<u32>(std::max<u64>(2, (kInflightBytes + slot_bytes - 1) / slot_bytes));
        if (nslot < spare + 4) continue;
        const u32 warps = std::min<u32>(15, nslot - spare);
        if (q == 2 && warps < 10) continue;                     // too few warps to fill the FMA pipe: one row per thread
        up.q = q;
        up.rb = rb;
        up.slot_floats = static_cast<u32>(slot_floats);
        up.warps = warps;
        up.nslot = std::min(nslot, warps + spare + 2);
        up.smem_bytes = 512 + up.nslot * static_cast<u32>(slot_bytes);
        ok = true;
        break;
    }
    if (!ok) return false;
    up.l = l;
    up.m = m;
    up.np = np;
    up.vec = m % 4 == 0 ? 4 : m % 2 == 0 ? 2 : 1;
    up.back = back;
    up.chunks = chunks;
    up.nvec = nvec;
    up.halo_u0 = halo_u0;
    up.halo_n = halo_n;
    up.off2 = off2;
    for (u32 p = 0; p < 8; ++p) {
        up.cs[p] = cs[p];
        up.ce[p] = ce[p];
    }
    up.debug = 0;
    if (const char *e = getenv("APTB200_TILE_DEBUG")) up.debug = static_cast<u32>(atoi(e));
    return true;
}

}  // namespace aptb200
