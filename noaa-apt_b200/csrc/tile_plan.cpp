// Host-side geometry of the tiled polyphase kernel (kernels_fast.cuh).
#include <algorithm>
#include <numeric>

#include "launch.hpp"

namespace aptb200 {

namespace {
constexpr u32 kR = 8, kQ = 4, kKS = 4, kQT = (32 / kKS) * kQ;
constexpr u32 kMaxGroups = 13;              // warps per CTA the kernel is compiled for (416 threads, 2 CTAs/SM)
constexpr u32 kSmemTwoCtas = 113 * 1024;    // per-CTA budget that still lets two CTAs share an SM
}  // namespace

bool make_tile_plan(u32 l, u32 m, const std::vector<float> &taps, TilePlan &tp, std::vector<float> &tile_taps,
                    std::vector<u32> &group_xs) {
    if (l < 2 || m == 0 || taps.empty()) return false;
    const u64 off2 = 2 * ((static_cast<u64>(taps.size()) - 1) / 2);
    const u32 groups = l / std::gcd(kR, l);
    if (groups > kMaxGroups) return false;
    const u64 p_out = static_cast<u64>(kR) * groups;
    const u64 p_in = p_out * m / l;          // exact: p_out is a multiple of l
    if (p_in % 4 != 0 || p_in > (1u << 20)) return false;

    group_xs.assign(groups, 0);
    u64 need = 0, max_xs = 0;
    for (u32 g = 0; g < groups; ++g) {
        const u64 k0 = static_cast<u64>(kR) * g;
        const u64 xs = (k0 * m + l - 1) / l;            // first sample output k0 touches
        const u64 xs4 = xs & ~static_cast<u64>(3);      // 16-byte aligned window start
        const u64 xmax = ((k0 + kR - 1) * m + off2) / l;   // last sample output k0+R-1 touches
        group_xs[g] = static_cast<u32>(xs4);
        need = std::max(need, xmax - xs4 + 1);
        max_xs = std::max(max_xs, xs4);
    }
    const u64 usteps = (need + 4 * kKS - 1) / (4 * kKS) * (4 * kKS);
    u64 row_len = (max_xs + usteps + 3) / 4 * 4;
    row_len = std::max(row_len, (p_in + 3) / 4 * 4);    // the halo output and edge fills index whole periods
    if ((row_len / 4) % 2 == 0) row_len += 4;           // odd pitch in 16-byte units: conflict-free LDS.128
    const u64 smem = 16 + groups * usteps * kR * 4 + static_cast<u64>(kQT) * row_len * 4;
    if (smem > kSmemTwoCtas) return false;
    if (static_cast<u64>(kQT) * p_out + 1 > static_cast<u64>(kQT) * row_len) return false;   // parked tile must fit

    tile_taps.assign(groups * usteps * kR, 0.f);
    for (u32 g = 0; g < groups; ++g) {
        for (u64 u = 0; u < usteps; ++u) {
            for (u32 r = 0; r < kR; ++r) {
                const long long idx = static_cast<long long>((group_xs[g] + u) * l) -
                                      static_cast<long long>((static_cast<u64>(kR) * g + r) * m);
                if (idx >= 0 && static_cast<u64>(idx) <= off2)
                    tile_taps[(static_cast<size_t>(g) * usteps + u) * kR + r] = taps[static_cast<size_t>(idx)];
            }
        }
    }
    tp.l = l;
    tp.m = m;
    tp.groups = groups;
    tp.p_out = static_cast<u32>(p_out);
    tp.p_in = static_cast<u32>(p_in);
    tp.usteps = static_cast<u32>(usteps);
    tp.row_len = static_cast<u32>(row_len);
    tp.qt = kQT;
    tp.smem_bytes = static_cast<u32>(smem);
    tp.off2 = off2;
    return true;
}

}  // namespace aptb200
