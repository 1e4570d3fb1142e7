// Polyphase resampler + envelope for LARGE interpolation factors (11025 / 22050 / 44100 Hz -> 12 480 Hz: L = 832 / 416 /
// 208, M = 735; fast_resampling dsp.rs:186-289 + demodulate dsp.rs:350-383), where the 14 057-tap filter gives every
// output only J = 17 / 34 / 68 taps but each of the L outputs of a period has its own tap set.
//
// Formulation (the same rows-by-phases view as kernels_ut.cuh, just with many phases and few taps):
//     y[L*q + r] = sum_{j < J} T[r][j] * X[M*q + xs[r] + j],   xs[r] = ceil(r*M / L),  T[r][j] = h[xs[r]*L - r*M + L*j]
// A warp's 32 lanes are 32 consecutive periods q, so the phase r -- and with it the tap set, the window start xs[r] and
// its 16-byte misalignment xs[r] & 3 -- is WARP-UNIFORM: taps are broadcast LDS.128 from the phase table in shared
// memory, every period's input span sits in its own 16-byte aligned shared-memory row (pitch = 4 mod 32 floats: the 32
// lanes' LDS.128 are conflict-free), and a warp-uniform switch on xs[r] & 3 picks one of four code variants with
// compile-time register offsets.  Per output: (JPAD + 4) / 4 window + JPAD / 4 tap LDS.128, J FFMA in the reference's
// ascending order.  PCM16 input is converted while the rows are staged (wav.rs:37 fused into the load).
//
// A warp owns a range of phases for the CTA's 32 periods; consecutive phases of a lane are consecutive outputs, so the
// envelope's predecessor is the previous loop iteration (one extra output per warp for the range start).  Outputs are
// transposed through a small per-warp stage so that every store is 32 consecutive floats of one period.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_generic.cuh"
#include "launch.hpp"

namespace aptb200 {

constexpr int kPhPeriods = static_cast<int>(kPhTilePeriods);   // periods per tile = lanes
constexpr int kPhWarps = 8;

struct PhGeom {
    u32 l, m;
    u32 j;             // taps per output
    u32 pitch;         // floats per input row (16-byte aligned rows; pitch % 32 == 4)
    u32 row_len;       // staged samples per row: m + jpad + 8 rounded up to 4
    u32 smem_bytes;
};

__device__ __forceinline__ void ph_cp_async4(float *dst, const float *src, bool valid) {
    const u32 bytes = valid ? 4u : 0u;            // src-size 0: nothing is read, the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(static_cast<u32>(__cvta_generic_to_shared(dst))), "l"(src),
                 "r"(bytes) : "memory");
}

// rows of a tile: f32 samples by 4-byte cp.async (a row starts at an arbitrary sample; all copies of a tile are in flight
// together), PCM16 samples through registers, eight loads at a time, converted on the way (wav.rs:37)
__device__ __forceinline__ void ph_stage_row(const float *signal, u64 len, long long gx0, u32 row_len, float *row, u32 lane) {
    if (gx0 >= 0 && static_cast<u64>(gx0) + row_len <= len) {
        const float *src = signal + gx0;
        for (u32 i = lane; i < row_len; i += 32) ph_cp_async4(row + i, src + i, true);
    } else {
        for (u32 i = lane; i < row_len; i += 32) {
            const long long gx = gx0 + i;
            const bool ok = gx >= 0 && static_cast<u64>(gx) < len;
            ph_cp_async4(row + i, ok ? signal + gx : signal, ok);
        }
    }
}
__device__ __forceinline__ void ph_stage_row(const int16_t *signal, u64 len, long long gx0, u32 row_len, float *row, u32 lane) {
    for (u32 i0 = lane; i0 < row_len; i0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long gx = gx0 + i0 + 32 * u;
            v[u] = (i0 + 32 * u < row_len && gx >= 0 && static_cast<u64>(gx) < len) ? load_sample(signal, static_cast<u64>(gx)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + 32 * u < row_len) row[i0 + 32 * u] = v[u];
    }
}

// one output of phase r for this lane's period: `row` points at the sample X[M*q - 4]
template <int JPAD, int SH>
__device__ __forceinline__ float ph_dot(const float *row, u32 x0a, const float4 *tp) {
    float w[JPAD + 4];
#pragma unroll
    for (int k = 0; k < (JPAD + 4) / 4; ++k) {
        const float4 q = *reinterpret_cast<const float4 *>(row + x0a + 4 * k);
        w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < JPAD / 4; ++k) {
        const float4 t = tp[k];                          // warp-uniform address: one broadcast wavefront
        acc = fmaf(t.x, w[SH + 4 * k], acc);
        acc = fmaf(t.y, w[SH + 4 * k + 1], acc);
        acc = fmaf(t.z, w[SH + 4 * k + 2], acc);
        acc = fmaf(t.w, w[SH + 4 * k + 3], acc);
    }
    return acc;
}

template <int JPAD>
__device__ __forceinline__ float ph_output(const float *row, u32 xs, const float *table, u32 r) {
    const u32 x = xs + 4;                                // row[0] is X[M*q - 4]
    const float4 *tp = reinterpret_cast<const float4 *>(table + r * JPAD);
    switch (x & 3) {                                     // warp-uniform
    case 0: return ph_dot<JPAD, 0>(row, x & ~3u, tp);
    case 1: return ph_dot<JPAD, 1>(row, x & ~3u, tp);
    case 2: return ph_dot<JPAD, 2>(row, x & ~3u, tp);
    default: return ph_dot<JPAD, 3>(row, x & ~3u, tp);
    }
}

template <typename InT, int JPAD>
__global__ void __launch_bounds__(32 * kPhWarps, 1)
k_polyphase_ph(const InT *__restrict__ signal, u64 len, const float *__restrict__ table_g, const unsigned short *__restrict__ xs_g,
               const PhGeom g, u64 nout, u64 tile_begin, u64 tile_end, int envelope, float cosphi2, float inv_sinphi,
               float *__restrict__ out) {
    extern __shared__ __align__(16) float ph_smem[];
    float *s_table = ph_smem;                                        // [l][JPAD]
    float *s_rows = s_table + static_cast<size_t>(g.l) * JPAD;       // [33][pitch]: row 0 = the period before the tile
    float *s_stage = s_rows + static_cast<size_t>(kPhPeriods + 1) * g.pitch;   // [warps][32][33]
    unsigned short *s_xs = reinterpret_cast<unsigned short *>(s_stage + kPhWarps * 32 * 33);   // [l]

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (u32 i = tid; i < g.l * JPAD / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(s_table)[i] = __ldg(reinterpret_cast<const float4 *>(table_g) + i);
    for (u32 i = tid; i < g.l; i += blockDim.x) s_xs[i] = xs_g[i];

    // phases of this warp
    const u32 per = (g.l + kPhWarps - 1) / kPhWarps;
    const u32 ra = min(warp * per, g.l), rb = min(ra + per, g.l);
    float *stage = s_stage + warp * (32 * 33);

    for (u64 tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const u64 q0 = tile * kPhPeriods;
        __syncthreads();                                             // the rows of the previous tile are no longer read
        // ---- stage the 33 rows: row p holds X[M*(q0 + p - 1) - 4 + i], zero outside [0, len) ----
        for (u32 p = warp; p <= kPhPeriods; p += kPhWarps) {
            const long long gx0 = static_cast<long long>(q0 + p) * g.m - static_cast<long long>(g.m) - 4;
            ph_stage_row(signal, len, gx0, g.row_len, s_rows + p * g.pitch, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (ra >= rb) continue;
        const float *row = s_rows + (lane + 1) * g.pitch;            // this lane's period q0 + lane
        const u64 kq = (q0 + lane) * g.l;                            // first output of the period
        float prev = 0.f;
        if (envelope) {
            // r[k-1] of the range's first output: phase ra-1 of the same period, or the last phase of the period before
            prev = ra > 0 ? ph_output<JPAD>(row, s_xs[ra - 1], s_table, ra - 1)
                          : ph_output<JPAD>(row - g.pitch, s_xs[g.l - 1], s_table, g.l - 1);
        }
        for (u32 r0 = ra; r0 < rb; r0 += 32) {
            const u32 nr = min(32u, rb - r0);
            u32 s = 0;
            for (; s + 1 < nr; s += 2) {                             // two phases per iteration: two independent FMA chains
                const u32 r = r0 + s;
                const float v0 = ph_output<JPAD>(row, s_xs[r], s_table, r);
                const float v1 = ph_output<JPAD>(row, s_xs[r + 1], s_table, r + 1);
                float o0 = v0, o1 = v1;
                if (envelope) {
                    o0 = (kq + r == 0) ? 0.f : envelope2_fast(prev, v0, cosphi2, inv_sinphi);   // dsp.rs:364: e[0] = 0
                    o1 = envelope2_fast(v0, v1, cosphi2, inv_sinphi);
                    prev = v1;
                }
                stage[lane * 33 + s] = o0;
                stage[lane * 33 + s + 1] = o1;
            }
            if (s < nr) {
                const u32 r = r0 + s;
                const float v = ph_output<JPAD>(row, s_xs[r], s_table, r);
                float o = v;
                if (envelope) {
                    o = (kq + r == 0) ? 0.f : envelope2_fast(prev, v, cosphi2, inv_sinphi);
                    prev = v;
                }
                stage[lane * 33 + s] = o;
            }
            __syncwarp();
            // 32 consecutive phases of one period per store
            for (u32 p = 0; p < kPhPeriods; ++p) {
                const u64 k = (q0 + p) * g.l + r0 + lane;
                if (lane < nr && k < nout) out[k] = stage[p * 33 + lane];
            }
            __syncwarp();
        }
    }
}

}  // namespace aptb200
