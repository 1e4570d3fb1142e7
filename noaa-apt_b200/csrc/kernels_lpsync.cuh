// Fused demodulation low-pass + sync cross-correlation (dsp::filter dsp.rs:386-410 with the 37..61-tap
// Lowpass of decode.rs:95-102, and the correlation loop of find_sync decode.rs:225-233).
//
//   f[i]    = sum_{j < NT, j < i} e[i-j] * c[j]                 (causal; e[0] is never read: strict i > j)
//   corr[i] = sum_{j < 38*PW} guard[j] * f[i+j],   i < n - 38*PW
//
// The template is +-1 in runs of 2*PW samples (19 runs: - | (-,+) x 7 | - - - -), so the correlation is
// computed from box sums B[n] = f[n] + ... + f[n+2*PW-1]:  corr[i] = sum_b sign_b * B[i + 2*PW*b]  -- 24 adds
// per output instead of 114.  The summation order differs from the reference's sequential loop, i.e. the
// values agree to fp32 rounding (~1e-7 relative), not bit for bit; the generic kernel keeps the exact order.
//
// One CTA handles 1856 consecutive positions: e tile -> f tile (registers: 8 outputs x NT taps per thread, taps
// are kernel parameters = constant-bank operands) -> box sums -> correlation (16 outputs per thread, each box
// value loaded once and applied to the <= 3 outputs it belongs to).  f and corr go to HBM once each; e is read
// once (+7 % halo).  Shared-memory tiles are skewed by 4 floats every 32 so that 8-float-strided LDS.128 is
// conflict-free.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "launch.hpp"

namespace aptb200 {

struct LpTaps {
    float c[64];      // zero-padded; the kernel is instantiated for the exact count
};

constexpr int kLpTile = 1856;   // (T + 38*PW - 1) / 8 <= 256 for PW <= 5: every phase is one pass of the CTA

__device__ __forceinline__ u32 skew(u32 n) { return n + ((n >> 5) << 2); }   // 4 floats of padding every 32

template <int NT, int PW>
__global__ void __launch_bounds__(256)
k_lowpass_corr(const float *__restrict__ e, u64 n, u64 ncorr, LpTaps taps, float *__restrict__ f_out,
               float *__restrict__ corr_out) {
    constexpr int T = kLpTile;
    constexpr int G = 38 * PW;                 // template length
    constexpr int BOX = 2 * PW;                // run length
    constexpr int FN = T + G - 1;              // f values needed by the tile's correlations
    constexpr int FNP = (FN + 7) / 8 * 8;      // computed in groups of 8
    constexpr int EN = FNP + NT - 1;           // e values needed (NT-1 before the first)
    constexpr int EOFF = (NT - 1 + 3) / 4 * 4; // e tile starts this many samples before the tile (multiple of 4)
    constexpr int ENP = FNP + EOFF;
    constexpr int BN = T + 18 * BOX;           // box sums needed
    __shared__ __align__(16) float s_e[(ENP + 32) + (ENP + 32) / 8 + 8];
    __shared__ __align__(16) float s_f[(FNP + BOX + 16) + (FNP + BOX + 16) / 8 + 8];
    __shared__ __align__(16) float s_b[(BN + 16) + (BN + 16) / 8 + 8];

    const u32 tid = threadIdx.x;
    for (u64 tile = blockIdx.x; tile * T < n; tile += gridDim.x) {
        const u64 i0 = tile * T;
        // ---- e tile: s_e[skew(m)] = e[i0 - EOFF + m]; indices < 1 or >= n read as zero (dsp.rs:399: the sum
        //      only takes signal[i-j] with i > j, so signal[0] and anything before it never contribute) ----
        for (u32 m4 = tid * 4; m4 < ENP; m4 += 256 * 4) {
            const long long g = static_cast<long long>(i0) - EOFF + m4;
            float4 v;
            if (g >= 1 && static_cast<u64>(g) + 3 < n) {
                v = __ldg(reinterpret_cast<const float4 *>(e + g));
            } else {
                v.x = g >= 1 && static_cast<u64>(g) < n ? __ldg(e + g) : 0.f;
                v.y = g + 1 >= 1 && static_cast<u64>(g + 1) < n ? __ldg(e + g + 1) : 0.f;
                v.z = g + 2 >= 1 && static_cast<u64>(g + 2) < n ? __ldg(e + g + 2) : 0.f;
                v.w = g + 3 >= 1 && static_cast<u64>(g + 3) < n ? __ldg(e + g + 3) : 0.f;
            }
            *reinterpret_cast<float4 *>(s_e + skew(m4)) = v;
        }
        __syncthreads();
        // ---- low-pass: 8 outputs per thread-item, window of 8 + NT - 1 samples in registers ----
        for (u32 item = tid; item < FNP / 8; item += 256) {
            const u32 o = item * 8;                       // first output (relative to i0)
            constexpr int WN = (8 + NT - 1 + 3 + 3) / 4 * 4;   // window registers (aligned start)
            const u32 wstart = o + EOFF - (NT - 1);       // tile index of the oldest sample needed
            const u32 wal = wstart & ~3u;                 // aligned down
            float w[WN];
#pragma unroll
            for (int k = 0; k < WN / 4; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(s_e + skew(wal + 4 * k));
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
            constexpr int SH = (EOFF - (NT - 1)) & 3;     // o is a multiple of 8, EOFF of 4: the shift is static
            float acc[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] = 0.f;
#pragma unroll
            for (int j = 0; j < NT; ++j)                  // ascending j like the reference (sum += x[i-j]*c[j])
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[v] = fmaf(w[SH + v + NT - 1 - j], taps.c[j], acc[v]);
            *reinterpret_cast<float4 *>(s_f + skew(o)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4 *>(s_f + skew(o + 4)) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            if (o < T) {                                  // the tile's own f values go to HBM (f_out is 16-byte aligned)
                const u64 gi = i0 + o;
                if (gi + 7 < n) {
                    *reinterpret_cast<float4 *>(f_out + gi) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4 *>(f_out + gi + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                } else {
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        if (gi + v < n) f_out[gi + v] = acc[v];
                }
            }
        }
        if (corr_out == nullptr) {
            __syncthreads();
            continue;
        }
        if (tid < BOX) s_f[skew(FNP + tid)] = 0.f;        // tail read by the last box sums, never used
        __syncthreads();
        // ---- box sums B[m] = f[m] + ... + f[m+BOX-1], 8 per thread-item ----
        for (u32 item = tid; item < (BN + 7) / 8; item += 256) {
            const u32 o = item * 8;
            constexpr int WB = (8 + BOX - 1 + 3) / 4 * 4;
            float w[WB];
#pragma unroll
            for (int k = 0; k < WB / 4; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(s_f + skew(o + 4 * k));
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
            float b[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float s = w[v];
#pragma unroll
                for (int t = 1; t < BOX; ++t) s += w[v + t];
                b[v] = s;
            }
            *reinterpret_cast<float4 *>(s_b + skew(o)) = make_float4(b[0], b[1], b[2], b[3]);
            *reinterpret_cast<float4 *>(s_b + skew(o + 4)) = make_float4(b[4], b[5], b[6], b[7]);
        }
        __syncthreads();
        // ---- correlation: 16 outputs per thread-item; box value m belongs to outputs m - BOX*b ----
        for (u32 item = tid; item < T / 16; item += 256) {
            const u32 o = item * 16;
            if (i0 + o >= ncorr) break;
            constexpr int NB = 16 + 18 * BOX;             // box values this item needs
            constexpr int NBP = (NB + 3) / 4 * 4;
            float acc[16];
#pragma unroll
            for (int v = 0; v < 16; ++v) acc[v] = 0.f;
#pragma unroll
            for (int k = 0; k < NBP / 4; ++k) {
                const float4 q = *reinterpret_cast<const float4 *>(s_b + skew(o + 4 * k));
                const float val[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int m = 4 * k + c;              // box index relative to o (compile time)
#pragma unroll
                    for (int b = 0; b < 19; ++b) {
                        const int v = m - BOX * b;        // output it contributes to with run b
                        if (v >= 0 && v < 16) {
                            // runs: b = 0 '-', then (-,+) x 7 for b = 1..14, then '-' x 4 (decode.rs:188-198)
                            const bool plus = b >= 1 && b <= 14 && (b % 2 == 0);
                            acc[v] = plus ? acc[v] + val[c] : acc[v] - val[c];
                        }
                    }
                }
            }
            const u64 gi = i0 + o;
            if (gi + 15 < ncorr) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<float4 *>(corr_out + gi + 4 * k) = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
            } else {
#pragma unroll
                for (int v = 0; v < 16; ++v)
                    if (gi + v < ncorr) corr_out[gi + v] = acc[v];
            }
        }
        __syncthreads();
    }
}

}  // namespace aptb200
