// Tiled polyphase resampler fused with the envelope demodulator -- the hot kernel of the path
// (fast_resampling dsp.rs:186-289 + demodulate dsp.rs:350-383), written for sm_100a.
//
// Formulation.  y[k] = sum_x h[x*L - k*M] * X[x].  Outputs k and k + P_out (P_out = lcm(R, L)) use the
// same taps on inputs shifted by P_in = P_out*M/L, so the work is a small dense product per "group"
//     acc[r][q] += T_g[u][r] * X[tile_x0 + q*P_in + xs_g + u]        r < R, q < QT, u < U
// where group g covers the R consecutive outputs R*g .. R*g+R-1 of a super-period, xs_g is the first
// input sample any of them touches (rounded down to a multiple of 4 for 16-byte loads) and T_g is the
// zero-padded slice of h those outputs see.  G = L/gcd(R, L) groups cover every phase.
//
// Mapping.  One persistent CTA per SM slot (2 per SM), one warp per group.  A warp's 32 lanes are
// KS=4 slices of the u range x 8 row lanes; each thread owns R=8 outputs x Q=4 rows (32 accumulators):
//   - samples: rows of the input tile in shared memory, one 16-byte LDS.128 per row per 4 taps; the
//     8 row lanes of a quarter-warp hit 8 distinct 16-byte bank groups because the row pitch/4 is odd;
//   - taps: 8 per u step by two warp-broadcast LDS.128 (4 distinct addresses per warp);
//   - 32 FFMA per u step per thread; fp32 accumulation in ascending u inside a slice, the 4 slices are
//     then combined by a shuffle reduce-scatter (24 SHFL per thread per tile).
// Input rows and the tap table are staged by 1-D TMA bulk copies (cp.async.bulk, one row per lane,
// completion on an mbarrier); the CTA's other resident CTA on the SM computes meanwhile.
// Epilogue: the tile of resampled values is parked in the (now free) row buffer, then every thread
// turns (r[k-1], r[k]) into the envelope and stores it coalesced -- r itself never reaches HBM.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_generic.cuh"
#include "launch.hpp"

namespace aptb200 {

// ---- mbarrier / TMA bulk-copy wrappers (PTX ISA: cp.async.bulk, mbarrier) -------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(void *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void *bar, u32 parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// global -> shared bulk copy, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, u32 bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kTileR = 8, kTileQ = 4, kTileKS = 4;
constexpr int kTileRowLanes = 32 / kTileKS;           // 8
constexpr int kTileQT = kTileRowLanes * kTileQ;       // 32 rows per tile

// group_xs[g] = xs'_g, the (4-aligned) first input sample of group g relative to its row.
// f32 instantiation of the tile kernel.  (PCM16 input goes through the generic kernel for now.)
template <bool ENVELOPE>
__global__ void __launch_bounds__(32 * 13, 2)
k_polyphase_tiled_f32(const float *__restrict__ signal, u64 len, const float *__restrict__ raw_taps,
                      const float *__restrict__ tile_taps, const u32 *__restrict__ group_xs, TilePlan tp, u64 nout,
                      u64 ntiles, float cosphi2, float sinphi, float *__restrict__ out) {
    constexpr int R = kTileR, Q = kTileQ, KS = kTileKS, QT = kTileQT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [mbarrier 16 B][taps G*U*R][rows QT*row_len]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *s_taps = reinterpret_cast<float *>(smem_raw + 16);
    const u32 taps_floats = tp.groups * tp.usteps * R;
    float *s_rows = s_taps + taps_floats;
    __shared__ float s_halo;                      // r[K0 - 1]

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ks = lane >> 3, ql = lane & 7;
    const u32 nthreads = blockDim.x;
    const u32 tile_out = QT * tp.p_out;

    u32 phase = 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // tap table: one bulk copy for the whole CTA lifetime
    if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, taps_floats * 4);
        // bulk copies are limited in size only by the tx-count field (2^20-1 bytes); split to be safe
        u32 done = 0;
        const u32 total = taps_floats * 4;
        while (done < total) {
            const u32 chunk = min(total - done, 32768u);
            tma_bulk_g2s(reinterpret_cast<unsigned char *>(s_taps) + done,
                         reinterpret_cast<const unsigned char *>(tile_taps) + done, chunk, bar);
            done += chunk;
        }
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    const u32 xs_g = warp < tp.groups ? group_xs[warp] : 0;
    const u32 ul = tp.usteps / KS;                // u steps per slice (multiple of 4)
    const bool aligned16 = (reinterpret_cast<uintptr_t>(signal) & 15) == 0 && (tp.p_in & 3) == 0;

    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 k_base = tile * tile_out;                       // first output of the tile
        const u64 x_base = tile * static_cast<u64>(QT) * tp.p_in; // first input sample of row 0
        // ---- stage the input rows ----
        const u64 x_last = x_base + static_cast<u64>(QT - 1) * tp.p_in + tp.row_len;   // one past the last sample needed
        const bool interior = aligned16 && x_last <= len;
        if (interior) {
            if (warp == 0) {
                fence_proxy_async();              // rows were last touched through the generic proxy
                if (lane == 0) mbar_expect_tx(bar, QT * tp.row_len * 4);
                __syncwarp();
                tma_bulk_g2s(s_rows + lane * tp.row_len, signal + x_base + static_cast<u64>(lane) * tp.p_in,
                             tp.row_len * 4, bar);
            }
        } else {
            for (u32 i = tid; i < QT * tp.row_len; i += nthreads) {
                const u32 q = i / tp.row_len, c = i - q * tp.row_len;
                const u64 x = x_base + static_cast<u64>(q) * tp.p_in + c;
                s_rows[i] = x < len ? __ldg(signal + x) : 0.f;    // past the end: signal.get(x) == None
            }
        }
        // ---- r[K0 - 1] for the envelope of the tile's first output (raw taps, straight from global) ----
        if (ENVELOPE && warp == tp.groups - 1) {
            float part = 0.f;
            if (k_base > 0) {
                const u64 k = k_base - 1;
                const u64 t0 = k * tp.m;
                u64 x = (t0 + tp.l - 1) / tp.l;
                u64 xe = (t0 + tp.off2) / tp.l;
                if (xe >= len) xe = len - 1;
                for (u64 xi = x + lane; xi <= xe; xi += 32)
                    part = fmaf(__ldg(raw_taps + (xi * tp.l - t0)), __ldg(signal + xi), part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_halo = part;
        }
        if (interior) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            __syncthreads();
        }

        // ---- the product: R x Q accumulators per thread over this lane's slice of u ----
        float acc[R][Q];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < Q; ++j) acc[r][j] = 0.f;
        if (warp < tp.groups) {
            const u32 u0 = ks * ul;
            const float4 *tap4 = reinterpret_cast<const float4 *>(s_taps + (static_cast<size_t>(warp) * tp.usteps + u0) * R);
            const float *row0 = s_rows + xs_g + u0;
            const float4 *rp[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) rp[j] = reinterpret_cast<const float4 *>(row0 + (ql + 8 * j) * tp.row_len);
            for (u32 c = 0; c < ul / 4; ++c) {
                float4 s[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j) s[j] = rp[j][c];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 ta = tap4[(c * 4 + i) * 2], tb = tap4[(c * 4 + i) * 2 + 1];
                    const float t[R] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        const float sv = i == 0 ? s[j].x : i == 1 ? s[j].y : i == 2 ? s[j].z : s[j].w;
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r][j] = fmaf(t[r], sv, acc[r][j]);
                    }
                }
            }
        }
        // ---- combine the KS slices: reduce-scatter over lanes ks (xor 16, xor 8) ----
        // flat index i = r*Q + j; after both rounds lane ks owns flat indices [ks*8, ks*8+8)
        float h1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float lo = acc[i / Q][i % Q], hi = acc[(i + 16) / Q][(i + 16) % Q];
            const float send = (ks & 2) ? lo : hi;
            const float keep = (ks & 2) ? hi : lo;
            h1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        float h2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = (ks & 1) ? h1[i] : h1[i + 8];
            const float keep = (ks & 1) ? h1[i + 8] : h1[i];
            h2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        __syncthreads();   // every warp is done reading the rows: park the resampled tile there
        float *s_r = s_rows;                       // s_r[1 + k_local], s_r[0] = halo
        if (warp < tp.groups) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u32 flat = ks * 8 + i;       // = r*Q + j
                const u32 r = flat / Q, j = flat % Q;
                const u32 q = ql + 8 * j;
                s_r[1 + q * tp.p_out + warp * R + r] = h2[i];
            }
        }
        if (tid == 0) s_r[0] = ENVELOPE ? s_halo : 0.f;
        __syncthreads();
        // ---- epilogue: envelope (dsp.rs:373) and coalesced store ----
        for (u32 kl = tid; kl < tile_out; kl += nthreads) {
            const u64 k = k_base + kl;
            if (k >= nout) break;
            float v;
            if (ENVELOPE) v = k == 0 ? 0.f : envelope2(s_r[kl], s_r[kl + 1], cosphi2, sinphi);
            else v = s_r[kl + 1];
            out[k] = v;
        }
        __syncthreads();   // rows buffer is free for the next tile's bulk copies
    }
}

}  // namespace aptb200
