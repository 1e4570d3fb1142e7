// Tiled polyphase resampler fused with the envelope demodulator -- the hot kernel of the path
// (fast_resampling dsp.rs:186-289 + demodulate dsp.rs:350-383), written for sm_100a.
//
// Formulation.  y[k] = sum_x h[x*L - k*M] * X[x].  Outputs k and k + P_out (P_out = lcm(8, L)) use the
// same taps on inputs shifted by P_in = P_out*M/L, so per "group" g (outputs 8g..8g+7 of a super-period)
//     acc[r][q] += T_g[u][r] * X[tile_x0 + q*P_in + w0_g + u]        r < 8, q < QT
// with w0_g the first input sample the group touches (rounded down to a multiple of 4 for 16-byte loads)
// and T_g the zero-padded slice of h its outputs see.  The group's two halves (outputs 0..3 / 4..7) have
// windows offset by `shift` samples, so the first shift/16 loop iterations skip half B and the last skip A.
//
// Mapping.  Persistent CTAs (2 per SM when shared memory allows), one warp per group.  A warp's 32 lanes
// are KS=4 interleaved slices of the sample axis x 8 row lanes; a thread owns 8 outputs x 4 rows
// (32 accumulators) and, per loop iteration, one 16-byte chunk of each of its rows:
//   - samples: LDS.128 from the row tile; the 8 row lanes of a quarter-warp hit 8 distinct 16-byte bank
//     groups because the row pitch/4 is odd;
//   - taps: 8 LDS.128 per iteration, warp-broadcast with 4 distinct addresses skewed across bank groups;
//   - 128 FFMA per iteration; fp32, ascending sample order within a slice.
// Input rows and the tap table arrive by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an
// mbarrier; the next tile's rows are prefetched into L2 meanwhile.  The four slices' partial sums meet in
// shared memory (the row buffer is free by then), and the epilogue turns (r[k-1], r[k]) into the envelope
// and stores float4 -- the resampled signal itself never reaches HBM.  r[K0-1] comes from one extra
// "virtual" row holding the window of the last group of the previous super-period.
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
// warm L2 with a span of global memory (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2(const void *src, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

constexpr int kTileR = 8, kTileH = 4, kTileQ = 4, kTileKS = 4;
constexpr int kTileRowLanes = 32 / kTileKS;           // 8
constexpr int kTileQT = kTileRowLanes * kTileQ;       // 32 rows per tile

// ---- packed fp32x2 helpers: one FFMA2 issue slot does two FMAs (each half rounded on its own) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float4 lds128(u32 addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// One half (4 outputs, as two packed pairs) x 4 rows x the 4 consecutive samples of one 16-byte chunk.
// acc[p][j] packs outputs (2p, 2p+1) of row j; the tap pair comes straight out of the LDS.128 register quad,
// the sample is duplicated into both lanes of the packed operand.
__device__ __forceinline__ void half_fma2(f32x2 (&acc)[2][kTileQ], u32 tap_addr, const float4 (&s)[kTileQ]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 tp = lds128(tap_addr + 16 * u);  // taps of outputs r = 0..3 for sample u of the chunk
        const f32x2 t01 = pack2(tp.x, tp.y), t23 = pack2(tp.z, tp.w);
#pragma unroll
        for (int j = 0; j < kTileQ; ++j) {
            const float sv = u == 0 ? s[j].x : u == 1 ? s[j].y : u == 2 ? s[j].z : s[j].w;
            const f32x2 sv2 = pack2(sv, sv);
            acc[0][j] = fma2(t01, sv2, acc[0][j]);
            acc[1][j] = fma2(t23, sv2, acc[1][j]);
        }
    }
}

// group_xs[g] = w0_g, the (4-aligned) first input sample of group g relative to its row.
template <bool ENVELOPE>
__global__ void __launch_bounds__(32 * 13, 2)
k_polyphase_tiled_f32(const float *__restrict__ signal, u64 len, const float *__restrict__ tile_taps,
                      const u32 *__restrict__ group_xs, TilePlan tp, u64 nout, u64 ntiles, float cosphi2,
                      float sinphi, float *__restrict__ out) {
    constexpr int H = kTileH, Q = kTileQ, KS = kTileKS, QT = kTileQT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [mbarrier 16 B][taps G*group_stride][rows / partial-sum planes][virtual halo row]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *s_taps = reinterpret_cast<float *>(smem_raw + 16);
    const u32 taps_floats = tp.groups * tp.group_stride;
    float *s_rows = s_taps + taps_floats;
    float *s_vrow = s_rows + tp.rows_floats;
    __shared__ float s_halo;                      // r[K0 - 1]

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ks = lane >> 3, ql = lane & 7;
    const u32 nthreads = blockDim.x;
    const u32 tile_out = QT * tp.p_out;
    const float inv_sinphi = 1.f / sinphi;

    u32 phase = 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // tap table: bulk copies once per CTA lifetime
    if (tid == 0) {
        fence_proxy_async();
        const u32 total = taps_floats * 4;
        mbar_expect_tx(bar, total);
        for (u32 done = 0; done < total; done += 32768u)
            tma_bulk_g2s(reinterpret_cast<unsigned char *>(s_taps) + done,
                         reinterpret_cast<const unsigned char *>(tile_taps) + done, min(total - done, 32768u), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    const u32 w0 = group_xs[warp];
    const u32 w0_last = group_xs[tp.groups - 1];
    const bool aligned16 = (reinterpret_cast<uintptr_t>(signal) & 15) == 0;
    const u32 it_a_end = tp.half_taps / 16;       // half A is active for iterations [0, it_a_end)
    const u32 it_b_begin = tp.shift / 16;         // half B for [it_b_begin, iters)

    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 k_base = tile * tile_out;                       // first output of the tile
        const u64 x_base = tile * static_cast<u64>(QT) * tp.p_in; // first input sample of row 0
        // ---- stage the input rows (+ the virtual halo row) ----
        const u64 x_end = x_base + static_cast<u64>(QT - 1) * tp.p_in + tp.row_len;   // one past the last sample
        const bool interior = aligned16 && x_end <= len;
        const bool want_halo = ENVELOPE && k_base > 0;
        const u64 x_halo = x_base - tp.p_in + w0_last;            // only meaningful when k_base > 0
        if (tp.debug == 2) {
            __syncthreads();
        } else if (interior) {
            if (tid == 0) {
                fence_proxy_async();              // the buffer was last touched through the generic proxy
                mbar_expect_tx(bar, (QT * tp.row_len + (want_halo ? tp.usteps : 0)) * 4);
            }
            if (lane == 0) {
                fence_proxy_async();
                for (u32 q = warp; q < QT; q += tp.groups) {
                    if (tp.debug >= 3) {   // timing experiment: 128-byte aligned source (wrong data)
                        const u64 xa = (x_base + static_cast<u64>(q) * tp.p_in) & ~static_cast<u64>(31);
                        tma_bulk_g2s(s_rows + q * tp.row_len, signal + xa, tp.row_len * 4, bar);
                    } else
                    tma_bulk_g2s(s_rows + q * tp.row_len, signal + x_base + static_cast<u64>(q) * tp.p_in,
                                 tp.row_len * 4, bar);
                }
                if (want_halo && warp == tp.groups - 1) tma_bulk_g2s(s_vrow, signal + x_halo, tp.usteps * 4, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            // this tile has landed: pull the rows of this CTA's next tile into L2 while we compute, so that
            // HBM stays busy during the FMA phase (CTAs run in lockstep; without this loads and math alternate)
            if (lane == 0 && tile + gridDim.x < ntiles) {
                for (u32 q = warp; q < QT; q += tp.groups) {
                    const u64 nx = x_base + (static_cast<u64>(gridDim.x) * QT + q) * tp.p_in;
                    if (nx + tp.p_in <= len) tma_prefetch_l2(signal + nx, tp.p_in * 4);
                }
            }
        } else {
            for (u32 i = tid; i < QT * tp.row_len; i += nthreads) {
                const u32 q = i / tp.row_len, c = i - q * tp.row_len;
                const u64 x = x_base + static_cast<u64>(q) * tp.p_in + c;
                s_rows[i] = x < len ? __ldg(signal + x) : 0.f;    // past the end: signal.get(x) == None
            }
            if (want_halo)
                for (u32 i = tid; i < tp.usteps; i += nthreads) s_vrow[i] = x_halo + i < len ? __ldg(signal + x_halo + i) : 0.f;
            __syncthreads();
        }

        // ---- the product: 8 x 4 accumulators per thread (16 packed pairs) over this lane's chunks ----
        f32x2 acc_a[2][Q], acc_b[2][Q];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < Q; ++j) acc_a[p][j] = acc_b[p][j] = 0ull;
        if (tp.debug != 1 && tp.debug != 3) {
            // 32-bit shared-memory byte addresses, advanced by constants: no per-iteration address math
            u32 tap_addr = smem_u32(s_taps) + (warp * tp.group_stride + ks * tp.slice_stride) * 4;
            const u32 tap_step = KS * tp.slice_stride * 4;
            u32 row_addr = smem_u32(s_rows) + (ql * tp.row_len + w0 + ks * 4) * 4;
            const u32 row_step8 = 8 * tp.row_len * 4;            // rows ql, ql+8, ql+16, ql+24
            u32 it = 0;
            for (; it < it_b_begin; ++it) {                        // half A only
                float4 s[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j) s[j] = lds128(row_addr + j * row_step8);
                half_fma2(acc_a, tap_addr, s);
                row_addr += KS * 16;
                tap_addr += tap_step;
            }
            for (; it < it_a_end; ++it) {                          // both halves
                float4 s[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j) s[j] = lds128(row_addr + j * row_step8);
                half_fma2(acc_a, tap_addr, s);
                half_fma2(acc_b, tap_addr + 64, s);
                row_addr += KS * 16;
                tap_addr += tap_step;
            }
            for (; it < tp.iters; ++it) {                          // half B only
                float4 s[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j) s[j] = lds128(row_addr + j * row_step8);
                half_fma2(acc_b, tap_addr + 64, s);
                row_addr += KS * 16;
                tap_addr += tap_step;
            }
        }
        __syncthreads();   // every warp is done reading the rows: the buffer now takes the partial sums

        // ---- partial sums of the 4 slices -> planes [ks][row][plane_pitch] ----
        float *s_plane = s_rows;
        const u32 plane_floats = QT * tp.plane_pitch;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            float *dst = s_plane + ks * plane_floats + (ql + 8 * j) * tp.plane_pitch + warp * kTileR;
            float4 va, vb;
            unpack2(acc_a[0][j], va.x, va.y);
            unpack2(acc_a[1][j], va.z, va.w);
            unpack2(acc_b[0][j], vb.x, vb.y);
            unpack2(acc_b[1][j], vb.z, vb.w);
            *reinterpret_cast<float4 *>(dst) = va;
            *reinterpret_cast<float4 *>(dst + 4) = vb;
        }
        __syncthreads();
        // ---- r[K0 - 1]: output 7 of the last group of the previous super-period, from the virtual row ----
        if (ENVELOPE && warp == tp.groups - 1) {
            float part = 0.f;
            if (want_halo) {
                const float *tg = s_taps + static_cast<size_t>(tp.groups - 1) * tp.group_stride;
                for (u32 u = tp.shift + lane; u < tp.usteps; u += 32) {
                    const u32 chunk = u >> 2, uu = u & 3;            // chunk = it*KS + ks
                    part = fmaf(tg[chunk * tp.slice_stride + 16 + uu * H + 3], s_vrow[u], part);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_halo = part;
        }

        // ---- reduce the planes into plane 0 (4 outputs per thread-iteration) ----
        const u32 vec_per_row = tp.p_out / 4;
        const u32 nvec = QT * vec_per_row;
        for (u32 v = tid; v < nvec; v += nthreads) {
            const u32 q = (v * tp.vec_magic) >> 16, c4 = v - q * vec_per_row;      // v / vec_per_row, exact for v < nvec
            float *p0 = s_plane + q * tp.plane_pitch + 4 * c4;
            const float4 a = *reinterpret_cast<const float4 *>(p0);
            const float4 b = *reinterpret_cast<const float4 *>(p0 + plane_floats);
            const float4 c = *reinterpret_cast<const float4 *>(p0 + 2 * plane_floats);
            const float4 d = *reinterpret_cast<const float4 *>(p0 + 3 * plane_floats);
            *reinterpret_cast<float4 *>(p0) = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y),
                                                          (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
        }
        __syncthreads();
        // ---- epilogue: envelope (dsp.rs:373) and 16-byte stores ----
        const bool full_tile = k_base + tile_out <= nout && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
        float *out_tile = out + k_base;
        for (u32 v = tid; v < nvec; v += nthreads) {
            const u32 q = (v * tp.vec_magic) >> 16, c4 = v - q * vec_per_row;
            const u32 kl = q * tp.p_out + 4 * c4;                    // output index inside the tile
            const float *cell = s_plane + q * tp.plane_pitch + 4 * c4;
            const float4 cur = *reinterpret_cast<const float4 *>(cell);
            float4 res = cur;
            if (ENVELOPE) {
                // previous output: left neighbour, last output of the previous row, or the halo for kl == 0
                const float prev = c4 > 0 ? cell[-1] : q > 0 ? cell[-5] : s_halo;
                res.x = envelope2_fast(prev, cur.x, cosphi2, inv_sinphi);
                res.y = envelope2_fast(cur.x, cur.y, cosphi2, inv_sinphi);
                res.z = envelope2_fast(cur.y, cur.z, cosphi2, inv_sinphi);
                res.w = envelope2_fast(cur.z, cur.w, cosphi2, inv_sinphi);
                if (kl == 0 && k_base == 0) res.x = 0.f;             // output[0] = 0 (dsp.rs:357)
            }
            if (full_tile) {
                *reinterpret_cast<float4 *>(out_tile + kl) = res;
            } else {
                const u64 k = k_base + kl;
                if (k < nout) out[k] = res.x;
                if (k + 1 < nout) out[k + 1] = res.y;
                if (k + 2 < nout) out[k + 2] = res.z;
                if (k + 3 < nout) out[k + 3] = res.w;
            }
        }
        __syncthreads();   // the buffer is free for the next tile's bulk copies
    }
}

}  // namespace aptb200
