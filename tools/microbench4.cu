// Inner loop of the tiled resampler in isolation: 13 compute warps, 1 CTA/SM, operands in shared memory.
// Variants of the FMA formulation; reports cycles per tile (7 iterations: 1 half, 5 full, 1 half).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench4 tools/microbench4.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef unsigned int u32;
typedef unsigned long long f32x2;
__device__ __forceinline__ u32 smem_u32(const void *p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float4 lds128(u32 addr) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v; }

constexpr int ROW_LEN = 484, GROUP_STRIDE = 928, SLICE_STRIDE = 232, ITERS = 7, TILES = 64;

__device__ __forceinline__ void half_fma2(f32x2 (&acc)[2][4], u32 tap_addr, const float4 (&s)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 tp = lds128(tap_addr + 16 * u);
        const f32x2 t01 = pack2(tp.x, tp.y), t23 = pack2(tp.z, tp.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float sv = u == 0 ? s[j].x : u == 1 ? s[j].y : u == 2 ? s[j].z : s[j].w;
            const f32x2 sv2 = pack2(sv, sv);
            acc[0][j] = fma2(t01, sv2, acc[0][j]);
            acc[1][j] = fma2(t23, sv2, acc[1][j]);
        }
    }
}
__device__ __forceinline__ void half_fma1(float (&acc)[4][4], u32 tap_addr, const float4 (&s)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 tp = lds128(tap_addr + 16 * u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float sv = u == 0 ? s[j].x : u == 1 ? s[j].y : u == 2 ? s[j].z : s[j].w;
            acc[0][j] = fmaf(tp.x, sv, acc[0][j]);
            acc[1][j] = fmaf(tp.y, sv, acc[1][j]);
            acc[2][j] = fmaf(tp.z, sv, acc[2][j]);
            acc[3][j] = fmaf(tp.w, sv, acc[3][j]);
        }
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) k_inner(float *out, long long *cyc, int warps_used) {
    extern __shared__ __align__(128) float sm[];
    float *s_taps = sm, *s_rows = sm + 13 * GROUP_STRIDE;
    for (int i = threadIdx.x; i < 13 * GROUP_STRIDE + 32 * ROW_LEN; i += blockDim.x) sm[i] = 1e-3f * (i % 97);
    __syncthreads();
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ks = lane >> 3, ql = lane & 7;
    if (warp >= (u32)warps_used) return;
    const u32 g = warp % 13;
    const u32 tap_base = smem_u32(s_taps) + (g * GROUP_STRIDE + ks * SLICE_STRIDE) * 4;
    const u32 row_base = smem_u32(s_rows) + (ql * ROW_LEN + g * 28 + ks * 4) * 4;
    const u32 row_step8 = 8 * ROW_LEN * 4;
    float total = 0.f;
    long long t0 = clock64();
    for (int tile = 0; tile < TILES; ++tile) {
        if (VARIANT == 1) {
            float aa[4][4], ab[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) aa[r][j] = ab[r][j] = 0.f;
            u32 tap_addr = tap_base, row_addr = row_base;
            for (int it = 0; it < ITERS; ++it) {
                float4 s[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = lds128(row_addr + j * row_step8);
                if (it < ITERS - 1) half_fma1(aa, tap_addr, s);
                if (it > 0) half_fma1(ab, tap_addr + 64, s);
                row_addr += 64; tap_addr += 128;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) total += aa[r][j] + ab[r][j];
        } else {
            f32x2 aa[2][4], ab[2][4];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) aa[p][j] = ab[p][j] = 0ull;
            u32 tap_addr = tap_base, row_addr = row_base;
            if (VARIANT == 0) {
                for (int it = 0; it < ITERS; ++it) {
                    float4 s[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) s[j] = lds128(row_addr + j * row_step8);
                    if (it < ITERS - 1) half_fma2(aa, tap_addr, s);
                    if (it > 0) half_fma2(ab, tap_addr + 64, s);
                    row_addr += 64; tap_addr += 128;
                }
            } else if (VARIANT == 3) {   // ping-pong sample buffers, next iteration's samples requested first
                float4 s0[4], s1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) s0[j] = lds128(row_addr + j * row_step8);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    row_addr += 64;
                    if (it + 1 < ITERS) {
                        if (it & 1) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) s0[j] = lds128(row_addr + j * row_step8);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) s1[j] = lds128(row_addr + j * row_step8);
                        }
                    }
                    if (it & 1) {
                        if (it < ITERS - 1) half_fma2(aa, tap_addr, s1);
                        if (it > 0) half_fma2(ab, tap_addr + 64, s1);
                    } else {
                        if (it < ITERS - 1) half_fma2(aa, tap_addr, s0);
                        if (it > 0) half_fma2(ab, tap_addr + 64, s0);
                    }
                    tap_addr += 128;
                }
            } else {   // VARIANT 2: samples of the next iteration are loaded before this iteration's FMAs
                float4 s[4], sn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = lds128(row_addr + j * row_step8);
                for (int it = 0; it < ITERS; ++it) {
                    row_addr += 64;
                    if (it + 1 < ITERS) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) sn[j] = lds128(row_addr + j * row_step8);
                    }
                    if (it < ITERS - 1) half_fma2(aa, tap_addr, s);
                    if (it > 0) half_fma2(ab, tap_addr + 64, s);
                    tap_addr += 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) s[j] = sn[j];
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) { float lo, hi; unpack2(aa[p][j], lo, hi); total += lo + hi; unpack2(ab[p][j], lo, hi); total += lo + hi; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
static void run(const char *name, int warps) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    long long *cyc; float *out;
    CK(cudaMalloc(&cyc, sms * sizeof(long long))); CK(cudaMalloc(&out, sms * 1024 * sizeof(float)));
    const size_t smem = (13 * GROUP_STRIDE + 32 * ROW_LEN) * 4;
    CK(cudaFuncSetAttribute(k_inner<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = ((warps + 0) * 32);
    k_inner<V><<<sms, threads, smem>>>(out, cyc, warps); CK(cudaDeviceSynchronize());
    k_inner<V><<<sms, threads, smem>>>(out, cyc, warps); CK(cudaDeviceSynchronize());
    static long long h[256]; CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const double fma = warps * 32.0 * 768 * TILES;   // useful lane-FMAs per CTA
    printf("%-34s %2d warps: %7.0f cycles/tile  %6.1f FMA/clk/SM\n", name, warps, avg / TILES, fma / avg);
    cudaFree(cyc); cudaFree(out);
}

int main() {
    run<0>("FFMA2 (tap pair x broadcast sample)", 13);
    run<1>("scalar FFMA", 13);
    run<2>("FFMA2 + sample prefetch", 13);
    run<0>("FFMA2", 16); run<1>("scalar FFMA", 16); run<2>("FFMA2 + prefetch", 16);
    run<0>("FFMA2", 26); run<1>("scalar FFMA", 26); run<2>("FFMA2 + prefetch", 26);
    run<0>("FFMA2", 8); run<1>("scalar FFMA", 8);
    run<3>("FFMA2 ping-pong prefetch", 13); run<3>("FFMA2 ping-pong prefetch", 16); run<3>("FFMA2 ping-pong prefetch", 26);
    return 0;
}
