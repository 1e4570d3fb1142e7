// Taps through the constant bank / uniform datapath: thread = 8 outputs x 1 row, warp = one group, all 32
// lanes share the tap (warp-uniform constant address), samples by LDS.128.  Reports FMA/clk/SM.
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

constexpr int ROW_LEN = 484, U = 112, TILES = 64;
__constant__ float c_taps[13 * U * 8];     // [group][u][r]  46.6 KB

// VARIANT 0: scalar FFMA, 1: FFMA2
template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) k_const(float *out, long long *cyc, int warps_used) {
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < 32 * ROW_LEN; i += blockDim.x) sm[i] = 1e-3f * (i % 97);
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for the compiler
    if (warp >= (u32)warps_used) return;
    const u32 g = warp % 13;
    const u32 row_base = smem_u32(sm) + (lane * ROW_LEN + g * 28) * 4;
    const float *tg = c_taps + g * U * 8;
    float total = 0.f;
    long long t0 = clock64();
    for (int tile = 0; tile < TILES; ++tile) {
        if (VARIANT == 0) {
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = 0.f;
#pragma unroll 2
            for (int c = 0; c < U / 4; ++c) {
                const float4 s = lds128(row_base + c * 16);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float sv = u == 0 ? s.x : u == 1 ? s.y : u == 2 ? s.z : s.w;
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = fmaf(tg[(c * 4 + u) * 8 + r], sv, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) total += acc[r];
        } else {
            f32x2 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = 0ull;
#pragma unroll 2
            for (int c = 0; c < U / 4; ++c) {
                const float4 s = lds128(row_base + c * 16);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float sv = u == 0 ? s.x : u == 1 ? s.y : u == 2 ? s.z : s.w;
                    const f32x2 sv2 = pack2(sv, sv);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        acc[r] = fma2(pack2(tg[(c * 4 + u) * 8 + 2 * r], tg[(c * 4 + u) * 8 + 2 * r + 1]), sv2, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) { float lo, hi; unpack2(acc[r], lo, hi); total += lo + hi; }
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
    const size_t smem = 32 * ROW_LEN * 4;
    CK(cudaFuncSetAttribute(k_const<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_const<V><<<sms, warps * 32, smem>>>(out, cyc, warps); CK(cudaDeviceSynchronize());
    k_const<V><<<sms, warps * 32, smem>>>(out, cyc, warps); CK(cudaDeviceSynchronize());
    static long long h[256]; CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const double fma = warps * 32.0 * 8 * U * TILES;
    printf("%-28s %2d warps: %7.0f cycles/tile  %6.1f FMA/clk/SM\n", name, warps, avg / TILES, fma / avg);
    cudaFree(cyc); cudaFree(out);
}

int main() {
    static float taps[13 * U * 8]; for (int i = 0; i < 13 * U * 8; ++i) taps[i] = 1.0f / (1 + i % 251);
    CK(cudaMemcpyToSymbol(c_taps, taps, sizeof(taps)));
    run<0>("const taps, scalar FFMA", 13); run<1>("const taps, FFMA2", 13);
    run<0>("const taps, scalar FFMA", 26); run<1>("const taps, FFMA2", 26);
    run<0>("const taps, scalar FFMA", 8); run<1>("const taps, FFMA2", 8);
    return 0;
}
