// Taps as a large kernel parameter (constant bank 0) read through the uniform datapath: thread = 1 row x 13 outputs
// (7 packed accumulators), warp-uniform tap index -> FFMA2 with uniform-register tap pairs.  FMA/clk/SM.
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

constexpr int U = 96, NP = 7, TILES = 64, P_IN = 50;
struct Taps { float t[U][2 * NP + 2]; };    // 96 x 16 floats = 6 KB

__global__ void __launch_bounds__(512, 1) k_uni(const __grid_constant__ Taps taps, float *out, long long *cyc, int rows_per_warp_dummy) {
    extern __shared__ __align__(16) float sm[];
    const int nrows = blockDim.x;
    for (int i = threadIdx.x; i < nrows * P_IN + U + 8; i += blockDim.x) sm[i] = 1e-3f * (i % 97);
    __syncthreads();
    const float *row = sm + threadIdx.x * P_IN;      // 8-byte aligned
    float total = 0.f;
    long long t0 = clock64();
    for (int tile = 0; tile < TILES; ++tile) {
        f32x2 acc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] = 0ull;
#pragma unroll 2
        for (int c = 0; c < U / 2; ++c) {
            const float2 s = *reinterpret_cast<const float2 *>(row + 2 * c);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float sv = u == 0 ? s.x : s.y;
                const f32x2 sv2 = pack2(sv, sv);
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    acc[p] = fma2(pack2(taps.t[2 * c + u][2 * p], taps.t[2 * c + u][2 * p + 1]), sv2, acc[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) { float lo, hi; unpack2(acc[p], lo, hi); total += lo + hi; }
        row += (tile & 1) ? -2 : 2;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    static Taps h; for (int u = 0; u < U; ++u) for (int r = 0; r < 16; ++r) h.t[u][r] = 1.0f / (1 + (u * 16 + r) % 251);
    long long *cyc; float *out;
    CK(cudaMalloc(&cyc, sms * sizeof(long long))); CK(cudaMalloc(&out, sms * 1024 * sizeof(float)));
    for (int threads : {128, 256, 384, 512}) {
        const size_t smem = (threads * P_IN + U + 16) * 4;
        CK(cudaFuncSetAttribute(k_uni, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_uni<<<sms, threads, smem>>>(h, out, cyc, 0); CK(cudaDeviceSynchronize());
        k_uni<<<sms, threads, smem>>>(h, out, cyc, 0); CK(cudaDeviceSynchronize());
        static long long hc[256]; CK(cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < sms; ++i) avg += hc[i]; avg /= sms;
        printf("uniform-tap FFMA2, %2d warps: %7.0f cycles/tile  %6.1f FMA/clk/SM\n", threads / 32, avg / TILES, threads * 2.0 * NP * U * TILES / avg);
    }
    return 0;
}
