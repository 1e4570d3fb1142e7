// Taps as a large kernel parameter (constant bank 0) read through the uniform datapath: thread = Q rows x 16 outputs
// (8 packed accumulators per row), warp-uniform tap index -> FFMA2 with uniform-register tap pairs.  FMA/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef unsigned int u32;
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int U = 96, TILES = 64, P_IN = 50;
template <int NP> struct Taps { float4 t[U][NP / 2]; };

template <int Q, int NP>
__global__ void __launch_bounds__(512, 1) k_uni(const __grid_constant__ Taps<NP> taps, float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    const int nrows = blockDim.x * Q;
    for (int i = threadIdx.x; i < nrows * P_IN + U + 8; i += blockDim.x) sm[i] = 1e-3f * (i % 97);
    __syncthreads();
    const float *row = sm + threadIdx.x * P_IN;      // 8-byte aligned; row q of this thread = + q * blockDim.x * P_IN
    const int qstride = blockDim.x * P_IN;
    float total = 0.f;
    long long t0 = clock64();
    for (int tile = 0; tile < TILES; ++tile) {
        f32x2 acc[Q][NP];
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[q][p] = 0ull;
#pragma unroll 1
        for (int c = 0; c < U / 2; ++c) {
            float2 s[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) s[q] = *reinterpret_cast<const float2 *>(row + q * qstride + 2 * c);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int h = 0; h < NP / 2; ++h) {
                    const float4 tv = taps.t[2 * c + u][h];
                    const f32x2 t0p = pack2(tv.x, tv.y), t1p = pack2(tv.z, tv.w);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float sv = u == 0 ? s[q].x : s[q].y;
                        const f32x2 sv2 = pack2(sv, sv);
                        acc[q][2 * h] = fma2(t0p, sv2, acc[q][2 * h]);
                        acc[q][2 * h + 1] = fma2(t1p, sv2, acc[q][2 * h + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int p = 0; p < NP; ++p) { float lo, hi; unpack2(acc[q][p], lo, hi); total += lo + hi; }
        row += (tile & 1) ? -2 : 2;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int Q, int NP> void run(int sms, float *out, long long *cyc) {
    static Taps<NP> h;
    float *hf = reinterpret_cast<float *>(&h);
    for (size_t i = 0; i < sizeof(h) / 4; ++i) hf[i] = 1.0f / (1 + i % 251);
    for (int threads : {128, 256, 384, 512}) {
        const size_t smem = (size_t)(threads * Q * P_IN + U + 16) * 4;
        if (smem > 227 * 1024) continue;
        CK(cudaFuncSetAttribute(k_uni<Q, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_uni<Q, NP><<<sms, threads, smem>>>(h, out, cyc); CK(cudaDeviceSynchronize());
        k_uni<Q, NP><<<sms, threads, smem>>>(h, out, cyc); CK(cudaDeviceSynchronize());
        static long long hc[256]; CK(cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < sms; ++i) avg += hc[i]; avg /= sms;
        printf("Q=%d NP=%d %2d warps: %7.0f cycles/tile  %6.1f FMA/clk/SM\n", Q, NP, threads / 32, avg / TILES, threads * 2.0 * NP * Q * U * TILES / avg);
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    long long *cyc; float *out;
    CK(cudaMalloc(&cyc, sms * sizeof(long long))); CK(cudaMalloc(&out, sms * 1024 * sizeof(float)));
    run<1, 8>(sms, out, cyc);
    run<2, 8>(sms, out, cyc);
    run<4, 8>(sms, out, cyc);
    run<2, 4>(sms, out, cyc);
    run<4, 4>(sms, out, cyc);
    return 0;
}
