// Shared-memory access-pattern costs and candidate inner loops for the tiled polyphase kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ITERS = 1024;

// MODE: 0 LDS.32 distinct, 1 LDS.32 bcast, 2 LDS.64 bcast, 3 LDS.128 bcast, 4..6 LDS.128 with 2/4/8 distinct
// addresses per warp, 7 LDS.64 distinct, 8 LDS.128 distinct stride 32B (2-way), 9 LDS.128 distinct contiguous,
// 10 LDS.128 13 distinct addresses
template <int MODE>
__global__ void __launch_bounds__(1024) k_lds(float *out, long long *cyc) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    int base;   // in floats
    switch (MODE) {
    case 0: base = lane; break;
    case 1: case 2: case 3: base = 0; break;
    case 4: base = (lane >> 4) * 4; break;
    case 5: base = (lane >> 3) * 4; break;
    case 6: base = (lane >> 2) * 4; break;
    case 7: base = lane * 2; break;
    case 8: base = lane * 8; break;
    case 9: base = lane * 4; break;
    default: base = (lane % 13) * 4; break;
    }
    base += warp * 64;
    const float *f = reinterpret_cast<const float *>(sm);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = (base + j * 256 + it * 4) & 8191 & ~3 | (base & 3);
            if (MODE == 0 || MODE == 1) acc += f[idx];
            else if (MODE == 2 || MODE == 7) { float2 v = *reinterpret_cast<const float2 *>(f + (idx & ~1)); acc += v.x + v.y; }
            else { float4 v = *reinterpret_cast<const float4 *>(f + (idx & ~3)); acc += v.x + v.y + v.z + v.w; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Candidate inner loop: thread tile R outputs x Q periods, scalar FFMA.  Samples S[c][q] (transposed, q fastest):
// Q/4 LDS.128 per u-step; taps T[u][r] broadcast: R/4 LDS.128 per u-step.
template <int R, int Q, int THREADS>
__global__ void __launch_bounds__(THREADS) k_tile(float *out, long long *cyc, int usteps) {
    extern __shared__ float4 dyn[];
    float4 *smp = dyn;            // 4096 float4 = 64 KB
    float4 *tap = dyn + 4096;     // 1024 float4 = 16 KB
    for (int i = threadIdx.x; i < 4096; i += THREADS) smp[i] = make_float4(i, i + 1, i + 2, i + 3);
    for (int i = threadIdx.x; i < 1024; i += THREADS) tap[i] = make_float4(i, i + .5f, i + 1, i + 2);
    __syncthreads();
    float acc[R][Q];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[r][q] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll 4
        for (int u = 0; u < usteps; ++u) {
            float s[Q], t[R];
#pragma unroll
            for (int q4 = 0; q4 < Q / 4; ++q4) {
                const float4 v = smp[((u + it) * 32 * (Q / 4) + q4 * 32 + lane) & 4095];
                s[4 * q4] = v.x; s[4 * q4 + 1] = v.y; s[4 * q4 + 2] = v.z; s[4 * q4 + 3] = v.w;
            }
#pragma unroll
            for (int r4 = 0; r4 < R / 4; ++r4) {
                const float4 v = tap[(warp * 16 + u * (R / 4) + r4) & 1023];
                t[4 * r4] = v.x; t[4 * r4 + 1] = v.y; t[4 * r4 + 2] = v.z; t[4 * r4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int q = 0; q < Q; ++q) acc[r][q] = fmaf(t[r], s[q], acc[r][q]);
        }
    }
    long long t1 = clock64();
    float sres = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < Q; ++q) sres += acc[r][q];
    out[blockIdx.x * THREADS + threadIdx.x] = sres;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, int threads, double fma_per_thread, double lds_per_warp, int ctas_per_sm = 1) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, grid = sms * ctas_per_sm;
    long long *cyc; float *out;
    CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    CK(cudaMalloc(&out, grid * 1024 * sizeof(float)));
    launch(grid, threads, out, cyc); CK(cudaDeviceSynchronize());
    launch(grid, threads, out, cyc); CK(cudaDeviceSynchronize());
    static long long h[4096]; CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    const double warps = threads / 32.0 * ctas_per_sm;
    printf("%-28s cycles/CTA %9.0f  %7.1f FMA/clk/SM  %6.2f clk per warp-LDS\n", name, avg,
           fma_per_thread * threads * ctas_per_sm / avg, lds_per_warp > 0 ? avg / (lds_per_warp * warps) : 0.0);
    cudaFree(cyc); cudaFree(out);
}

template <int R, int Q, int THREADS>
static void run_tile(const char *name, int usteps, int ctas) {
    auto kern = k_tile<R, Q, THREADS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    run(name, [&](int g, int t, float *o, long long *c) { kern<<<g, t, 80 * 1024>>>(o, c, usteps); }, THREADS,
        double(R) * Q * usteps * (ITERS / 16), double(Q / 4 + R / 4) * usteps * (ITERS / 16), ctas);
}

int main() {
    const char *names[] = {"LDS.32 distinct", "LDS.32 bcast", "LDS.64 bcast", "LDS.128 bcast", "LDS.128 2 addr", "LDS.128 4 addr",
                           "LDS.128 8 addr", "LDS.64 distinct", "LDS.128 stride32B", "LDS.128 distinct", "LDS.128 13 addr"};
#define RUNL(M) run(names[M], [&](int g, int t, float *o, long long *c) { k_lds<M><<<g, t>>>(o, c); }, 1024, 0, 8.0 * ITERS);
    RUNL(0) RUNL(1) RUNL(2) RUNL(3) RUNL(4) RUNL(5) RUNL(6) RUNL(7) RUNL(8) RUNL(9) RUNL(10)
    run_tile<4, 4, 512>("tile R4 Q4 512thr x1", 88, 1);
    run_tile<8, 4, 512>("tile R8 Q4 512thr x1", 104, 1);
    run_tile<8, 4, 256>("tile R8 Q4 256thr x2", 104, 2);
    run_tile<4, 8, 256>("tile R4 Q8 256thr x2", 88, 2);
    run_tile<8, 8, 256>("tile R8 Q8 256thr x1", 104, 1);
    run_tile<8, 8, 256>("tile R8 Q8 256thr x2", 104, 2);
    run_tile<8, 8, 128>("tile R8 Q8 128thr x4", 104, 4);
    run_tile<16, 4, 256>("tile R16 Q4 256thr x2", 136, 2);
    run_tile<12, 8, 128>("tile R12 Q8 128thr x3", 120, 3);
    return 0;
}
