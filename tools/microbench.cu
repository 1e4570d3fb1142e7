// Micro-benchmarks that size the fused resample->envelope kernel (SURVEY.md F9 / §7 hard parts):
// fp32 FMA issue rate per SM for scalar FFMA, packed FFMA2 (fma.rn.f32x2), constant-bank operands,
// and shared-memory operand bandwidth.  One CTA of 1024 threads per SM; cycles from clock64().
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__constant__ float c_taps[1024];

constexpr int ITERS = 2048;

__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// scalar FFMA, 3 register operands, 16 independent chains
__global__ void __launch_bounds__(1024) k_ffma(float *out, float a, float b, long long *cyc) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FIR-like scalar FFMA: acc[i] += tap * x[i]   (tap shared by 8 accumulators; distinct x registers)
__global__ void __launch_bounds__(1024) k_ffma_fir(float *out, const float *in, long long *cyc) {
    float acc[16], x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = 0.f; x[i] = in[threadIdx.x + i]; }
    float t = in[threadIdx.x + 100];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(t, x[i], acc[i]);
        t += 1.0f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed FFMA2: acc2[i] = fma2(t2, x2[i], acc2[i])
__global__ void __launch_bounds__(1024) k_ffma2_fir(float *out, const float *in, long long *cyc) {
    unsigned long long acc[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = f2_pack(0.f, 0.f); x[i] = f2_pack(in[threadIdx.x + 2 * i], in[threadIdx.x + 2 * i + 1]); }
    float t = in[threadIdx.x + 100];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        unsigned long long t2 = f2_pack(t, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma2(t2, x[i], acc[i]);
        t += 1.0f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 with 16 packed accumulators (32 fp32 accumulators), two tap pairs
__global__ void __launch_bounds__(512) k_ffma2_wide(float *out, const float *in, long long *cyc) {
    unsigned long long acc[16], x[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = f2_pack(in[threadIdx.x + 2 * i], in[threadIdx.x + 2 * i + 1]);
    float t = in[threadIdx.x + 100];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        unsigned long long ta = f2_pack(t, t), tb = f2_pack(t + 1.f, t + 1.f), tc = f2_pack(t + 2.f, t + 2.f), td = f2_pack(t + 3.f, t + 3.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[4 * i + 0] = fma2(ta, x[i], acc[4 * i + 0]);
            acc[4 * i + 1] = fma2(tb, x[i], acc[4 * i + 1]);
            acc[4 * i + 2] = fma2(tc, x[i], acc[4 * i + 2]);
            acc[4 * i + 3] = fma2(td, x[i], acc[4 * i + 3]);
        }
        t += 1.0f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// scalar FFMA with a constant-bank operand (taps in __constant__, compile-time indices)
__global__ void __launch_bounds__(1024) k_ffma_const(float *out, const float *in, long long *cyc) {
    float acc[16], x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = 0.f; x[i] = in[threadIdx.x + i]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(c_taps[j * 16 + i], x[i], acc[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += 1.f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared memory: LDS.128 conflict-free, all lanes distinct (samples) -> bytes/clk/SM
__global__ void __launch_bounds__(1024) k_lds128(float *out, long long *cyc) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 v = sm[(idx + j * 32) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 7;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared memory: LDS.128 broadcast (all lanes the same address: the tap path)
__global__ void __launch_bounds__(1024) k_lds128_bcast(float *out, long long *cyc) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 v = sm[(idx + j) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 8;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the candidate inner loop: per u-step 1 LDS.128 (4 samples of 4 q's) + 2 broadcast LDS.128 (4 taps,
// duplicated into pairs) + 8 FFMA2 = 16 FMA
__global__ void __launch_bounds__(512) k_inner_4x4(float *out, long long *cyc) {
    __shared__ float4 smp[2048];   // samples [c][q], 32 KB
    __shared__ float4 tap[512];   // duplicated taps
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) smp[i] = make_float4(i, i + 1, i + 2, i + 3);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tap[i] = make_float4(i, i, i + 1, i + 1);
    __syncthreads();
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = f2_pack(0.f, 0.f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll 8
        for (int u = 0; u < 88; ++u) {
            const float4 s = smp[((u + it) * 32 + lane) & 2047];
            const float4 ta = tap[(warp * 8 + 2 * u) & 511];
            const float4 tb = tap[(warp * 8 + 2 * u + 1) & 511];
            const unsigned long long s01 = f2_pack(s.x, s.y), s23 = f2_pack(s.z, s.w);
            acc[0] = fma2(f2_pack(ta.x, ta.y), s01, acc[0]);
            acc[1] = fma2(f2_pack(ta.x, ta.y), s23, acc[1]);
            acc[2] = fma2(f2_pack(ta.z, ta.w), s01, acc[2]);
            acc[3] = fma2(f2_pack(ta.z, ta.w), s23, acc[3]);
            acc[4] = fma2(f2_pack(tb.x, tb.y), s01, acc[4]);
            acc[5] = fma2(f2_pack(tb.x, tb.y), s23, acc[5]);
            acc[6] = fma2(f2_pack(tb.z, tb.w), s01, acc[6]);
            acc[7] = fma2(f2_pack(tb.z, tb.w), s23, acc[7]);
        }
    }
    long long t1 = clock64();
    float sres = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); sres += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sres;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, int threads, double fma_per_thread, double bytes_per_thread) {
    int sms = 148;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); sms = p.multiProcessorCount;
    long long *cyc; float *out;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&out, sms * 1024 * sizeof(float)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(sms, threads, out, cyc); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    launch(sms, threads, out, cyc);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%-18s cycles/CTA %10.0f  ms %.3f  => %7.1f FMA/clk/SM  %7.1f B/clk/SM  (clock ~%.0f MHz)\n", name, avg, ms,
           fma_per_thread * threads / avg, bytes_per_thread * threads / avg, avg / (ms * 1e3));
    cudaFree(cyc); cudaFree(out);
}

int main() {
    float *in; CK(cudaMalloc(&in, 4096 * sizeof(float))); CK(cudaMemset(in, 0, 4096 * sizeof(float)));
    float taps[1024]; for (int i = 0; i < 1024; ++i) taps[i] = 1.0f / (i + 1);
    CK(cudaMemcpyToSymbol(c_taps, taps, sizeof(taps)));
    run("ffma_rrr", [&](int g, int t, float *o, long long *c) { k_ffma<<<g, t>>>(o, 1.0001f, 0.5f, c); }, 1024, 16.0 * ITERS, 0);
    run("ffma_fir", [&](int g, int t, float *o, long long *c) { k_ffma_fir<<<g, t>>>(o, in, c); }, 1024, 16.0 * ITERS, 0);
    run("ffma2_fir", [&](int g, int t, float *o, long long *c) { k_ffma2_fir<<<g, t>>>(o, in, c); }, 1024, 16.0 * ITERS, 0);
    run("ffma2_wide512", [&](int g, int t, float *o, long long *c) { k_ffma2_wide<<<g, t>>>(o, in, c); }, 512, 32.0 * ITERS, 0);
    run("ffma_const", [&](int g, int t, float *o, long long *c) { k_ffma_const<<<g, t>>>(o, in, c); }, 1024, 16.0 * ITERS, 0);
    run("lds128", [&](int g, int t, float *o, long long *c) { k_lds128<<<g, t>>>(o, c); }, 1024, 0, 128.0 * ITERS);
    run("lds128_bcast", [&](int g, int t, float *o, long long *c) { k_lds128_bcast<<<g, t>>>(o, c); }, 1024, 0, 128.0 * ITERS);
    run("inner_4x4_512", [&](int g, int t, float *o, long long *c) { k_inner_4x4<<<g, t>>>(o, c); }, 512, 16.0 * 88 * (ITERS / 8), 48.0 * 88 * (ITERS / 8));
    return 0;
}
