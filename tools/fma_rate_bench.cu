// Issue rate of the FP32 FMA forms the kernels use, per SM sub-partition (run on the GPU box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fma_rate_bench tools/fma_rate_bench.cu
//   A  FFMA2  R(pair)  x UR(pair) + R(pair)      k_lowpass_records / k_gather_rows_lp   (packed window x warp-uniform tap pair)
//   B  FFMA2  R(.F32)  x UR(pair) + R(pair)      k_polyphase_ut                         (broadcast sample x warp-uniform tap pair)
//   C  FFMA2  R(pair)  x R(pair)  + R(pair)      all-register packed
//   D  FFMA   R x UR + R                         scalar, uniform operand
//   E  FFMA   R x R  + R                         scalar, three registers
//   F  FFMA2  R(pair)  x (-1)     + R(pair)      packed add/subtract written as an FMA with an immediate (correlation signs)
//   G  FADD2  R(pair)  + R(pair)                 packed add
//   H  FADD   R + R                              scalar add
//   I  FFMA2  R(pair)  x R(.F32)  + R(pair)      k_polyphase_ph (tap pair from shared memory x broadcast sample)
// Each thread keeps 16 independent accumulator chains; one CTA per SM, W warps per CTA (W/4 per sub-partition); the
// reported figure is cycles per warp-instruction per sub-partition (1.0 = one instruction per clock), and FMA/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

struct Taps { float2 t[16]; };
constexpr int CH = 16, ITER = 4096;

template <int FORM>
__global__ void __launch_bounds__(1024, 1) k_rate(const __grid_constant__ Taps taps, const float *__restrict__ in, float *out, long long *cycles) {
    const int tid = threadIdx.x;
    f32x2 acc2[CH];
    float acc[CH];
    f32x2 x2[CH];
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        x[i] = in[(tid + 32 * i) & 1023];
        x2[i] = pack2(x[i], x[i] + 1.f);
        acc2[i] = 0ull;
        acc[i] = 0.f;
    }
    f32x2 r2[4];
    float r1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r1[i] = in[(tid * 7 + i) & 1023]; r2[i] = pack2(r1[i], r1[i] * 0.5f); }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const f32x2 u2 = pack2(taps.t[i].x, taps.t[i].y);
            if (FORM == 0) acc2[i] = fma2(x2[i], u2, acc2[i]);
            if (FORM == 1) acc2[i] = fma2(pack2(x[i], x[i]), u2, acc2[i]);
            if (FORM == 2) acc2[i] = fma2(x2[i], r2[i & 3], acc2[i]);
            if (FORM == 3) acc[i] = fmaf(x[i], taps.t[i].x, acc[i]);
            if (FORM == 4) acc[i] = fmaf(x[i], r1[i & 3], acc[i]);
            if (FORM == 5) acc2[i] = fma2(x2[i], pack2(-1.f, -1.f), acc2[i]);
            if (FORM == 6) acc2[i] = add2(acc2[i], x2[i]);
            if (FORM == 8) acc2[i] = fma2(x2[i], pack2(r1[i & 3], r1[i & 3]), acc2[i]);
            if (FORM == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(x[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float lo, hi;
        unpack2(acc2[i], lo, hi);
        s += lo + hi + acc[i];
    }
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int FORM>
static void run(const char *name, int fma_per_inst, const float *in, float *out, long long *cyc, int sms) {
    Taps t;
    for (int i = 0; i < 16; ++i) t.t[i] = make_float2(1.0f + 1e-3f * i, 1.0f - 1e-3f * i);
    for (int warps : {4, 8, 16, 32}) {
        k_rate<FORM><<<sms, 32 * warps>>>(t, in, out, cyc);
        cudaDeviceSynchronize();
        k_rate<FORM><<<sms, 32 * warps>>>(t, in, out, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed\n", name); return; }
        long long h[256];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += h[i];
        avg /= sms;
        const double inst_per_smsp = static_cast<double>(ITER) * CH * warps / 4.0;
        printf("%-44s %2d warps/SM: %.2f clk per warp-instruction per sub-partition, %.1f FMA/clk/SM\n", name, warps, avg / inst_per_smsp,
               inst_per_smsp * 4 * 32 * fma_per_inst / avg);
    }
}

// same-address global atomics (returning): every warp's lane 0 adds to ONE counter `per_warp` times in a dependent chain
__global__ void k_atomic(unsigned *counter, unsigned *sink, int per_warp) {
    if ((threadIdx.x & 31) != 0) return;
    unsigned v = 0;
    for (int i = 0; i < per_warp; ++i) v += atomicAdd(counter, 1u + (v & 0u));
    sink[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = v;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096);
    cudaMalloc(&out, sizeof(float) * 1024 * sms);
    cudaMalloc(&cyc, sizeof(long long) * 256);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 1e-4f * i;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    {
        unsigned *ctr, *sink;
        cudaMalloc(&ctr, 256);
        cudaMalloc(&sink, sizeof(unsigned) * 64 * sms);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int warps : {1, 4, 16, 32}) {
            const int per_warp = 64;
            cudaMemset(ctr, 0, 256);
            k_atomic<<<sms, 32 * warps>>>(ctr, sink, per_warp);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_atomic<<<sms, 32 * warps>>>(ctr, sink, per_warp);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double total = static_cast<double>(sms) * warps * per_warp;
            printf("same-address atomicAdd (returning), %4d warps in flight: %.0f atomics in %.1f us = %.2f ns each\n", sms * warps, total,
                   ms * 1e3, ms * 1e6 / total);
        }
    }
    run<0>("A FFMA2 R.pair x UR.pair + R.pair", 2, in, out, cyc, sms);
    run<1>("B FFMA2 R.F32 (broadcast) x UR.pair + R.pair", 2, in, out, cyc, sms);
    run<2>("C FFMA2 R.pair x R.pair + R.pair", 2, in, out, cyc, sms);
    run<3>("D FFMA  R x UR + R", 1, in, out, cyc, sms);
    run<4>("E FFMA  R x R + R", 1, in, out, cyc, sms);
    run<5>("F FFMA2 R.pair x (-1) + R.pair", 2, in, out, cyc, sms);
    run<6>("G FADD2 R.pair + R.pair", 2, in, out, cyc, sms);
    run<7>("H FADD  R + R", 1, in, out, cyc, sms);
    run<8>("I FFMA2 R.pair x R.F32 (broadcast) + R.pair", 2, in, out, cyc, sms);
    return 0;
}
