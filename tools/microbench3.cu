// TMA 1-D bulk copy (cp.async.bulk) streaming throughput vs copy size / copies in flight, against a plain
// LDG.128 streaming read.  One CTA per SM, each CTA streams its own contiguous slice of a 1 GiB buffer.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef unsigned int u32;
__device__ __forceinline__ u32 smem_u32(const void *p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(void *bar, u32 parity) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, u32 bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// STAGES buffers of `chunk` bytes; each filled by `copies` bulk copies of chunk/copies bytes; one thread drives.
__global__ void k_tma(const char *src, size_t per_cta, u32 chunk, u32 copies, u32 stages, int *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    unsigned char *buf = sm + 1024;
    if (threadIdx.x == 0) {
        for (u32 s = 0; s < stages; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const char *base = src + static_cast<size_t>(blockIdx.x) * per_cta;
    const u32 n = static_cast<u32>(per_cta / chunk);
    const u32 piece = chunk / copies;
    int acc = 0;
    for (u32 i = 0; i < n + stages; ++i) {
        if (i >= stages) {     // wait for chunk i - stages
            const u32 j = i - stages, s = j % stages;
            while (!mbar_try_wait(bars + s, (j / stages) & 1)) {}
            acc += buf[s * chunk];
        }
        if (i < n) {
            const u32 s = i % stages;
            mbar_expect_tx(bars + s, chunk);
            for (u32 c = 0; c < copies; ++c) tma_bulk_g2s(buf + s * chunk + c * piece, base + static_cast<size_t>(i) * chunk + c * piece, piece, bars + s);
        }
    }
    sink[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(512) k_ldg(const float4 *src, size_t n4_per_cta, float *sink) {
    const float4 *p = src + static_cast<size_t>(blockIdx.x) * n4_per_cta;
    float a = 0;
    for (size_t i = threadIdx.x; i < n4_per_cta; i += blockDim.x * 4) {
        float4 v0 = p[i], v1 = i + blockDim.x < n4_per_cta ? p[i + blockDim.x] : v0;
        float4 v2 = i + 2 * blockDim.x < n4_per_cta ? p[i + 2 * blockDim.x] : v0, v3 = i + 3 * blockDim.x < n4_per_cta ? p[i + 3 * blockDim.x] : v0;
        a += v0.x + v1.y + v2.z + v3.w;
    }
    if (a == 123.456f) sink[0] = a;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const size_t per_cta = 4u << 20;            // 4 MiB per CTA -> 592 MiB total, > L2
    const size_t total = per_cta * sms;
    char *src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
    int *sink; CK(cudaMalloc(&sink, sms * sizeof(int)));
    float *fs; CK(cudaMalloc(&fs, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    auto run = [&](u32 chunk, u32 copies, u32 stages) {
        const size_t smem = 1024 + static_cast<size_t>(chunk) * stages;
        k_tma<<<sms, 32, smem>>>(src, per_cta, chunk, copies, stages, sink); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_tma<<<sms, 32, smem>>>(src, per_cta, chunk, copies, stages, sink);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("tma chunk %6u B x %2u copies, %u stages: %7.1f GB/s (%.1f us)\n", chunk, copies, stages, total / ms / 1e6, ms * 1e3);
    };
    run(65536, 1, 1); run(65536, 1, 2); run(65536, 1, 3); run(65536, 32, 2); run(32768, 1, 2); run(32768, 1, 4); run(32768, 1, 6);
    run(16384, 1, 4); run(16384, 1, 8); run(16384, 1, 12); run(8192, 1, 8); run(8192, 1, 16); run(8192, 1, 24); run(2048, 1, 32); run(2048, 1, 64);
    k_ldg<<<sms, 512>>>(reinterpret_cast<const float4 *>(src), per_cta / 16, fs); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_ldg<<<sms, 512>>>(reinterpret_cast<const float4 *>(src), per_cta / 16, fs);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.128 512 thr x1 CTA/SM: %7.1f GB/s\n", total / ms / 1e6);
    k_ldg<<<sms * 4, 512>>>(reinterpret_cast<const float4 *>(src), per_cta / 16 / 4, fs); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_ldg<<<sms * 4, 512>>>(reinterpret_cast<const float4 *>(src), per_cta / 16 / 4, fs);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.128 512 thr x4 CTA/SM: %7.1f GB/s\n", total / ms / 1e6);
    return 0;
}
