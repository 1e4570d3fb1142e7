// Host-side copy microbenchmark (run on the GPU box): how fast can a pageable buffer reach the GPU?
//   a) cudaMemcpy from pageable memory (driver staging)
//   b) cudaMemcpy from pinned memory
//   c) T threads memcpy pageable -> pinned ring, H2D of chunk c overlapped with the memcpy of chunk c+1
//   d) cudaHostRegister + copy + unregister
// nvcc -O2 -o tools/hostcopy_bench tools/hostcopy_bench.cu -lpthread
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void par_memcpy(char *dst, const char *src, size_t bytes, int threads) {
    if (threads <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    size_t per = (bytes / threads + 4095) & ~size_t(4095);
    for (int t = 0; t < threads; ++t) {
        size_t a = per * t, b = std::min(bytes, a + per);
        if (a >= b) break;
        th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
    }
    for (auto &t : th) t.join();
}

int main(int argc, char **argv) {
    const size_t bytes = 172800000;
    char *page = (char *)malloc(bytes);
    memset(page, 1, bytes);
    char *pinned; cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault); memset(pinned, 2, bytes);
    char *dev; cudaMalloc(&dev, bytes);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now(); cudaMemcpy(dev, page, bytes, cudaMemcpyHostToDevice); double t1 = now();
        printf("pageable cudaMemcpy      : %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, bytes / (t1 - t0) / 1e9);
        t0 = now(); cudaMemcpy(dev, pinned, bytes, cudaMemcpyHostToDevice); t1 = now();
        printf("pinned cudaMemcpy        : %.2f ms  %.1f GB/s\n", (t1 - t0) * 1e3, bytes / (t1 - t0) / 1e9);
    }
    for (int threads : {1, 2, 4, 8, 16, 32}) {
        double t0 = now(); par_memcpy(pinned, page, bytes, threads); double t1 = now();
        printf("memcpy page->pinned %2d thr (spawned): %.2f ms  %.1f GB/s\n", threads, (t1 - t0) * 1e3, bytes / (t1 - t0) / 1e9);
    }
    // persistent worker pool + ring pipeline
    for (int threads : {2, 4, 8, 16}) for (size_t chunk : {size_t(4) << 20, size_t(16) << 20}) {
        const int NR = 4;
        char *ring[NR]; cudaEvent_t ev[NR];
        for (int i = 0; i < NR; ++i) { cudaHostAlloc(&ring[i], chunk, cudaHostAllocDefault); memset(ring[i], 0, chunk); cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); }
        std::atomic<long> gen{0}; std::atomic<int> done{0}; std::atomic<bool> quit{false};
        const char *cur_src = nullptr; char *cur_dst = nullptr; size_t cur_bytes = 0;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back([&, t] {
            long seen = 0;
            for (;;) {
                while (gen.load(std::memory_order_acquire) == seen) { if (quit.load()) return; }
                seen = gen.load(std::memory_order_acquire);
                size_t per = (cur_bytes / threads + 4095) & ~size_t(4095);
                size_t a = per * t, b = std::min(cur_bytes, a + per);
                if (a < b) memcpy(cur_dst + a, cur_src + a, b - a);
                done.fetch_add(1, std::memory_order_release);
            }
        });
        for (int rep = 0; rep < 3; ++rep) {
            double t0 = now();
            size_t off = 0; int c = 0;
            while (off < bytes) {
                size_t nb = std::min(chunk, bytes - off);
                int r = c % NR;
                if (c >= NR) cudaEventSynchronize(ev[r]);
                cur_src = page + off; cur_dst = ring[r]; cur_bytes = nb; done.store(0);
                gen.fetch_add(1, std::memory_order_release);
                while (done.load(std::memory_order_acquire) < threads) {}
                cudaMemcpyAsync(dev + off, ring[r], nb, cudaMemcpyHostToDevice, s);
                cudaEventRecord(ev[r], s);
                off += nb; ++c;
            }
            cudaStreamSynchronize(s);
            double t1 = now();
            if (rep == 2) printf("pipeline %2d thr chunk %2zu MiB : %.2f ms  %.1f GB/s\n", threads, chunk >> 20, (t1 - t0) * 1e3, bytes / (t1 - t0) / 1e9);
        }
        quit.store(true);
        for (auto &t : pool) t.join();
        for (int i = 0; i < NR; ++i) { cudaFreeHost(ring[i]); cudaEventDestroy(ev[i]); }
    }
    {
        double t0 = now(); cudaHostRegister(page, bytes, cudaHostRegisterDefault); double t1 = now();
        cudaMemcpy(dev, page, bytes, cudaMemcpyHostToDevice); double t2 = now();
        cudaHostUnregister(page); double t3 = now();
        printf("cudaHostRegister %.2f ms, copy %.2f ms, unregister %.2f ms\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3);
    }
    // D2H 15 MB into pageable
    {
        const size_t ob = 15000000; char *op = (char *)malloc(ob); memset(op, 0, ob);
        for (int rep = 0; rep < 2; ++rep) {
            double t0 = now(); cudaMemcpy(op, dev, ob, cudaMemcpyDeviceToHost); double t1 = now();
            printf("D2H 15 MB pageable: %.2f ms; ", (t1 - t0) * 1e3);
            t0 = now(); cudaMemcpy(pinned, dev, ob, cudaMemcpyDeviceToHost); t1 = now(); memcpy(op, pinned, ob); double t2 = now();
            printf("pinned: %.2f ms + memcpy %.2f ms\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3);
        }
    }
    return 0;
}
