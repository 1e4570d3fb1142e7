// Sync-frame peak picking (decode.rs:204-263) and row alignment (decode.rs:122-134,158-159).
//
// The reference's picker is a sequential state machine over the correlation.  It is
// restated here in a form that parallelises (derivation in DESIGN.md "Peak picker"):
//
//   D   = min_distance = row*8/10                      (decode.rs:216)
//   root(p)      <=>  no corr[j] > corr[p] for j in (p, p+D]        (p "survives" the else-if at :250)
//   firstroot(s) =    smallest root >= s     == the peak the picker ends on when it starts at s
//   start s' after a peak p found from start s:  s' = max(p + D + 1, row*(s/row + 1))   (:241-246)
//   pushes at s': (s'/row - len) copies of s', the last of which is refined to firstroot(s')
//
// so the peak list is the orbit of  F(s) = max(firstroot(s) + D + 1, row*(s/row + 1)).
// k_roots finds every root with a van Herk / Gil-Werman sliding maximum (one CTA per block
// of D positions); the walk over the orbit touches only the root lists.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_generic.cuh"
#include "launch.hpp"

namespace aptb200 {

constexpr u32 kNoSeed = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// k_roots: block b owns positions [b*D, (b+1)*D).  For p in the block the window (p, p+D] splits
// into the rest of the block (suffix maximum) and a prefix of the next block (prefix maximum).
// Writes the block's roots, ascending, to root_list[b*D ...] and their count to root_count[b].
// Dynamic shared memory: 2*D floats.
// ---------------------------------------------------------------------------------------------
template <int THREADS, int CHUNK>
__global__ void __launch_bounds__(THREADS)
k_roots(const float *__restrict__ corr, u64 ncorr, u32 dist, u32 *__restrict__ root_list,
        u32 *__restrict__ root_count, SyncResult *__restrict__ result) {
    extern __shared__ float sm[];
    float *a = sm;             // a[0..D): this block, a[D..2D): next block (later: its prefix maxima)
    __shared__ float s_suffix[THREADS];   // max of chunks strictly to the right, within this block
    __shared__ float s_prefix[THREADS];   // max of chunks strictly to the left, within the next block
    __shared__ u32 s_count[THREADS];
    __shared__ u32 s_seed;

    const float NEG = -INFINITY;
    const u32 tid = threadIdx.x;
    const u64 base = static_cast<u64>(blockIdx.x) * dist;

    for (u32 i = tid; i < 2 * dist; i += THREADS) {
        const u64 g = base + i;
        a[i] = g < ncorr ? __ldg(corr + g) : NEG;
    }
    if (tid == 0) s_seed = kNoSeed;
    __syncthreads();

    // seed of the peak list: (0, 0.0) is replaced by the first corr[i] > 0.0 with i <= D
    // (decode.rs:208-209 with the else-if at :250 while i - 0 <= D).
    if (blockIdx.x == 0) {
        for (u32 i = tid; i <= dist && i < 2 * dist; i += THREADS)
            if (a[i] > 0.f) atomicMin(&s_seed, i);
    }

    const u32 lo = tid * CHUNK;
    const u32 hi = min(lo + CHUNK, dist);   // chunk [lo, hi) of the block (may be empty)

    float cmax_a = NEG, cmax_b = NEG;
    for (u32 i = lo; i < hi; ++i) {
        cmax_a = fmaxf(cmax_a, a[i]);
        cmax_b = fmaxf(cmax_b, a[dist + i]);
    }
    s_suffix[tid] = cmax_a;
    s_prefix[tid] = cmax_b;
    __syncthreads();

    // exclusive suffix / prefix maxima over chunks (log-step scans in shared memory)
    for (u32 step = 1; step < THREADS; step <<= 1) {
        const float sv = tid + step < THREADS ? s_suffix[tid + step] : NEG;
        const float pv = tid >= step ? s_prefix[tid - step] : NEG;
        __syncthreads();
        s_suffix[tid] = fmaxf(s_suffix[tid], sv);
        s_prefix[tid] = fmaxf(s_prefix[tid], pv);
        __syncthreads();
    }
    const float right = tid + 1 < THREADS ? s_suffix[tid + 1] : NEG;   // chunks > tid of this block
    const float left = tid > 0 ? s_prefix[tid - 1] : NEG;              // chunks < tid of the next block

    // prefix maxima of the next block's chunk, kept in registers (a[dist + i] is read by this thread only)
    float pm[CHUNK];
    {
        float run = left;
#pragma unroll
        for (int c = 0; c < CHUNK; ++c) {
            const u32 i = lo + c;
            if (i < hi) run = fmaxf(run, a[dist + i]);
            pm[c] = run;
        }
    }
    // walk the chunk right-to-left with the running suffix maximum
    u32 flags = 0;
    {
        float run = right;
#pragma unroll
        for (int c = CHUNK - 1; c >= 0; --c) {
            const u32 i = lo + c;
            if (i < hi) {
                const float v = a[i];
                const float wmax = fmaxf(run, pm[c]);       // max of corr over (p, p+D]
                if (base + i < ncorr && !(wmax > v)) flags |= 1u << c;
                run = fmaxf(run, v);
            }
        }
    }
    s_count[tid] = __popc(flags);
    __syncthreads();
    // exclusive scan of the per-thread counts
    for (u32 step = 1; step < THREADS; step <<= 1) {
        const u32 v = tid >= step ? s_count[tid - step] : 0;
        __syncthreads();
        s_count[tid] += v;
        __syncthreads();
    }
    u32 w = tid > 0 ? s_count[tid - 1] : 0;
    u32 *list = root_list + base;
#pragma unroll
    for (int c = 0; c < CHUNK; ++c)
        if (flags & (1u << c)) list[w++] = static_cast<u32>(base + lo + c);
    if (tid == THREADS - 1) root_count[blockIdx.x] = s_count[THREADS - 1];
    if (blockIdx.x == 0 && tid == 0) result->seed_index = s_seed;
}

// Smallest root >= s.  Binary search in the block of s, then the first root of the following
// blocks (the last correlation index is always a root, so the search terminates).
__device__ __forceinline__ u32 first_root(u32 s, u32 dist, const u32 *__restrict__ root_list,
                                          const u32 *__restrict__ root_count, u32 nblocks) {
    u32 b = s / dist;
    {
        const u32 *list = root_list + static_cast<u64>(b) * dist;
        u32 lo = 0, hi = root_count[b];
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (list[mid] < s) lo = mid + 1; else hi = mid;
        }
        if (lo < root_count[b]) return list[lo];
    }
    for (++b; b < nblocks; ++b)
        if (root_count[b] > 0) return root_list[static_cast<u64>(b) * dist];
    return 0xFFFFFFFFu;   // unreachable for s < ncorr
}

// ---------------------------------------------------------------------------------------------
// k_pick_sequential: one thread walks the orbit of F.  O(rows * log) dependent loads -- the
// always-correct fallback (and the path for pathological inputs with millions of roots).
// ---------------------------------------------------------------------------------------------
__global__ void k_pick_sequential(u64 ncorr, u64 nwork, u32 row, u32 dist, const u32 *__restrict__ root_list,
                                  const u32 *__restrict__ root_count, u32 nblocks, u32 *__restrict__ positions,
                                  u32 max_positions, SyncResult *__restrict__ result) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 len = 1;
    // peak #1: the seed (0, 0.0), refined if some corr[i] > 0 turns up within D of position 0
    const u32 seed = result->seed_index;
    u32 p = seed == kNoSeed ? 0u : first_root(seed, dist, root_list, root_count, nblocks);
    positions[0] = p;
    u64 s = max(static_cast<u64>(p) + dist + 1, 2ull * row);
    while (s < ncorr) {
        const u32 target = static_cast<u32>(s / row);     // peaks.len() after the pushes at s
        for (; len + 1 < target && len < max_positions; ++len) positions[len] = static_cast<u32>(s);
        p = first_root(static_cast<u32>(s), dist, root_list, root_count, nblocks);
        if (len < max_positions) positions[len] = p;
        ++len;
        s = max(static_cast<u64>(p) + dist + 1, static_cast<u64>(row) * (s / row + 1));
    }
    if (len > max_positions) len = max_positions;
    // rows: every peak but the last, as long as a whole row fits (decode.rs:125-127);
    // positions are non-decreasing so the passing ones are a prefix.
    u32 rows = 0;
    for (u32 i = 0; i + 1 < len; ++i)
        if (static_cast<u64>(positions[i]) + row < nwork) ++rows; else break;
    result->n_peaks = len;
    result->n_rows = rows;
    result->status = len < 5 ? 3u /* APT_ERR_FEW_SYNC_FRAMES */ : 0u;
    u32 total = 0;
    for (u32 b = 0; b < nblocks; ++b) total += root_count[b];
    result->n_roots = total;
}

// ---------------------------------------------------------------------------------------------
// k_gather_rows: aligned rows + final NoFilter/decimate stage fused (decode.rs:122-134, 158-159).
//   out[j*px + c] = f[pos[j] + c*dec]   for j < n_rows, c < px   (px = 2080, dec = work_rate/4160)
// Element 0 of the whole output is 0: dsp::filter with NoFilter never reads signal[0] (dsp.rs:399).
// positions == nullptr: the --no-sync branch, pos[j] = j*row (decode.rs:141-147).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gather_rows(const float *__restrict__ f, const u32 *__restrict__ positions,
              const SyncResult *__restrict__ result, u32 fixed_rows, u32 row, u32 px, u32 dec,
              float *__restrict__ out) {
    const u32 n_rows = positions ? result->n_rows : fixed_rows;
    for (u32 j = blockIdx.x; j < n_rows; j += gridDim.x) {
        const u64 p = positions ? positions[j] : static_cast<u64>(j) * row;
        for (u32 c = threadIdx.x; c < px; c += blockDim.x) {
            const float v = __ldg(f + p + static_cast<u64>(c) * dec);
            out[static_cast<u64>(j) * px + c] = (j == 0 && c == 0) ? 0.f : v;
        }
    }
}

}  // namespace aptb200
