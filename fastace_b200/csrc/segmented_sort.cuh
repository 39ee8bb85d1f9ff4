// Stable counting sort of a request list by the firm on the other side — the "segmented sort" of the large-economy
// path (large_economy.cuh): after it, the events of firm f are the contiguous segment [seg[f], seg[f+1]) in the
// reference's event order (visiting rank, jobs before goods, slot), which is simply the input order.
//
// The keys are small (firm id < F, or 0xFFFF = "no request", sorted behind everything) and the input is already in
// event order, so one counting pass does it; no radix digits, no library:
//   sort_chunk_hist     one warp per chunk of `chunk` consecutive elements: histogram over the F + 1 bins in shared
//                       memory -> hist[chunk][bin]
//   sort_bin_scan       one warp per bin: exclusive prefix of its column over the chunks (+ the bin's segment start),
//                       written back in place: where the chunk's first element of that bin goes
//   sort_chunk_scatter  one warp per chunk, 32 elements at a time in input order: rank among equal keys of the batch by
//                       match.any + popc, running per-bin offsets in shared memory -> stable scatter of (key, value)
// `seg` (exclusive prefix of the per-firm totals; seg[F] = number of real requests) comes from large_scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

struct SortParams {
    const uint16_t* key_in;    // [n] firm id or 0xFFFF
    const uint32_t* val_in;    // [n]
    uint16_t* key_out;         // [n]
    uint32_t* val_out;         // [n]
    const uint32_t* seg;       // [F + 1]
    uint32_t* hist;            // [chunks][F + 1]
    int n, F, chunk, chunks;
};

__device__ __forceinline__ int sort_bin_of(uint32_t key, int F) { return key == 0xFFFFu ? F : (int)key; }

__global__ void __launch_bounds__(32) sort_chunk_hist(const SortParams sp) {
    FASTACE_DYN_SMEM(smem);
    uint32_t* h = reinterpret_cast<uint32_t*>(smem);
    const int lane = threadIdx.x, c = blockIdx.x, NB = sp.F + 1;
    for (int b = lane; b < NB; b += 32) h[b] = 0u;
    __syncwarp();
    const int lo = c * sp.chunk, hi = min(sp.n, lo + sp.chunk);
    for (int k = lo + lane; k < hi; k += 32) atomicAdd(&h[sort_bin_of(sp.key_in[k], sp.F)], 1u);
    __syncwarp();
    uint32_t* out = sp.hist + (size_t)c * NB;
    for (int b = lane; b < NB; b += 32) out[b] = h[b];
}

constexpr int kSortScanWarps = 4;

__global__ void __launch_bounds__(32 * kSortScanWarps) sort_bin_scan(const SortParams sp) {
    const int lane = threadIdx.x & 31, NB = sp.F + 1;
    const int b = blockIdx.x * kSortScanWarps + (threadIdx.x >> 5);
    if (b >= NB) return;
    uint32_t run = sp.seg[b];                      // bin F ("no request") starts where the real requests end
    for (int c0 = 0; c0 < sp.chunks; c0 += 32) {
        const int c = c0 + lane;
        const uint32_t v = c < sp.chunks ? sp.hist[(size_t)c * NB + b] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (c < sp.chunks) sp.hist[(size_t)c * NB + b] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void __launch_bounds__(32) sort_chunk_scatter(const SortParams sp) {
    FASTACE_DYN_SMEM(smem);
    uint32_t* off = reinterpret_cast<uint32_t*>(smem);
    const int lane = threadIdx.x, c = blockIdx.x, NB = sp.F + 1;
    const uint32_t* start = sp.hist + (size_t)c * NB;
    for (int b = lane; b < NB; b += 32) off[b] = start[b];
    __syncwarp();
    const int lo = c * sp.chunk, hi = min(sp.n, lo + sp.chunk);
    constexpr int kAhead = 4;                     // batches whose loads are in flight before the first is ranked
    for (int g0 = lo; g0 < hi; g0 += 32 * kAhead) {
        uint32_t keys[kAhead], vals[kAhead];
#pragma unroll
        for (int j = 0; j < kAhead; j++) {
            const int k = g0 + 32 * j + lane;
            keys[j] = k < hi ? (uint32_t)sp.key_in[k] : 0xFFFFFFFFu;      // idle lanes: a key nobody else has
            vals[j] = k < hi ? sp.val_in[k] : 0u;
        }
#pragma unroll
        for (int j = 0; j < kAhead; j++) {
            if (g0 + 32 * j >= hi) break;                                // uniform
            const bool valid = g0 + 32 * j + lane < hi;
            const uint32_t key = keys[j];
            const unsigned peers = __match_any_sync(0xffffffffu, key);   // lanes of the batch with the same key
            const int rank = __popc(peers & ((1u << lane) - 1u));        // input order within the batch = lane order
            const int leader = __ffs((int)peers) - 1;
            uint32_t base = 0;
            if (valid && lane == leader) {
                const int b = sort_bin_of(key, sp.F);
                base = off[b];
                off[b] = base + (uint32_t)__popc(peers);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) {
                sp.key_out[base + rank] = (uint16_t)key;
                sp.val_out[base + rank] = vals[j];
            }
            __syncwarp();
        }
    }
}

}  // namespace fastace
