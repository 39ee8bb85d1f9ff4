// Visiting orders on the device: Economy::time_step's two std::shuffle calls (/root/reference/src/base/economy.cpp:110-111),
// bit for bit as libstdc++ executes them with the economy's std::default_random_engine (= std::minstd_rand0).
//
// What has to be reproduced (libstdc++, bits/stl_algo.h `std::shuffle`, bits/uniform_int_dist.h):
//   * minstd_rand0: x <- 16807 * x mod (2^31 - 1), min() = 1, max() = 2^31 - 2, so urngrange = 2^31 - 3;
//     seed(s): x = s mod m, or 1 if that is 0.
//   * uniform_int_distribution{0, r}(g) with a generator range that is NOT 2^k - 1 takes the classic path:
//       uerange = r + 1; scaling = urngrange / uerange; past = uerange * scaling;
//       do ret = g() - 1; while (ret >= past);   return ret / scaling;
//   * std::shuffle over n elements: if urngrange / n >= n, elements are swapped in PAIRS from one draw:
//       if n is even: swap(a[1], a[uniform{0,1}]) first;  then for i (odd count left), with b = i + 1:
//       x = uniform{0, b * (b + 1) - 1};  swap(a[i], a[x / (b + 1)]);  swap(a[i + 1], a[x % (b + 1)]);  i += 2
//     else (n > 46340) one draw per element: swap(a[i], a[uniform{0, i}]).
// KAT (g++ 13.3): seed 1234, 0..9 -> 5 0 4 8 1 2 7 6 3 9 -> 8 7 0 3 4 1 9 5 2 6.
//
// One thread per economy (the generator is a serial chain); a block stages its economies' arrays in shared memory
// as 16-bit ids when they fit, so that HBM sees coalesced loads and stores only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

constexpr uint32_t kMinstdM = 2147483647u;         // 2^31 - 1
constexpr uint32_t kUrngRange = 2147483645u;       // max() - min()

struct Minstd0Dev {
    uint32_t x;
    __device__ __forceinline__ uint32_t next() {
        const uint64_t pr = (uint64_t)x * 16807ull;                         // < 2^46
        uint32_t r = (uint32_t)(pr & 0x7FFFFFFFull) + (uint32_t)(pr >> 31); // Mersenne reduction
        if (r >= kMinstdM) r -= kMinstdM;
        x = r;
        return r;
    }
    // uniform_int_distribution<unsigned long>{0, r}(*this), r < urngrange
    __device__ __forceinline__ uint32_t uniform(uint32_t r) {
        const uint32_t uerange = r + 1u;
        const uint32_t scaling = kUrngRange / uerange;
        const uint32_t past = uerange * scaling;
        uint32_t ret;
        do { ret = next() - 1u; } while (ret >= past);
        return ret / scaling;
    }
};

template <typename T>
__device__ __forceinline__ void shuffle_array(T* a, int stride, int n, Minstd0Dev& g) {
    if (n <= 1) return;
    auto swap = [&](int i, int j) { const T t = a[(size_t)i * stride]; a[(size_t)i * stride] = a[(size_t)j * stride]; a[(size_t)j * stride] = t; };
    if (kUrngRange / (uint32_t)n >= (uint32_t)n) {
        int i = 1;
        if ((n % 2) == 0) { swap(i, (int)g.uniform(1u)); i++; }
        while (i < n) {
            const uint32_t b0 = (uint32_t)i + 1u, b1 = b0 + 1u;
            const uint32_t x = g.uniform(b0 * b1 - 1u);
            swap(i, (int)(x / b1));
            swap(i + 1, (int)(x % b1));
            i += 2;
        }
    } else {
        for (int i = 1; i < n; i++) swap(i, (int)g.uniform((uint32_t)i));
    }
}

struct ShuffleParams {
    int E, P, F;
    uint32_t seed;
    int restart;            // 1: economy e seeds minstd_rand0(seed + e) and starts from the identity order
    int steps;              // cumulative shuffles of this launch (>= 1); step t writes slice t of the outputs
    uint64_t* rng_state;    // [E]      the engines, carried between launches
    int32_t* state_person;  // [E][P]   the cumulative orders, carried between launches
    int32_t* state_firm;    // [E][F]
    int32_t* out_person;    // [steps][E][P]  int32 form (or null)
    int32_t* out_firm;      // [steps][E][F]
    uint16_t* out_person16; // the same in 16-bit form (or null)
    uint16_t* out_firm16;
    int use_smem;           // 1: a block's economies are staged in shared memory as u16 ids
};

constexpr int kShuffleThreads = 64;   // economies per block

__global__ void __launch_bounds__(kShuffleThreads) shuffle_orders_kernel(const ShuffleParams sp) {
    FASTACE_DYN_SMEM(smem);
    const int e0 = blockIdx.x * kShuffleThreads, t = threadIdx.x, e = e0 + t;
    const int E = sp.E, P = sp.P, F = sp.F;
    const int ne = min(kShuffleThreads, E - e0);     // economies of this block
    Minstd0Dev g;
    g.x = 1u;
    if (e < E) {
        if (sp.restart) {
            const uint32_t s0 = (sp.seed + (uint32_t)e) % kMinstdM;       // linear_congruential_engine::seed
            g.x = s0 == 0u ? 1u : s0;
        } else {
            g.x = (uint32_t)sp.rng_state[e];
        }
    }
    const size_t EP = (size_t)E * P, EF = (size_t)E * F;
    if (sp.use_smem) {
        // element i of economy (e0 + k) lives at s[i * kShuffleThreads + k]: conflict-free for one thread per economy;
        // HBM sees only coalesced loads / stores of the block's contiguous [ne][P] and [ne][F] slabs
        uint16_t* sp_ = reinterpret_cast<uint16_t*>(smem);
        uint16_t* sf_ = sp_ + (size_t)P * kShuffleThreads;
        for (int k = t; k < ne * P; k += kShuffleThreads)
            sp_[(k % P) * kShuffleThreads + k / P] = sp.restart ? (uint16_t)(k % P) : (uint16_t)sp.state_person[(size_t)e0 * P + k];
        for (int k = t; k < ne * F; k += kShuffleThreads)
            sf_[(k % F) * kShuffleThreads + k / F] = sp.restart ? (uint16_t)(k % F) : (uint16_t)sp.state_firm[(size_t)e0 * F + k];
        __syncthreads();
        for (int step = 0; step < sp.steps; step++) {
            if (e < E) {
                shuffle_array<uint16_t>(sp_ + t, kShuffleThreads, P, g);      // persons first, then firms (economy.cpp:110-111)
                shuffle_array<uint16_t>(sf_ + t, kShuffleThreads, F, g);
            }
            __syncthreads();
            const bool last = step == sp.steps - 1;
            for (int k = t; k < ne * P; k += kShuffleThreads) {
                const uint16_t v = sp_[(k % P) * kShuffleThreads + k / P];
                if (sp.out_person16) sp.out_person16[(size_t)step * EP + (size_t)e0 * P + k] = v;
                if (sp.out_person) sp.out_person[(size_t)step * EP + (size_t)e0 * P + k] = (int32_t)v;
                if (last) sp.state_person[(size_t)e0 * P + k] = (int32_t)v;
            }
            for (int k = t; k < ne * F; k += kShuffleThreads) {
                const uint16_t v = sf_[(k % F) * kShuffleThreads + k / F];
                if (sp.out_firm16) sp.out_firm16[(size_t)step * EF + (size_t)e0 * F + k] = v;
                if (sp.out_firm) sp.out_firm[(size_t)step * EF + (size_t)e0 * F + k] = (int32_t)v;
                if (last) sp.state_firm[(size_t)e0 * F + k] = (int32_t)v;
            }
            __syncthreads();
        }
    } else if (e < E) {
        // large economies: in place in global memory (one economy = one thread = one serial chain); the host copies
        // the state arrays to the output slices after each launch (steps == 1 here)
        int32_t* pp = sp.state_person + (size_t)e * P;
        int32_t* pf = sp.state_firm + (size_t)e * F;
        if (sp.restart) {
            for (int i = 0; i < P; i++) pp[i] = i;
            for (int i = 0; i < F; i++) pf[i] = i;
        }
        shuffle_array<int32_t>(pp, 1, P, g);
        shuffle_array<int32_t>(pf, 1, F, g);
    }
    if (e < E) sp.rng_state[e] = (uint64_t)g.x;
}

}  // namespace fastace
