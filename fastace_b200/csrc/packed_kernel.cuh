// The packed host encoding (fastace_actions_packed_t) expanded on the device into the compact encoding the step kernels
// read: one thread per (economy, agent) unpacks the agent's S bit fields into S index bytes and a take mask.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

struct ExpandParams {
    int agents;            // E * P  (or E * F)
    int S, bits, bytes;    // field width and byte-string length per agent
    const uint8_t* packed; // [agents][bytes]
    uint8_t* idx;          // [agents][S]   compact, agent-major
    uint16_t* take;        // [agents]
};

__global__ void expand_packed_kernel(const ExpandParams a, const ExpandParams b) {
    // two lists per launch (a person's job and goods lists; for firms b.agents = 0)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int which = 0; which < 2; which++) {
        const ExpandParams& x = which == 0 ? a : b;
        if (t >= x.agents) continue;
        const uint8_t* src = x.packed + (size_t)t * x.bytes;
        uint8_t* dst = x.idx + (size_t)t * x.S;
        const uint32_t none = (1u << x.bits) - 1u;
        uint32_t take = 0;
        for (int i = 0; i < x.S; i++) {
            const int bit = i * x.bits, byte = bit >> 3;
            uint32_t w = src[byte];
            if (byte + 1 < x.bytes) w |= (uint32_t)src[byte + 1] << 8;
            const uint32_t v = (w >> (bit & 7)) & none;
            dst[i] = (uint8_t)(v == none ? 0xFFu : v);
            take |= (v != none ? 1u : 0u) << i;
        }
        x.take[t] = (uint16_t)take;
    }
}

}  // namespace fastace
