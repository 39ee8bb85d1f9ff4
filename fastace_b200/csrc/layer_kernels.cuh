// Element-wise halves of one training-time layer of the decision networks (policy side):
//     forward   y = [x +] tanh(z + b)            z = x W^T from cuBLAS (a plain library GEMM)
//     backward  dz = dy * (1 - t^2)              t = tanh(z + b) kept from the forward
// Eager autograd spends three element-wise passes per layer forward (bias add, tanh, residual add) and as many
// backward; these two kernels make it one each.  fp32 throughout, accurate tanhf: the reference-pinned gradient bar
// holds for this path too (tests/test_trainer.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fastace {

// n = rows*H elements; H % 4 == 0 takes the float4 path (all pointers 16-byte aligned: torch allocations)
__global__ void layer_forward_kernel(const float* __restrict__ z, const float* __restrict__ bias, const float* __restrict__ x,
                                     float* __restrict__ y, float* __restrict__ t, long long n, int H) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if ((H & 3) == 0 && i4 + 3 < n) {
        const float4 zv = *reinterpret_cast<const float4*>(z + i4);
        const float4 bv = *reinterpret_cast<const float4*>(bias + (int)(i4 % H));
        float4 tv = make_float4(tanhf(zv.x + bv.x), tanhf(zv.y + bv.y), tanhf(zv.z + bv.z), tanhf(zv.w + bv.w));
        *reinterpret_cast<float4*>(t + i4) = tv;
        if (x) {
            const float4 xv = *reinterpret_cast<const float4*>(x + i4);
            tv.x += xv.x; tv.y += xv.y; tv.z += xv.z; tv.w += xv.w;
        }
        *reinterpret_cast<float4*>(y + i4) = tv;
    } else {
        for (long long i = i4; i < n && i < i4 + 4; i++) {
            const float tv = tanhf(z[i] + bias[(int)(i % H)]);
            t[i] = tv;
            y[i] = x ? x[i] + tv : tv;
        }
    }
}

__global__ void layer_backward_kernel(const float* __restrict__ dy, const float* __restrict__ t, float* __restrict__ dz, long long n) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 3 < n && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0) {
        const float4 g = *reinterpret_cast<const float4*>(dy + i4);
        const float4 tv = *reinterpret_cast<const float4*>(t + i4);
        *reinterpret_cast<float4*>(dz + i4) = make_float4(g.x * (1.f - tv.x * tv.x), g.y * (1.f - tv.y * tv.y),
                                                          g.z * (1.f - tv.z * tv.z), g.w * (1.f - tv.w * tv.w));
    } else {
        for (long long i = i4; i < n && i < i4 + 4; i++) dz[i] = dy[i] * (1.f - t[i] * t[i]);
    }
}

}  // namespace fastace
