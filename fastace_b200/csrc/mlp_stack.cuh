// Fused residual tanh stack of the decision networks (policy side):
//     x <- x + tanh(x W_l^T + b_l)   for l = 0 .. L-1            (decisionNets.cpp:66-70, 135-139, 186-190, ...)
// for every row (agent) of a [rows, H] activation matrix, H <= 128 (default hidden size 100, 12 layers).
//
// Eager torch runs each layer as GEMM + bias add + tanh + residual add, four trips through HBM per layer
// (profiles/r01_summary.md: 76 % of a policy step).  Here a warp owns 16 rows and keeps the fp32 residual stream
// in registers for the whole stack: per layer it re-packs the stream to bf16 A fragments (the m16n8k16 accumulator
// layout of two adjacent 8-column tiles IS the A-fragment layout of one 16-wide k-tile, so no shuffle or shared
// memory is involved), multiplies by the layer's bf16 weights staged in shared memory (cp.async, double-buffered,
// row stride padded to a conflict-free 4*odd words) with mma.sync, and applies bias + tanh.approx + residual in the
// accumulator registers.  HBM traffic: one read and one write of the activations for the whole stack.
// bf16 operands / fp32 accumulate and residual: a rollout-only fast mode (no autograd), ~1e-2 absolute on the nets'
// outputs against the fp32 path (tests/test_policy.py).
//
// Why mma.sync and not tcgen05: H = 100 gives 16 x 104 x 112 tiles per warp-layer whose operand A lives in
// registers between layers; tcgen05 wants A in shared memory / TMEM and 128-row tiles, i.e. a round trip of the
// stream through shared memory per layer.  The kernel is bound by shared-memory fragment loads, not tensor throughput.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fastace {

constexpr int kMlpWarps = 6;
constexpr int kMlpBlocksPerSM = 2;   // two independent CTAs per SM: their mma / tanh phases interleave
constexpr int kMlpThreads = kMlpWarps * 32;

struct MlpParams {
    const float* x;        // [rows][H] stack input (when there is no first layer)
    float* y;              // [rows][H] stack output (may alias x), or null
    long long rows;
    int H, L;
    const uint16_t* w;     // [L][NT*8][KP] bf16, W_l[n][k] zero-padded
    const float* bias;     // [L][NT*8] zero-padded
    // optional first layer  x <- tanh(x0 W0^T + b0)  (decisionNets.cpp:63, 133, 184, ...)
    const float* x0;       // [rows][K0] or null
    int K0, K0P;           // K0P = 16*ceil(K0/16) + 8: padded row of W0
    const uint16_t* w0;    // [NT*8][K0P] bf16
    const float* b0;       // [NT*8]
    // optional last layer  out <- act(x Wl^T + bl), at most 16 outputs (decisionNets.cpp:76, 143, 195, ...)
    float* out;            // [rows][NOUT] or null
    int NOUT, act;         // act: 0 none, 1 sigmoid, 2 tanh
    const uint16_t* wl;    // [16][KP] bf16
    const float* bl;       // [16]
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int NT>
struct MlpTile {
    static constexpr int KT = (NT + 1) / 2;          // 16-wide k tiles
    static constexpr int NP = NT * 8;                // padded outputs
    static constexpr int KP = KT * 16 + 8;           // padded row of W in bf16 elements (stride = 4*odd words)
    static constexpr int W_BYTES = NP * KP * 2;
    static constexpr int B_BYTES = NP * 4;
    static constexpr int STAGE = W_BYTES + B_BYTES;  // multiple of 16
    static constexpr int SMEM = 2 * STAGE;
    static constexpr int LAST_BYTES = 16 * KP * 2 + 16 * 4;   // last layer: 16 padded outputs + bias
};

template <int NT>
__global__ void __launch_bounds__(kMlpThreads, kMlpBlocksPerSM) mlp_residual_stack_kernel(const MlpParams mp) {
    using T = MlpTile<NT>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int H = mp.H, L = mp.L;
    const long long rows_per_block = (long long)kMlpWarps * 16;
    const long long nblocks = (mp.rows + rows_per_block - 1) / rows_per_block;

    // shared memory: two stages of hidden weights | last layer | first layer
    uint16_t* s_wl = reinterpret_cast<uint16_t*>(smem + T::SMEM);
    float* s_bl = reinterpret_cast<float*>(smem + T::SMEM + 16 * T::KP * 2);
    uint16_t* s_w0 = reinterpret_cast<uint16_t*>(smem + T::SMEM + T::LAST_BYTES);
    float* s_b0 = reinterpret_cast<float*>(smem + T::SMEM + T::LAST_BYTES + (size_t)T::NP * mp.K0P * 2);
    if (mp.out) {
        for (int c = threadIdx.x; c < 16 * T::KP * 2 / 16; c += kMlpThreads)
            cp_async16(reinterpret_cast<unsigned char*>(s_wl) + c * 16, reinterpret_cast<const unsigned char*>(mp.wl) + c * 16);
        for (int c = threadIdx.x; c < 4; c += kMlpThreads)
            cp_async16(reinterpret_cast<unsigned char*>(s_bl) + c * 16, reinterpret_cast<const unsigned char*>(mp.bl) + c * 16);
    }
    if (mp.x0) {
        const int wbytes = T::NP * mp.K0P * 2;
        for (int c = threadIdx.x; c < wbytes / 16; c += kMlpThreads)
            cp_async16(reinterpret_cast<unsigned char*>(s_w0) + c * 16, reinterpret_cast<const unsigned char*>(mp.w0) + c * 16);
        for (int c = threadIdx.x; c < T::B_BYTES / 16; c += kMlpThreads)
            cp_async16(reinterpret_cast<unsigned char*>(s_b0) + c * 16, reinterpret_cast<const unsigned char*>(mp.b0) + c * 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    auto stage_w = [&](int s) { return reinterpret_cast<uint16_t*>(smem + (size_t)s * T::STAGE); };
    auto stage_b = [&](int s) { return reinterpret_cast<float*>(smem + (size_t)s * T::STAGE + T::W_BYTES); };
    auto prefetch = [&](int layer, int s) {
        const unsigned char* gw = reinterpret_cast<const unsigned char*>(mp.w) + (size_t)layer * T::W_BYTES;
        const unsigned char* gb = reinterpret_cast<const unsigned char*>(mp.bias) + (size_t)layer * T::B_BYTES;
        unsigned char* sw = smem + (size_t)s * T::STAGE;
        for (int c = threadIdx.x; c < T::W_BYTES / 16; c += kMlpThreads) cp_async16(sw + c * 16, gw + c * 16);
        for (int c = threadIdx.x; c < T::B_BYTES / 16; c += kMlpThreads) cp_async16(sw + T::W_BYTES + c * 16, gb + c * 16);
        cp_async_commit();
    };

    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long long row0 = blk * rows_per_block + warp * 16 + g;   // this thread's rows: row0 and row0 + 8
        // residual stream in accumulator layout: r[nt] = (row0, 8nt+2t), (row0, 8nt+2t+1), (row0+8, ..), (row0+8, ..)
        float r[NT][4];
        if (mp.x0) {
            // first layer: A fragments straight from the fp32 input rows (scalar loads: K0 may be odd), no residual
            const int K0 = mp.K0, K0P = mp.K0P;
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
            const float* xa = mp.x0 + row0 * K0;           // row g
            const float* xb = xa + 8 * (long long)K0;      // row g + 8
            const bool va = row0 < mp.rows, vb = row0 + 8 < mp.rows;
            for (int kt = 0; kt * 16 < K0; kt++) {
                const int k = kt * 16 + 2 * t;
                auto ld = [&](const float* base, bool ok, int kk) { return (ok && kk < K0) ? base[kk] : 0.f; };
                const uint32_t a0 = pack_bf16(ld(xa, va, k), ld(xa, va, k + 1));
                const uint32_t a1 = pack_bf16(ld(xb, vb, k), ld(xb, vb, k + 1));
                const uint32_t a2 = pack_bf16(ld(xa, va, k + 8), ld(xa, va, k + 9));
                const uint32_t a3 = pack_bf16(ld(xb, vb, k + 8), ld(xb, vb, k + 9));
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    const uint16_t* wr = s_w0 + (size_t)(nt * 8 + g) * K0P + kt * 16 + 2 * t;
                    mma_bf16(acc[nt], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const float2 bb = *reinterpret_cast<const float2*>(s_b0 + nt * 8 + 2 * t);
                r[nt][0] = tanh_fast(acc[nt][0] + bb.x);
                r[nt][1] = tanh_fast(acc[nt][1] + bb.y);
                r[nt][2] = tanh_fast(acc[nt][2] + bb.x);
                r[nt][3] = tanh_fast(acc[nt][3] + bb.y);
            }
            // padded columns (>= H) come out as tanh(0 + 0) = 0: W0 rows and b0 are zero there
        } else {
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const int c = nt * 8 + 2 * t;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const long long row = row0 + 8 * h;
                    float2 v = make_float2(0.f, 0.f);
                    if (row < mp.rows) {
                        if (c + 1 < H) v = *reinterpret_cast<const float2*>(mp.x + row * H + c);
                        else if (c < H) v.x = mp.x[row * H + c];
                    }
                    r[nt][2 * h] = v.x; r[nt][2 * h + 1] = v.y;
                }
            }
        }
        __syncthreads();            // every warp is done with both stages of the previous row block
        if (L > 0) prefetch(0, 0);
        for (int l = 0; l < L; l++) {
            const int s = l & 1;
            if (l + 1 < L) { prefetch(l + 1, s ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncthreads();        // layer l's weights are visible to all warps
            const uint16_t* w = stage_w(s);
            const float* b = stage_b(s);
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
            for (int kt = 0; kt < T::KT; kt++) {
                const uint32_t a0 = pack_bf16(r[2 * kt][0], r[2 * kt][1]);
                const uint32_t a1 = pack_bf16(r[2 * kt][2], r[2 * kt][3]);
                uint32_t a2 = 0u, a3 = 0u;
                if (2 * kt + 1 < NT) {
                    a2 = pack_bf16(r[2 * kt + 1][0], r[2 * kt + 1][1]);
                    a3 = pack_bf16(r[2 * kt + 1][2], r[2 * kt + 1][3]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    const uint16_t* wr = w + (size_t)(nt * 8 + g) * T::KP + kt * 16 + 2 * t;
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
                    mma_bf16(acc[nt], a0, a1, a2, a3, b0, b1);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const float2 bb = *reinterpret_cast<const float2*>(b + nt * 8 + 2 * t);
                r[nt][0] += tanh_fast(acc[nt][0] + bb.x);
                r[nt][1] += tanh_fast(acc[nt][1] + bb.y);
                r[nt][2] += tanh_fast(acc[nt][2] + bb.x);
                r[nt][3] += tanh_fast(acc[nt][3] + bb.y);
            }
            __syncthreads();        // all warps have consumed stage s before layer l+2 is prefetched into it
        }
        if (mp.y) {
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const int c = nt * 8 + 2 * t;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const long long row = row0 + 8 * h;
                    if (row < mp.rows) {
                        if (c + 1 < H) *reinterpret_cast<float2*>(mp.y + row * H + c) = make_float2(r[nt][2 * h], r[nt][2 * h + 1]);
                        else if (c < H) mp.y[row * H + c] = r[nt][2 * h];
                    }
                }
            }
        }
        if (mp.out) {
            // last layer: two 8-column tiles of outputs from the stream still in registers
            float o[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; nt++) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
            for (int kt = 0; kt < T::KT; kt++) {
                const uint32_t a0 = pack_bf16(r[2 * kt][0], r[2 * kt][1]);
                const uint32_t a1 = pack_bf16(r[2 * kt][2], r[2 * kt][3]);
                uint32_t a2 = 0u, a3 = 0u;
                if (2 * kt + 1 < NT) {
                    a2 = pack_bf16(r[2 * kt + 1][0], r[2 * kt + 1][1]);
                    a3 = pack_bf16(r[2 * kt + 1][2], r[2 * kt + 1][3]);
                }
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    const uint16_t* wr = s_wl + (size_t)(nt * 8 + g) * T::KP + kt * 16 + 2 * t;
                    mma_bf16(o[nt], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
                }
            }
            const int NOUT = mp.NOUT;
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = nt * 8 + 2 * t + (q & 1);
                    const long long row = row0 + 8 * (q >> 1);
                    if (c < NOUT && row < mp.rows) {
                        float v = o[nt][q] + s_bl[c];
                        if (mp.act == 1) v = 1.f / (1.f + __expf(-v));
                        else if (mp.act == 2) v = tanh_fast(v);
                        mp.out[row * NOUT + c] = v;
                    }
                }
            }
        }
    }
}

}  // namespace fastace
