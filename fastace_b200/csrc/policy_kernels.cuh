// Sampling rules of the decision nets as two element-wise kernels (rollout-only fast path of fastace_b200/policy.py;
// the differentiable path stays in torch).  Reference: DecisionNetHandler's sample_normal / sample_logitNormal /
// sample_logNormal (/root/reference/src/neural/decisionNetHandler.cpp:27-46) and the Bernoulli takes with their
// log-probabilities (:368-387, 398-403, 476-480).  Inputs are agent-major rows as the nets emit them; outputs go
// straight into the env's action layout ([E][slot | good][agent]), so no permute / cast kernels follow.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fastace {

struct BernoulliParams {
    const float* probas;      // [E*A][S]
    const float* uniforms;    // [E*A][S]
    const int64_t* idx;       // [E*A][S]  offer indices of the draws
    const uint8_t* valid;     // [E]       the economy's book is not empty
    int E, A, S;
    int invalid_nan;          // log-probability of an agent facing an empty book: 1 = NaN ("no decision"), 0 = 0.0
    int32_t* out_idx;         // [E][S][A]
    uint8_t* out_take;        // [E][S][A]
    float* out_logp;          // [E][A]
};

// take_i = u_i < p_i ; log pi = sum_i log p_i or log(1 - p_i)
__global__ void __launch_bounds__(256) policy_bernoulli_kernel(const BernoulliParams q) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)q.E * q.A) return;
    const int e = (int)(t / q.A), a = (int)(t - (long long)e * q.A);
    const bool valid = q.valid[e] != 0;
    const float* p = q.probas + t * q.S;
    const float* u = q.uniforms + t * q.S;
    const int64_t* ix = q.idx + t * q.S;
    float lp = 0.f;
    for (int s = 0; s < q.S; s++) {
        const float ps = p[s];
        const bool take = u[s] < ps;
        lp += take ? logf(ps) : logf(1.f - ps);
        const size_t o = ((size_t)e * q.S + s) * q.A + a;
        q.out_take[o] = (take && valid) ? 1 : 0;
        q.out_idx[o] = (int32_t)ix[s];
    }
    q.out_logp[t] = valid ? lp : (q.invalid_nan ? __int_as_float(0x7fc00000) : 0.f);
}

struct NormalParams {
    const float* params;      // [E*A][C] groups of `stride` floats; (mu, log sigma) at `offset`, `offset + 1`
    const float* noise;       // [E*A][C] standard normals
    int stride, offset;
    int E, A, C;
    int kind;                 // 0: logit-normal (sigmoid), 1: log-normal (exp)
    int accumulate;           // add the log-density to out_logp instead of storing it
    float* out_x;             // [E][C][A]
    float* out_logp;          // [E][A]
    float log_sqrt2pi_scale;  // SQRT2PI of neuralConstants.h:10, passed by the host so that both paths use one value
};

__global__ void __launch_bounds__(256) policy_normal_kernel(const NormalParams q) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)q.E * q.A) return;
    const int e = (int)(t / q.A), a = (int)(t - (long long)e * q.A);
    float lp = 0.f;
    for (int c = 0; c < q.C; c++) {
        const float* pr = q.params + ((size_t)t * q.C + c) * q.stride + q.offset;
        const float mu = pr[0], sigma = expf(pr[1]);
        const float x = q.noise[(size_t)t * q.C + c] * sigma + mu;
        const float z = (x - mu) / sigma;
        lp += -0.5f * (z * z) - logf(sigma * q.log_sqrt2pi_scale);
        q.out_x[((size_t)e * q.C + c) * q.A + a] = q.kind == 0 ? 1.f / (1.f + expf(-x)) : expf(x);
    }
    q.out_logp[t] = q.accumulate ? q.out_logp[t] + lp : lp;
}

}  // namespace fastace
