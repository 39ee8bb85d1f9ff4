// Shared device helpers of the fastACE step kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fastace_b200.h"

// dynamic shared memory of a kernel (tests/emu/warp_emu.h, which runs this source on the CPU for the
// parity tests of the GPU-less container, substitutes its own block buffer)
#ifndef FASTACE_DYN_SMEM
#define FASTACE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// Programmatic dependent launch (sm_90+): a kernel launched with programmatic stream serialization may start while
// its predecessor in the stream is still draining; it must not touch the predecessor's outputs before
// grid_dependency_wait(), and the predecessor lets it be scheduled early with grid_launch_dependents().
#ifndef FASTACE_HAVE_SMEM_OPS
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Publishing a finished unit of work to a kernel that runs concurrently (the completion queue between match_kernel
// and update_kernel), with ONE fence on either side: writes -> fence.acq_rel.gpu -> strong (relaxed) store of the flag;
// strong (relaxed) polling loads of the flag -> fence.acq_rel.gpu -> reads.
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void store_relaxed_u32(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t load_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ticket_add(uint32_t* p) { return atomicAdd(p, 1u); }
__device__ __forceinline__ void backoff_ns(unsigned ns) { __nanosleep(ns); }
#else
inline void grid_dependency_wait() {}
inline void grid_launch_dependents() {}
inline void fence_gpu() {}
inline void store_relaxed_u32(uint32_t* p, uint32_t v) { *p = v; }
inline uint32_t load_relaxed_u32(const uint32_t* p) { return *p; }
inline uint32_t ticket_add(uint32_t* p) { return (*p)++; }
inline void backoff_ns(unsigned) {}
#endif

// event counters of the CPU emulation build (how many rounds / re-scans / slow paths a workload takes); nothing
// on the device
#ifndef FASTACE_STAT
#define FASTACE_STAT(which, n)
#endif

namespace fastace {

// Shared-memory accesses of the hot loops by explicit 32-bit shared address + immediate offset: one LDS / STS each,
// no generic-pointer arithmetic for the compiler to re-derive under the register cap.  (The CPU emulation build
// supplies the same functions over its own block buffer.)
#ifndef FASTACE_HAVE_SMEM_OPS
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// a value the compiler must keep in a register instead of re-deriving it from its operands at every use
__device__ __forceinline__ uint32_t keep_u32(uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
#define FASTACE_SMEM_LD(NAME, T, PTX, C)                                                                   \
    template <int OFF = 0> __device__ __forceinline__ T NAME(uint32_t a) {                                 \
        T v; asm volatile("ld.shared." PTX " %0, [%1+%2];" : "=" C(v) : "r"(a), "n"(OFF)); return v;       \
    }
#define FASTACE_SMEM_ST(NAME, T, PTX, C)                                                                   \
    template <int OFF = 0> __device__ __forceinline__ void NAME(uint32_t a, T v) {                         \
        asm volatile("st.shared." PTX " [%0+%1], %2;" :: "r"(a), "n"(OFF), C(v));                          \
    }
FASTACE_SMEM_LD(lds_u8, uint32_t, "u8", "r")
FASTACE_SMEM_LD(lds_u16, uint32_t, "u16", "r")
FASTACE_SMEM_LD(lds_u32, uint32_t, "u32", "r")
FASTACE_SMEM_LD(lds_f64, double, "f64", "d")
FASTACE_SMEM_ST(sts_u8, uint32_t, "u8", "r")
FASTACE_SMEM_ST(sts_u16, uint32_t, "u16", "r")
FASTACE_SMEM_ST(sts_u32, uint32_t, "u32", "r")
FASTACE_SMEM_ST(sts_f64, double, "f64", "d")
#undef FASTACE_SMEM_LD
#undef FASTACE_SMEM_ST
template <int OFF = 0> __device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF)); return v;
}
template <int OFF = 0> __device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" :: "r"(a), "n"(OFF), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void reds_or_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// hint: bring the line that holds *p towards this SM (no register, no dependency)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#endif

enum { kStatWindows, kStatRounds, kStatRescans, kStatRiskyWalks, kStatSalesWindows, kStatDeadExits, kStatFirmSerial, kStatRoundsW0, kStatRoundsW1, kStatRoundsW2, kStatRoundsW3, kStatRescansW0, kStatRiskyCoop, kStatRiskyKept, kStatCount };

struct StepParams {
    int E, P, F, S;
    uint32_t flags;
    uint32_t time_before;
    int util_kind, prod_kind;       // fastace_function_kind_t of the persons' utility / the firms' production
    int compact;                    // 1: decisions come from `cz` (index/take/order arrays) + the float arrays of `ac`
    fastace_state_t st;
    fastace_actions_t ac;           // continuous actions are always read from here
    fastace_actions_compact_t cz;
    fastace_step_out_t out;
};

__device__ __forceinline__ int perm_firm_at(const StepParams& p, size_t k) {
    return p.compact ? (int)p.cz.perm_firm[k] : p.ac.perm_firm[k];
}

constexpr double kEps = 1e-8;            // constants::eps (base/constants.h:9)
constexpr double kLargeNumber = 1e8;     // constants::largeNumber (base/constants.h:10)
constexpr double kAmountPerOffer = 1.0;  // neural/neuralFirmDecisionMaker.cpp:6
constexpr double kLaborPerOffer = 0.5;   // neural/neuralFirmDecisionMaker.cpp:7
constexpr int kNone = 0xFF;
constexpr int kMaxStack = FASTACE_MAX_STACK;

// (int)double as the x86-64 reference binary does it (cvttsd2si): out-of-range and NaN
// give INT_MIN, which fails the `numOffers > 0` tests.
__device__ __forceinline__ int x86_double_to_int(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(x);
}

// Index mapping (decisionNetHandler.cpp:327-365 draws indices in [0,count)).  count <= 254 here, so
// the modulo of a 32-bit draw is done as two exact 32-bit "fastmod" steps (Lemire): with
// M = floor(2^32 / c) + 1, (x mod c) = mulhi(M * x mod 2^32, c) for every x < 2^24.
struct IndexMap {
    uint32_t magic, count, k16;   // k16 = 65536 mod count
    bool modulo;
    __device__ __forceinline__ IndexMap(int cnt, uint32_t flags) : IndexMap(cnt, (flags & FASTACE_IDX_MODULO) != 0) {}
    // (the reciprocal and the remainder are only computed for the modulo mapping)
    __device__ __forceinline__ IndexMap(int cnt, bool mod)
        : magic(mod && cnt > 0 ? 0xFFFFFFFFu / (uint32_t)cnt + 1u : 0u), count(cnt > 0 ? (uint32_t)cnt : 0u),
          k16(mod && cnt > 0 ? 65536u % (uint32_t)cnt : 0u), modulo(mod) {}
    __device__ __forceinline__ uint32_t mod24(uint32_t x) const { return __umulhi(magic * x, count); }
    // precondition: count > 0 (callers skip mapping for an empty book)
    __device__ __forceinline__ int operator()(int raw) const {
        const uint32_t u = (uint32_t)raw;
        if (modulo) return (int)mod24(mod24(u >> 16) * k16 + (u & 0xFFFFu));
        return (u >= count) ? kNone : raw;
    }
};

// pow for the reward path (utilities are rewards: 1e-5 relative tolerance, they never feed back into state).
// x^y = exp(y * log x) in fp64 with short polynomials: log by 2*atanh((m-1)/(m+1)) on m in [sqrt(1/2), sqrt(2)) up to
// s^13 (truncation < 2e-12 relative), exp by a degree-10 Taylor polynomial on |f| <= ln2/2 (< 1e-12); measured
// against libm pow: <= 2e-11 relative over x in [1e-9, 1e9], |y| <= 30, |y ln x| < 690 (tests/test_emu_kernels.py).  ~45 fp64
// instructions instead of ~90 for exp(y*log(x)) with the library routines, which remain the path for anything that
// is not a positive normal number or whose exponent leaves [-700, 700].
__device__ __forceinline__ double fast_rcp(double d) {
#ifndef FASTACE_HAVE_SMEM_OPS
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));     // ~2^-23 relative
    double t = fma(-d, r, 1.0);
    r = fma(r, t, r);                                          // 2^-46
    t = fma(-d, r, 1.0);
    return fma(r, t, r);                                       // full precision
#else
    return 1.0 / d;
#endif
}
__device__ __forceinline__ double pow_reward(double x, double y) {
    if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return exp(y * log(x));
    const long long b = __double_as_longlong(x);
    int e = (int)(b >> 52) - 1023;
    double m = __longlong_as_double((b & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);      // [1, 2)
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double s = (m - 1.0) * fast_rcp(m + 1.0);
    const double s2 = s * s;
    double p = 1.0 / 13.0;
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    p = fma(p, s2, 1.0);
    const double lx = fma((double)e, 0.6931471805599453, 2.0 * s * p);                        // log x
    const double z = y * lx;
    if (!(fabs(z) < 700.0)) return exp(z);
    const double kf = rint(z * 1.4426950408889634);
    double f = fma(kf, -0.6931471803691238, z);                // ln2 split: hi has 32 significant bits
    f = fma(kf, -1.9082149292705877e-10, f);
    double q = 1.0 / 3628800.0;
    q = fma(q, f, 1.0 / 362880.0);
    q = fma(q, f, 1.0 / 40320.0);
    q = fma(q, f, 1.0 / 5040.0);
    q = fma(q, f, 1.0 / 720.0);
    q = fma(q, f, 1.0 / 120.0);
    q = fma(q, f, 1.0 / 24.0);
    q = fma(q, f, 1.0 / 6.0);
    q = fma(q, f, 0.5);
    q = fma(q, f, 1.0);
    q = fma(q, f, 1.0);
    return __longlong_as_double(__double_as_longlong(q) + ((long long)(int)kf << 52));       // q * 2^k, result normal
}

// One goods request by a buyer whose money/inventory live at (money, inv[g*istride]).
// Agent::respond_to_offer -> review_offer_response -> accept_offer_response
// (base/agent.cpp:99-161), fp64 updates in the reference's order.

// VecToScalar::f of a function family over N inputs (functions/vecToScalar.cpp:30-32, 45-47, 67-69, 80-82,
// 112-118), fp64, in the reference's operation order.  kReward selects the cheaper exp(y*log x) power for
// CES utilities (reward only, 1e-5 tolerance); every other use takes the accurate fp64 pow.
template <int N, bool kReward>
__device__ __forceinline__ double eval_function(int kind, double tfp, const double (&share)[N], const double (&theta)[N],
                                                double rho, const double (&x)[N]) {
    if (kind == FASTACE_FN_CES) {
        double inner = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) inner += share[i] * (kReward ? pow_reward(x[i] + kEps, rho) : pow(x[i] + kEps, rho));
        return tfp * (kReward ? pow_reward(inner, 1 / rho) : pow(inner, 1 / rho));
    }
    if (kind == FASTACE_FN_LINEAR) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) s += share[i] * x[i];
        return s;
    }
    if (kind == FASTACE_FN_LEONTIEF) {
        double m = x[0] * share[0];
#pragma unroll
        for (int i = 1; i < N; i++) { const double v = x[i] * share[i]; if (v < m) m = v; }
        return m;
    }
    // Cobb-Douglas / Stone-Geary
    double p = 1.0;
#pragma unroll
    for (int i = 0; i < N; i++) p *= pow(kind == FASTACE_FN_STONE_GEARY ? x[i] - theta[i] : x[i], share[i]);
    return tfp * p;
}

// floor(x) as a lot count: how many times `inventory >= 1.0; inventory -= 1.0` succeeds
// (agent.cpp:140,156; the subtraction is exact for x < 2^53).  NaN never compares "short".
__device__ __forceinline__ uint32_t unit_sales_possible(double x) {
    if (x < 1.0) return 0u;
    if (!(x < 4294967296.0)) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

// optional outputs: per-request success flags of one person (jobs in bits 0..15, goods in 16..31)
__device__ __forceinline__ void write_person_ok(const StepParams& p, int e, int pid, uint32_t okm) {
    if (p.out.p_job_ok == nullptr && p.out.p_good_ok == nullptr) return;
    for (int i = 0; i < p.S; i++) {
        const size_t k = ((size_t)e * p.S + i) * p.P + pid;
        if (p.out.p_job_ok) p.out.p_job_ok[k] = (okm >> i) & 1u;
        if (p.out.p_good_ok) p.out.p_good_ok[k] = (okm >> (16 + i)) & 1u;
    }
}

}  // namespace fastace
