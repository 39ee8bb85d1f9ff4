// TEST INFRASTRUCTURE ONLY — a SIMT emulator that runs the CUDA kernels' *source* on the CPU.
//
// The step kernels of fastace_b200/csrc are plain C++ plus warp collectives.  This header gives g++
// the handful of device intrinsics they use and runs every CUDA thread of a block as a ucontext fiber:
// a fiber runs until it reaches a warp collective (__ballot_sync, __shfl_*_sync, __syncwarp, ...) or
// __syncthreads, parks there, and is released when every live thread of its warp (block) has arrived —
// the semantics of full-mask collectives in converged code, which is all the kernels use.
// It exists so that the kernel logic can be checked against the oracle in this GPU-less container
// (tests/test_emu_kernels.py); nothing in the product package includes it.
#pragma once
#include <cuda_runtime.h>   // vector types, host-side declarations only (g++)
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace emu {

struct Thread {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    int waiting = 0;          // 0 runnable, 1 parked at a warp collective, 2 parked at __syncthreads
    uint3 tid{0, 0, 0};
};

struct Warp {
    uint64_t val[2][32];
    int arrived = 0;
    int epoch = 0;
};

struct Block {
    std::vector<Thread> threads;
    std::vector<Warp> warps;
    std::vector<unsigned char> smem;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    int cur = -1;
    int barrier_arrived = 0;
    ucontext_t sched;
    std::function<void()> body;
};

inline Block* g_block = nullptr;            // the emulator is single-threaded
inline unsigned char* g_smem = nullptr;
inline size_t g_smem_size = 0;
inline Block*& cur_block() { return g_block; }

inline uint3 tid() { Block* b = cur_block(); return b->threads[b->cur].tid; }
inline uint3 bid() { return cur_block()->bid; }
inline dim3 bdim() { return cur_block()->bdim; }
inline dim3 gdim() { return cur_block()->gdim; }
inline unsigned char* dyn_smem() { return g_smem; }

inline void yield_to_scheduler() {
    Block* b = cur_block();
    swapcontext(&b->threads[b->cur].ctx, &b->sched);
}

inline int live_in_warp(Block* b, int w) {
    int n = 0;
    const int lo = w * 32, hi = std::min<int>(lo + 32, (int)b->threads.size());
    for (int t = lo; t < hi; t++) n += !b->threads[t].done;
    return n;
}

// All live threads of the warp exchange one 64-bit value; returns the buffer of this collective.
inline const uint64_t* warp_exchange(uint64_t mine) {
    Block* b = cur_block();
    const int t = b->cur, w = t / 32, lane = t % 32;
    Warp& W = b->warps[w];
    const int e = W.epoch & 1;
    W.val[e][lane] = mine;
    W.arrived++;
    b->threads[t].waiting = 1;
    yield_to_scheduler();
    return W.val[e];
}

inline unsigned live_mask() {
    Block* b = cur_block();
    const int w = b->cur / 32;
    unsigned m = 0;
    for (int l = 0; l < 32; l++) {
        const int t = w * 32 + l;
        if (t < (int)b->threads.size() && !b->threads[t].done) m |= 1u << l;
    }
    return m;
}

inline void trampoline() {
    Block* b = cur_block();
    b->body();
    b->threads[b->cur].done = true;
    swapcontext(&b->threads[b->cur].ctx, &b->sched);
}

inline void run_block(Block& b) {
    cur_block() = &b;
    g_smem = b.smem.data();
    g_smem_size = b.smem.size();
    const int n = (int)b.threads.size();
    for (int t = 0; t < n; t++) {
        Thread& T = b.threads[t];
        T.stack.resize(256 * 1024);
        getcontext(&T.ctx);
        T.ctx.uc_stack.ss_sp = T.stack.data();
        T.ctx.uc_stack.ss_size = T.stack.size();
        T.ctx.uc_link = &b.sched;
        makecontext(&T.ctx, (void (*)())trampoline, 0);
    }
    int remaining = n;
    while (remaining > 0) {
        bool progress = false;
        for (int t = 0; t < n; t++) {
            Thread& T = b.threads[t];
            if (T.done || T.waiting) continue;
            b.cur = t;
            swapcontext(&b.sched, &T.ctx);
            progress = true;
            if (T.done) remaining--;
        }
        // release warps / the block whose live threads have all arrived
        for (size_t w = 0; w < b.warps.size(); w++) {
            Warp& W = b.warps[w];
            const int live = live_in_warp(&b, (int)w);
            if (W.arrived > 0 && W.arrived == live) {
                W.arrived = 0;
                W.epoch++;
                for (int l = 0; l < 32; l++) {
                    const int t = (int)w * 32 + l;
                    if (t < n && b.threads[t].waiting == 1) b.threads[t].waiting = 0;
                }
                progress = true;
            }
        }
        {
            int live = 0;
            for (auto& T : b.threads) live += !T.done;
            if (b.barrier_arrived > 0 && b.barrier_arrived == live) {
                b.barrier_arrived = 0;
                for (auto& T : b.threads) if (T.waiting == 2) T.waiting = 0;
                progress = true;
            }
        }
        if (!progress) {
            std::fprintf(stderr, "warp_emu: deadlock (divergent collective?) in block %u\n", b.bid.x);
            std::abort();
        }
    }
    cur_block() = nullptr;
}

// order in which the blocks of a launch run: 0 ascending, 1 descending, 2 a fixed scramble (tests of code whose result
// must not depend on which block finishes first, e.g. the completion queue between match_kernel and update_kernel)
inline int& block_order() { static int o = 0; return o; }

// kernel<<<grid, block, smem>>>(params): blocks run one after the other
template <typename Kernel, typename Params>
void launch(Kernel kernel, unsigned grid, unsigned block, size_t smem_bytes, const Params& params) {
    for (unsigned k = 0; k < grid; k++) {
        unsigned bx = k;
        if (block_order() == 1) bx = grid - 1 - k;
        else if (block_order() == 2) {          // k -> k * odd + offset (mod the next power of two), skipping values >= grid
            unsigned n = 1; while (n < grid) n <<= 1;
            static thread_local unsigned cursor; if (k == 0) cursor = 0;
            do { bx = (cursor * 0x9E3779B1u + 12345u) & (n - 1); cursor++; } while (bx >= grid);
        }
        Block b;
        b.threads.resize(block);
        b.warps.resize((block + 31) / 32);
        b.smem.assign(smem_bytes + 16, 0xCD);   // poison: reads of unwritten shared memory show up
        b.bid = uint3{bx, 0, 0};
        b.bdim = dim3(block, 1, 1);
        b.gdim = dim3(grid, 1, 1);
        for (unsigned t = 0; t < block; t++) b.threads[t].tid = uint3{t, 0, 0};
        b.body = [&]() { kernel(params); };
        run_block(b);
    }
}

template <typename Kernel, typename P1, typename P2>
void launch(Kernel kernel, unsigned grid, unsigned block, size_t smem_bytes, const P1& p1, const P2& p2) {
    struct Both { const P1* a; const P2* b; } both{&p1, &p2};
    auto call = [kernel](const Both& x) { kernel(*x.a, *x.b); };
    launch(call, grid, block, smem_bytes, both);
}

template <typename T> inline uint64_t to_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T from_bits(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

}  // namespace emu

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())
#define FASTACE_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
// the explicit shared-memory accessors of common.cuh over the emulated block buffer (address = offset in it)
#define FASTACE_HAVE_SMEM_OPS
namespace fastace {
inline uint32_t smem_addr(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - emu::dyn_smem()); }
inline uint32_t keep_u32(uint32_t x) { return x; }
// every explicit shared-memory access is bounds- and alignment-checked against the block's dynamic allocation
inline void emu_check(uint32_t a, size_t n) {
    if ((size_t)a + n > emu::g_smem_size || (a % n) != 0) {
        std::fprintf(stderr, "emulator: shared-memory access of %zu bytes at offset %u outside the %zu-byte allocation (or misaligned)\n", n, a, emu::g_smem_size);
        std::abort();
    }
}
template <typename T> inline T emu_ld(uint32_t a) { emu_check(a, sizeof(T)); T v; std::memcpy(&v, emu::dyn_smem() + a, sizeof(T)); return v; }
template <typename T> inline void emu_st(uint32_t a, T v) { emu_check(a, sizeof(T)); std::memcpy(emu::dyn_smem() + a, &v, sizeof(T)); }
template <int OFF = 0> inline uint32_t lds_u8(uint32_t a) { return emu_ld<uint8_t>(a + OFF); }
template <int OFF = 0> inline uint32_t lds_u16(uint32_t a) { return emu_ld<uint16_t>(a + OFF); }
template <int OFF = 0> inline uint32_t lds_u32(uint32_t a) { return emu_ld<uint32_t>(a + OFF); }
template <int OFF = 0> inline double lds_f64(uint32_t a) { return emu_ld<double>(a + OFF); }
template <int OFF = 0> inline uint4 lds_v4(uint32_t a) { return emu_ld<uint4>(a + OFF); }
template <int OFF = 0> inline void sts_u8(uint32_t a, uint32_t v) { emu_st<uint8_t>(a + OFF, (uint8_t)v); }
template <int OFF = 0> inline void sts_u16(uint32_t a, uint32_t v) { emu_st<uint16_t>(a + OFF, (uint16_t)v); }
template <int OFF = 0> inline void sts_u32(uint32_t a, uint32_t v) { emu_st<uint32_t>(a + OFF, v); }
template <int OFF = 0> inline void sts_f64(uint32_t a, double v) { emu_st<double>(a + OFF, v); }
template <int OFF = 0> inline void sts_v4(uint32_t a, uint4 v) { emu_st<uint4>(a + OFF, v); }
inline void reds_or_u32(uint32_t a, uint32_t v) { emu_st<uint32_t>(a, emu_ld<uint32_t>(a) | v); }
inline void prefetch_l1(const void*) {}
}
namespace emu { inline unsigned long long* stats() { static unsigned long long s[16] = {}; return s; } }
#define FASTACE_STAT(which, n) (emu::stats()[which] += (n))

inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
inline void __syncthreads() {
    emu::Block* b = emu::cur_block();
    b->barrier_arrived++;
    b->threads[b->cur].waiting = 2;
    emu::yield_to_scheduler();
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned live = emu::live_mask();
    const uint64_t* v = emu::warp_exchange(pred ? 1 : 0);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && v[l]) m |= 1u << l;
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { const unsigned live = emu::live_mask(); return __ballot_sync(m, pred) == live; }
inline unsigned __activemask() { return emu::live_mask(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = (int)(emu::tid().x % 32);
    const uint64_t* b = emu::warp_exchange(emu::to_bits(v));
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return emu::from_bits<T>(b[s]);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = (int)(emu::tid().x % 32);
    const uint64_t* b = emu::warp_exchange(emu::to_bits(v));
    const int s = lane - (int)delta;
    return (s < (lane & ~(width - 1))) ? v : emu::from_bits<T>(b[s]);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = (int)(emu::tid().x % 32);
    const uint64_t* b = emu::warp_exchange(emu::to_bits(v));
    const int s = lane + (int)delta;
    return (s > (lane | (width - 1))) ? v : emu::from_bits<T>(b[s]);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    const int lane = (int)(emu::tid().x % 32);
    const uint64_t* b = emu::warp_exchange(emu::to_bits(v));
    (void)width;
    return emu::from_bits<T>(b[lane ^ mask]);
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    const unsigned live = emu::live_mask();
    const int lane = (int)(emu::tid().x % 32);
    const uint64_t* b = emu::warp_exchange(v);
    const uint64_t mine = b[lane];
    unsigned m = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && b[l] == mine) m |= 1u << l;
    return m;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    const unsigned live = emu::live_mask();
    const uint64_t* b = emu::warp_exchange(v);
    unsigned s = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) s += (unsigned)b[l];
    return s;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    const unsigned live = emu::live_mask();
    const uint64_t* b = emu::warp_exchange(v);
    unsigned s = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) s |= (unsigned)b[l];
    return s;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const unsigned live = emu::live_mask();
    const uint64_t* b = emu::warp_exchange(v);
    unsigned s = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) s = std::max(s, (unsigned)b[l]);
    return s;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xFF);
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline int __double2hiint(double x) { return (int)(emu::to_bits(x) >> 32); }
inline int __double2loint(double x) { return (int)(emu::to_bits(x) & 0xFFFFFFFFu); }
inline double __hiloint2double(int hi, int lo) { return emu::from_bits<double>(((uint64_t)(unsigned)hi << 32) | (unsigned)lo); }
inline int __double2int_rz(double x) { return (int)x; }
inline long long __double_as_longlong(double x) { return (long long)emu::to_bits(x); }
inline double __longlong_as_double(long long x) { return emu::from_bits<double>((uint64_t)x); }
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }

// CUDA's global-namespace min / max overloads (mixed signedness promotes like the device ones)
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline double min(double a, double b) { return std::fmin(a, b); }
inline double max(double a, double b) { return std::fmax(a, b); }
