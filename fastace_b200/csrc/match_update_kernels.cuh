// Economy::time_step for E independent economies — default path: two sm_100a kernels.
//
//   match_kernel<G>   (match_kernel.cuh) one warp per economy: ALL matching of the step (person phase, then
//                     firm phase) first-come-first-served in the visiting order, offer books in shared memory.
//   update_kernel<G>  (below) fully parallel element-wise work that depends on the matching result:
//                     one thread per person (apply purchases, consume, CES utility = reward);
//                     one lane per (firm, output good) (CES production, decode of the new goods /
//                     job offers, warp scan -> new books in market order).
//
// Reference code restated (paths under /root/reference/src): see step_kernel.cuh header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "match_kernel.cuh"

namespace fastace {

// ---- UtilMaxer::consume_goods (utilMaxer.cpp:88-92) + choose_goods_to_consume
//      (neuralPersonDecisionMaker.cpp:93-111) + UtilMaxer::u (utilMaxer.cpp:54-62)
// CES: both function families are the reference's default CES — the other families' code (accurate pow chains of
// Cobb-Douglas / Stone-Geary, theta loads) is compiled out of the specialised kernel
// FULL: the call is a full step (no phase flag): the trade-only / consume-only variants are compiled out
template <int G, bool CES, bool FULL>
__device__ __forceinline__ void update_person(const StepParams& p, const uint8_t* scr_pnh, const uint8_t* scr_pnb, int e, int pid) {
    const int util_kind = CES ? (int)FASTACE_FN_CES : p.util_kind;
    const int P = p.P;
    const size_t t = (size_t)e * P + pid;
    const double labor = kLaborPerOffer * (double)scr_pnh[t];       // exact: 0, 0.5 or 1.0
    if (!FULL && (p.flags & FASTACE_STEP_PERSONS_TRADE)) {
        // trade-only call: purchases and labour go into the state, consumption follows in its own call
        // (it touches nobody but the person, so deferring it changes nothing: utilMaxer.cpp:88-92)
#pragma unroll
        for (int g = 0; g < G; g++) {
            const size_t k = ((size_t)e * G + g) * P + pid;
            double v = p.st.p_inv[k];
            const int nb = scr_pnb[k];
            for (int q = 0; q < nb; q++) v += kAmountPerOffer;         // agent.cpp:109, one add per purchase
            p.st.p_inv[k] = v;
        }
        p.st.p_labor[t] = labor;
        return;
    }
    const bool applied = !FULL && (p.flags & FASTACE_STEP_PERSONS_CONSUME) != 0;   // purchases are already in p_inv
    double x[G + 1], inv[G];
    x[0] = 1 - labor;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const size_t k = ((size_t)e * G + g) * P + pid;
        double v = p.st.p_inv[k];
        const int nb = applied ? 0 : scr_pnb[k];
        for (int q = 0; q < nb; q++) v += kAmountPerOffer;             // agent.cpp:109, one add per purchase
        const double c = v * (double)p.ac.p_consume[k];                 // neuralPersonDecisionMaker.cpp:99
        x[g + 1] = c;
        inv[g] = v - c;                                                 // utilMaxer.cpp:91
    }
    double share[G + 1], theta[G + 1];
#pragma unroll
    for (int i = 0; i <= G; i++) {
        share[i] = p.st.p_util_share[((size_t)e * (G + 1) + i) * P + pid];
        theta[i] = (util_kind == FASTACE_FN_STONE_GEARY) ? p.st.p_util_theta[((size_t)e * (G + 1) + i) * P + pid] : 0.0;
    }
    p.out.p_reward[t] = eval_function<G + 1, true>(util_kind, p.st.p_util_tfp[t], share, theta, p.st.p_util_rho[t], x);
    p.st.p_labor[t] = labor;
#pragma unroll
    for (int g = 0; g < G; g++) p.st.p_inv[((size_t)e * G + g) * P + pid] = inv[g];
}

// ---- firms: produce (profitMaxer.cpp:68-72), sell_goods / search_for_laborers decode
//      (neuralFirmDecisionMaker.cpp:111-180), new books in market order (economy.cpp:52-59, 125-126); one warp
template <int G, bool CES>
__device__ __forceinline__ void update_firms(const StepParams& p, int e, int lane) {
    const int prod_kind = CES ? (int)FASTACE_FN_CES : p.prod_kind;
    const int F = p.F;
    const int cap = F * G;
    const size_t eF = (size_t)e * F, eCap = (size_t)e * cap;
    int base_m = 0, base_j = 0;
    // whole firms per pass (FPC firms x G goods <= 32 lanes), so that all lanes reading a firm's
    // inventories / labour do so before any lane of the same pass overwrites them
    const int FPC = 32 / G;
    for (int r0 = 0; r0 < F; r0 += FPC) {
        const int r = r0 + lane / G, g = lane % G;
        const bool active = lane < FPC * G && r < F;
        int lots = 0, jlots = 0, f = 0;
        double price = 0.0, jwage = 0.0, newinv = 0.0;
        if (active) {
            f = perm_firm_at(p, eF + r);
            double in[G + 1];
            in[0] = p.st.f_labor[eF + f];
            double xg = 0.0, invg = 0.0;
#pragma unroll
            for (int k = 0; k < G; k++) {
                const size_t a = ((size_t)e * G + k) * F + f;
                const double iv = p.st.f_inv[a];
                const double xk = iv * (double)p.ac.f_prod[a];              // neuralFirmDecisionMaker.cpp:101
                in[k + 1] = xk;
                if (k == g) { xg = xk; invg = iv; }
            }
            const size_t ag = ((size_t)e * G + g) * F + f;
            double share[G + 1], theta[G + 1];
#pragma unroll
            for (int i = 0; i <= G; i++) {
                const size_t k = (((size_t)e * G + g) * (G + 1) + i) * F + f;
                share[i] = p.st.f_prod_share[k];
                theta[i] = (prod_kind == FASTACE_FN_STONE_GEARY) ? p.st.f_prod_theta[k] : 0.0;
            }
            const double outg = eval_function<G + 1, false>(prod_kind, p.st.f_prod_tfp[ag], share, theta, p.st.f_prod_rho[ag], in);
            newinv = invg + (outg - xg);                                    // profitMaxer.cpp:71
            // decisionNetHandler.cpp:591 amounts = proportion * inventory; neuralFirmDecisionMaker.cpp:129
            const double amount = (double)p.ac.f_offer_amt[ag] * newinv;
            lots = x86_double_to_int(amount / kAmountPerOffer);
            price = (double)p.ac.f_offer_price[ag] / kAmountPerOffer;
            if (g == 0) {
                // neuralFirmDecisionMaker.cpp:164-176, decisionNetHandler.cpp:631-635
                double wage = (double)p.ac.f_job_wage[eF + f];
                if (wage > kLargeNumber) wage = kLargeNumber;
                jlots = x86_double_to_int((double)p.ac.f_job_labor[eF + f] / kLaborPerOffer);
                jwage = wage / kLaborPerOffer;
            }
        }
        __syncwarp();   // every lane has read its firm's inventories and labour before any is overwritten
        const unsigned mmask = __ballot_sync(0xffffffffu, active && lots > 0);
        const unsigned jmask = __ballot_sync(0xffffffffu, active && g == 0 && jlots > 0);
        if (active) {
            const size_t ag = ((size_t)e * G + g) * F + f;
            p.st.f_inv[ag] = newinv;
            if (g == 0) p.st.f_labor[eF + f] = 0.0;                          // firm.cpp:41
            if (lots > 0) {
                const int slot = base_m + __popc(mmask & ((1u << lane) - 1u));
                p.st.m_owner[eCap + slot] = f;
                p.st.m_good[eCap + slot] = g;
                p.st.m_left[eCap + slot] = (uint32_t)lots;
                p.st.m_taken[eCap + slot] = 0;
                p.st.m_price[eCap + slot] = price;
            }
            if (g == 0 && jlots > 0) {
                const int slot = base_j + __popc(jmask & ((1u << lane) - 1u));
                p.st.j_owner[eF + slot] = f;
                p.st.j_left[eF + slot] = (uint32_t)jlots;
                p.st.j_taken[eF + slot] = 0;
                p.st.j_wage[eF + slot] = jwage;
            }
        }
        base_m += __popc(mmask);
        base_j += __popc(jmask);
    }
    if (lane == 0) {
        p.st.m_count[e] = base_m;
        p.st.j_count[e] = base_j;
    }
}

// -------------------------------------------------------------------------------------------------
// update_kernel, two work assignments:
//   * whole-grid (phase-wise calls, `done_list` null): blocks [0, firm_blocks) handle firms (one warp per economy,
//     lanes = (visiting rank, output good) pairs; scheduled first because their fp64 pow chains are the longest), the
//     remaining blocks handle persons (one thread each); everything waits for match_kernel to complete.
//   * completion queue (full steps): the same firm warps and person threads, but assigned to queue slots instead of
//     economies — slot k is the k-th economy match_kernel finishes — and each waits only for its own slot
//     (match_kernel.cuh: completion queue).
struct UpdateParams {
    StepParams sp;
    const uint8_t* scr_pnh;
    const uint8_t* scr_pnb;
    int firm_blocks;
    const uint32_t* done_list;
    uint32_t done_tag;
    int group_person_blocks;         // queue mode: person blocks per group of kQueueGroup slots
    volatile uint32_t* dev_err;
};

constexpr int kQueueGroup = 32;

constexpr int kUpdateThreads = 128;
constexpr unsigned kQueuePollNs = 128;
constexpr uint32_t kQueuePollCap = 1u << 23;     // x 128 ns: about a second, then kDevErrQueue

// QUEUE: the instance of the full step (completion queue, no phase flags); otherwise the whole-grid assignment
template <int G, bool CES = false, bool QUEUE = false>
__global__ void __launch_bounds__(kUpdateThreads, 8) update_kernel(const UpdateParams up) {
    const StepParams& p = up.sp;
    const int P = p.P, F = p.F;
    const int lane = threadIdx.x & 31;
    int e = 0, pid = 0, pid_end = 0, pid_step = 1;
    bool firms_here = false;
    if (QUEUE) {
        grid_launch_dependents();
        // blocks come in groups that consume kQueueGroup consecutive queue slots: first the firm blocks (a warp per
        // slot), then the person blocks (a thread per person of those slots), so that blocks are dispatched in the
        // order their economies finish
        const int firm_blocks = kQueueGroup / (kUpdateThreads / 32);
        const int per_group = firm_blocks + up.group_person_blocks;
        const int grp = (int)blockIdx.x / per_group, idx = (int)blockIdx.x % per_group;
        int slot;
        bool work = true;
        if (idx < firm_blocks) {
            slot = grp * kQueueGroup + idx * (kUpdateThreads / 32) + (int)(threadIdx.x >> 5);
            firms_here = true;
        } else {
            const int v = (idx - firm_blocks) * kUpdateThreads + (int)threadIdx.x;
            work = v < kQueueGroup * P;
            slot = grp * kQueueGroup + (work ? v / P : 0);
            pid = work ? v % P : 0;
        }
        work = work && slot < p.E;
        // every thread polls the slot it needs (the lanes of a warp read one or two words); the previous step's update
        // has completed before any warp of match_kernel took a ticket, so a matching tag is this step's entry
        uint32_t v = 0, polls = 0;
        while (work) {
            v = load_relaxed_u32(up.done_list + slot);
            if ((v >> kQueueTagShift) == up.done_tag) break;
            if (++polls >= kQueuePollCap) { up.dev_err[kDevErrQueue] = 1u; work = false; }
            else backoff_ns(kQueuePollNs);
        }
        fence_gpu();                          // acquire: the economy's matching results are visible from here on
        e = (int)(v & kQueueEconMask);
        firms_here = firms_here && work;      // warp-uniform: the lanes of a firm warp share one slot
        pid_end = (work && !firms_here && idx >= firm_blocks) ? pid + 1 : pid;
    } else {
        grid_dependency_wait();        // the matching results of this step (programmatic dependent launch)
        grid_launch_dependents();
        if ((int)blockIdx.x >= up.firm_blocks) {
            const size_t t = (size_t)((int)blockIdx.x - up.firm_blocks) * kUpdateThreads + threadIdx.x;
            if (t >= (size_t)p.E * P) return;
            e = (int)(t / P); pid = (int)(t % P); pid_end = pid + 1;
        } else {
            e = (int)blockIdx.x * (kUpdateThreads / 32) + (int)(threadIdx.x >> 5);
            if (e >= p.E) return;
            firms_here = true;
        }
    }
    if (firms_here) update_firms<G, CES>(p, e, lane);
    for (; pid < pid_end; pid += pid_step) update_person<G, CES, QUEUE>(p, up.scr_pnh, up.scr_pnb, e, pid);
    // queue mode: the block of the last economy to finish keeps this grid open until match_kernel has completed as a
    // grid, so that whatever follows in the stream is ordered after both kernels
    if (QUEUE && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) grid_dependency_wait();
}

}  // namespace fastace
