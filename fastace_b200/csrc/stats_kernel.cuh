// Per-step market statistics as a device reduction: what the reference's `run` prints after every step
// (print_info / print_offer_info / print_jobOffer_info, /root/reference/src/pybindings.cpp:20-75) for every economy of the
// env, without copying the books to the host.  One thread per economy walks its two books in market order, so the
// floating-point sums are accumulated in exactly the order of the reference's loops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

struct StatsParams {
    int E, F, G;
    fastace_state_t st;
    fastace_market_stats_t out;
};

__global__ void market_stats_kernel(const StatsParams sp) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sp.E) return;
    const int F = sp.F, G = sp.G, cap = F * G;
    const size_t eCap = (size_t)e * cap, eF = (size_t)e * F;
    double sum_price[FASTACE_MAX_GOODS];
    uint32_t count[FASTACE_MAX_GOODS], lots[FASTACE_MAX_GOODS];
    for (int g = 0; g < G; g++) { sum_price[g] = 0.0; count[g] = 0; lots[g] = 0; }
    const int nm = sp.st.m_count[e];
    for (int n = 0; n < nm; n++) {
        const int good = sp.st.m_good[eCap + n];
        for (int g = 0; g < G; g++)                                   // an Offer of the shipped plugin: one unit of one good
            if (g == good) {
                sum_price[g] += kAmountPerOffer / sp.st.m_price[eCap + n];   // quantities(i) / price, pybindings.cpp:31
                count[g] += 1;
                lots[g] += sp.st.m_left[eCap + n];
            }
    }
    for (int g = 0; g < G; g++) {
        if (sp.out.sum_quantity_per_price) sp.out.sum_quantity_per_price[(size_t)e * G + g] = sum_price[g];
        if (sp.out.offers) sp.out.offers[(size_t)e * G + g] = count[g];
        if (sp.out.lots) sp.out.lots[(size_t)e * G + g] = lots[g];
    }
    const int nj = sp.st.j_count[e];
    double sum_wage = 0.0;
    uint32_t jlots = 0;
    for (int n = 0; n < nj; n++) {
        sum_wage += sp.st.j_wage[eF + n] / kLaborPerOffer;            // wage / labor, pybindings.cpp:52
        jlots += sp.st.j_left[eF + n];
    }
    if (sp.out.sum_wage_per_labor) sp.out.sum_wage_per_labor[e] = sum_wage;
    if (sp.out.job_offers) sp.out.job_offers[e] = (uint32_t)nj;
    if (sp.out.job_lots) sp.out.job_lots[e] = jlots;
}

}  // namespace fastace
