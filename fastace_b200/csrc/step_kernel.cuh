// Economy::time_step for E independent economies — sm_100a kernel.
//
// Mapping: ONE WARP PER ECONOMY (one 32-thread CTA each; 4096 economies = 4096 CTAs, all
// co-resident on 148 SMs at <= 32 CTAs/SM).  The economy's offer books, firm state and the
// persons' compacted request lists are staged in shared memory for the whole step.
//
//   phase 0  lanes own agents / book slots: coalesced loads of state + actions from HBM,
//            index mapping, request lists -> smem
//   phase 1  PERSON MATCHING: lane 0 walks persons in visiting order and resolves every
//            job / goods request first-come-first-served against the smem books, applying
//            each fp64 update in exactly the reference's order (bit-exact by construction)
//   phase 2  lanes own persons: apply purchases, consume, CES utility, coalesced stores
//   phase 3  FIRM MATCHING: lane 0 walks firms in visiting order: stale-offer check, profit
//            record, goods purchases (incl. self-purchase), withdrawal of last step's offers
//   phase 4  lanes own firms (by visiting rank): production (CES per output good), decode of
//            goods / job offers, warp scan -> new books in market order, coalesced stores
//
// Reference code restated (paths under /root/reference/src):
//   Economy::time_step base/economy.cpp:95-139 | Person::time_step base/person.cpp:19-33
//   respond_to_jobOffer base/person.cpp:36-54 | review/accept_jobOffer_response base/firm.cpp:56-113
//   respond_to_offer / review / accept base/agent.cpp:99-161 | check_my_offers base/agent.cpp:54-97
//   Firm::time_step base/firm.cpp:23-46 | produce/sell/search_for_laborers firms/profitMaxer.cpp:68-100
//   CES::f functions/vecToScalar.cpp:112-118 | consume + reward neural/neuralPersonDecisionMaker.cpp:93-111
//   offer decode neural/neuralFirmDecisionMaker.cpp:111-180 | profit neural/neuralFirmDecisionMaker.cpp:65-74
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fastace_b200.h"
#include "common.cuh"

namespace fastace {

// shared-memory carve-up for one economy (sizes in bytes, 8-aligned sections)
struct SmemLayout {
    int off_pmoney, off_fmoney, off_finv, off_mprice, off_jwage;               // double
    int off_mleft, off_mtaken, off_jleft, off_jtaken, off_fnh, off_fok; // u32
    int off_permp, off_permf;                                                    // u16
    int off_att, off_fatt, off_mowner, off_mgood, off_jowner, off_pnh, off_pnb, off_ffirst, off_fcnt; // u8
    int att_stride;             // bytes per person in s_att: jobs at [0,S4), goods at [S4,2*S4)
    int total;
};

__host__ __device__ inline SmemLayout make_layout(int P, int F, int G, int S) {
    SmemLayout L;
    const int cap = F * G;
    const int S4 = (S + 3) & ~3;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 7) & ~7; return r; };
    L.att_stride = 2 * S4;
    L.off_pmoney = take(8 * P);
    L.off_fmoney = take(8 * F);
    L.off_finv = take(8 * G * F);
    L.off_mprice = take(8 * cap);
    L.off_jwage = take(8 * F);
    L.off_mleft = take(4 * cap);
    L.off_mtaken = take(4 * cap);
    L.off_jleft = take(4 * F);
    L.off_jtaken = take(4 * F);
    L.off_fnh = take(4 * F);
    L.off_fok = take(4 * F);
    L.off_permp = take(2 * P);
    L.off_permf = take(2 * F);
    L.off_att = take(P * L.att_stride);
    L.off_fatt = take(F * S);
    L.off_mowner = take(cap);
    L.off_mgood = take(cap);
    L.off_jowner = take(F);
    L.off_pnh = take(P);
    L.off_pnb = take(G * P);
    L.off_ffirst = take(F);
    L.off_fcnt = take(F);
    L.total = o;
    return L;
}

template <int G>
__device__ __forceinline__ bool request_good(int n, double& money, double* s_fmoney, double* s_finv, int F,
                                             uint32_t* s_mleft, uint32_t* s_mtaken, const double* s_mprice,
                                             const uint8_t* s_mowner, const uint8_t* s_mgood) {
    const double price = s_mprice[n];
    if (!(money >= price)) return false;          // agent.cpp:102
    const uint32_t left = s_mleft[n];
    if (!(left > 0)) return false;                // agent.cpp:124
    const int s = s_mowner[n], good = s_mgood[n];
    bool short_ = false;                          // agent.cpp:140
#pragma unroll
    for (int g = 0; g < G; g++) {
        const double q = (g == good) ? kAmountPerOffer : 0.0;
        if (s_finv[g * F + s] < q) short_ = true;
    }
    if (short_) { s_mleft[n] = 0; return false; } // agent.cpp:143
    s_fmoney[s] += price;                         // agent.cpp:155-160 (seller)
    s_finv[good * F + s] -= kAmountPerOffer;
    s_mleft[n] = left - 1;
    s_mtaken[n] += 1;
    money -= price;                               // agent.cpp:108 (buyer; inventory applied by caller)
    return true;
}

// Fused single-kernel step with SERIAL matching (kernel v1, FASTACE_STEP_SERIAL): persons and
// firms are matched by a walk on lane 0, every fp64 update (firm money included) in exactly the
// reference's order.  The default path is match_kernel + update_kernel (match_update_kernels.cuh).
template <int G>
__global__ void __launch_bounds__(32, 32) step_kernel(const StepParams p) {
    FASTACE_DYN_SMEM(smem);
    const int e = blockIdx.x;
    const int lane = threadIdx.x;
    const int P = p.P, F = p.F, S = p.S;
    const int cap = F * G;
    const SmemLayout L = make_layout(P, F, G, S);

    double* s_pmoney = reinterpret_cast<double*>(smem + L.off_pmoney);
    double* s_fmoney = reinterpret_cast<double*>(smem + L.off_fmoney);
    double* s_finv = reinterpret_cast<double*>(smem + L.off_finv);
    double* s_mprice = reinterpret_cast<double*>(smem + L.off_mprice);
    double* s_jwage = reinterpret_cast<double*>(smem + L.off_jwage);
    uint32_t* s_mleft = reinterpret_cast<uint32_t*>(smem + L.off_mleft);
    uint32_t* s_mtaken = reinterpret_cast<uint32_t*>(smem + L.off_mtaken);
    uint32_t* s_jleft = reinterpret_cast<uint32_t*>(smem + L.off_jleft);
    uint32_t* s_jtaken = reinterpret_cast<uint32_t*>(smem + L.off_jtaken);
    uint32_t* s_fnh = reinterpret_cast<uint32_t*>(smem + L.off_fnh);
    uint32_t* s_fok = reinterpret_cast<uint32_t*>(smem + L.off_fok);
    uint16_t* s_permp = reinterpret_cast<uint16_t*>(smem + L.off_permp);
    uint16_t* s_permf = reinterpret_cast<uint16_t*>(smem + L.off_permf);
    uint8_t* s_att = smem + L.off_att;
    uint8_t* s_fatt = smem + L.off_fatt;
    uint8_t* s_mowner = smem + L.off_mowner;
    uint8_t* s_mgood = smem + L.off_mgood;
    uint8_t* s_jowner = smem + L.off_jowner;
    uint8_t* s_pnh = smem + L.off_pnh;
    uint8_t* s_pnb = smem + L.off_pnb;
    uint8_t* s_ffirst = smem + L.off_ffirst;
    uint8_t* s_fcnt = smem + L.off_fcnt;
    const int AS = L.att_stride, S4 = AS >> 1;

    const size_t eP = (size_t)e * P, eF = (size_t)e * F, eCap = (size_t)e * cap;
    const int NM = p.st.m_count[e];
    const int NJ = p.st.j_count[e];

    // ------------------------------ phase 0: stage ------------------------------------
    {
        const IndexMap mapJ(NJ, p.flags), mapM(NM, p.flags);
        const bool hasJ = NJ > 0, hasM = NM > 0;
        for (int pid = lane; pid < P; pid += 32) {
            s_pmoney[pid] = p.st.p_money[eP + pid];
            s_permp[pid] = (uint16_t)p.ac.perm_person[eP + pid];
            s_pnh[pid] = 0;
#pragma unroll
            for (int g = 0; g < G; g++) s_pnb[g * P + pid] = 0;
            const size_t k0 = (size_t)e * S * P + pid;
            const int32_t* ji = p.ac.p_job_idx + k0;
            const int32_t* gi = p.ac.p_good_idx + k0;
            const uint8_t* jt = p.ac.p_job_take + k0;
            const uint8_t* gt = p.ac.p_good_take + k0;
            uint8_t* att = s_att + pid * AS;
            for (int i = 0; i < S; i++) {
                const int j = hasJ ? mapJ(ji[(size_t)i * P]) : kNone, g = hasM ? mapM(gi[(size_t)i * P]) : kNone;
                att[i] = (uint8_t)((jt[(size_t)i * P] && hasJ) ? j : kNone);
                att[S4 + i] = (uint8_t)((gt[(size_t)i * P] && hasM) ? g : kNone);
            }
            for (int i = S; i < S4; i++) { att[i] = (uint8_t)kNone; att[S4 + i] = (uint8_t)kNone; }
        }
        for (int f = lane; f < F; f += 32) {
            s_fmoney[f] = p.st.f_money[eF + f];
            s_permf[f] = (uint16_t)p.ac.perm_firm[eF + f];
            s_fnh[f] = 0;
            s_fok[f] = 0;
            s_fcnt[f] = 0;
            s_ffirst[f] = 0;
#pragma unroll
            for (int g = 0; g < G; g++) s_finv[g * F + f] = p.st.f_inv[((size_t)e * G + g) * F + f];
            const size_t k0 = (size_t)e * S * F + f;
            for (int i = 0; i < S; i++) {
                const int g = hasM ? mapM(p.ac.f_good_idx[k0 + (size_t)i * F]) : kNone;
                s_fatt[f * S + i] = (uint8_t)(p.ac.f_good_take[k0 + (size_t)i * F] ? g : kNone);
            }
        }
    }
    for (int n = lane; n < NJ; n += 32) {
        s_jowner[n] = (uint8_t)p.st.j_owner[eF + n];
        s_jleft[n] = p.st.j_left[eF + n];
        s_jtaken[n] = p.st.j_taken[eF + n];
        s_jwage[n] = p.st.j_wage[eF + n];
    }
    __syncwarp();
    for (int n = lane; n < NM; n += 32) {
        const int owner = p.st.m_owner[eCap + n];
        s_mowner[n] = (uint8_t)owner;
        s_mgood[n] = (uint8_t)p.st.m_good[eCap + n];
        s_mleft[n] = p.st.m_left[eCap + n];
        s_mtaken[n] = p.st.m_taken[eCap + n];
        s_mprice[n] = p.st.m_price[eCap + n];
        // a firm's entries are contiguous in market order (it posts all goods in one turn)
        const int prev = (n > 0) ? p.st.m_owner[eCap + n - 1] : -1;
        if (owner != prev) s_ffirst[owner] = (uint8_t)n;
    }
    __syncwarp();
    for (int n = lane; n < NM; n += 32) {
        // count of own entries: last index - first + 1
        const int owner = s_mowner[n];
        const int next = (n + 1 < NM) ? s_mowner[n + 1] : -1;
        if (owner != next) s_fcnt[owner] = (uint8_t)(n + 1 - s_ffirst[owner]);
    }
    __syncwarp();

    // ------------------------------ phase 1: persons ----------------------------------
    // kernel v1: serial walk
    if (lane == 0) {
        for (int r = 0; r < P; r++) {
            const int pid = s_permp[r];
            double money = s_pmoney[pid];
            int nh = 0;
            uint32_t ok = 0;
            const uint8_t* att = s_att + pid * AS;
            // UtilMaxer::search_for_jobs (persons/utilMaxer.cpp:76-85)
            for (int i = 0; i < S; i++) {
                const int n = att[i];
                if (n == kNone) continue;
                // person.cpp:39  laborSupplied + labor <= 1, laborSupplied = 0.5*nh exactly
                if (!(kLaborPerOffer * nh + kLaborPerOffer <= 1)) continue;
                const uint32_t left = s_jleft[n];
                if (!(left > 0)) continue;                             // firm.cpp:64
                const int f = s_jowner[n];
                const double wage = s_jwage[n];
                const double fm = s_fmoney[f];
                if (fm < wage) { s_jleft[n] = 0; continue; }           // firm.cpp:80-84
                s_fmoney[f] = fm - wage;                               // firm.cpp:108
                s_fnh[f] += 1;                                         // laborHired += 0.5, applied in phase 4
                s_jleft[n] = left - 1;
                s_jtaken[n] += 1;
                nh++;                                                  // person.cpp:48
                money += wage;                                         // person.cpp:49
                ok |= 1u << i;
            }
            // UtilMaxer::buy_goods (persons/utilMaxer.cpp:64-73)
            for (int i = 0; i < S; i++) {
                const int n = att[S4 + i];
                if (n == kNone) continue;
                if (request_good<G>(n, money, s_fmoney, s_finv, F, s_mleft, s_mtaken, s_mprice, s_mowner, s_mgood)) {
                    s_pnb[s_mgood[n] * P + pid] += 1;                  // inventory += quantities, applied in phase 2
                    ok |= 1u << (16 + i);
                }
            }
            s_pmoney[pid] = money;
            s_pnh[pid] = (uint8_t)nh;
            write_person_ok(p, e, pid, ok);
        }
    }
    __syncwarp();

    // ------------------------------ phase 2: consume ----------------------------------
    for (int pid = lane; pid < P; pid += 32) {
        const double labor = kLaborPerOffer * (double)s_pnh[pid];      // exact: 0, 0.5 or 1.0
        double x[G + 1], c[G], inv[G];
        x[0] = 1 - labor;                                              // utilMaxer.cpp:56
#pragma unroll
        for (int g = 0; g < G; g++) {
            double v = p.st.p_inv[((size_t)e * G + g) * P + pid];
            const int nb = s_pnb[g * P + pid];
            for (int k = 0; k < nb; k++) v += kAmountPerOffer;         // agent.cpp:109, one add per purchase
            c[g] = v * (double)p.ac.p_consume[((size_t)e * G + g) * P + pid];  // neuralPersonDecisionMaker.cpp:99
            x[g + 1] = c[g];
            inv[g] = v - c[g];                                         // utilMaxer.cpp:91
        }
        double share[G + 1], theta[G + 1];
#pragma unroll
        for (int i = 0; i <= G; i++) {
            share[i] = p.st.p_util_share[((size_t)e * (G + 1) + i) * P + pid];
            theta[i] = (p.util_kind == FASTACE_FN_STONE_GEARY) ? p.st.p_util_theta[((size_t)e * (G + 1) + i) * P + pid] : 0.0;
        }
        const double util = eval_function<G + 1, true>(p.util_kind, p.st.p_util_tfp[eP + pid], share, theta, p.st.p_util_rho[eP + pid], x);
        p.out.p_reward[eP + pid] = util;
        p.st.p_money[eP + pid] = s_pmoney[pid];
        p.st.p_labor[eP + pid] = labor;
#pragma unroll
        for (int g = 0; g < G; g++) p.st.p_inv[((size_t)e * G + g) * P + pid] = inv[g];
    }
    // job counters are final after the person phase
    if (p.out.old_j_left) for (int n = lane; n < NJ; n += 32) p.out.old_j_left[eF + n] = s_jleft[n];
    if (p.out.old_j_taken) for (int n = lane; n < NJ; n += 32) p.out.old_j_taken[eF + n] = s_jtaken[n];

    // ------------------------------ phase 3: firms, serial part -----------------------
    if (lane == 0) {
        for (int r = 0; r < F; r++) {
            const int f = s_permf[r];
            const int first = s_ffirst[f], cnt = s_fcnt[f];
            // Agent::check_my_offers (base/agent.cpp:54-97): running inventoryLeft over own entries
            {
                double invLeft[G];
#pragma unroll
                for (int g = 0; g < G; g++) invLeft[g] = s_finv[g * F + f];
                for (int n = first; n < first + cnt; n++) {
                    const int good = s_mgood[n];
                    uint32_t left = s_mleft[n];
                    double delta = kAmountPerOffer * (double)left;     // agent.cpp:73 (other goods: 0*left = 0)
                    for (;;) {
                        bool okk = !(delta > invLeft[good]);
#pragma unroll
                        for (int g = 0; g < G; g++) if (g != good && 0.0 > invLeft[g]) okk = false;
                        if (okk || left == 0) break;                   // left==0 guard: see SURVEY.md B.2
                        delta -= kAmountPerOffer;                      // agent.cpp:79-80
                        left--;
                    }
                    s_mleft[n] = left;
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == good) invLeft[g] -= delta;  // agent.cpp:83
                }
            }
            // first decision: profit of the previous step (neuralFirmDecisionMaker.cpp:65-74)
            {
                const double m = s_fmoney[f];
                const double last = p.st.f_last_money[eF + f];
                p.out.f_profit[eF + f] = (p.time_before > 0) ? (m - last) : 0.0;
                p.st.f_last_money[eF + f] = m;
            }
            // ProfitMaxer::buy_goods (firms/profitMaxer.cpp:102-111)
            uint32_t ok = 0;
            for (int i = 0; i < S; i++) {
                const int n = s_fatt[f * S + i];
                if (n == kNone) continue;
                double money = s_fmoney[f];
                const double price = s_mprice[n];
                if (!(money >= price)) continue;
                const uint32_t left = s_mleft[n];
                if (!(left > 0)) continue;
                const int s = s_mowner[n], good = s_mgood[n];
                bool short_ = false;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const double q = (g == good) ? kAmountPerOffer : 0.0;
                    if (s_finv[g * F + s] < q) short_ = true;
                }
                if (short_) { s_mleft[n] = 0; continue; }
                s_fmoney[s] += price;                                  // seller first (may be f itself)
                s_finv[good * F + s] -= kAmountPerOffer;
                s_mleft[n] = left - 1;
                s_mtaken[n] += 1;
                s_fmoney[f] -= price;                                  // then buyer
                s_finv[good * F + f] += kAmountPerOffer;
                ok |= 1u << i;
            }
            s_fok[f] = ok;
            // ProfitMaxer::sell_goods withdraws last step's offers (firms/profitMaxer.cpp:79-81);
            // nothing between buy_goods and that point touches another agent.
            for (int n = first; n < first + cnt; n++) {
                if (p.out.old_m_left) p.out.old_m_left[eCap + n] = s_mleft[n];
                if (p.out.old_m_taken) p.out.old_m_taken[eCap + n] = s_mtaken[n];
                s_mleft[n] = 0;
            }
        }
    }
    __syncwarp();

    // ------------------------------ phase 4: produce + post ---------------------------
    int base_m = 0, base_j = 0;
    for (int r0 = 0; r0 < F; r0 += 32) {
        const int r = r0 + lane;
        const bool active = r < F;
        int lots[G];
        double price[G];
        int nposted = 0, jlots = 0;
        double jwage = 0.0;
        int f = 0;
        if (active) {
            f = s_permf[r];
            double labor = p.st.f_labor[eF + f];
            const uint32_t nh = s_fnh[f];
            for (uint32_t k = 0; k < nh; k++) labor += kLaborPerOffer;  // firm.cpp:109, one add per hire
            double in[G + 1], x[G], inv[G];
            in[0] = labor;
#pragma unroll
            for (int g = 0; g < G; g++) {
                inv[g] = s_finv[g * F + f];
                x[g] = inv[g] * (double)p.ac.f_prod[((size_t)e * G + g) * F + f];  // neuralFirmDecisionMaker.cpp:101
                in[g + 1] = x[g];
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                double share[G + 1], theta[G + 1];
#pragma unroll
                for (int i = 0; i <= G; i++) {
                    const size_t k = (((size_t)e * G + g) * (G + 1) + i) * F + f;
                    share[i] = p.st.f_prod_share[k];
                    theta[i] = (p.prod_kind == FASTACE_FN_STONE_GEARY) ? p.st.f_prod_theta[k] : 0.0;
                }
                const double outg = eval_function<G + 1, false>(p.prod_kind, p.st.f_prod_tfp[((size_t)e * G + g) * F + f], share, theta,
                                                                p.st.f_prod_rho[((size_t)e * G + g) * F + f], in);
                inv[g] += (outg - x[g]);                               // profitMaxer.cpp:71
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                // decisionNetHandler.cpp:591 amounts = proportion * inventory; neuralFirmDecisionMaker.cpp:129
                const double amount = (double)p.ac.f_offer_amt[((size_t)e * G + g) * F + f] * inv[g];
                lots[g] = x86_double_to_int(amount / kAmountPerOffer);
                price[g] = (double)p.ac.f_offer_price[((size_t)e * G + g) * F + f] / kAmountPerOffer;
                if (lots[g] > 0) nposted++;
                p.st.f_inv[((size_t)e * G + g) * F + f] = inv[g];
            }
            p.st.f_money[eF + f] = s_fmoney[f];
            p.st.f_labor[eF + f] = 0.0;                                // firm.cpp:41
            // neuralFirmDecisionMaker.cpp:164-176, decisionNetHandler.cpp:631-635
            double wage = (double)p.ac.f_job_wage[eF + f];
            if (wage > kLargeNumber) wage = kLargeNumber;
            jlots = x86_double_to_int((double)p.ac.f_job_labor[eF + f] / kLaborPerOffer);
            jwage = wage / kLaborPerOffer;
            if (p.out.f_good_ok) {
                const uint32_t ok = s_fok[f];
                for (int i = 0; i < S; i++) p.out.f_good_ok[((size_t)e * S + i) * F + f] = (ok >> i) & 1u;
            }
        }
        // market order = visiting rank, goods ascending, lots > 0 (economy.cpp:52-59, util.h:50-65)
        int incl = nposted;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        int slot = base_m + incl - nposted;
        const unsigned jmask = __ballot_sync(0xffffffffu, active && jlots > 0);
        const int jslot = base_j + __popc(jmask & ((1u << lane) - 1u));
        if (active) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (lots[g] > 0) {
                    p.st.m_owner[eCap + slot] = f;
                    p.st.m_good[eCap + slot] = g;
                    p.st.m_left[eCap + slot] = (uint32_t)lots[g];
                    p.st.m_taken[eCap + slot] = 0;
                    p.st.m_price[eCap + slot] = price[g];
                    slot++;
                }
            }
            if (jlots > 0) {
                p.st.j_owner[eF + jslot] = f;
                p.st.j_left[eF + jslot] = (uint32_t)jlots;
                p.st.j_taken[eF + jslot] = 0;
                p.st.j_wage[eF + jslot] = jwage;
            }
        }
        base_m += __shfl_sync(0xffffffffu, incl, 31);
        base_j += __popc(jmask);
    }
    if (lane == 0) {
        p.st.m_count[e] = base_m;
        p.st.j_count[e] = base_j;
    }
}

}  // namespace fastace
