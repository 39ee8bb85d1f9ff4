// Host-side pieces of the C ABI that need no GPU: scenario initial-state draws, the
// per-step visiting orders, and the two struct-returning legacy entry points.
//
// Reference behaviour followed here:
//   create_scenario_params / create_training_params   src/pybindings.cpp:8-18
//   CustomScenarioParams / TrainingParams defaults     src/neural/neuralScenarios.h:49-186,
//                                                      src/neural/neuralConstants.h:13-29
//   CustomScenario::setup (initial-state distributions) src/neural/neuralScenarios.cpp:93-161
//   CES constructor normalisation                      src/functions/vecToScalar.cpp:105-110
//   util::make_nonnegative / make_positive             src/base/util.h:86-104
//   Economy::time_step's two std::shuffle calls        src/base/economy.cpp:110-111
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "../../include/fastace_b200.h"
#include "fastace_internal.h"

static_assert(sizeof(fastace_custom_scenario_params_t) == 344, "must match neural::CustomScenarioParams");
static_assert(sizeof(fastace_training_params_t) == 136, "must match neural::TrainingParams");

namespace fastace {
thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace fastace

extern "C" {

int fastace_abi_version(void) { return FASTACE_ABI_VERSION; }
const char* fastace_last_error(void) { return fastace::g_last_error.c_str(); }

fastace_custom_scenario_params_t create_scenario_params(unsigned int numPeople, unsigned int numFirms) {
    fastace_custom_scenario_params_t p;
    p.numPeople = numPeople; p.numFirms = numFirms;
    p.money_mu = 10.0; p.money_sigma = 2.0;
    p.good1_mu = 10.0; p.good1_sigma = 2.0;
    p.good2_mu = 1.0; p.good2_sigma = 0.2;
    p.labor_share_mu = 0.4; p.labor_share_sigma = 0.1;
    p.good1_share_mu = 0.4; p.good1_share_sigma = 0.1;
    p.good2_share_mu = 0.1; p.good2_share_sigma = 0.02;
    p.discount_mu = 2.0; p.discount_sigma = 1.0;
    p.elasticity_mu = 10.0; p.elasticity_sigma = 2.5;
    p.firm_money_mu = 50.0; p.firm_money_sigma = 10.0;
    p.firm_good1_mu = 10.0; p.firm_good1_sigma = 4.0;
    p.firm_good2_mu = 30.0; p.firm_good2_sigma = 5.0;
    p.firm_tfp1_mu = 1.0; p.firm_tfp1_sigma = 0.2;
    p.firm_tfp2_mu = 1.0; p.firm_tfp2_sigma = 0.2;
    p.firm_labor_share1_mu = 0.4; p.firm_labor_share1_sigma = 0.05;
    p.firm_good1_share1_mu = 0.1; p.firm_good1_share1_sigma = 0.02;
    p.firm_good2_share1_mu = 0.4; p.firm_good2_share1_sigma = 0.02;
    p.firm_labor_share2_mu = 0.4; p.firm_labor_share2_sigma = 0.05;
    p.firm_good1_share2_mu = 0.1; p.firm_good1_share2_sigma = 0.02;
    p.firm_good2_share2_mu = 0.4; p.firm_good2_share2_sigma = 0.05;
    p.firm_elasticity1_mu = 10.0; p.firm_elasticity1_sigma = 2.5;
    p.firm_elasticity2_mu = 10.0; p.firm_elasticity2_sigma = 2.5;
    return p;
}

fastace_training_params_t create_training_params(void) {
    fastace_training_params_t t;
    t.numEpisodes = 100; t.episodeLength = 20; t.updateEveryNEpisodes = 10; t.checkpointEveryNEpisodes = 10;
    t.stackSize = 10; t.encodingSize = 10; t.hiddenSize = 100; t.nHidden = 12; t.nHiddenSmall = 6;
    t.purchaseNetLR = t.firmPurchaseNetLR = t.laborSearchNetLR = t.consumptionNetLR = 1e-5;
    t.productionNetLR = t.offerNetLR = t.jobOfferNetLR = t.valueNetLR = t.firmValueNetLR = 1e-5;
    t.episodeBatchSizeForLRDecay = 10; t.patienceForLRDecay = 5;
    t.multiplierForLRDecay = 0.5; t.reverseAnnealingPeriod = 3;
    return t;
}

static inline double make_nonnegative(double x) { return x < 0 ? 0.0 : x; }
static inline double make_positive(double x) { return x <= 0 ? 1e-8 : x; }

// Draw order is fixed here to SOURCE order of the arguments in CustomScenario::setup.
// (In the reference the order of the randn(rng) calls inside one constructor-call
// expression is unspecified by C++, so its initial states are not portable across
// compilers; SURVEY.md B.5.  Distributions, clamps and the good-1 sigma quirk of
// neuralScenarios.cpp:130 are the reference's.)
int fastace_scenario_custom_init(const fastace_dims_t* dims, const fastace_custom_scenario_params_t* sp,
                                 uint32_t seed, fastace_state_t* st, double* p_discount) {
    if (!dims || !sp || !st) { fastace::set_error("null argument"); return FASTACE_ERR_INVALID; }
    const int E = dims->num_econ, P = dims->num_persons, F = dims->num_firms, G = dims->num_goods;
    if (G != 2) { fastace::set_error("CustomScenario has exactly 2 goods (neuralScenarios.cpp:94-96)"); return FASTACE_ERR_INVALID; }
    const size_t cap = (size_t)F * G;
    for (int e = 0; e < E; e++) {
        std::minstd_rand0 rng(seed + (uint32_t)e);  // std::default_random_engine in libstdc++
        std::normal_distribution<double> randn(0, 1);
        for (int p = 0; p < P; p++) {
            const size_t ip = (size_t)e * P + p;
            double g1 = make_nonnegative(sp->good1_mu + sp->good1_sigma * randn(rng));
            double g2 = make_nonnegative(sp->good2_mu + sp->good2_sigma * randn(rng));
            double money = make_nonnegative(sp->money_mu + sp->money_sigma * randn(rng));
            double s0 = sp->labor_share_mu + sp->labor_share_sigma * randn(rng);
            double s1 = sp->good1_share_mu + sp->good1_share_sigma * randn(rng);
            double s2 = sp->good2_share_mu + sp->good2_share_sigma * randn(rng);
            double elast = make_positive(sp->elasticity_mu + sp->elasticity_sigma * randn(rng));
            double disc = 1.0 / (1.0 + std::exp(-sp->discount_mu - sp->discount_sigma * randn(rng)));
            st->p_money[ip] = money;
            st->p_inv[((size_t)e * G + 0) * P + p] = g1;
            st->p_inv[((size_t)e * G + 1) * P + p] = g2;
            st->p_labor[ip] = 0.0;
            st->p_util_tfp[ip] = 1.0;
            double sum = 0.0;  // CES::CES: shareParams / shareParams.sum(), rho = 1/(1-elasticity)
            sum += s0; sum += s1; sum += s2;
            st->p_util_share[((size_t)e * 3 + 0) * P + p] = s0 / sum;
            st->p_util_share[((size_t)e * 3 + 1) * P + p] = s1 / sum;
            st->p_util_share[((size_t)e * 3 + 2) * P + p] = s2 / sum;
            st->p_util_rho[ip] = 1 / (1 - elast);
            if (p_discount) p_discount[ip] = disc;
        }
        for (int f = 0; f < F; f++) {
            const size_t jf = (size_t)e * F + f;
            double g1 = make_nonnegative(sp->firm_good1_mu + sp->firm_good2_sigma * randn(rng));  // sic, :130
            double g2 = make_nonnegative(sp->firm_good2_mu + sp->firm_good2_sigma * randn(rng));
            double money = make_nonnegative(sp->firm_money_mu + sp->firm_money_sigma * randn(rng));
            double tfp[2], sh[2][3], el[2];
            tfp[0] = make_nonnegative(sp->firm_tfp1_mu + sp->firm_tfp1_sigma * randn(rng));
            tfp[1] = make_nonnegative(sp->firm_tfp2_mu + sp->firm_tfp2_sigma * randn(rng));
            sh[0][0] = sp->firm_labor_share1_mu + sp->firm_labor_share1_sigma * randn(rng);
            sh[0][1] = sp->firm_good1_share1_mu + sp->firm_good1_share1_sigma * randn(rng);
            sh[0][2] = sp->firm_good2_share1_mu + sp->firm_good2_share1_sigma * randn(rng);
            sh[1][0] = sp->firm_labor_share2_mu + sp->firm_labor_share2_sigma * randn(rng);
            sh[1][1] = sp->firm_good1_share2_mu + sp->firm_good1_share2_sigma * randn(rng);
            sh[1][2] = sp->firm_good2_share2_mu + sp->firm_good2_share2_sigma * randn(rng);
            el[0] = make_positive(sp->firm_elasticity1_mu + sp->firm_elasticity1_sigma * randn(rng));
            el[1] = make_positive(sp->firm_elasticity2_mu + sp->firm_elasticity2_sigma * randn(rng));
            st->f_money[jf] = money;
            st->f_inv[((size_t)e * G + 0) * F + f] = g1;
            st->f_inv[((size_t)e * G + 1) * F + f] = g2;
            st->f_labor[jf] = 0.0;
            st->f_last_money[jf] = 0.0;
            for (int g = 0; g < 2; g++) {
                st->f_prod_tfp[((size_t)e * G + g) * F + f] = tfp[g];
                double sum = 0.0;
                sum += sh[g][0]; sum += sh[g][1]; sum += sh[g][2];
                for (int i = 0; i < 3; i++)
                    st->f_prod_share[(((size_t)e * G + g) * 3 + i) * F + f] = sh[g][i] / sum;
                st->f_prod_rho[((size_t)e * G + g) * F + f] = 1 / (1 - el[g]);
            }
        }
        st->m_count[e] = 0;
        st->j_count[e] = 0;
        for (size_t n = 0; n < cap; n++) {
            st->m_owner[e * cap + n] = 0; st->m_good[e * cap + n] = 0; st->m_left[e * cap + n] = 0;
            st->m_taken[e * cap + n] = 0; st->m_price[e * cap + n] = 0.0;
        }
        for (int n = 0; n < F; n++) {
            st->j_owner[(size_t)e * F + n] = 0; st->j_left[(size_t)e * F + n] = 0;
            st->j_taken[(size_t)e * F + n] = 0; st->j_wage[(size_t)e * F + n] = 0.0;
        }
    }
    return FASTACE_OK;
}

// Economy::time_step shuffles `persons` then `firms` IN PLACE with the economy's own
// engine (economy.cpp:110-111), so each step's order is the previous order permuted
// again.  std::default_random_engine is std::minstd_rand0 in libstdc++: x <- 16807 x mod
// (2^31 - 1); its whole state is that one word, carried between calls in rng_state.
// Minstd0 exposes the same min/max/result_type, so std::shuffle consumes it identically
// (checked against std::minstd_rand0 itself in tests/test_host_logic.py).
namespace {
struct Minstd0 {
    typedef std::minstd_rand0::result_type result_type;
    uint64_t x;
    static constexpr result_type min() { return 1u; }
    static constexpr result_type max() { return 2147483646u; }
    result_type operator()() { x = (x * 16807ull) % 2147483647ull; return (result_type)x; }
};
}  // namespace

int fastace_shuffle_orders(const fastace_dims_t* dims, uint32_t seed, uint64_t* rng_state,
                           int32_t* perm_person, int32_t* perm_firm, int first_call) {
    if (!dims || !rng_state || !perm_person || !perm_firm) { fastace::set_error("null argument"); return FASTACE_ERR_INVALID; }
    const int E = dims->num_econ, P = dims->num_persons, F = dims->num_firms;
    for (int e = 0; e < E; e++) {
        Minstd0 rng;
        int32_t* pp = perm_person + (size_t)e * P;
        int32_t* pf = perm_firm + (size_t)e * F;
        if (first_call) {
            // linear_congruential_engine::seed(s): state = s mod m, or 1 if that is 0
            uint64_t s0 = (uint64_t)(seed + (uint32_t)e) % 2147483647ull;
            rng.x = s0 == 0 ? 1 : s0;
            for (int i = 0; i < P; i++) pp[i] = i;
            for (int i = 0; i < F; i++) pf[i] = i;
        } else {
            rng.x = rng_state[e];
        }
        std::shuffle(pp, pp + P, rng);
        std::shuffle(pf, pf + F, rng);
        rng_state[e] = rng.x;
    }
    return FASTACE_OK;
}

}  // extern "C"
