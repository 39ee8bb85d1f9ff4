// TEST INFRASTRUCTURE ONLY — harness around the reference's UNMODIFIED trainer
// (/root/reference/src/neural/advantageActorCritic.cpp) and decision-net handler
// (decisionNetHandler.cpp, neuralEconomy.cpp, decisionNets.cpp), compiled against the LibTorch of
// the torch wheel into oracle/_ref/libfastace_refa2c.so.  It pins fastace_b200/trainer.py: same
// episode loss, same gradients, same parameters after the Adam steps, same LR schedule.
//
// OURS: a driver that fills one episode of the handler's history through the handler's own
// public decision API (the calls NeuralPersonDecisionMaker / NeuralFirmDecisionMaker make, in their
// order) from injected agent states, and export entry points.  The nets, the sampling, the
// log-probability bookkeeping, the loss, backward, Adam and the schedulers are the reference's.
// Random draws come from torch's default CPU generator, which the calling python process shares
// (same libtorch), so the test reproduces them by seeding and replaying the same call sequence.
#include <torch/torch.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "advantageActorCritic.h"
#include "neuralEconomy.h"
#include "vecToScalar.h"
#include "vecToVec.h"

using namespace neural;

namespace {

class IdlePerson : public PersonDecisionMaker {   // never asked: the harness does not run Economy::time_step
public:
    std::vector<Order<JobOffer>> choose_jobs() override { return {}; }
    std::vector<Order<Offer>> choose_goods() override { return {}; }
    Eigen::ArrayXd choose_goods_to_consume() override { return Eigen::ArrayXd::Zero(1); }
};
class IdleFirm : public FirmDecisionMaker {
public:
    std::vector<Order<Offer>> choose_goods() override { return {}; }
    Eigen::ArrayXd choose_production_inputs() override { return Eigen::ArrayXd::Zero(1); }
    std::vector<std::shared_ptr<Offer>> choose_good_offers() override { return {}; }
    std::vector<std::shared_ptr<JobOffer>> choose_job_offers() override { return {}; }
};

struct RefA2C {
    int G, S, P, F;
    std::shared_ptr<NeuralEconomy> economy;
    std::shared_ptr<DecisionNetHandler> handler;
    std::shared_ptr<AdvantageActorCritic> trainer;
    std::vector<std::shared_ptr<UtilMaxer>> persons;
    std::vector<std::shared_ptr<ProfitMaxer>> firms;
    std::vector<Eigen::ArrayXd> utilParams, prodParams;
    std::vector<std::shared_ptr<Offer>> offers;          // keep-alive for the weak_ptrs in the market
    std::vector<std::shared_ptr<JobOffer>> jobOffers;
    std::vector<std::pair<std::string, torch::Tensor>> flat;
    int steps = 0;
};

Eigen::ArrayXd arr(const double* p, int n) {
    Eigen::ArrayXd a(n);
    for (int i = 0; i < n; i++) a(i) = p[i];
    return a;
}

}  // namespace

extern "C" {

// util_params [P][G+3] = tfp, G+1 normalised shares, rho ; prod_params [F][G][G+3] likewise per output good
RefA2C* refa2c_create(int G, int S, int enc, int hidden, int nHidden, int nHiddenSmall, int P, int F, double discount,
                      const double* util_params, const double* prod_params, const double* lrs9, unsigned episodeBatch,
                      unsigned patience, double multiplier, unsigned cosinePeriod, uint64_t seed) {
    torch::manual_seed(seed);
    auto* r = new RefA2C;
    r->G = G; r->S = S; r->P = P; r->F = F;
    r->economy = NeuralEconomy::init_dummy(G);
    r->handler = std::make_shared<DecisionNetHandler>(r->economy, S, enc, hidden, nHidden, nHiddenSmall);
    r->economy->handler = r->handler;
    const int U = G + 3;
    for (int p = 0; p < P; p++) {
        const double* u = util_params + (size_t)p * U;
        auto ces = std::make_shared<CES>(u[0], arr(u + 1, G + 1), 2.0);
        ces->shareParams = arr(u + 1, G + 1);
        ces->substitutionParam = u[U - 1];
        r->persons.push_back(UtilMaxer::init(r->economy.get(), Eigen::ArrayXd::Zero(G), 0.0, ces, discount,
                                             std::make_shared<IdlePerson>()));
        r->utilParams.push_back(arr(u, U));
    }
    for (int f = 0; f < F; f++) {
        std::vector<std::shared_ptr<VecToVec>> inner(G);
        for (int g = 0; g < G; g++) {
            const double* u = prod_params + ((size_t)f * G + g) * U;
            auto ces = std::make_shared<CES>(u[0], arr(u + 1, G + 1), 2.0);
            ces->shareParams = arr(u + 1, G + 1);
            ces->substitutionParam = u[U - 1];
            inner[g] = std::make_shared<VToVFromVToS<CES>>(ces, G, g);
        }
        r->firms.push_back(ProfitMaxer::init(r->economy.get(), std::vector<std::shared_ptr<Agent>>(),
                                             Eigen::ArrayXd::Zero(G), 0.0, std::make_shared<SumOfVecToVec>(inner),
                                             std::make_shared<IdleFirm>()));
        r->prodParams.push_back(arr(prod_params + (size_t)f * G * U, G * U));   // get_prodFuncParams column order
    }
    r->trainer = std::make_shared<AdvantageActorCritic>(r->handler, lrs9[0], lrs9[1], lrs9[2], lrs9[3], lrs9[4], lrs9[5],
                                                         lrs9[6], lrs9[7], lrs9[8], episodeBatch, patience, multiplier,
                                                         cosinePeriod);
    auto& h = *r->handler;
    std::vector<std::pair<std::string, torch::nn::Module*>> all = {
        {"offerEncoder", h.offerEncoder.get()}, {"jobOfferEncoder", h.jobOfferEncoder.get()},
        {"purchaseNet", h.purchaseNet.get()}, {"firmPurchaseNet", h.firmPurchaseNet.get()},
        {"laborSearchNet", h.laborSearchNet.get()}, {"consumptionNet", h.consumptionNet.get()},
        {"productionNet", h.productionNet.get()}, {"offerNet", h.offerNet.get()},
        {"jobOfferNet", h.jobOfferNet.get()}, {"valueNet", h.valueNet.get()}, {"firmValueNet", h.firmValueNet.get()}};
    for (auto& kv : all)
        for (auto& np : kv.second->named_parameters(/*recurse=*/true))
            r->flat.push_back({kv.first + "/" + np.key(), np.value()});
    return r;
}

void refa2c_destroy(RefA2C* r) { delete r; }

// posts nM goods offers (one unit of good[i] at price[i]) and nJ job offers (0.5 labour at wage[j]) from firm 0
// (Economy::add_offer / add_jobOffer, base.h:108-109) — what update_encodedOffers will see
void refa2c_post_market(RefA2C* r, int nM, const int* good, const double* price, int nJ, const double* wage) {
    for (int i = 0; i < nM; i++) {
        Eigen::ArrayXd q = Eigen::ArrayXd::Zero(r->G);
        q(good[i]) = 1.0;
        auto o = std::make_shared<Offer>(r->firms[0], 3, q, price[i]);
        r->offers.push_back(o);
        r->economy->add_offer(o);
    }
    for (int j = 0; j < nJ; j++) {
        auto o = std::make_shared<JobOffer>(r->firms[0], 3, 0.5, wage[j]);
        r->jobOffers.push_back(o);
        r->economy->add_jobOffer(o);
    }
}

// new episode: DecisionNetHandler::reset (decisionNetHandler.cpp:310-325)
void refa2c_reset(RefA2C* r) {
    r->handler->reset(r->economy);
    r->steps = 0;
}

// One simulated economy step: DecisionNetHandler::time_step(), then for every person and every firm the
// handler calls their decision maker would issue, in its order (neuralPersonDecisionMaker.cpp:18-111,
// neuralFirmDecisionMaker.cpp:20-180).  States are injected: p_* [P], p_inv [P][G], f_* [F], f_inv [F][G].
// p_reward [P] is the utility recorded for this step; f_profit [F] is recorded for the PREVIOUS step when t>0.
void refa2c_step(RefA2C* r, const double* p_money, const double* p_labor, const double* p_inv, const double* p_reward,
                 const double* f_money, const double* f_labor, const double* f_inv, const double* f_profit) {
    auto& h = *r->handler;
    h.time_step();
    const int G = r->G;
    for (int p = 0; p < r->P; p++) {
        Agent* who = r->persons[p].get();
        auto inv = arr(p_inv + (size_t)p * G, G);
        auto idxM = h.generate_offerIndices();
        auto idxJ = h.generate_jobOfferIndices();
        h.record_value(who, idxM, idxJ, r->utilParams[p], p_money[p], p_labor[p], inv);
        h.get_joboffers_to_request(who, idxJ, r->utilParams[p], p_money[p], p_labor[p], inv);
        h.get_offers_to_request(who, idxM, r->utilParams[p], p_money[p], p_labor[p], inv);
        h.get_consumption_proportions(who, r->utilParams[p], p_money[p], p_labor[p], inv);
        h.record_reward(who, p_reward[p]);
    }
    for (int f = 0; f < r->F; f++) {
        Agent* who = r->firms[f].get();
        auto inv = arr(f_inv + (size_t)f * G, G);
        auto idxM = h.firm_generate_offerIndices();
        auto idxJ = h.firm_generate_jobOfferIndices();
        h.firm_record_value(who, idxM, idxJ, r->prodParams[f], f_money[f], f_labor[f], inv);
        if (r->steps > 0) h.record_reward(who, f_profit[f], 1);
        h.firm_get_offers_to_request(who, idxM, r->prodParams[f], f_money[f], f_labor[f], inv);
        h.get_production_proportions(who, r->prodParams[f], f_money[f], f_labor[f], inv);
        h.choose_offers(who, idxM, r->prodParams[f], f_money[f], f_labor[f], inv);
        h.choose_job_offers(who, idxJ, r->prodParams[f], f_money[f], f_labor[f], inv);
    }
    r->steps++;
}

// what the handler recorded: kind 0..6 = purchase, firmPurchase, laborSearch, consumption, production, offer,
// jobOffer log-probabilities; 7 = value; 8 = reward.  agent: person index, or P + firm index.
double refa2c_record(RefA2C* r, int kind, int t, int agent) {
    auto& h = *r->handler;
    DecisionNetHandler::VecMapTensor* maps[] = {&h.purchaseNetLogProba, &h.firmPurchaseNetLogProba,
        &h.laborSearchNetLogProba, &h.consumptionNetLogProba, &h.productionNetLogProba, &h.offerNetLogProba,
        &h.jobOfferNetLogProba, &h.values, &h.rewards};
    Agent* who = agent < r->P ? (Agent*)r->persons[agent].get() : (Agent*)r->firms[agent - r->P].get();
    auto& m = (*maps[kind])[t];
    auto it = m.find(who);
    if (it == m.end()) return -12345.0;
    return it->second.item<double>();
}

double refa2c_train_on_episode(RefA2C* r) { return r->trainer->train_on_episode(); }

void refa2c_lrs(RefA2C* r, double* out9) {
    auto& t = *r->trainer;
    LRScheduler* s[] = {&t.purchaseNetScheduler, &t.firmPurchaseNetScheduler, &t.laborSearchNetScheduler,
                        &t.consumptionNetScheduler, &t.productionNetScheduler, &t.offerNetScheduler,
                        &t.jobOfferNetScheduler, &t.valueNetScheduler, &t.firmValueNetScheduler};
    for (int i = 0; i < 9; i++) out9[i] = s[i]->get_lr();
}

int refa2c_num_params(RefA2C* r) { return (int)r->flat.size(); }
int refa2c_param_info(RefA2C* r, int i, char* name, int name_cap, int64_t* shape2) {
    std::snprintf(name, name_cap, "%s", r->flat[i].first.c_str());
    const auto& t = r->flat[i].second;
    shape2[0] = t.size(0);
    shape2[1] = t.dim() > 1 ? t.size(1) : 0;
    return (int)t.numel();
}
// grad != 0: the .grad left by the last train_on_episode (zeros if none)
void refa2c_param_data(RefA2C* r, int i, int grad, float* dst) {
    torch::NoGradGuard g;
    torch::Tensor t = r->flat[i].second;
    if (grad) t = t.grad().defined() ? t.grad() : torch::zeros_like(t);
    auto c = t.contiguous().to(torch::kFloat32);
    std::memcpy(dst, c.data_ptr<float>(), sizeof(float) * c.numel());
}

}  // extern "C"
