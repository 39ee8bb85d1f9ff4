// TEST INFRASTRUCTURE ONLY — harness around the UNMODIFIED reference sources.
//
// Built by oracle/Makefile into oracle/_ref/libfastace_ref.so together with
// /root/reference/src/{base,functions,persons,firms}/*.cpp compiled where they lie
// (with oracle/shim/Eigen/Dense standing in for Eigen and oracle/ref_prelude.h selecting
// the reference's own single-threaded branch).  Nothing in the product imports this.
//
// What is OURS here (and why it is needed):
//   * SeededEconomy   — subclass that seeds the protected Economy::rng (base.h:127; the
//                       reference seeds it from the clock, util.cpp:6-9) and exposes the
//                       protected agent vectors so the visiting order can be exported.
//   * ReplayPerson / ReplayFirm — implementations of the reference's plugin interfaces
//                       PersonDecisionMaker (utilMaxer.h:11-28) and FirmDecisionMaker
//                       (profitMaxer.h:11-24) that replay injected action arrays instead
//                       of running the LibTorch nets.  They follow the shipped neural
//                       plugins line by line for everything that is NOT a net forward:
//                         snapshot rule            decisionNetHandler.cpp:236-275, 303-308
//                         order construction       decisionNetHandler.cpp:368-387, 446-466
//                         consumption + reward     neuralPersonDecisionMaker.cpp:93-111
//                         production inputs        neuralFirmDecisionMaker.cpp:95-108
//                         goods-offer decode       neuralFirmDecisionMaker.cpp:111-146, decisionNetHandler.cpp:586-592
//                         job-offer decode         neuralFirmDecisionMaker.cpp:149-180, decisionNetHandler.cpp:626-635
//                         profit reward            neuralFirmDecisionMaker.cpp:65-74
//   Everything inside Economy::time_step (shuffles, FCFS matching, stale-offer checks,
//   production/utility functions, flushes) is executed by the reference's own code.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

#include "base.h"
#include "profitMaxer.h"
#include "utilMaxer.h"
#include "vecToScalar.h"
#include "vecToVec.h"

#include "../include/fastace_b200.h"

namespace {

const double AMOUNT_PER_OFFER = 1.0;        // neuralFirmDecisionMaker.cpp:6
const double LABOR_AMOUNT_PER_OFFER = 0.5;  // neuralFirmDecisionMaker.cpp:7

struct RefEconomyBox;

class SeededEconomy : public Economy {
public:
    SeededEconomy(std::vector<std::string> goods, unsigned int seed) : Economy(goods) {
        rng = std::default_random_engine(seed);
    }
    const std::vector<std::shared_ptr<Person>>& person_order() const { return persons; }
    const std::vector<std::shared_ptr<Firm>>& firm_order() const { return firms; }
};

// Pointers into the current step's injected arrays for ONE economy (already offset to
// that economy), plus the market snapshot shared by all of its agents.
struct StepContext {
    int P = 0, F = 0, G = 0, S = 0;
    uint32_t flags = 0;
    const int32_t* p_job_idx = nullptr;
    const uint8_t* p_job_take = nullptr;
    const int32_t* p_good_idx = nullptr;
    const uint8_t* p_good_take = nullptr;
    const float* p_consume = nullptr;
    const int32_t* f_good_idx = nullptr;
    const uint8_t* f_good_take = nullptr;
    const float* f_prod = nullptr;
    const float* f_offer_amt = nullptr;
    const float* f_offer_price = nullptr;
    const float* f_job_labor = nullptr;
    const float* f_job_wage = nullptr;
    double* p_reward = nullptr;
    double* f_profit = nullptr;

    // snapshot (DecisionNetHandler::offers / jobOffers, neuralEconomy.h:218-224)
    unsigned int snapshot_time = 0;
    std::vector<std::weak_ptr<const Offer>> offers;
    std::vector<std::weak_ptr<const JobOffer>> jobOffers;
    // last-seen counters of snapshot entries (sampled while the objects are alive)
    bool sample = false;
    std::vector<uint32_t> m_left, m_taken, j_left, j_taken;
    std::vector<uint8_t> m_final;  // 1 once the owner has withdrawn the entry (left value frozen)

    void sample_counters() {
        if (!sample) return;
        for (size_t n = 0; n < offers.size(); n++) {
            if (m_final[n]) continue;
            auto o = offers[n].lock();
            if (o) { m_left[n] = o->amountLeft; m_taken[n] = o->amountTaken; }
            else   { m_left[n] = 0; }  // destroyed by its owner's flush (agent.cpp:21)
        }
        for (size_t n = 0; n < jobOffers.size(); n++) {
            auto o = jobOffers[n].lock();
            // frozen at the end of the person phase (see ReplayFirm::confirm_synchronized)
            if (o && !jobs_frozen) { j_left[n] = o->amountLeft; j_taken[n] = o->amountTaken; }
        }
    }
    bool jobs_frozen = false;
};

// DecisionNetHandler::synchronize_time → time_step → update_encodedOffers /
// update_encodedJobOffers (decisionNetHandler.cpp:303-308, 294-301, 236-275): the first
// decision call of a step copies both market vectors.
void synchronize(StepContext* ctx, Economy* economy) {
    if (economy->get_time() > ctx->snapshot_time) {
        ctx->offers = economy->get_market();
        ctx->jobOffers = economy->get_jobMarket();
        ctx->snapshot_time = economy->get_time();
        ctx->m_left.assign(ctx->offers.size(), 0);
        ctx->m_taken.assign(ctx->offers.size(), 0);
        ctx->m_final.assign(ctx->offers.size(), 0);
        ctx->j_left.assign(ctx->jobOffers.size(), 0);
        ctx->j_taken.assign(ctx->jobOffers.size(), 0);
        ctx->jobs_frozen = false;
        ctx->sample_counters();
    }
}

inline bool map_index(int32_t raw, size_t count, uint32_t flags, size_t* out) {
    if (count == 0) return false;
    if (flags & FASTACE_IDX_MODULO) {
        *out = (size_t)((uint32_t)raw % (uint32_t)count);
        return true;
    }
    if (raw < 0 || (size_t)raw >= count) return false;
    *out = (size_t)raw;
    return true;
}

class ReplayPerson : public PersonDecisionMaker {
public:
    ReplayPerson(StepContext* ctx, int id) : ctx(ctx), id(id) {}

    // get_joboffers_to_request + create_joboffer_requests (decisionNetHandler.cpp:446-493)
    std::vector<Order<JobOffer>> choose_jobs() override {
        auto parent_ = parent.lock();
        synchronize(ctx, parent_->get_economy());
        std::vector<Order<JobOffer>> toRequest;
        if (ctx->jobOffers.empty()) return toRequest;  // :476-480
        for (int i = 0; i < ctx->S; i++) {
            if (ctx->p_job_take[i * ctx->P + id]) {
                size_t n;
                if (map_index(ctx->p_job_idx[i * ctx->P + id], ctx->jobOffers.size(), ctx->flags, &n))
                    toRequest.push_back(Order<JobOffer>(ctx->jobOffers[n], 1));
            }
        }
        return toRequest;
    }

    // get_offers_to_request + create_offer_requests (decisionNetHandler.cpp:368-416)
    std::vector<Order<Offer>> choose_goods() override {
        auto parent_ = parent.lock();
        synchronize(ctx, parent_->get_economy());
        ctx->sample_counters();
        std::vector<Order<Offer>> toRequest;
        if (ctx->offers.empty()) return toRequest;  // :398-403
        for (int i = 0; i < ctx->S; i++) {
            if (ctx->p_good_take[i * ctx->P + id]) {
                size_t n;
                if (map_index(ctx->p_good_idx[i * ctx->P + id], ctx->offers.size(), ctx->flags, &n))
                    toRequest.push_back(Order<Offer>(ctx->offers[n], 1));
            }
        }
        return toRequest;
    }

    // NeuralPersonDecisionMaker::choose_goods_to_consume (neuralPersonDecisionMaker.cpp:93-111)
    Eigen::ArrayXd choose_goods_to_consume() override {
        auto parent_ = parent.lock();
        synchronize(ctx, parent_->get_economy());
        ctx->sample_counters();
        Eigen::ArrayXd proportions(ctx->G);
        for (int g = 0; g < ctx->G; g++) proportions(g) = (double)ctx->p_consume[g * ctx->P + id];
        Eigen::ArrayXd to_consume = parent_->get_inventory() * proportions;
        double util = parent_->u(to_consume);
        ctx->p_reward[id] = util;
        return to_consume;
    }

    StepContext* ctx;
    int id;
};

class ReplayFirm : public FirmDecisionMaker {
public:
    ReplayFirm(StepContext* ctx, int id) : ctx(ctx), id(id) {}

    // NeuralFirmDecisionMaker::confirm_synchronized + record_profit
    // (neuralFirmDecisionMaker.cpp:20-33, 65-74)
    void confirm_synchronized() {
        auto parent_ = parent.lock();
        synchronize(ctx, parent_->get_economy());
        // Job counters are final once the person phase is over; a firm's own
        // check_myJobOffers (firm.cpp:93-103, runs before its first decision) may raise
        // amountLeft again, so they are only sampled from person callbacks.
        ctx->jobs_frozen = true;
        if (parent_->get_time() > time) {
            if (time > 0) {
                ctx->f_profit[id] = parent_->get_money() - last_money;
            } else {
                ctx->f_profit[id] = 0.0;
            }
            last_money = parent_->get_money();
            time++;
        }
    }

    // firm_get_offers_to_request (decisionNetHandler.cpp:419-443)
    std::vector<Order<Offer>> choose_goods() override {
        confirm_synchronized();
        ctx->sample_counters();
        std::vector<Order<Offer>> toRequest;
        if (ctx->offers.empty()) return toRequest;
        for (int i = 0; i < ctx->S; i++) {
            if (ctx->f_good_take[i * ctx->F + id]) {
                size_t n;
                if (map_index(ctx->f_good_idx[i * ctx->F + id], ctx->offers.size(), ctx->flags, &n))
                    toRequest.push_back(Order<Offer>(ctx->offers[n], 1));
            }
        }
        return toRequest;
    }

    // NeuralFirmDecisionMaker::choose_production_inputs (neuralFirmDecisionMaker.cpp:95-108)
    Eigen::ArrayXd choose_production_inputs() override {
        confirm_synchronized();
        ctx->sample_counters();
        auto parent_ = parent.lock();
        Eigen::ArrayXd proportions(ctx->G);
        for (int g = 0; g < ctx->G; g++) proportions(g) = (double)ctx->f_prod[g * ctx->F + id];
        return parent_->get_inventory() * proportions;
    }

    // NeuralFirmDecisionMaker::choose_good_offers (neuralFirmDecisionMaker.cpp:111-146)
    // with DecisionNetHandler::choose_offers' amount = proportion * inventory (:586-592)
    std::vector<std::shared_ptr<Offer>> choose_good_offers() override {
        confirm_synchronized();
        auto parent_ = parent.lock();
        // sell_goods() withdraws the firm's old offers right after this call
        // (profitMaxer.cpp:74-86): freeze their final counters now.
        if (ctx->sample) {
            ctx->sample_counters();
            for (size_t n = 0; n < ctx->offers.size(); n++) {
                auto o = ctx->offers[n].lock();
                if (o && o->offerer.lock().get() == static_cast<Agent*>(parent_.get())) ctx->m_final[n] = 1;
            }
        }
        Eigen::ArrayXd amtProp(ctx->G), prices(ctx->G);
        for (int g = 0; g < ctx->G; g++) {
            amtProp(g) = (double)ctx->f_offer_amt[g * ctx->F + id];
            prices(g) = (double)ctx->f_offer_price[g * ctx->F + id];
        }
        Eigen::ArrayXd amounts = amtProp * parent_->get_inventory();
        Eigen::ArrayXi numOffers = (amounts / AMOUNT_PER_OFFER).cast<int>();
        int numGoods = amounts.size();
        std::vector<std::shared_ptr<Offer>> offers;
        for (int i = 0; i < numGoods; i++) {
            if (numOffers(i) > 0) {
                Eigen::ArrayXd quantities = Eigen::ArrayXd::Zero(numGoods);
                quantities(i) = AMOUNT_PER_OFFER;
                offers.push_back(std::make_shared<Offer>(parent, numOffers(i), quantities, prices(i) / AMOUNT_PER_OFFER));
            }
        }
        return offers;
    }

    // NeuralFirmDecisionMaker::choose_job_offers (neuralFirmDecisionMaker.cpp:149-180)
    // with the wage clip of DecisionNetHandler::choose_job_offers (:626-635)
    std::vector<std::shared_ptr<JobOffer>> choose_job_offers() override {
        confirm_synchronized();
        double laborAmount = (double)ctx->f_job_labor[id];
        double wage = (double)ctx->f_job_wage[id];
        if (wage > constants::largeNumber) wage = constants::largeNumber;
        int numOffers = laborAmount / LABOR_AMOUNT_PER_OFFER;
        if (numOffers > 0) {
            std::vector<std::shared_ptr<JobOffer>> offers = {
                std::make_shared<JobOffer>(parent, numOffers, LABOR_AMOUNT_PER_OFFER, wage / LABOR_AMOUNT_PER_OFFER)};
            return offers;
        }
        return {};
    }

    StepContext* ctx;
    int id;
    double last_money = 0.0;
    unsigned int time = 0;
};

struct RefEconomyBox {
    std::shared_ptr<SeededEconomy> economy;
    StepContext ctx;
    std::vector<std::shared_ptr<UtilMaxer>> persons;   // creation order = agent id
    std::vector<std::shared_ptr<ProfitMaxer>> firms;
    std::vector<std::shared_ptr<ReplayFirm>> firmDMs;
    std::vector<double> scratch_reward, scratch_profit;
    mutable std::unordered_map<const void*, int> person_ids, firm_ids;   // agent object -> creation index (built on first use)
};

}  // namespace

struct fastace_ref {
    fastace_dims_t dims;
    std::vector<RefEconomyBox> econ;
};

static int person_id(const RefEconomyBox& b, const Person* p) {
    if (b.person_ids.empty())
        for (size_t i = 0; i < b.persons.size(); i++) b.person_ids[static_cast<const Person*>(b.persons[i].get())] = (int)i;
    auto it = b.person_ids.find(p);
    return it == b.person_ids.end() ? -1 : it->second;
}
static int firm_id(const RefEconomyBox& b, const Agent* f) {
    if (b.firm_ids.empty())
        for (size_t i = 0; i < b.firms.size(); i++) b.firm_ids[static_cast<const Agent*>(b.firms[i].get())] = (int)i;
    auto it = b.firm_ids.find(f);
    return it == b.firm_ids.end() ? -1 : it->second;
}

static int g_ref_util_kind = FASTACE_FN_CES, g_ref_prod_kind = FASTACE_FN_CES;

// the reference's own function objects for a family, parameters written through public members
static std::shared_ptr<VecToScalar> make_function(int kind, double tfp, const Eigen::ArrayXd& share,
                                                  const Eigen::ArrayXd& theta, double rho) {
    switch (kind) {
        case FASTACE_FN_LINEAR: return std::make_shared<Linear>(share);
        case FASTACE_FN_COBB_DOUGLAS: return std::make_shared<CobbDouglas>(tfp, share);
        case FASTACE_FN_STONE_GEARY: return std::make_shared<StoneGeary>(tfp, share, theta);
        case FASTACE_FN_LEONTIEF: return std::make_shared<Leontief>(share);
        default: {
            auto ces = std::make_shared<CES>(1.0, share, 0.5);
            ces->tfp = tfp; ces->shareParams = share; ces->substitutionParam = rho;
            return ces;
        }
    }
}

extern "C" {

void fastace_ref_set_function_kinds(int util_kind, int prod_kind) { g_ref_util_kind = util_kind; g_ref_prod_kind = prod_kind; }

// Builds E reference economies from a host fastace_state_t (markets must be empty: the
// reference has no way to load a mid-episode book).  Economy e's rng is seeded with
// seed + e.  CES objects are constructed and then given the stored normalised
// shares / rho verbatim through their public members (vecToScalar.h:89-91).
fastace_ref* fastace_ref_create(const fastace_dims_t* dims, const fastace_state_t* st, uint32_t seed) {
    const int E = dims->num_econ, P = dims->num_persons, F = dims->num_firms, G = dims->num_goods;
    auto* h = new fastace_ref;
    h->dims = *dims;
    h->econ.resize(E);
    for (int e = 0; e < E; e++) {
        RefEconomyBox& b = h->econ[e];
        std::vector<std::string> goods(G);
        for (int g = 0; g < G; g++) goods[g] = "good" + std::to_string(g);
        b.economy = std::make_shared<SeededEconomy>(goods, seed + (uint32_t)e);
        b.ctx.P = P; b.ctx.F = F; b.ctx.G = G; b.ctx.S = dims->stack_size;
        for (int p = 0; p < P; p++) {
            Eigen::ArrayXd inv(G), share(G + 1);
            for (int g = 0; g < G; g++) inv(g) = st->p_inv[((size_t)e * G + g) * P + p];
            for (int i = 0; i <= G; i++) share(i) = st->p_util_share[((size_t)e * (G + 1) + i) * P + p];
            Eigen::ArrayXd theta = Eigen::ArrayXd::Zero(G + 1);
            if (st->p_util_theta)
                for (int i = 0; i <= G; i++) theta(i) = st->p_util_theta[((size_t)e * (G + 1) + i) * P + p];
            std::shared_ptr<VecToScalar> ces = make_function(g_ref_util_kind, st->p_util_tfp[(size_t)e * P + p], share, theta,
                                                             st->p_util_rho[(size_t)e * P + p]);
            auto person = UtilMaxer::init(
                b.economy.get(), inv, st->p_money[(size_t)e * P + p], ces, 0.9,
                std::make_shared<ReplayPerson>(&b.ctx, p));
            b.persons.push_back(person);
        }
        for (int f = 0; f < F; f++) {
            Eigen::ArrayXd inv(G);
            for (int g = 0; g < G; g++) inv(g) = st->f_inv[((size_t)e * G + g) * F + f];
            std::vector<double> tfps(G, 1.0), elast(G, 0.5);
            std::vector<Eigen::ArrayXd> shares;
            for (int g = 0; g < G; g++) {
                Eigen::ArrayXd s(G + 1);
                for (int i = 0; i <= G; i++) s(i) = st->f_prod_share[(((size_t)e * G + g) * (G + 1) + i) * F + f];
                shares.push_back(s);
            }
            // one VecToScalar per output good, summed (create_CES_VecToVec pattern, vecToVec.cpp:34-53)
            std::vector<std::shared_ptr<VecToVec>> inner(G);
            for (int g = 0; g < G; g++) {
                Eigen::ArrayXd theta = Eigen::ArrayXd::Zero(G + 1);
                if (st->f_prod_theta)
                    for (int i = 0; i <= G; i++) theta(i) = st->f_prod_theta[(((size_t)e * G + g) * (G + 1) + i) * F + f];
                inner[g] = std::make_shared<VToVFromVToS<VecToScalar>>(
                    make_function(g_ref_prod_kind, st->f_prod_tfp[((size_t)e * G + g) * F + f], shares[g], theta,
                                  st->f_prod_rho[((size_t)e * G + g) * F + f]), G, g);
            }
            auto prod = std::make_shared<SumOfVecToVec>(inner);
            auto dm = std::make_shared<ReplayFirm>(&b.ctx, f);
            auto firm = ProfitMaxer::init(
                b.economy.get(), std::vector<std::shared_ptr<Agent>>(), inv,
                st->f_money[(size_t)e * F + f], prod, dm);
            b.firms.push_back(firm);
            b.firmDMs.push_back(dm);
        }
        b.scratch_reward.assign(P, 0.0);
        b.scratch_profit.assign(F, 0.0);
    }
    return h;
}

void fastace_ref_destroy(fastace_ref* h) { delete h; }

static void step_one(fastace_ref* h, int e, const fastace_actions_t* a, const fastace_step_out_t* out,
                     uint32_t flags, int32_t* perm_person_out, int32_t* perm_firm_out) {
    const int P = h->dims.num_persons, F = h->dims.num_firms, G = h->dims.num_goods, S = h->dims.stack_size;
    RefEconomyBox& b = h->econ[e];
    StepContext& c = b.ctx;
    c.flags = flags;
    c.p_job_idx = a->p_job_idx + (size_t)e * S * P;
    c.p_job_take = a->p_job_take + (size_t)e * S * P;
    c.p_good_idx = a->p_good_idx + (size_t)e * S * P;
    c.p_good_take = a->p_good_take + (size_t)e * S * P;
    c.p_consume = a->p_consume + (size_t)e * G * P;
    c.f_good_idx = a->f_good_idx + (size_t)e * S * F;
    c.f_good_take = a->f_good_take + (size_t)e * S * F;
    c.f_prod = a->f_prod + (size_t)e * G * F;
    c.f_offer_amt = a->f_offer_amt + (size_t)e * G * F;
    c.f_offer_price = a->f_offer_price + (size_t)e * G * F;
    c.f_job_labor = a->f_job_labor + (size_t)e * F;
    c.f_job_wage = a->f_job_wage + (size_t)e * F;
    c.p_reward = (out && out->p_reward) ? out->p_reward + (size_t)e * P : b.scratch_reward.data();
    c.f_profit = (out && out->f_profit) ? out->f_profit + (size_t)e * F : b.scratch_profit.data();
    c.sample = out && (out->old_m_left || out->old_m_taken || out->old_j_left || out->old_j_taken);

    b.economy->time_step();  // <- the reference's hot path, unmodified (economy.cpp:95-139)

    if (perm_person_out) {
        const auto& order = b.economy->person_order();
        for (int r = 0; r < P; r++) perm_person_out[(size_t)e * P + r] = person_id(b, order[r].get());
    }
    if (perm_firm_out) {
        const auto& order = b.economy->firm_order();
        for (int r = 0; r < F; r++) perm_firm_out[(size_t)e * F + r] = firm_id(b, order[r].get());
    }
    if (c.sample) {
        for (size_t n = 0; n < c.offers.size(); n++) {
            if (out->old_m_left) out->old_m_left[(size_t)e * F * G + n] = c.m_left[n];
            if (out->old_m_taken) out->old_m_taken[(size_t)e * F * G + n] = c.m_taken[n];
        }
        for (size_t n = 0; n < c.jobOffers.size(); n++) {
            if (out->old_j_left) out->old_j_left[(size_t)e * F + n] = c.j_left[n];
            if (out->old_j_taken) out->old_j_taken[(size_t)e * F + n] = c.j_taken[n];
        }
    }
}

// One Economy::time_step() per economy.  The visiting orders the reference chose (its own
// std::shuffle on its own seeded rng) are written to perm_person_out [E][P] /
// perm_firm_out [E][F] so that the same orders can be injected elsewhere; the perm_*
// members of `a` are ignored.  Economies are independent and are spread over `nthreads`
// host threads (each economy itself runs the single-threaded branch).
int fastace_ref_step(fastace_ref* h, const fastace_actions_t* a, const fastace_step_out_t* out,
                     uint32_t flags, int32_t* perm_person_out, int32_t* perm_firm_out, int nthreads) {
    const int E = h->dims.num_econ;
    if (nthreads <= 1) {
        for (int e = 0; e < E; e++) step_one(h, e, a, out, flags, perm_person_out, perm_firm_out);
        return 0;
    }
    std::vector<std::thread> threads;
    for (int t = 0; t < nthreads; t++) {
        threads.emplace_back([=]() {
            for (int e = t; e < E; e += nthreads) step_one(h, e, a, out, flags, perm_person_out, perm_firm_out);
        });
    }
    for (auto& t : threads) t.join();
    return 0;
}

// Exports the reference objects' state through the reference's public read API
// (base.h:90-106, 147-150, 205, 245) into a host fastace_state_t.  NULL members skipped.
int fastace_ref_get_state(const fastace_ref* h, fastace_state_t* st, uint32_t* time_out) {
    const int E = h->dims.num_econ, P = h->dims.num_persons, F = h->dims.num_firms, G = h->dims.num_goods;
    for (int e = 0; e < E; e++) {
        const RefEconomyBox& b = h->econ[e];
        for (int p = 0; p < P; p++) {
            const auto& person = b.persons[p];
            if (st->p_money) st->p_money[(size_t)e * P + p] = person->get_money();
            if (st->p_labor) st->p_labor[(size_t)e * P + p] = person->get_laborSupplied();
            for (int g = 0; g < G; g++)
                if (st->p_inv) st->p_inv[((size_t)e * G + g) * P + p] = person->get_inventory()(g);
            auto ces = std::dynamic_pointer_cast<const CES>(person->get_utilFunc());
            if (ces) {
                if (st->p_util_tfp) st->p_util_tfp[(size_t)e * P + p] = ces->tfp;
                if (st->p_util_rho) st->p_util_rho[(size_t)e * P + p] = ces->substitutionParam;
                for (int i = 0; i <= G; i++)
                    if (st->p_util_share) st->p_util_share[((size_t)e * (G + 1) + i) * P + p] = ces->shareParams(i);
            }
        }
        for (int f = 0; f < F; f++) {
            const auto& firm = b.firms[f];
            if (st->f_money) st->f_money[(size_t)e * F + f] = firm->get_money();
            if (st->f_labor) st->f_labor[(size_t)e * F + f] = firm->get_laborHired();
            if (st->f_last_money) st->f_last_money[(size_t)e * F + f] = b.firmDMs[f]->last_money;
            for (int g = 0; g < G; g++)
                if (st->f_inv) st->f_inv[((size_t)e * G + g) * F + f] = firm->get_inventory()(g);
            auto prod = std::static_pointer_cast<const SumOfVecToVec>(firm->get_prodFunc());
            for (int g = 0; g < G; g++) {
                auto ces = std::dynamic_pointer_cast<CES>(
                    std::static_pointer_cast<const VToVFromVToS<VecToScalar>>(prod->innerFunctions[g])->vecToScalar);
                if (!ces) continue;
                if (st->f_prod_tfp) st->f_prod_tfp[((size_t)e * G + g) * F + f] = ces->tfp;
                if (st->f_prod_rho) st->f_prod_rho[((size_t)e * G + g) * F + f] = ces->substitutionParam;
                for (int i = 0; i <= G; i++)
                    if (st->f_prod_share)
                        st->f_prod_share[(((size_t)e * G + g) * (G + 1) + i) * F + f] = ces->shareParams(i);
            }
        }
        const auto& market = b.economy->get_market();
        if (st->m_count) st->m_count[e] = (int32_t)market.size();
        for (size_t n = 0; n < market.size() && n < (size_t)(F * G); n++) {
            auto o = market[n].lock();
            size_t k = (size_t)e * F * G + n;
            int good = -1;
            for (int g = 0; g < G; g++) if (o->quantities(g) > 0) good = g;
            if (st->m_owner) st->m_owner[k] = firm_id(b, o->offerer.lock().get());
            if (st->m_good) st->m_good[k] = good;
            if (st->m_left) st->m_left[k] = o->amountLeft;
            if (st->m_taken) st->m_taken[k] = o->amountTaken;
            if (st->m_price) st->m_price[k] = o->price;
        }
        const auto& jobs = b.economy->get_jobMarket();
        if (st->j_count) st->j_count[e] = (int32_t)jobs.size();
        for (size_t n = 0; n < jobs.size() && n < (size_t)F; n++) {
            auto o = jobs[n].lock();
            size_t k = (size_t)e * F + n;
            if (st->j_owner) st->j_owner[k] = firm_id(b, o->offerer.lock().get());
            if (st->j_left) st->j_left[k] = o->amountLeft;
            if (st->j_taken) st->j_taken[k] = o->amountTaken;
            if (st->j_wage) st->j_wage[k] = o->wage;
        }
    }
    if (time_out) *time_out = h->econ.empty() ? 0 : h->econ[0].economy->get_time();
    return 0;
}

// Known-answer hooks for the function plugins (vecToScalar.cpp:45-47, 105-118).
double fastace_ref_ces_f(double tfp, const double* share_raw, double elasticity, const double* x, int n) {
    Eigen::ArrayXd s(n), q(n);
    for (int i = 0; i < n; i++) { s(i) = share_raw[i]; q(i) = x[i]; }
    CES ces(tfp, s, elasticity);
    return ces.f(q);
}
// exports the constructor's normalisation so it can be compared with ours
void fastace_ref_ces_params(const double* share_raw, double elasticity, int n, double* share_out, double* rho_out) {
    Eigen::ArrayXd s(n);
    for (int i = 0; i < n; i++) s(i) = share_raw[i];
    CES ces(1.0, s, elasticity);
    for (int i = 0; i < n; i++) share_out[i] = ces.shareParams(i);
    *rho_out = ces.substitutionParam;
}
double fastace_ref_cobb_douglas_f(double tfp, const double* elast, const double* x, int n) {
    Eigen::ArrayXd e(n), q(n);
    for (int i = 0; i < n; i++) { e(i) = elast[i]; q(i) = x[i]; }
    CobbDouglas cd(tfp, e);
    return cd.f(q);
}
// std::shuffle / default_random_engine KAT (economy.cpp:110-111 uses exactly these)
void fastace_ref_shuffle_kat(unsigned int seed, int n, int rounds, int32_t* out) {
    std::default_random_engine rng(seed);
    std::vector<int32_t> v(n);
    for (int i = 0; i < n; i++) v[i] = i;
    for (int r = 0; r < rounds; r++) {
        std::shuffle(v.begin(), v.end(), rng);
        for (int i = 0; i < n; i++) out[r * n + i] = v[i];
    }
}

}  // extern "C"
