// fastace_b200.hpp — C++ host side above the C ABI (include/fastace_b200.h), header-only.
//
// Mirrors the reference's C++ surface for the time-step path, batched over E economies, with the
// reference's names, argument meaning and conventions:
//
//   fastace::BatchedEconomy               <->  Economy                      src/base/base.h:80-133
//     ::init(goods, E, P, F, S, device)        Economy(goods) + util::create<T>(...) for every agent
//                                              (economy.cpp:3-27, util.h:41-46); factory, like T::init
//     ::time_step()  -> bool                   Economy::time_step (economy.cpp:95-139): false = did not act
//     ::get_time / get_goods / get_numGoods    base.h:90-98
//     ::get_market / get_jobMarket             base.h:99-100 (market order, snapshot of the posted book)
//     ::person_* / firm_* getters              Agent::get_money / get_inventory (base.h:147-150),
//                                              Person::get_laborSupplied (:205), Firm::get_laborHired (:245)
//   fastace::BatchedPersonDecisionMaker   <->  PersonDecisionMaker         src/persons/utilMaxer.h:11-28
//     choose_jobs / choose_goods / choose_goods_to_consume
//   fastace::BatchedFirmDecisionMaker     <->  FirmDecisionMaker           src/firms/profitMaxer.h:11-24
//     choose_goods / choose_production_inputs / choose_good_offers / choose_job_offers
//
// Difference that batching forces: the reference calls a decision maker from inside one
// agent's turn and gets that agent's orders back; here each plugin method is called ONCE per step and fills
// the decisions of every agent of every economy (the action arrays of fastace_actions_t), before the fused
// step runs on the GPU.  A decision maker belongs to exactly one economy (utilMaxer.cpp:34).
// Conventions kept: no exceptions on the step path; bool "did not act" returns; last_error() says why.
// There is no CPU path: init() returns nullptr when no GPU / library is present.
#ifndef FASTACE_B200_HPP
#define FASTACE_B200_HPP

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "fastace_b200.h"

namespace fastace {

class BatchedEconomy;

// host-side action arrays of one step, layouts of fastace_actions_t
struct StepActions {
    std::vector<int32_t> perm_person, perm_firm, p_job_idx, p_good_idx, f_good_idx;
    std::vector<uint8_t> p_job_take, p_good_take, f_good_take;
    std::vector<float> p_consume, f_prod, f_offer_amt, f_offer_price, f_job_labor, f_job_wage;
    void resize(const fastace_dims_t& d) {
        const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods, S = d.stack_size;
        perm_person.assign(E * P, 0); perm_firm.assign(E * F, 0);
        p_job_idx.assign(E * S * P, 0); p_job_take.assign(E * S * P, 0);
        p_good_idx.assign(E * S * P, 0); p_good_take.assign(E * S * P, 0);
        p_consume.assign(E * G * P, 0.f);
        f_good_idx.assign(E * S * F, 0); f_good_take.assign(E * S * F, 0);
        f_prod.assign(E * G * F, 0.f); f_offer_amt.assign(E * G * F, 0.f); f_offer_price.assign(E * G * F, 0.f);
        f_job_labor.assign(E * F, 0.f); f_job_wage.assign(E * F, 0.f);
    }
    fastace_actions_t view() const {
        return fastace_actions_t{perm_person.data(), perm_firm.data(), p_job_idx.data(), p_job_take.data(),
                                 p_good_idx.data(), p_good_take.data(), p_consume.data(), f_good_idx.data(),
                                 f_good_take.data(), f_prod.data(), f_offer_amt.data(), f_offer_price.data(),
                                 f_job_labor.data(), f_job_wage.data()};
    }
};

// host mirror of the whole state (fastace_state_t layouts)
struct HostState {
    std::vector<double> p_money, p_inv, p_labor, p_util_tfp, p_util_share, p_util_rho, f_money, f_inv, f_labor,
        f_last_money, f_prod_tfp, f_prod_share, f_prod_rho, m_price, j_wage, p_util_theta, f_prod_theta;
    std::vector<int32_t> m_count, m_owner, m_good, j_count, j_owner;
    std::vector<uint32_t> m_left, m_taken, j_left, j_taken;
    void resize(const fastace_dims_t& d) {
        const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods;
        p_money.assign(E * P, 0); p_inv.assign(E * G * P, 0); p_labor.assign(E * P, 0);
        p_util_tfp.assign(E * P, 0); p_util_share.assign(E * (G + 1) * P, 0); p_util_rho.assign(E * P, 0);
        f_money.assign(E * F, 0); f_inv.assign(E * G * F, 0); f_labor.assign(E * F, 0); f_last_money.assign(E * F, 0);
        f_prod_tfp.assign(E * G * F, 0); f_prod_share.assign(E * G * (G + 1) * F, 0); f_prod_rho.assign(E * G * F, 0);
        m_count.assign(E, 0); m_owner.assign(E * F * G, 0); m_good.assign(E * F * G, 0);
        m_left.assign(E * F * G, 0); m_taken.assign(E * F * G, 0); m_price.assign(E * F * G, 0);
        j_count.assign(E, 0); j_owner.assign(E * F, 0); j_left.assign(E * F, 0); j_taken.assign(E * F, 0);
        j_wage.assign(E * F, 0);
        p_util_theta.assign(E * (G + 1) * P, 0); f_prod_theta.assign(E * G * (G + 1) * F, 0);
    }
    fastace_state_t view() {
        return fastace_state_t{p_money.data(), p_inv.data(), p_labor.data(), p_util_tfp.data(), p_util_share.data(),
                               p_util_rho.data(), f_money.data(), f_inv.data(), f_labor.data(), f_last_money.data(),
                               f_prod_tfp.data(), f_prod_share.data(), f_prod_rho.data(), m_count.data(),
                               m_owner.data(), m_good.data(), m_left.data(), m_taken.data(), m_price.data(),
                               j_count.data(), j_owner.data(), j_left.data(), j_taken.data(), j_wage.data(),
                               p_util_theta.data(), f_prod_theta.data()};
    }
};

// Offer / JobOffer as a caller of Economy::get_market() sees them (base.h:21-66)
struct OfferView { int offerer; unsigned amountLeft, amountTaken; int good; double price; };
struct JobOfferView { int offerer; unsigned amountLeft, amountTaken; double labor, wage; };

class BatchedPersonDecisionMaker {   // PersonDecisionMaker, utilMaxer.h:11-28
public:
    virtual ~BatchedPersonDecisionMaker() {}
    // fill a.p_job_idx / a.p_job_take for every person of every economy   (choose_jobs)
    virtual void choose_jobs(StepActions& a) = 0;
    // fill a.p_good_idx / a.p_good_take                                     (choose_goods)
    virtual void choose_goods(StepActions& a) = 0;
    // fill a.p_consume: proportion of each good's inventory to consume      (choose_goods_to_consume)
    virtual void choose_goods_to_consume(StepActions& a) = 0;
    std::weak_ptr<BatchedEconomy> parent;
};

class BatchedFirmDecisionMaker {     // FirmDecisionMaker, profitMaxer.h:11-24
public:
    virtual ~BatchedFirmDecisionMaker() {}
    virtual void choose_goods(StepActions& a) = 0;               // a.f_good_idx / a.f_good_take
    virtual void choose_production_inputs(StepActions& a) = 0;   // a.f_prod
    virtual void choose_good_offers(StepActions& a) = 0;         // a.f_offer_amt / a.f_offer_price
    virtual void choose_job_offers(StepActions& a) = 0;          // a.f_job_labor / a.f_job_wage
    std::weak_ptr<BatchedEconomy> parent;
};

class BatchedEconomy : public std::enable_shared_from_this<BatchedEconomy> {
public:
    // goods.size() = numGoods, like Economy(std::vector<std::string> goods) (economy.cpp:3-6)
    static std::shared_ptr<BatchedEconomy> init(std::vector<std::string> goods, unsigned numEconomies,
                                                unsigned numPersons, unsigned numFirms, unsigned stackSize,
                                                int device = 0, uint32_t seed = 0) {
        std::shared_ptr<BatchedEconomy> self(new BatchedEconomy());
        self->goods_ = goods;
        self->dims_ = fastace_dims_t{(int32_t)numEconomies, (int32_t)numPersons, (int32_t)numFirms,
                                     (int32_t)goods.size(), (int32_t)stackSize};
        self->seed_ = seed;
        if (fastace_env_create(&self->dims_, device, &self->env_) != FASTACE_OK) {
            last_error_() = fastace_last_error();
            return nullptr;
        }
        self->state_.resize(self->dims_);
        self->actions_.resize(self->dims_);
        self->rng_state_.assign(numEconomies, 0);
        self->p_reward_.assign((size_t)numEconomies * numPersons, 0.0);
        self->f_profit_.assign((size_t)numEconomies * numFirms, 0.0);
        return self;
    }
    ~BatchedEconomy() { if (env_) fastace_env_destroy(env_); }
    BatchedEconomy(const BatchedEconomy&) = delete;
    BatchedEconomy& operator=(const BatchedEconomy&) = delete;

    static const std::string& last_error() { return last_error_(); }

    // CustomScenario::setup (src/neural/neuralScenarios.cpp:93-161) for every economy; numGoods must be 2
    bool setup(const fastace_custom_scenario_params_t& params, std::vector<double>* discount = nullptr) {
        fastace_state_t v = state_.view();
        std::vector<double> disc((size_t)dims_.num_econ * dims_.num_persons);
        if (!ok(fastace_scenario_custom_init(&dims_, &params, seed_, &v, disc.data()))) return false;
        if (discount) *discount = disc;
        return set_state(state_);
    }
    // arbitrary initial state (what util::create<UtilMaxer/ProfitMaxer>(...) arguments carry in the reference)
    bool set_state(HostState& s, uint32_t time = 0) {
        fastace_state_t v = s.view();
        if (!ok(fastace_env_set_state(env_, &v, time))) return false;
        if (&s != &state_) state_ = s;
        fresh_ = true;
        first_step_ = (time == 0);
        return true;
    }
    void set_decision_makers(std::shared_ptr<BatchedPersonDecisionMaker> p, std::shared_ptr<BatchedFirmDecisionMaker> f) {
        person_dm_ = p; firm_dm_ = f;
        p->parent = shared_from_this(); f->parent = shared_from_this();
    }
    bool set_function_kinds(fastace_function_kind_t util, fastace_function_kind_t prod) {
        return ok(fastace_env_set_function_kinds(env_, util, prod));
    }

    // Economy::time_step (economy.cpp:95-139): shuffle persons, then firms; every agent decides and acts; flush.
    // Returns false (and changes nothing) if there is no decision maker or the device call failed.
    bool time_step(uint32_t flags = FASTACE_IDX_ABSOLUTE) {
        if (!person_dm_ || !firm_dm_) { last_error_() = "no decision makers"; return false; }
        if (!ok(fastace_shuffle_orders(&dims_, seed_, rng_state_.data(), actions_.perm_person.data(),
                                       actions_.perm_firm.data(), first_step_ ? 1 : 0))) return false;
        first_step_ = false;
        // the order the reference's agents ask their decision makers in (person.cpp:19-33, firm.cpp:23-46)
        person_dm_->choose_jobs(actions_);
        person_dm_->choose_goods(actions_);
        person_dm_->choose_goods_to_consume(actions_);
        firm_dm_->choose_goods(actions_);
        firm_dm_->choose_production_inputs(actions_);
        firm_dm_->choose_good_offers(actions_);
        firm_dm_->choose_job_offers(actions_);
        return step_with(actions_, flags);
    }
    // The same step taken phase by phase, so that every decision maker is asked when the reference would ask it and
    // sees the state the reference's plugin sees at that point through the read API: persons choose jobs and goods
    // (state at the start of the step), trade; choose_goods_to_consume sees money / laborSupplied / inventory after
    // the trades (neuralPersonDecisionMaker.cpp:93-111); the firm plugins see the economy after the person phase
    // (economy.cpp:118-123).  Bit-identical to time_step() for the same decisions.
    bool time_step_phased(uint32_t flags = FASTACE_IDX_ABSOLUTE) {
        if (!person_dm_ || !firm_dm_) { last_error_() = "no decision makers"; return false; }
        if (!ok(fastace_shuffle_orders(&dims_, seed_, rng_state_.data(), actions_.perm_person.data(),
                                       actions_.perm_firm.data(), first_step_ ? 1 : 0))) return false;
        first_step_ = false;
        const fastace_actions_t av = actions_.view();     // the vectors are sized once: pointers stay valid
        fastace_step_out_t none{};
        person_dm_->choose_jobs(actions_);
        person_dm_->choose_goods(actions_);
        if (!ok(fastace_env_step_host(env_, &av, &none, flags | FASTACE_STEP_PERSONS_TRADE))) return false;
        fresh_ = false;
        person_dm_->choose_goods_to_consume(actions_);
        fastace_step_out_t pout{};
        pout.p_reward = p_reward_.data();
        if (!ok(fastace_env_step_host(env_, &av, &pout, flags | FASTACE_STEP_PERSONS_CONSUME))) return false;
        fresh_ = false;
        firm_dm_->choose_goods(actions_);
        firm_dm_->choose_production_inputs(actions_);
        firm_dm_->choose_good_offers(actions_);
        firm_dm_->choose_job_offers(actions_);
        fastace_step_out_t fout{};
        fout.f_profit = f_profit_.data();
        if (!ok(fastace_env_step_host(env_, &av, &fout, flags | FASTACE_STEP_FIRMS))) return false;
        fresh_ = false;
        return true;
    }
    // the same step with caller-provided decisions and visiting orders
    bool step_with(const StepActions& a, uint32_t flags = FASTACE_IDX_ABSOLUTE) {
        fastace_actions_t av = a.view();
        fastace_step_out_t out{};
        out.p_reward = p_reward_.data();
        out.f_profit = f_profit_.data();
        if (!ok(fastace_env_step_host(env_, &av, &out, flags))) return false;
        fresh_ = false;
        return true;
    }

    // ---- read API ----
    unsigned get_time() const { uint32_t t = 0; fastace_env_time(env_, &t); return t; }
    const std::vector<std::string>& get_goods() const { return goods_; }
    unsigned get_numGoods() const { return (unsigned)dims_.num_goods; }
    unsigned get_numEconomies() const { return (unsigned)dims_.num_econ; }
    unsigned get_numPersons() const { return (unsigned)dims_.num_persons; }
    unsigned get_numFirms() const { return (unsigned)dims_.num_firms; }
    const fastace_dims_t& dims() const { return dims_; }
    fastace_env_t* handle() { return env_; }
    const StepActions& last_actions() const { return actions_; }

    double person_money(unsigned e, unsigned p) { sync(); return state_.p_money[(size_t)e * dims_.num_persons + p]; }
    double person_inventory(unsigned e, unsigned p, unsigned g) { sync(); return state_.p_inv[((size_t)e * dims_.num_goods + g) * dims_.num_persons + p]; }
    double person_laborSupplied(unsigned e, unsigned p) { sync(); return state_.p_labor[(size_t)e * dims_.num_persons + p]; }
    double firm_money(unsigned e, unsigned f) { sync(); return state_.f_money[(size_t)e * dims_.num_firms + f]; }
    double firm_inventory(unsigned e, unsigned f, unsigned g) { sync(); return state_.f_inv[((size_t)e * dims_.num_goods + g) * dims_.num_firms + f]; }
    double firm_laborHired(unsigned e, unsigned f) { sync(); return state_.f_labor[(size_t)e * dims_.num_firms + f]; }
    // rewards of the last step: utility (neuralPersonDecisionMaker.cpp:107-108), profit (neuralFirmDecisionMaker.cpp:65-74)
    double person_reward(unsigned e, unsigned p) const { return p_reward_[(size_t)e * dims_.num_persons + p]; }
    double firm_profit(unsigned e, unsigned f) const { return f_profit_[(size_t)e * dims_.num_firms + f]; }

    std::vector<OfferView> get_market(unsigned e) {
        sync();
        const size_t cap = (size_t)dims_.num_firms * dims_.num_goods;
        std::vector<OfferView> v((size_t)state_.m_count[e]);
        for (size_t n = 0; n < v.size(); n++)
            v[n] = OfferView{state_.m_owner[e * cap + n], state_.m_left[e * cap + n], state_.m_taken[e * cap + n],
                             state_.m_good[e * cap + n], state_.m_price[e * cap + n]};
        return v;
    }
    std::vector<JobOfferView> get_jobMarket(unsigned e) {
        sync();
        const size_t cap = (size_t)dims_.num_firms;
        std::vector<JobOfferView> v((size_t)state_.j_count[e]);
        for (size_t n = 0; n < v.size(); n++)
            v[n] = JobOfferView{state_.j_owner[e * cap + n], state_.j_left[e * cap + n], state_.j_taken[e * cap + n], 0.5,
                                state_.j_wage[e * cap + n]};
        return v;
    }
    // whole host mirror (refreshed from the device if a step ran since the last read)
    HostState& state() { sync(); return state_; }

private:
    BatchedEconomy() {}
    static std::string& last_error_() { static thread_local std::string s; return s; }
    bool ok(int status) {
        if (status == FASTACE_OK) return true;
        last_error_() = fastace_last_error();
        return false;
    }
    void sync() {
        if (fresh_) return;
        fastace_state_t v = state_.view();
        if (ok(fastace_env_get_state(env_, &v))) fresh_ = true;
    }

    fastace_env_t* env_ = nullptr;
    fastace_dims_t dims_{};
    std::vector<std::string> goods_;
    uint32_t seed_ = 0;
    bool fresh_ = true, first_step_ = true;
    HostState state_;
    StepActions actions_;
    std::vector<uint64_t> rng_state_;
    std::vector<double> p_reward_, f_profit_;
    std::shared_ptr<BatchedPersonDecisionMaker> person_dm_;
    std::shared_ptr<BatchedFirmDecisionMaker> firm_dm_;
};

}  // namespace fastace

#endif  // FASTACE_B200_HPP
