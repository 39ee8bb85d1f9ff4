// Parity test of the C++ host mirror (include/fastace_b200.hpp) written the way a test of the reference's own
// Economy would read: build economies, attach decision makers, call time_step(), read agents back.
// The checker is the CPU oracle (oracle/liboracle.so — test infrastructure), stepped on a host copy with the very
// same decisions and visiting orders.  Integers / counters / books bit-exact, fp64 state within 1e-5 relative
// (1e-9 in practice).  Needs a GPU: run by tests/test_cpp_host.py under the `gpu` marker.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/fastace_b200.hpp"
#include "../../oracle/fastace_oracle.h"

using namespace fastace;

namespace {

struct Lcg {   // tiny deterministic generator for the plugin decisions
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ULL + 3037000493ULL) {}
    uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
    float uniform() { return (float)(next() & 0xFFFFFF) / 16777216.0f; }
};

// persons: request random posted offers with probability 1/2, consume a random share
class RandomPerson : public BatchedPersonDecisionMaker {
public:
    explicit RandomPerson(uint64_t seed) : rng(seed) {}
    void choose_jobs(StepActions& a) override { draw(a.p_job_idx, a.p_job_take, /*jobs=*/true); }
    void choose_goods(StepActions& a) override { draw(a.p_good_idx, a.p_good_take, false); }
    void choose_goods_to_consume(StepActions& a) override {
        auto econ = parent.lock();
        const auto& d = econ->dims();
        for (auto& x : a.p_consume) x = rng.uniform();
        // in phased mode this is called after the trades: what the plugin reads is the post-trade inventory
        seen_inventory = econ->person_inventory(0, 0, 0) + econ->person_inventory(d.num_econ - 1, d.num_persons - 1, 1);
        seen_labor = 0.0;
        for (int p = 0; p < d.num_persons; p++) seen_labor += econ->person_laborSupplied(0, p);
    }
    double seen_inventory = 0.0, seen_labor = 0.0;
private:
    void draw(std::vector<int32_t>& idx, std::vector<uint8_t>& take, bool jobs) {
        auto econ = parent.lock();
        const auto& d = econ->dims();
        for (int e = 0; e < d.num_econ; e++) {
            // generate_offerIndices: randint(0, numOffers) (decisionNetHandler.cpp:327-365); empty market = no orders
            const int count = jobs ? (int)econ->get_jobMarket(e).size() : (int)econ->get_market(e).size();
            for (int i = 0; i < d.stack_size * d.num_persons; i++) {
                const size_t at = (size_t)e * d.stack_size * d.num_persons + i;
                idx[at] = count ? (int32_t)(rng.next() % (uint32_t)count) : 0;
                take[at] = count ? (uint8_t)(rng.next() & 1) : 0;
            }
        }
    }
    Lcg rng;
};

class RandomFirm : public BatchedFirmDecisionMaker {
public:
    explicit RandomFirm(uint64_t seed) : rng(seed) {}
    void choose_goods(StepActions& a) override {
        auto econ = parent.lock();
        const auto& d = econ->dims();
        for (int e = 0; e < d.num_econ; e++) {
            const int count = (int)econ->get_market(e).size();
            for (int i = 0; i < d.stack_size * d.num_firms; i++) {
                const size_t at = (size_t)e * d.stack_size * d.num_firms + i;
                a.f_good_idx[at] = count ? (int32_t)(rng.next() % (uint32_t)count) : 0;
                a.f_good_take[at] = count ? (uint8_t)(rng.next() & 1) : 0;
            }
        }
    }
    void choose_production_inputs(StepActions& a) override { for (auto& x : a.f_prod) x = 0.5f * rng.uniform(); }
    void choose_good_offers(StepActions& a) override {
        for (auto& x : a.f_offer_amt) x = rng.uniform();
        for (auto& x : a.f_offer_price) x = 0.5f + 2.0f * rng.uniform();
    }
    void choose_job_offers(StepActions& a) override {
        for (auto& x : a.f_job_labor) x = 10.0f * rng.uniform();
        for (auto& x : a.f_job_wage) x = 0.05f + 0.2f * rng.uniform();
    }
private:
    Lcg rng;
};

int failures = 0;
void expect(bool cond, const char* what, int step) {
    if (!cond) { std::printf("FAIL step %d: %s\n", step, what); failures++; }
}
template <typename T>
bool same(const std::vector<T>& a, const std::vector<T>& b) { return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0; }
bool close(const std::vector<double>& a, const std::vector<double>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) {
        const double tol = 1e-5 * std::fmax(std::fabs(b[i]), 1e-9);
        if (!(std::fabs(a[i] - b[i]) <= tol)) return false;
    }
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    const unsigned E = argc > 1 ? (unsigned)std::atoi(argv[1]) : 24, steps = argc > 2 ? (unsigned)std::atoi(argv[2]) : 8;
    const unsigned P = 37, F = 6, S = 7;

    // no decision makers: time_step() does not act and says why (bool convention, economy.cpp:97-106)
    auto economy = BatchedEconomy::init({"bread", "capital"}, E, P, F, S, /*device=*/0, /*seed=*/77);
    if (!economy) { std::printf("init failed: %s\n", BatchedEconomy::last_error().c_str()); return 2; }
    expect(!economy->time_step(), "time_step without decision makers must return false", 0);
    expect(economy->get_time() == 0, "time must not advance when the step did not act", 0);
    expect(economy->get_numGoods() == 2 && economy->get_goods()[1] == "capital", "goods", 0);

    fastace_custom_scenario_params_t params = create_scenario_params(P, F);
    std::vector<double> discount;
    expect(economy->setup(params, &discount), "setup", 0);
    auto persons = std::make_shared<RandomPerson>(1);
    economy->set_decision_makers(persons, std::make_shared<RandomFirm>(2));

    // the checker's copy of the world
    HostState ref = economy->state();
    const fastace_dims_t dims = economy->dims();
    std::vector<double> ref_reward((size_t)E * P), ref_profit((size_t)E * F);
    double money0 = 0;
    for (double m : ref.p_money) money0 += m;
    for (double m : ref.f_money) money0 += m;

    unsigned long trades = 0;
    double labour_seen_by_consumption = 0.0;
    for (unsigned t = 0; t < steps; t++) {
        // odd steps phase by phase (plugins asked when the reference would ask them), even steps in one call
        expect((t & 1) ? economy->time_step_phased() : economy->time_step(), "time_step", (int)t);
        if (t & 1) labour_seen_by_consumption += persons->seen_labor;
        expect(economy->get_time() == t + 1, "get_time", (int)t);
        // oracle on the same decisions and visiting orders
        fastace_actions_t av = economy->last_actions().view();
        fastace_state_t rv = ref.view();
        fastace_step_out_t out{};
        out.p_reward = ref_reward.data(); out.f_profit = ref_profit.data();
        expect(fastace_oracle_step(&dims, &rv, &av, &out, FASTACE_IDX_ABSOLUTE, t, 0, (int)E) == 0, "oracle step", (int)t);
        HostState& got = economy->state();
        expect(same(got.m_count, ref.m_count) && same(got.j_count, ref.j_count), "book sizes", (int)t);
        // books are compared over their live prefix (entries past m_count / j_count are not part of the market)
        bool goods_book = true, job_book = true, prices = true;
        const size_t capM = (size_t)F * 2, capJ = F;
        for (unsigned e = 0; e < E; e++) {
            for (size_t n = 0; n < (size_t)ref.m_count[e]; n++) {
                const size_t k = e * capM + n;
                goods_book &= got.m_owner[k] == ref.m_owner[k] && got.m_good[k] == ref.m_good[k] &&
                              got.m_left[k] == ref.m_left[k] && got.m_taken[k] == ref.m_taken[k];
                prices &= std::fabs(got.m_price[k] - ref.m_price[k]) <= 1e-5 * std::fabs(ref.m_price[k]);
            }
            for (size_t n = 0; n < (size_t)ref.j_count[e]; n++) {
                const size_t k = e * capJ + n;
                job_book &= got.j_owner[k] == ref.j_owner[k] && got.j_left[k] == ref.j_left[k] && got.j_taken[k] == ref.j_taken[k];
                prices &= std::fabs(got.j_wage[k] - ref.j_wage[k]) <= 1e-5 * std::fabs(ref.j_wage[k]);
            }
        }
        expect(goods_book, "goods book (bit-exact)", (int)t);
        expect(job_book, "job book (bit-exact)", (int)t);
        expect(prices, "prices / wages", (int)t);
        expect(same(got.p_labor, ref.p_labor) && same(got.f_labor, ref.f_labor), "labour (bit-exact: multiples of 0.5)", (int)t);
        expect(close(got.p_money, ref.p_money) && close(got.f_money, ref.f_money), "money", (int)t);
        expect(close(got.p_inv, ref.p_inv) && close(got.f_inv, ref.f_inv), "inventories", (int)t);
        bool rewards = true;
        for (unsigned e = 0; e < E && rewards; e++) {
            for (unsigned p = 0; p < P; p++) rewards &= std::fabs(economy->person_reward(e, p) - ref_reward[(size_t)e * P + p]) <= 1e-5 * std::fmax(std::fabs(ref_reward[(size_t)e * P + p]), 1e-9);
            for (unsigned f = 0; f < F; f++) rewards &= std::fabs(economy->firm_profit(e, f) - ref_profit[(size_t)e * F + f]) <= 1e-5 * std::fmax(std::fabs(ref_profit[(size_t)e * F + f]), 1e-6);
        }
        expect(rewards, "rewards / profits", (int)t);
        // read API: the posted book as Economy::get_market() would list it
        auto market = economy->get_market(0);
        for (const auto& o : market) expect(o.amountLeft > 0 && o.offerer >= 0 && o.offerer < (int)F && o.good >= 0 && o.good < 2, "get_market entry", (int)t);
        for (size_t n = 0; n < got.m_taken.size(); n++) trades += got.m_taken[n];
        expect(economy->person_money(0, 3) == got.p_money[3] && economy->firm_laborHired(E - 1, F - 1) == got.f_labor[(size_t)(E - 1) * F + F - 1], "getters", (int)t);
    }
    double money1 = 0;
    for (double m : economy->state().p_money) money1 += m;
    for (double m : economy->state().f_money) money1 += m;
    expect(std::fabs(money1 - money0) <= 1e-9 * money0, "money is conserved", (int)steps);
    expect(economy->get_jobMarket(0).size() > 0 || economy->get_market(0).size() > 0, "markets populated", (int)steps);
    expect(labour_seen_by_consumption > 0.0, "phased mode: choose_goods_to_consume saw this step's laborSupplied", (int)steps);

    if (failures) { std::printf("%d failure(s)\n", failures); return 1; }
    std::printf("OK: %u economies x %u steps through fastace::BatchedEconomy matched the oracle\n", E, steps);
    return 0;
}
