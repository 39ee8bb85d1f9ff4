/*
 * TEST INFRASTRUCTURE ONLY — plain-C, single-threaded restatement of fastACE's
 * Economy::time_step hot path.  It is the CHECKER for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product never routes through it (there is no CPU fallback).
 *
 * Parity pin: the reference ships no tests and no golden vectors,
 * so this restatement is pinned against the reference ITSELF — the unmodified sources
 * compiled into oracle/_ref/libfastace_ref.so (oracle/Makefile, oracle/ref_harness.cpp).
 * tests/test_oracle_vs_ref.py steps both with identical injected actions and the
 * reference's own shuffles and requires every double to be BIT-identical; the vectors
 * generated from the reference that way are committed under tests/golden/ so that the
 * pin travels to machines where /root/reference does not exist.
 *
 * Each function cites the reference code it follows (paths relative to
 * /root/reference/src).  Arithmetic is written operation by operation in the
 * reference's order; compile with -ffp-contract=off so no FMA contraction changes it.
 */
#include "fastace_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-8                      /* constants::eps            base/constants.h:9  */
#define LARGE_NUMBER 1e8              /* constants::largeNumber    base/constants.h:10 */
#define AMOUNT_PER_OFFER 1.0          /* neural/neuralFirmDecisionMaker.cpp:6 */
#define LABOR_AMOUNT_PER_OFFER 0.5    /* neural/neuralFirmDecisionMaker.cpp:7 */
#define MAXG FASTACE_MAX_GOODS

/* (int)double exactly as the reference binary does it on x86-64 (cvttsd2si): values
 * that do not fit, and NaN, give INT_MIN ("integer indefinite"), which then fails the
 * `numOffers > 0` tests (neuralFirmDecisionMaker.cpp:135, 169). */
int32_t fastace_oracle_double_to_int(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return (int32_t)x;
}

/* CES::CES — shares normalised to sum 1, rho = 1/(1-elasticity)  functions/vecToScalar.cpp:105-110 */
void fastace_oracle_ces_params(const double* share_raw, double elasticity, int n, double* share_out, double* rho_out) {
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += share_raw[i];
    for (int i = 0; i < n; i++) share_out[i] = share_raw[i] / sum;
    *rho_out = 1 / (1 - elasticity);
}

/* CES::get_inner_sum + CES::f  functions/vecToScalar.cpp:112-118
 *   inner = sum_i share_i * pow(x_i + eps, rho);  f = tfp * pow(inner, 1/rho)
 * `share` is read with the given stride so SoA storage can be used directly. */
double fastace_oracle_ces_f(double tfp, const double* share, double rho, const double* x, int n, int stride) {
    double inner = 0.0;
    for (int i = 0; i < n; i++) inner += share[(size_t)i * stride] * pow(x[i] + EPS, rho);
    return tfp * pow(inner, 1 / rho);
}

/* CobbDouglas::f  functions/vecToScalar.cpp:45-47 :  tfp * prod_i pow(x_i, e_i) */
double fastace_oracle_cobb_douglas_f(double tfp, const double* elast, const double* x, int n) {
    double prod = 1.0;
    for (int i = 0; i < n; i++) prod *= pow(x[i], elast[i]);
    return tfp * prod;
}

static int g_util_kind = FASTACE_FN_CES, g_prod_kind = FASTACE_FN_CES;
void fastace_oracle_set_function_kinds(int util_kind, int prod_kind) { g_util_kind = util_kind; g_prod_kind = prod_kind; }

/* VecToScalar::f by family, each in the reference's operation order (Eigen reductions over a
 * handful of elements evaluate left to right):
 *   Linear::f (productivities*quantities).sum()                      functions/vecToScalar.cpp:30-32
 *   CobbDouglas::f tfp*pow(quantities, elasticities).prod()            :45-47
 *   StoneGeary::f  tfp*pow(quantities - thresholds, elasticities).prod() :67-69
 *   Leontief::f    (quantities*productivities).minCoeff()              :80-82
 *   CES::f                                                             :112-118 */
double fastace_oracle_function_f(int kind, double tfp, const double* share, const double* theta, double rho,
                                 const double* x, int n, int stride) {
    switch (kind) {
        case FASTACE_FN_LINEAR: {
            double s = 0.0;
            for (int i = 0; i < n; i++) s += share[(size_t)i * stride] * x[i];
            return s;
        }
        case FASTACE_FN_COBB_DOUGLAS: {
            double p = 1.0;
            for (int i = 0; i < n; i++) p *= pow(x[i], share[(size_t)i * stride]);
            return tfp * p;
        }
        case FASTACE_FN_STONE_GEARY: {
            double p = 1.0;
            for (int i = 0; i < n; i++) p *= pow(x[i] - (theta ? theta[(size_t)i * stride] : 0.0), share[(size_t)i * stride]);
            return tfp * p;
        }
        case FASTACE_FN_LEONTIEF: {
            double m = x[0] * share[0];
            for (int i = 1; i < n; i++) { double v = x[i] * share[(size_t)i * stride]; if (v < m) m = v; }
            return m;
        }
        default:
            return fastace_oracle_ces_f(tfp, share, rho, x, n, stride);
    }
}

typedef struct {
    int P, F, G, S, cap; /* cap = F*G */
    uint32_t flags;
    /* per-economy views */
    double *p_money, *p_inv, *p_labor, *p_tfp, *p_share, *p_rho;
    double *f_money, *f_inv, *f_labor, *f_last, *f_tfp, *f_share, *f_rho;
    int32_t *m_count, *m_owner, *m_good; uint32_t *m_left, *m_taken; double* m_price;
    int32_t *j_count, *j_owner; uint32_t *j_left, *j_taken; double* j_wage;
} econ_t;

static int map_index(int32_t raw, int count, uint32_t flags, int* out) {
    if (count <= 0) return 0;
    if (flags & FASTACE_IDX_MODULO) { *out = (int)((uint32_t)raw % (uint32_t)count); return 1; }  /* FASTACE_STEP_SERIAL has no meaning here */
    if (raw < 0 || raw >= count) return 0;
    *out = raw;
    return 1;
}

/* Buyer `money`/`inv` (stride `bstride` between goods) requests one unit of goods entry n.
 *   Agent::respond_to_offer        base/agent.cpp:99-116
 *   Agent::review_offer_response   base/agent.cpp:118-150
 *   Agent::accept_offer_response   base/agent.cpp:152-161
 * An entry that its owner already withdrew or that was flushed away (weak_ptr null,
 * agent.cpp:101) has amountLeft == 0 here; both cases refuse without side effects. */
static int request_good(econ_t* ec, int n, double* money, double* inv, int bstride) {
    const int G = ec->G, F = ec->F;
    double price = ec->m_price[n];
    if (!(*money >= price)) return 0;                        /* agent.cpp:102 */
    if (!(ec->m_left[n] > 0)) return 0;                      /* agent.cpp:124 is_available */
    int s = ec->m_owner[n], good = ec->m_good[n];
    int short_ = 0;                                          /* agent.cpp:140 (inventory < quantities).any() */
    for (int g = 0; g < G; g++) {
        double q = (g == good) ? AMOUNT_PER_OFFER : 0.0;
        if (ec->f_inv[(size_t)g * F + s] < q) short_ = 1;
    }
    if (short_) { ec->m_left[n] = 0; return 0; }             /* agent.cpp:143 */
    /* accept_offer_response, seller side first (agent.cpp:152-161) */
    ec->f_money[s] += price;
    for (int g = 0; g < G; g++) {
        double q = (g == good) ? AMOUNT_PER_OFFER : 0.0;
        ec->f_inv[(size_t)g * F + s] -= q;
    }
    ec->m_left[n]--;
    ec->m_taken[n]++;
    /* buyer side (agent.cpp:105-111) */
    *money -= price;
    for (int g = 0; g < G; g++) {
        double q = (g == good) ? AMOUNT_PER_OFFER : 0.0;
        inv[(size_t)g * bstride] += q;
    }
    return 1;
}

/* Person p requests one lot of job entry n.
 *   Person::respond_to_jobOffer      base/person.cpp:36-54
 *   Firm::review_jobOffer_response   base/firm.cpp:56-90
 *   Firm::accept_jobOffer_response   base/firm.cpp:106-113 */
static int request_job(econ_t* ec, int n, int p) {
    double labor = LABOR_AMOUNT_PER_OFFER, wage = ec->j_wage[n];
    if (!(ec->p_labor[p] + labor <= 1)) return 0;            /* person.cpp:39 */
    if (!(ec->j_left[n] > 0)) return 0;                      /* firm.cpp:64 */
    int f = ec->j_owner[n];
    if (ec->f_money[f] < wage) { ec->j_left[n] = 0; return 0; } /* firm.cpp:80-84 */
    ec->f_money[f] -= wage;                                  /* firm.cpp:108-111 */
    ec->f_labor[f] += labor;
    ec->j_left[n]--;
    ec->j_taken[n]++;
    ec->p_labor[p] += labor;                                 /* person.cpp:48-49 */
    ec->p_money[p] += wage;
    return 1;
}

/* Agent::check_my_offers + update_offer_amount_left + check_inventory_delta
 * (base/agent.cpp:54-97) for firm f: walk own entries in list order with a running
 * inventoryLeft, shrinking amountLeft until quantities*amountLeft fits. */
static void check_my_offers(econ_t* ec, int f) {
    const int G = ec->G, F = ec->F;
    double invLeft[MAXG];
    for (int g = 0; g < G; g++) invLeft[g] = ec->f_inv[(size_t)g * F + f];
    for (int n = 0; n < *ec->m_count; n++) {
        if (ec->m_owner[n] != f) continue;
        int good = ec->m_good[n];
        double q[MAXG], delta[MAXG];
        for (int g = 0; g < G; g++) {
            q[g] = (g == good) ? AMOUNT_PER_OFFER : 0.0;
            delta[g] = q[g] * (double)ec->m_left[n];         /* agent.cpp:73 */
        }
        for (;;) {
            int ok = 1;                                      /* agent.cpp:54-65 */
            for (int g = 0; g < G; g++) if (delta[g] > invLeft[g]) ok = 0;
            if (ok) break;
            /* The reference would underflow the unsigned counter here if an inventory
             * component were negative (unreachable single-threaded); stop at 0. */
            if (ec->m_left[n] == 0) break;
            for (int g = 0; g < G; g++) delta[g] -= q[g];    /* agent.cpp:79-80 */
            ec->m_left[n]--;
        }
        for (int g = 0; g < G; g++) invLeft[g] -= delta[g];  /* agent.cpp:83 */
    }
}

static void step_economy(const fastace_dims_t* d, fastace_state_t* st, const fastace_actions_t* a,
                         const fastace_step_out_t* out, uint32_t flags, uint32_t time_before, int e) {
    const int P = d->num_persons, F = d->num_firms, G = d->num_goods, S = d->stack_size;
    const int cap = F * G;
    econ_t ec;
    ec.P = P; ec.F = F; ec.G = G; ec.S = S; ec.cap = cap; ec.flags = flags;
    ec.p_money = st->p_money + (size_t)e * P;
    ec.p_inv = st->p_inv + (size_t)e * G * P;
    ec.p_labor = st->p_labor + (size_t)e * P;
    ec.p_tfp = st->p_util_tfp + (size_t)e * P;
    ec.p_share = st->p_util_share + (size_t)e * (G + 1) * P;
    ec.p_rho = st->p_util_rho + (size_t)e * P;
    ec.f_money = st->f_money + (size_t)e * F;
    ec.f_inv = st->f_inv + (size_t)e * G * F;
    ec.f_labor = st->f_labor + (size_t)e * F;
    ec.f_last = st->f_last_money + (size_t)e * F;
    ec.f_tfp = st->f_prod_tfp + (size_t)e * G * F;
    ec.f_share = st->f_prod_share + (size_t)e * G * (G + 1) * F;
    ec.f_rho = st->f_prod_rho + (size_t)e * G * F;
    ec.m_count = st->m_count + e;
    ec.m_owner = st->m_owner + (size_t)e * cap;
    ec.m_good = st->m_good + (size_t)e * cap;
    ec.m_left = st->m_left + (size_t)e * cap;
    ec.m_taken = st->m_taken + (size_t)e * cap;
    ec.m_price = st->m_price + (size_t)e * cap;
    ec.j_count = st->j_count + e;
    ec.j_owner = st->j_owner + (size_t)e * F;
    ec.j_left = st->j_left + (size_t)e * F;
    ec.j_taken = st->j_taken + (size_t)e * F;
    ec.j_wage = st->j_wage + (size_t)e * F;

    const int32_t* perm_p = a->perm_person + (size_t)e * P;
    const int32_t* perm_f = a->perm_firm + (size_t)e * F;
    const int32_t* p_job_idx = a->p_job_idx + (size_t)e * S * P;
    const uint8_t* p_job_take = a->p_job_take + (size_t)e * S * P;
    const int32_t* p_good_idx = a->p_good_idx + (size_t)e * S * P;
    const uint8_t* p_good_take = a->p_good_take + (size_t)e * S * P;
    const float* p_consume = a->p_consume + (size_t)e * G * P;
    const int32_t* f_good_idx = a->f_good_idx + (size_t)e * S * F;
    const uint8_t* f_good_take = a->f_good_take + (size_t)e * S * F;
    const float* f_prod = a->f_prod + (size_t)e * G * F;
    const float* f_offer_amt = a->f_offer_amt + (size_t)e * G * F;
    const float* f_offer_price = a->f_offer_price + (size_t)e * G * F;
    const float* f_job_labor = a->f_job_labor + (size_t)e * F;
    const float* f_job_wage = a->f_job_wage + (size_t)e * F;

    uint8_t* p_job_ok = (out && out->p_job_ok) ? out->p_job_ok + (size_t)e * S * P : NULL;
    uint8_t* p_good_ok = (out && out->p_good_ok) ? out->p_good_ok + (size_t)e * S * P : NULL;
    uint8_t* f_good_ok = (out && out->f_good_ok) ? out->f_good_ok + (size_t)e * S * F : NULL;
    if (p_job_ok) memset(p_job_ok, 0, (size_t)S * P);
    if (p_good_ok) memset(p_good_ok, 0, (size_t)S * P);
    if (f_good_ok) memset(f_good_ok, 0, (size_t)S * F);

    /* The snapshot the step trades against (decisionNetHandler.cpp:236-275, 303-308) is
     * the book as it stands now: every entry was posted during the previous step, and
     * entries posted during this step go to a separate new book (never visible in the
     * same step). */
    const int NM = *ec.m_count, NJ = *ec.j_count;

    /* ---- persons, in visiting order (economy.cpp:118-120; Person::time_step person.cpp:19-33) ---- */
    for (int r = 0; r < P; r++) {
        const int p = perm_p[r];
        /* Agent::time_step (agent.cpp:15-28): persons own no offers -> check_my_offers is a no-op */
        ec.p_labor[p] = 0.0;                                              /* person.cpp:24 */
        /* UtilMaxer::search_for_jobs (utilMaxer.cpp:76-85), orders of amount 1 in stack order */
        for (int i = 0; i < S; i++) {
            int n;
            if (p_job_take[(size_t)i * P + p] && map_index(p_job_idx[(size_t)i * P + p], NJ, flags, &n)) {
                int ok = request_job(&ec, n, p);
                if (p_job_ok) p_job_ok[(size_t)i * P + p] = (uint8_t)ok;
            }
        }
        /* UtilMaxer::buy_goods (utilMaxer.cpp:64-73) */
        for (int i = 0; i < S; i++) {
            int n;
            if (p_good_take[(size_t)i * P + p] && map_index(p_good_idx[(size_t)i * P + p], NM, flags, &n)) {
                int ok = request_good(&ec, n, &ec.p_money[p], &ec.p_inv[p], P);
                if (p_good_ok) p_good_ok[(size_t)i * P + p] = (uint8_t)ok;
            }
        }
        /* UtilMaxer::consume_goods (utilMaxer.cpp:88-92) with
         * NeuralPersonDecisionMaker::choose_goods_to_consume (neuralPersonDecisionMaker.cpp:93-111)
         * and UtilMaxer::u (utilMaxer.cpp:54-62): inputs = (1 - laborSupplied, to_consume) */
        double x[MAXG + 1], c[MAXG];
        x[0] = 1 - ec.p_labor[p];
        for (int g = 0; g < G; g++) {
            c[g] = ec.p_inv[(size_t)g * P + p] * (double)p_consume[(size_t)g * P + p];
            x[g + 1] = c[g];
        }
        double util = fastace_oracle_function_f(g_util_kind, ec.p_tfp[p], ec.p_share + p,
                                                st->p_util_theta ? st->p_util_theta + (size_t)e * (G + 1) * P + p : NULL,
                                                ec.p_rho[p], x, G + 1, P);
        out->p_reward[(size_t)e * P + p] = util;
        for (int g = 0; g < G; g++) ec.p_inv[(size_t)g * P + p] -= c[g];
    }

    /* job counters are final after the person phase (firms never take jobs) */
    if (out && out->old_j_left) memcpy(out->old_j_left + (size_t)e * F, ec.j_left, sizeof(uint32_t) * NJ);
    if (out && out->old_j_taken) memcpy(out->old_j_taken + (size_t)e * F, ec.j_taken, sizeof(uint32_t) * NJ);

    /* new books, filled in posting order (economy.cpp:52-59) */
    int32_t* nm_owner = (int32_t*)malloc(sizeof(int32_t) * (cap > 0 ? cap : 1));
    int32_t* nm_good = (int32_t*)malloc(sizeof(int32_t) * (cap > 0 ? cap : 1));
    uint32_t* nm_left = (uint32_t*)malloc(sizeof(uint32_t) * (cap > 0 ? cap : 1));
    double* nm_price = (double*)malloc(sizeof(double) * (cap > 0 ? cap : 1));
    int32_t* nj_owner = (int32_t*)malloc(sizeof(int32_t) * (F > 0 ? F : 1));
    uint32_t* nj_left = (uint32_t*)malloc(sizeof(uint32_t) * (F > 0 ? F : 1));
    double* nj_wage = (double*)malloc(sizeof(double) * (F > 0 ? F : 1));
    int nm = 0, nj = 0;

    /* ---- firms, in visiting order (economy.cpp:121-123; Firm::time_step firm.cpp:23-46) ---- */
    for (int r = 0; r < F; r++) {
        const int f = perm_f[r];
        /* Agent::time_step: check_my_offers, then flush own dead offers (agent.cpp:15-28).
         * A flushed entry simply stays at amountLeft == 0. */
        check_my_offers(&ec, f);
        /* Firm::check_myJobOffers (firm.cpp:93-103) can only RAISE amountLeft of the firm's
         * own job offers, after every person has already acted, and those offers are
         * withdrawn later in this same firm step: no observable effect. */

        /* first decision of the step: record last step's profit
         * (neuralFirmDecisionMaker.cpp:20-33, 65-74) */
        double profit = 0.0;
        if (time_before > 0) profit = ec.f_money[f] - ec.f_last[f];
        ec.f_last[f] = ec.f_money[f];
        out->f_profit[(size_t)e * F + f] = profit;

        /* ProfitMaxer::buy_goods (profitMaxer.cpp:102-111); self-purchase is possible */
        for (int i = 0; i < S; i++) {
            int n;
            if (f_good_take[(size_t)i * F + f] && map_index(f_good_idx[(size_t)i * F + f], NM, flags, &n)) {
                int ok = request_good(&ec, n, &ec.f_money[f], &ec.f_inv[f], F);
                if (f_good_ok) f_good_ok[(size_t)i * F + f] = (uint8_t)ok;
            }
        }
        /* ProfitMaxer::produce (profitMaxer.cpp:68-72) with choose_production_inputs
         * (neuralFirmDecisionMaker.cpp:95-108), ProfitMaxer::f (profitMaxer.cpp:45-49),
         * SumOfVecToVec::f / VToVFromVToS::f (vecToVec.cpp:17-23, vecToVec.h:39-43):
         * one CES per output good over inputs (laborHired, x). */
        double in[MAXG + 1], x[MAXG], prod[MAXG];
        in[0] = ec.f_labor[f];
        for (int g = 0; g < G; g++) {
            x[g] = ec.f_inv[(size_t)g * F + f] * (double)f_prod[(size_t)g * F + f];
            in[g + 1] = x[g];
        }
        for (int g = 0; g < G; g++)
            prod[g] = fastace_oracle_function_f(g_prod_kind, ec.f_tfp[(size_t)g * F + f], ec.f_share + (size_t)g * (G + 1) * F + f,
                                                st->f_prod_theta ? st->f_prod_theta + ((size_t)e * G + g) * (G + 1) * F + f : NULL,
                                                ec.f_rho[(size_t)g * F + f], in, G + 1, F);
        for (int g = 0; g < G; g++) ec.f_inv[(size_t)g * F + f] += (prod[g] - x[g]);

        /* ProfitMaxer::sell_goods (profitMaxer.cpp:74-86): choose_good_offers first
         * (neuralFirmDecisionMaker.cpp:111-146; amounts = proportion * inventory,
         * decisionNetHandler.cpp:586-592), then withdraw old offers, then post. */
        int32_t lots[MAXG];
        for (int g = 0; g < G; g++) {
            double amount = (double)f_offer_amt[(size_t)g * F + f] * ec.f_inv[(size_t)g * F + f];
            lots[g] = fastace_oracle_double_to_int(amount / AMOUNT_PER_OFFER);
        }
        for (int n = 0; n < NM; n++) {
            if (ec.m_owner[n] != f) continue;
            if (out && out->old_m_left) out->old_m_left[(size_t)e * cap + n] = ec.m_left[n];
            if (out && out->old_m_taken) out->old_m_taken[(size_t)e * cap + n] = ec.m_taken[n];
            ec.m_left[n] = 0;                                              /* profitMaxer.cpp:79-81 */
        }
        for (int g = 0; g < G; g++) {
            if (lots[g] > 0) {                                             /* neuralFirmDecisionMaker.cpp:135 */
                nm_owner[nm] = f; nm_good[nm] = g; nm_left[nm] = (uint32_t)lots[g];
                nm_price[nm] = (double)f_offer_price[(size_t)g * F + f] / AMOUNT_PER_OFFER;
                nm++;
            }
        }
        /* pay_dividends is a no-op; laborHired reset (firm.cpp:40-41) */
        ec.f_labor[f] = 0.0;
        /* ProfitMaxer::search_for_laborers (profitMaxer.cpp:88-100) with choose_job_offers
         * (neuralFirmDecisionMaker.cpp:149-180) and the wage clip (decisionNetHandler.cpp:631-635).
         * Withdrawal of the old job offer needs no action: the old job book is dropped below. */
        double laborAmount = (double)f_job_labor[f];
        double wage = (double)f_job_wage[f];
        if (wage > LARGE_NUMBER) wage = LARGE_NUMBER;
        int32_t jlots = fastace_oracle_double_to_int(laborAmount / LABOR_AMOUNT_PER_OFFER);
        if (jlots > 0) {
            nj_owner[nj] = f; nj_left[nj] = (uint32_t)jlots; nj_wage[nj] = wage / LABOR_AMOUNT_PER_OFFER;
            nj++;
        }
    }

    /* util::flush(market); util::flush(jobMarket) (economy.cpp:125-126, util.h:50-65): every
     * old entry has been withdrawn by its owner, every new entry has amountLeft > 0, and
     * std::remove_if is stable -> the markets are exactly the new books. */
    for (int n = 0; n < nm; n++) {
        ec.m_owner[n] = nm_owner[n]; ec.m_good[n] = nm_good[n]; ec.m_left[n] = nm_left[n];
        ec.m_taken[n] = 0; ec.m_price[n] = nm_price[n];
    }
    for (int n = nm; n < cap; n++) { ec.m_owner[n] = 0; ec.m_good[n] = 0; ec.m_left[n] = 0; ec.m_taken[n] = 0; ec.m_price[n] = 0.0; }
    *ec.m_count = nm;
    for (int n = 0; n < nj; n++) { ec.j_owner[n] = nj_owner[n]; ec.j_left[n] = nj_left[n]; ec.j_taken[n] = 0; ec.j_wage[n] = nj_wage[n]; }
    for (int n = nj; n < F; n++) { ec.j_owner[n] = 0; ec.j_left[n] = 0; ec.j_taken[n] = 0; ec.j_wage[n] = 0.0; }
    *ec.j_count = nj;
    free(nm_owner); free(nm_good); free(nm_left); free(nm_price); free(nj_owner); free(nj_left); free(nj_wage);
}

int fastace_oracle_step(const fastace_dims_t* d, fastace_state_t* st, const fastace_actions_t* a,
                        const fastace_step_out_t* out, uint32_t flags, uint32_t time_before,
                        int first_econ, int count) {
    if (!d || !st || !a || !out || !out->p_reward || !out->f_profit) return -1;
    if (d->num_goods < 1 || d->num_goods > MAXG || d->stack_size < 0 || d->num_persons < 0 || d->num_firms < 0) return -1;
    for (int e = first_econ; e < first_econ + count && e < d->num_econ; e++)
        step_economy(d, st, a, out, flags, time_before, e);
    return 0;
}

typedef struct {
    const fastace_dims_t* d; fastace_state_t* st; const fastace_actions_t* a; const fastace_step_out_t* out;
    uint32_t flags, time_before; int first, count;
} mt_job_t;

static void* mt_main(void* arg) {
    mt_job_t* j = (mt_job_t*)arg;
    fastace_oracle_step(j->d, j->st, j->a, j->out, j->flags, j->time_before, j->first, j->count);
    return NULL;
}

int fastace_oracle_step_mt(const fastace_dims_t* d, fastace_state_t* st, const fastace_actions_t* a,
                           const fastace_step_out_t* out, uint32_t flags, uint32_t time_before, int nthreads) {
    if (nthreads <= 1) return fastace_oracle_step(d, st, a, out, flags, time_before, 0, d->num_econ);
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    mt_job_t jobs[256];
    int E = d->num_econ, per = (E + nthreads - 1) / nthreads, n = 0;
    for (int t = 0; t < nthreads; t++) {
        int first = t * per;
        if (first >= E) break;
        jobs[n] = (mt_job_t){d, st, a, out, flags, time_before, first, per};
        pthread_create(&th[n], NULL, mt_main, &jobs[n]);
        n++;
    }
    for (int t = 0; t < n; t++) pthread_join(th[t], NULL);
    return 0;
}
