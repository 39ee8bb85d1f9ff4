// Economy::time_step for ONE LARGE economy (BASELINE config D: 100 000 persons, 5 000 firms, 8 goods) — books far
// beyond shared memory, and only one sequential visiting order to parallelise.
//
// Formulation.  The reference walks persons then firms in visiting order (economy.cpp:117-124); every request is
// decided first-come-first-served against finite lots (agent.cpp:99-161, firm.cpp:56-113).  The outcome of a request
// depends only on (i) the requester's own earlier requests of the step (its money / labour), and (ii) the earlier
// events at the firm on the other side (lots left, inventory, and — for jobs — the firm's money, which earlier hires
// lower and earlier sales raise).  So the step is the unique fixed point of two maps that are each embarrassingly
// parallel:
//   requester pass  one thread per person: walk the own S job + S goods requests in order, given which of them the
//                   other side granted  ->  `want` (own-side conditions hold: labour <= 1, money >= price)
//   firm pass       one warp per firm: ALL events at this firm in the reference's order (visiting rank, jobs before
//                   goods, slot), given `want`  ->  `ok`; this is the segmented, stably sorted event list: events
//                   keyed by firm, sorted stably by firm with the hand-written counting sort of segmented_sort.cuh
//                   whose input is enumerated in (rank, phase, slot) order.  Goods events are per-good prefix counts
//                   of `want` bits; only the firm's money (sales, hires) is walked sequentially.
// iterated until no `ok` flag changes, inside ONE cooperative launch (grid.sync between the passes).  By induction
// over the global event order a fixed point IS the sequential result, and round k finalises every event whose
// dependency chain crosses agents at most k times; measured 8 rounds at config D.  Each agent's fp64 money /
// inventory is updated in the reference's order, so unlike the warp-per-economy kernel there is no rounding-order
// caveat: firm money is bit-identical.
// The firm phase (firms buying from firms, firm.cpp:23-46) is the same iteration over F*S requests, with the rule
// that an offer is withdrawn once its owner's turn has passed (profitMaxer.cpp:79-81).
//
// Preconditions (true for every book the step itself produces): at most one live goods offer per (firm, good) and one
// job offer per firm; inventories are non-negative.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace fastace {

constexpr uint32_t kNoEntry = 0xFFFFFFFFu;
constexpr uint32_t kReqMask = (1u << 27) - 1u;   // request id in the sort value; bits 27.. = 0 job, 1+g goods of good g
constexpr int kLargeThreads = 256;
// first guess of the iteration: false = every request is granted; true = every taken request is made and the firms
// answer that (capacity-limited availability), which the requesters then refine
constexpr bool kPessimisticStart = false;   // measured at config D: 9 rounds + the extra pass vs 8 rounds

struct LargeScratch {
    int32_t* rank_f;       // [F]   visiting rank of firm f this step
    int32_t* own_offer;    // [F*G] goods-book entry of (f,g) or -1
    int32_t* own_job;      // [F]   job-book entry of f or -1
    // person phase, request id = (phase*S + i)*P + p
    uint32_t* req_n;       // [2SP] book entry the request refers to, or kNoEntry
    uint8_t* want;         // [2SP]
    uint8_t* ok;           // [2SP]
    uint16_t *key_in, *key_out;   // [2SP] firm on the other side (0xFFFF = no request)
    uint32_t *val_in, *val_out;   // [2SP]
    uint32_t* hist;        // [F]
    uint32_t* seg;         // [F+1]
    double* pm_money;      // [P] person results of the last requester pass
    uint8_t* pm_hires;     // [P]
    uint8_t* pm_bought;    // [G][P]
    // state of every firm after the person phase / after its own purchases
    double* fm_money;      // [F]
    double* fm_labor;      // [F]
    double* fm_inv;        // [G][F]
    uint32_t *fm_left, *fm_taken;    // [F*G] by (f,g)
    uint32_t *fm_jleft, *fm_jtaken;  // [F]
    // firm phase, request id = i*F + f
    uint32_t* freq_n;      // [SF]
    uint8_t *fwant, *fok;  // [SF]
    uint16_t *fkey_in, *fkey_out;
    uint32_t *fval_in, *fval_out;
    uint32_t* fhist;       // [F]
    uint32_t* fseg;        // [F+1]
    double* ff_profit;     // [F]   firm results of the last firm-phase pass (after its own purchases)
    double* ff_money;      // [F]
    double* ff_last;       // [F]   money at the first decision of the step
    double* ff_inv;        // [G][F]
    uint32_t *ff_left, *ff_taken;    // [F*G]
    int32_t* post_lots;    // [F*G]
    int32_t* post_jlots;   // [F]
    uint16_t* req_firm;    // [2SP] firm on the other side of the request
    uint32_t* req_pos;     // [2SP] position of the request in the sorted event list
    uint8_t *want_sorted, *ok_sorted;   // [2SP] the same flags in event-list order: the firm pass streams them
    uint8_t* dirty_person; // [P]   some `ok` of this person changed since its last requester pass
    uint8_t* dirty_firm;   // [F]   some `want` at this firm changed since its last firm pass
    uint32_t *post_base_m, *post_base_j;   // [F] by visiting rank: first new-book slot of that firm
    int* changed;
};

struct LargeParams {
    StepParams sp;     // state / action / out pointers already offset to the economy being stepped (E = 1 view)
    LargeScratch sc;
    int G;
    volatile uint32_t* dev_err;   // device error words of the env (match_kernel.cuh)
};

__device__ __forceinline__ uint32_t large_map_index(int32_t raw, int count, uint32_t flags) {
    if (count <= 0) return kNoEntry;
    if (flags & FASTACE_IDX_MODULO) return (uint32_t)raw % (uint32_t)count;
    if (raw < 0 || raw >= count) return kNoEntry;
    return (uint32_t)raw;
}

// ---- step prologue: who owns which entry, visiting ranks, cleared histograms ---------------------------------
__global__ void large_index_books(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, G = lp.G, cap = F * G;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cap) lp.sc.own_offer[t] = -1;
    if (t < F) { lp.sc.own_job[t] = -1; lp.sc.hist[t] = 0; lp.sc.dirty_firm[t] = 1; }
    if (t < p.P) lp.sc.dirty_person[t] = 1;
}
__global__ void large_index_books2(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, G = lp.G;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.st.m_count[0]) lp.sc.own_offer[p.st.m_owner[t] * G + p.st.m_good[t]] = t;
    if (t < p.st.j_count[0]) lp.sc.own_job[p.st.j_owner[t]] = t;
    (void)F;
}

// firm-phase prologue: visiting ranks of the firms, cleared histogram
__global__ void large_index_firms(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.F) { lp.sc.fhist[t] = 0; lp.sc.rank_f[p.ac.perm_firm[t]] = t; }
}
// after the person phase of a phase-wise step: the firms as they stand now, visible in the state arrays
__global__ void large_publish_firms(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, G = lp.G;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < F) { p.st.f_money[t] = lp.sc.fm_money[t]; p.st.f_labor[t] = lp.sc.fm_labor[t]; }
    if (t < F * G) p.st.f_inv[t] = lp.sc.fm_inv[t];
}

// ---- person phase: requests enumerated in the reference's order (rank, jobs before goods, slot) --------------
__global__ void large_prep_persons(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int P = p.P, S = p.S;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)2 * S * P) return;
    const int r = (int)(t / (2 * S)), rem = (int)(t % (2 * S)), phase = rem / S, i = rem % S;
    const int pid = p.ac.perm_person[r];
    const size_t a = (size_t)i * P + pid;
    const uint32_t req = (uint32_t)((size_t)(phase * S + i) * P + pid);
    uint32_t n = kNoEntry;
    if (phase == 0) { if (p.ac.p_job_take[a]) n = large_map_index(p.ac.p_job_idx[a], p.st.j_count[0], p.flags); }
    else            { if (p.ac.p_good_take[a]) n = large_map_index(p.ac.p_good_idx[a], p.st.m_count[0], p.flags); }
    lp.sc.req_n[req] = n;
    lp.sc.want[req] = (n != kNoEntry) && kPessimisticStart;
    lp.sc.ok[req] = (n != kNoEntry);     // optimistic start (replaced by a firm pass when kPessimisticStart)
    if (n == kNoEntry) { lp.sc.key_in[t] = 0xFFFFu; lp.sc.val_in[t] = req; return; }
    const int firm = phase == 0 ? p.st.j_owner[n] : p.st.m_owner[n];
    const uint32_t type = phase == 0 ? 0u : 1u + (uint32_t)p.st.m_good[n];
    lp.sc.key_in[t] = (uint16_t)firm;
    lp.sc.val_in[t] = req | (type << 27);
    lp.sc.req_firm[req] = (uint16_t)firm;
    atomicAdd(&lp.sc.hist[firm], 1u);
}

// exclusive scan of hist[0..n) into seg[0..n], one block
__global__ void large_scan(const uint32_t* hist, uint32_t* seg, int n) {
    __shared__ uint32_t part[1024];
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    uint32_t s = 0;
    for (int k = lo; k < hi; k++) s += hist[k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < (int)blockDim.x; d <<= 1) {
        const uint32_t v = threadIdx.x >= (unsigned)d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    for (int k = lo; k < hi; k++) { seg[k] = run; run += hist[k]; }
    if (threadIdx.x == blockDim.x - 1) seg[n] = part[blockDim.x - 1];
}

// event-list positions of the requests and the flags in event-list order (after the sort): the firm pass then reads
// its segment as contiguous bytes instead of gathering one byte per event through the request id
__global__ void large_index_events(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const size_t n = (size_t)2 * p.S * p.P;
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const bool valid = lp.sc.key_out[pos] != 0xFFFFu;
    lp.sc.want_sorted[pos] = valid && kPessimisticStart;
    lp.sc.ok_sorted[pos] = valid;
    if (valid) lp.sc.req_pos[lp.sc.val_out[pos] & kReqMask] = (uint32_t)pos;
}

// requester pass (Person::time_step person.cpp:19-33; respond_to_jobOffer :36-54; respond_to_offer agent.cpp:99-116).
// Persons none of whose requests changed outcome since their last pass are skipped (their results stand).  All
// gathers of a chain are issued before the chain is walked, so the walk itself runs on registers.
template <int G>
__device__ __forceinline__ void large_person_body(const LargeParams& lp, int pid) {
    const StepParams& p = lp.sp;
    const int P = p.P, S = p.S;
    if (!lp.sc.dirty_person[pid]) return;
    lp.sc.dirty_person[pid] = 0;
    double money = p.st.p_money[pid];
    double labor = 0.0;                                                    // person.cpp:24
    int hires = 0;
    uint8_t bought[G];
#pragma unroll
    for (int g = 0; g < G; g++) bought[g] = 0;
#pragma unroll
    for (int phase = 0; phase < 2; phase++) {
        uint32_t n[kMaxStack];
        double val[kMaxStack];        // wage (jobs) / price (goods) of the entry the slot refers to
        int good[kMaxStack];
        uint32_t okbits = 0, wantbits = 0;
#pragma unroll
        for (int i = 0; i < kMaxStack; i++) {
            n[i] = kNoEntry;
            if (i < S) {
                const size_t req = (size_t)(phase * S + i) * P + pid;
                n[i] = lp.sc.req_n[req];
                okbits |= (uint32_t)(lp.sc.ok[req] != 0) << i;
                wantbits |= (uint32_t)(lp.sc.want[req] != 0) << i;
            }
        }
#pragma unroll
        for (int i = 0; i < kMaxStack; i++) {
            val[i] = 0.0; good[i] = 0;
            if (n[i] != kNoEntry) {
                if (phase == 0) val[i] = p.st.j_wage[n[i]];
                else { val[i] = p.st.m_price[n[i]]; good[i] = p.st.m_good[n[i]]; }
            }
        }
        uint32_t neww = 0;
#pragma unroll
        for (int i = 0; i < kMaxStack; i++) {
            if (n[i] == kNoEntry) continue;
            if (phase == 0) {
                const bool w = labor + kLaborPerOffer <= 1;                // person.cpp:39
                neww |= (uint32_t)w << i;
                if (w && ((okbits >> i) & 1u)) { labor += kLaborPerOffer; money += val[i]; hires++; }   // person.cpp:48-49
            } else {
                const bool w = money >= val[i];                            // agent.cpp:102
                neww |= (uint32_t)w << i;
                if (w && ((okbits >> i) & 1u)) {
                    money -= val[i];                                       // agent.cpp:105-111
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == good[i]) bought[g]++;
                }
            }
        }
        uint32_t diff = neww ^ wantbits;
        while (diff) {
            const int i = __ffs(diff) - 1;
            diff &= diff - 1;
            const size_t req = (size_t)(phase * S + i) * P + pid;
            lp.sc.want[req] = (neww >> i) & 1u;
            lp.sc.want_sorted[lp.sc.req_pos[req]] = (neww >> i) & 1u;
            lp.sc.dirty_firm[lp.sc.req_firm[req]] = 1;
        }
    }
    lp.sc.pm_money[pid] = money;
    lp.sc.pm_hires[pid] = (uint8_t)hires;
#pragma unroll
    for (int g = 0; g < G; g++) lp.sc.pm_bought[(size_t)g * P + pid] = bought[g];
}

// (inventory < quantities).any() for one unit of `good` (agent.cpp:140)
template <int G>
__device__ __forceinline__ bool large_short(const double (&inv)[G], int good) {
    bool s = false;
#pragma unroll
    for (int g = 0; g < G; g++) s |= inv[g] < (g == good ? kAmountPerOffer : 0.0);
    return s;
}

// firm pass of the person phase: every event at firm f in the reference's order
// (review_jobOffer_response firm.cpp:56-90, accept :106-113; review_offer_response agent.cpp:118-150, accept :152-161).
// One warp per firm, 32 events (sorted value + the requester's `want`) per batch:
//   * goods events never look at the firm's money: a made request succeeds iff fewer than
//     cap_g = min(amountLeft, units the inventory covers) made requests for the same good came before it — a prefix
//     count of `want` bits per good (ballot + popc), no serial walk; the counters after the pass follow in closed form;
//   * the firm's money is the one sequential quantity (sales raise it, hires lower it, a hire needs money >= wage):
//     the warp walks, uniformly in all lanes, only the successful sales and the job events of the batch, in order,
//     adding in fp64 exactly as the reference does.
// Requests their requester does not currently make are answered hypothetically (no side effects).  Firms none of
// whose events changed `want` since their last pass are skipped.
template <int G>
__device__ __forceinline__ bool large_firm_body(const LargeParams& lp, int f, int lane) {
    const StepParams& p = lp.sp;
    const int F = p.F, P = p.P;
    if (!lp.sc.dirty_firm[f]) return false;
    __syncwarp();
    if (lane == 0) lp.sc.dirty_firm[f] = 0;
    // per-good state, uniform in all lanes (static indices: the loops over g are unrolled)
    double inv[G];
    uint32_t left[G], taken[G], cap[G], made[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        inv[g] = p.st.f_inv[(size_t)g * F + f];
        const int n = lp.sc.own_offer[f * G + g];
        left[g] = n >= 0 ? p.st.m_left[n] : 0u;
        taken[g] = n >= 0 ? p.st.m_taken[n] : 0u;
        made[g] = 0;
    }
    // lane g keeps the price of good g for the money walk
    double myprice = 0.0;
    if (lane < G) { const int n = lp.sc.own_offer[f * G + lane]; myprice = n >= 0 ? p.st.m_price[n] : 0.0; }
#pragma unroll
    for (int g = 0; g < G; g++) {
        // a good whose inventory is negative makes every sale of ANOTHER good short (agent.cpp:140 compares all
        // goods); inventories only fall by sold units here, so that set is fixed for the pass
        bool other_neg = false;
#pragma unroll
        for (int h = 0; h < G; h++) if (h != g) other_neg |= inv[h] < 0.0;
        cap[g] = other_neg ? 0u : min(left[g], unit_sales_possible(inv[g]));
    }
    double money = p.st.f_money[f], labor = p.st.f_labor[f];
    const int nj = lp.sc.own_job[f];
    uint32_t jleft = 0, jtaken = 0;
    double wage = 0.0;
    if (nj >= 0) { jleft = p.st.j_left[nj]; jtaken = p.st.j_taken[nj]; wage = p.st.j_wage[nj]; }
    const uint32_t lt = (1u << lane) - 1u;
    bool changed = false;
    const uint32_t lo = lp.sc.seg[f], hi = lp.sc.seg[f + 1];
    // the next batch's three loads are in flight while the current batch is walked
    uint32_t nv = 0u; uint8_t nw = 0, nold = 0;
    if (lo + lane < hi) { nv = lp.sc.val_out[lo + lane]; nw = lp.sc.want_sorted[lo + lane]; nold = lp.sc.ok_sorted[lo + lane]; }
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t pos = base + lane;
        const bool live = pos < hi;
        const uint32_t v = nv;
        const uint8_t myw = nw, old = nold;
        if (pos + 32 < hi) { nv = lp.sc.val_out[pos + 32]; nw = lp.sc.want_sorted[pos + 32]; nold = lp.sc.ok_sorted[pos + 32]; }
        const uint32_t req = v & kReqMask;
        const uint32_t type = live ? (v >> 27) : 0xFFu;
        const uint32_t wmask = __ballot_sync(0xffffffffu, live && myw);
        uint32_t okmask = 0, salemask = 0;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const uint32_t m = __ballot_sync(0xffffffffu, type == (uint32_t)(g + 1));
            if (m == 0) continue;
            const uint32_t wm = m & wmask;
            const uint32_t before = made[g] + __popc(wm & lt);             // made requests for this good ahead of mine
            const bool avail = before < cap[g];                            // agent.cpp:124, 140
            const uint32_t am = __ballot_sync(0xffffffffu, avail) & m;
            okmask |= am;
            salemask |= am & wm;
            made[g] += __popc(wm);
        }
        // the money walk: successful sales and job events of the batch, in order
        const uint32_t jobmask = __ballot_sync(0xffffffffu, type == 0u);
        uint32_t walk = salemask | jobmask;
        while (walk) {
            const int k = __ffs(walk) - 1;
            walk &= walk - 1;
            const uint32_t tk = __shfl_sync(0xffffffffu, type, k);
            if (tk != 0u) {
                money += __shfl_sync(0xffffffffu, myprice, (int)tk - 1);   // agent.cpp:153
            } else if (jleft > 0) {                                        // firm.cpp:64
                const bool is_made = (wmask >> k) & 1u;
                if (money < wage) { if (is_made) jleft = 0; }              // firm.cpp:80-84
                else {
                    okmask |= 1u << k;
                    if (is_made) { money -= wage; labor += kLaborPerOffer; jleft--; jtaken++; }   // firm.cpp:108-111
                }
            }
        }
        if (live) {
            const uint8_t o = (okmask >> lane) & 1u;
            if (o != old) { lp.sc.ok_sorted[pos] = o; lp.sc.ok[req] = o; lp.sc.dirty_person[req % (uint32_t)P] = 1; changed = true; }
        }
    }
    if (lane == 0) {
        lp.sc.fm_money[f] = money;
        lp.sc.fm_labor[f] = labor;
        lp.sc.fm_jleft[f] = jleft;
        lp.sc.fm_jtaken[f] = jtaken;
#pragma unroll
        for (int g = 0; g < G; g++) {
            // closed form of the walk: the first cap[g] made requests are sales; one more made request finds the offer
            // exhausted (amountLeft 0) or the inventory short (amountLeft <- 0, agent.cpp:143)
            const uint32_t sold = min(made[g], cap[g]);
            lp.sc.fm_inv[(size_t)g * F + f] = inv[g] - (double)sold;        // sold exact subtractions of 1.0
            lp.sc.fm_left[f * G + g] = made[g] > cap[g] ? 0u : left[g] - sold;
            lp.sc.fm_taken[f * G + g] = taken[g] + sold;
        }
    }
    return changed;
}

// The whole person-phase iteration as ONE cooperative launch: requester pass, grid sync, firm pass, grid sync, until a
// round changes no flag.  No host round trip, no launch gaps between rounds.  flags: sc.changed[0..2] rotate (a flag
// is cleared two rounds after it was last read), sc.changed[3] receives the number of rounds.
template <int G>
__global__ void __launch_bounds__(kLargeThreads, 3) large_iterate_persons(const LargeParams lp, int max_rounds) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int warp = tid >> 5, nwarps = nthreads >> 5, lane = threadIdx.x & 31;
    const int P = lp.sp.P, F = lp.sp.F;
    volatile int* flags = lp.sc.changed;
    int round = 0;
    if (kPessimisticStart) {
        for (int f = warp; f < F; f += nwarps) { lp.sc.dirty_firm[f] = 1; large_firm_body<G>(lp, f, lane); }
        grid.sync();
    }
    while (round < max_rounds) {
        if (tid == 0) flags[(round + 1) % 3] = 0;
        for (int pid = tid; pid < P; pid += nthreads) large_person_body<G>(lp, pid);
        grid.sync();
        bool ch = false;
        for (int f = warp; f < F; f += nwarps) ch |= large_firm_body<G>(lp, f, lane);
        if (ch) flags[round % 3] = 1;
        grid.sync();
        const int again = flags[round % 3];
        round++;
        if (!again) break;
        if (round >= max_rounds && tid == 0) lp.dev_err[kDevErrLargeRounds] = 1u;   // not converged: an error, not a result
    }
    if (tid == 0) flags[3] = round;
}

// consume_goods + utility (utilMaxer.cpp:88-92, 54-62; neuralPersonDecisionMaker.cpp:93-111), persons' new state
template <int G>
__global__ void large_finalize_persons(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int P = p.P;
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= P) return;
    const double labor = kLaborPerOffer * (double)lp.sc.pm_hires[pid];
    if (p.flags & FASTACE_STEP_PERSONS_TRADE) {      // trades into the state; consumption follows in its own call
#pragma unroll
        for (int g = 0; g < G; g++) {
            const size_t k = (size_t)g * P + pid;
            double v = p.st.p_inv[k];
            const int nb = lp.sc.pm_bought[k];
            for (int q = 0; q < nb; q++) v += kAmountPerOffer;
            p.st.p_inv[k] = v;
        }
        p.st.p_money[pid] = lp.sc.pm_money[pid];
        p.st.p_labor[pid] = labor;
        return;
    }
    const bool applied = (p.flags & FASTACE_STEP_PERSONS_CONSUME) != 0;   // purchases are already in p_inv
    double x[G + 1], inv[G];
    x[0] = 1 - labor;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const size_t k = (size_t)g * P + pid;
        double v = p.st.p_inv[k];
        const int nb = applied ? 0 : lp.sc.pm_bought[k];
        for (int q = 0; q < nb; q++) v += kAmountPerOffer;
        const double c = v * (double)p.ac.p_consume[k];
        x[g + 1] = c;
        inv[g] = v - c;
    }
    double share[G + 1], theta[G + 1];
#pragma unroll
    for (int i = 0; i <= G; i++) {
        share[i] = p.st.p_util_share[(size_t)i * P + pid];
        theta[i] = (p.util_kind == FASTACE_FN_STONE_GEARY) ? p.st.p_util_theta[(size_t)i * P + pid] : 0.0;
    }
    p.out.p_reward[pid] = eval_function<G + 1, true>(p.util_kind, p.st.p_util_tfp[pid], share, theta, p.st.p_util_rho[pid], x);
    p.st.p_money[pid] = lp.sc.pm_money[pid];
    p.st.p_labor[pid] = labor;
#pragma unroll
    for (int g = 0; g < G; g++) p.st.p_inv[(size_t)g * P + pid] = inv[g];
}

// success flags of the person phase: a request was transacted iff it was made and granted
__global__ void large_person_flags(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const size_t n = (size_t)p.S * p.P;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (p.out.p_job_ok) p.out.p_job_ok[t] = lp.sc.want[t] & lp.sc.ok[t];
    if (p.out.p_good_ok) p.out.p_good_ok[t] = lp.sc.want[n + t] & lp.sc.ok[n + t];
}

// job counters are final after the person phase
__global__ void large_old_jobs(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.st.j_count[0]) return;
    const int f = p.st.j_owner[t];
    if (p.out.old_j_left) p.out.old_j_left[t] = lp.sc.fm_jleft[f];
    if (p.out.old_j_taken) p.out.old_j_taken[t] = lp.sc.fm_jtaken[f];
}

// ---- firm phase ----------------------------------------------------------------------------------------------
__global__ void large_prep_firms(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, S = p.S;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * F) return;
    const int q = t / S, i = t % S;
    const int f = p.ac.perm_firm[q];
    const uint32_t freq = (uint32_t)(i * F + f);
    uint32_t n = kNoEntry;
    if (p.ac.f_good_take[freq]) n = large_map_index(p.ac.f_good_idx[freq], p.st.m_count[0], p.flags);
    lp.sc.freq_n[freq] = n;
    lp.sc.fwant[freq] = 0;
    lp.sc.fok[freq] = (n != kNoEntry);
    if (n == kNoEntry) { lp.sc.fkey_in[t] = 0xFFFFu; lp.sc.fval_in[t] = freq; return; }
    const int seller = p.st.m_owner[n];
    lp.sc.fkey_in[t] = (uint16_t)seller;
    lp.sc.fval_in[t] = freq | ((1u + (uint32_t)p.st.m_good[n]) << 27);
    atomicAdd(&lp.sc.fhist[seller], 1u);
}

// One firm's turn given what the other firms currently claim (Firm::time_step firm.cpp:23-46): sales to the firms that
// come earlier in the visiting order, check_my_offers (agent.cpp:54-97), profit record
// (neuralFirmDecisionMaker.cpp:65-74), own purchases (profitMaxer.cpp:102-111; self-purchase is possible).
template <int G>
__device__ __forceinline__ bool large_firm_phase_body(const LargeParams& lp, int f) {
    const StepParams& p = lp.sp;
    const int F = p.F, S = p.S;
    const int qf = lp.sc.rank_f[f];
    double money = lp.sc.fm_money[f];
    double inv[G], price[G];
    uint32_t left[G], taken[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        inv[g] = lp.sc.fm_inv[(size_t)g * F + f];
        left[g] = lp.sc.fm_left[f * G + g];
        taken[g] = lp.sc.fm_taken[f * G + g];
        const int n = lp.sc.own_offer[f * G + g];
        price[g] = n >= 0 ? p.st.m_price[n] : 0.0;
    }
    bool changed = false;
    const uint32_t lo = lp.sc.fseg[f], hi = lp.sc.fseg[f + 1];
    for (uint32_t pos = lo; pos < hi; pos++) {
        const uint32_t v = lp.sc.fval_out[pos];
        const uint32_t freq = v & kReqMask;
        const int good = (int)(v >> 27) - 1;
        const int qb = lp.sc.rank_f[freq % (uint32_t)F];
        if (qb == qf) continue;                       // own request: decided in the own turn below
        uint8_t o = 0;
        if (qb < qf && lp.sc.fwant[freq]) {           // later buyers find the offer withdrawn (profitMaxer.cpp:79-81)
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (g != good) continue;
                if (left[g] > 0) {
                    if (large_short<G>(inv, g)) left[g] = 0;
                    else { money += price[g]; inv[g] -= kAmountPerOffer; left[g]--; taken[g]++; o = 1; }
                }
            }
        }
        if (lp.sc.fok[freq] != o) { lp.sc.fok[freq] = o; changed = true; }
    }
    // check_my_offers: shrink amountLeft until quantities*amountLeft fits the inventory (agent.cpp:54-97)
#pragma unroll
    for (int g = 0; g < G; g++) {
        if ((double)left[g] > inv[g]) left[g] = inv[g] >= 0.0 ? (uint32_t)inv[g] : 0u;
    }
    // first decision of the step: last step's profit (neuralFirmDecisionMaker.cpp:20-33, 65-74)
    const double profit = p.time_before > 0 ? money - p.st.f_last_money[f] : 0.0;
    const double last = money;
    for (int i = 0; i < S; i++) {
        const uint32_t freq = (uint32_t)(i * F + f);
        const uint32_t n = lp.sc.freq_n[freq];
        if (n == kNoEntry) continue;
        const int seller = p.st.m_owner[n], good = p.st.m_good[n];
        const double pr = p.st.m_price[n];
        const bool w = money >= pr;                                        // agent.cpp:102
        if (seller == f) {
            uint8_t o = 0;
            if (w) {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (g != good) continue;
                    if (left[g] > 0) {
                        if (large_short<G>(inv, g)) left[g] = 0;
                        else {
                            money += pr; inv[g] -= kAmountPerOffer; left[g]--; taken[g]++;   // seller side (agent.cpp:152-161)
                            money -= pr; inv[g] += kAmountPerOffer;                          // buyer side (agent.cpp:105-111)
                            o = 1;
                        }
                    }
                }
            }
            if (lp.sc.fok[freq] != o) { lp.sc.fok[freq] = o; changed = true; }
        } else {
            if (lp.sc.fwant[freq] != (uint8_t)w) { lp.sc.fwant[freq] = (uint8_t)w; changed = true; }
            if (w && lp.sc.fok[freq]) {
                money -= pr;
#pragma unroll
                for (int g = 0; g < G; g++) if (g == good) inv[g] += kAmountPerOffer;
            }
        }
    }
    // results of the turn so far; production and posting follow once the iteration has settled
    lp.sc.ff_profit[f] = profit;
    lp.sc.ff_money[f] = money;
    lp.sc.ff_last[f] = last;
#pragma unroll
    for (int g = 0; g < G; g++) {
        lp.sc.ff_inv[(size_t)g * F + f] = inv[g];
        lp.sc.ff_left[f * G + g] = left[g];
        lp.sc.ff_taken[f * G + g] = taken[g];
    }
    return changed;
}

// the firm-phase iteration as one cooperative launch (flags as in large_iterate_persons; rounds -> sc.changed[7])
template <int G>
__global__ void __launch_bounds__(kLargeThreads) large_iterate_firms(const LargeParams lp, int max_rounds) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int F = lp.sp.F;
    volatile int* flags = lp.sc.changed + 4;
    int round = 0;
    while (round < max_rounds) {
        if (tid == 0) flags[(round + 1) % 3] = 0;
        bool ch = false;
        for (int f = tid; f < F; f += nthreads) ch |= large_firm_phase_body<G>(lp, f);
        if (ch) flags[round % 3] = 1;
        grid.sync();
        const int again = flags[round % 3];
        round++;
        if (!again) break;
        if (round >= max_rounds && tid == 0) lp.dev_err[kDevErrLargeRounds] = 1u;
    }
    if (tid == 0) flags[3] = round;
}

// produce (profitMaxer.cpp:68-72), sell_goods / search_for_laborers decode (neuralFirmDecisionMaker.cpp:111-180);
// one thread per (firm, output good)
template <int G>
__global__ void large_finalize_firms(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F * G) return;
    const int g = t / F, f = t % F;
    double in[G + 1];
    in[0] = lp.sc.fm_labor[f];                                             // laborHired after the person phase
    double inv_g = 0.0, x_g = 0.0;
#pragma unroll
    for (int k = 0; k < G; k++) {
        const double iv = lp.sc.ff_inv[(size_t)k * F + f];
        const double xk = iv * (double)p.ac.f_prod[(size_t)k * F + f];    // neuralFirmDecisionMaker.cpp:101
        in[k + 1] = xk;
        if (k == g) { inv_g = iv; x_g = xk; }
    }
    const size_t ag = (size_t)g * F + f;
    double share[G + 1], theta[G + 1];
#pragma unroll
    for (int i = 0; i <= G; i++) {
        const size_t k = ((size_t)g * (G + 1) + i) * F + f;
        share[i] = p.st.f_prod_share[k];
        theta[i] = (p.prod_kind == FASTACE_FN_STONE_GEARY) ? p.st.f_prod_theta[k] : 0.0;
    }
    const double outg = eval_function<G + 1, false>(p.prod_kind, p.st.f_prod_tfp[ag], share, theta, p.st.f_prod_rho[ag], in);
    const double newinv = inv_g + (outg - x_g);                            // profitMaxer.cpp:71
    p.st.f_inv[ag] = newinv;
    const double amount = (double)p.ac.f_offer_amt[ag] * newinv;           // decisionNetHandler.cpp:591
    lp.sc.post_lots[f * G + g] = x86_double_to_int(amount / kAmountPerOffer);
    const int n = lp.sc.own_offer[f * G + g];
    if (n >= 0) {   // counters of the old entry just before its owner withdraws it (profitMaxer.cpp:79-81)
        if (p.out.old_m_left) p.out.old_m_left[n] = lp.sc.ff_left[f * G + g];
        if (p.out.old_m_taken) p.out.old_m_taken[n] = lp.sc.ff_taken[f * G + g];
    }
    if (g == 0) {
        p.st.f_money[f] = lp.sc.ff_money[f];
        p.st.f_last_money[f] = lp.sc.ff_last[f];
        p.st.f_labor[f] = 0.0;                                             // firm.cpp:41
        p.out.f_profit[f] = lp.sc.ff_profit[f];
        lp.sc.post_jlots[f] = x86_double_to_int((double)p.ac.f_job_labor[f] / kLaborPerOffer);
    }
}

// new books in market order: firms in visiting order, goods ascending, lots > 0 (economy.cpp:52-59, 125-126).
// One block scans the per-rank offer counts ...
__global__ void large_post_scan(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, G = lp.G;
    __shared__ uint32_t part_m[1024], part_j[1024];
    const int per = (F + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min(F, (int)threadIdx.x * per), hi = min(F, lo + per);
    uint32_t sm = 0, sj = 0;
    for (int q = lo; q < hi; q++) {
        const int f = p.ac.perm_firm[q];
        for (int g = 0; g < G; g++) sm += lp.sc.post_lots[f * G + g] > 0;
        sj += lp.sc.post_jlots[f] > 0;
    }
    part_m[threadIdx.x] = sm; part_j[threadIdx.x] = sj;
    __syncthreads();
    for (int d = 1; d < (int)blockDim.x; d <<= 1) {
        const uint32_t a = threadIdx.x >= (unsigned)d ? part_m[threadIdx.x - d] : 0u;
        const uint32_t b = threadIdx.x >= (unsigned)d ? part_j[threadIdx.x - d] : 0u;
        __syncthreads();
        part_m[threadIdx.x] += a; part_j[threadIdx.x] += b;
        __syncthreads();
    }
    uint32_t bm = threadIdx.x ? part_m[threadIdx.x - 1] : 0u, bj = threadIdx.x ? part_j[threadIdx.x - 1] : 0u;
    for (int q = lo; q < hi; q++) {
        const int f = p.ac.perm_firm[q];
        lp.sc.post_base_m[q] = bm; lp.sc.post_base_j[q] = bj;
        for (int g = 0; g < G; g++) bm += lp.sc.post_lots[f * G + g] > 0;
        bj += lp.sc.post_jlots[f] > 0;
    }
    if (threadIdx.x == blockDim.x - 1) { p.st.m_count[0] = (int32_t)part_m[blockDim.x - 1]; p.st.j_count[0] = (int32_t)part_j[blockDim.x - 1]; }
}
// ... then every firm writes its entries, and what the new books do not cover is cleared (the old books are dead).
__global__ void large_post_write(const LargeParams lp) {
    const StepParams& p = lp.sp;
    const int F = p.F, G = lp.G, cap = F * G;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nm = p.st.m_count[0], nj = p.st.j_count[0];
    if (t >= nm && t < cap) { p.st.m_owner[t] = 0; p.st.m_good[t] = 0; p.st.m_left[t] = 0; p.st.m_taken[t] = 0; p.st.m_price[t] = 0.0; }
    if (t >= nj && t < F) { p.st.j_owner[t] = 0; p.st.j_left[t] = 0; p.st.j_taken[t] = 0; p.st.j_wage[t] = 0.0; }
    if (t >= F) return;
    const int q = t, f = p.ac.perm_firm[q];
    uint32_t bm = lp.sc.post_base_m[q];
    for (int g = 0; g < G; g++) {
        const int lots = lp.sc.post_lots[f * G + g];
        if (lots > 0) {                                                    // neuralFirmDecisionMaker.cpp:135
            p.st.m_owner[bm] = f; p.st.m_good[bm] = g; p.st.m_left[bm] = (uint32_t)lots; p.st.m_taken[bm] = 0;
            p.st.m_price[bm] = (double)p.ac.f_offer_price[(size_t)g * F + f] / kAmountPerOffer;
            bm++;
        }
    }
    const int jl = lp.sc.post_jlots[f];
    if (jl > 0) {
        double wage = (double)p.ac.f_job_wage[f];
        if (wage > kLargeNumber) wage = kLargeNumber;                       // decisionNetHandler.cpp:631-635
        const uint32_t bj = lp.sc.post_base_j[q];
        p.st.j_owner[bj] = f; p.st.j_left[bj] = (uint32_t)jl; p.st.j_taken[bj] = 0; p.st.j_wage[bj] = wage / kLaborPerOffer;
    }
}

}  // namespace fastace
