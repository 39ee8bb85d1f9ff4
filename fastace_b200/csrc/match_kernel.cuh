// match_kernel<G,SMAX> — all matching of one Economy::time_step, one warp per economy (kernel v4).
//
// What it computes (reference, paths under /root/reference/src): the person phase and the firm phase of
// Economy::time_step (base/economy.cpp:113-123) as far as they touch OTHER agents — labour acceptance
// (persons/utilMaxer.cpp:76-85, base/person.cpp:36-54, base/firm.cpp:56-113), goods purchases of persons and
// firms (utilMaxer.cpp:64-73, firms/profitMaxer.cpp:102-111, base/agent.cpp:99-161), the firms' stale-offer check
// (agent.cpp:54-97), profit record (neural/neuralFirmDecisionMaker.cpp:65-74) and withdrawal of last step's offers
// (profitMaxer.cpp:79-81, 93-95).  Everything that touches only the agent itself (consumption, utility,
// production, the new offers) is update_kernel's.
//
// Shared memory holds the economy's two books (a 32-byte record and three 32-byte lane vectors per offer) and the
// firms' money / inventories.  Person
// state and requests are never staged: a person's money and its 2S request slots are gathered straight from HBM
// (L1/L2-resident after the first window) into the registers of the lane that owns it.
//
// Persons, lane-parallel and exact.  A window = 32 persons with consecutive visiting ranks, lane = rank in the
// window.  For offer row R the ordinal number of an ELIGIBLE request (the person-side test passed:
// laborSupplied + 0.5 <= 1, person.cpp:39 / money >= price, agent.cpp:102) is
//     ord = (# eligible requests on R of lower lanes) + (own earlier ones),
// and the request succeeds iff ord < D[R], the row's death ordinal of the window: lots left for a job offer,
// lowered to the first request its firm cannot pay (firm.cpp:80-84: that request kills the offer);
// min(lots left, floor(seller inventory)) for a goods offer (agent.cpp:124, 140-143).  Offers only ever lose
// availability, so the serial first-come-first-served outcome is the unique fixed point of
//     evaluate the chains (lanes = persons)  ->  room[R][lane] = D[R] - prefix over lower lanes (scan)  -> ...
// Byte matrices per offer row: cnt[R][lane] = eligible requests of the lane in the current evaluation,
// room[R][lane] = how many of them can succeed, prev[R][lane] = the counts the rooms were computed for.  A row
// whose demand fits (tot <= D) keeps its initial rooms (= D for every lane: nothing can fail); only rows that are
// over-subscribed AND whose counts moved since their last scan are re-scanned (lanes = persons, one shuffle scan
// per row).  The iteration stops when no re-scan changes min(cnt, room) in any cell: the next evaluation would
// reproduce this one.  Lane k is exact two rounds after its lower lanes at the latest.
//
// Floating point.  Every fp64 quantity is updated in the reference's order, one operation per event:
//   person money    in the person's own request order (registers);
//   seller inventory  unit subtractions (exact in fp64);
//   firm money      hire by hire / sale by sale in the visiting order.  A window without sales folds `money -=
//                   wage` once per hire; a window with sales sorts its events by firm (counting sort over
//                   (firm, lane), stable in request order) and each firm folds its own list.  The "firm cannot
//                   pay" test is made on the exactly ordered running money.
// So money, labour and lot counters are bit-identical to the reference's; only pow() results differ in the last
// ulp (update_kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

// ---- offer rows ------------------------------------------------------------------------------------
// Every offer has a MATRIX row (three 32-byte lane vectors) and a RECORD row.  Job offers are rows 0..F
// (row NJ = "no request"), goods offers rows F+1.. (row F+1+NM likewise).
constexpr int kMatBytes = 96;
constexpr int kMatCnt = 0;      // u8[32]  eligible requests per lane, current evaluation
constexpr int kMatRoom = 32;    // u8[32]  how many of a lane's eligible requests can succeed
constexpr int kMatPrev = 64;    // u8[32]  the counts the rooms were last scanned for
constexpr int kRecBytes = 32;
constexpr int kRecValue = 0;    // f64     wage per lot (job row) / price (goods row)
constexpr int kRecLeft = 8;     // u32     BaseOffer::amountLeft
constexpr int kRecTaken = 12;   // u32     BaseOffer::amountTaken
constexpr int kRecD = 16;       // i32     death ordinal of the window; firm phase: amountLeft at withdrawal
constexpr int kRecTot = 20;     // i32     eligible requests of the window; after the commit: successes
constexpr int kRecMeta = 24;    // u32     owner | good << 8 | rescanned << 16
constexpr int kRoomMax = 127;   // rooms are clamped here (a lane has at most FASTACE_MAX_STACK requests)
constexpr int kRoundCap = 200;  // > 2 * 32 + 2, the proven bound: reaching it raises kDevErrRounds
constexpr int kEvPerLane = 2 + FASTACE_MAX_STACK;   // successes of one person: <= 2 hires + S purchases

// device error words (fastace_env_t::dev_err, host-mapped memory), surfaced as FASTACE_ERR_NOT_CONVERGED by
// fastace_env_sync, fastace_env_get_state and the next step call
constexpr int kDevErrRounds = 0;        // the window iteration hit kRoundCap (would break bit-exactness)
constexpr int kDevErrLargeRounds = 1;   // the large-economy iteration hit its round cap

struct MatchLayout {
    int off_mat;                        // matrix rows; between windows' evaluations the first 96*F bytes double as the
                                        // counting-sort scratch of the firm-money fold (evcnt u8 [F][32], evpos u16 [F][32])
    int off_rec;                        // record rows
    int off_fmoney, off_finv, off_flast;                  // double
    int off_fnh, off_fok, off_flive;                      // u32
    int off_evbase, off_evtot, off_permf;                 // u16
    int off_fatt;                                         // u8 [F][16]
    int off_fjob, off_ffirst, off_fcnt, off_frisk;        // u8 [F]
    int off_evlist;                                       // u8 [32][kEvPerLane]
    int total;
};

__host__ __device__ inline MatchLayout make_match_layout(int P, int F, int G, int S) {
    (void)P; (void)S;
    MatchLayout L;
    const int rows = F + 1 + F * G + 1;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.off_mat = take(kMatBytes * rows);
    L.off_rec = take(kRecBytes * rows);
    L.off_fmoney = take(8 * F);
    L.off_finv = take(8 * G * F);
    L.off_flast = take(8 * F);
    L.off_fnh = take(4 * F);
    L.off_fok = take(4 * F);
    L.off_flive = take(4 * F);
    L.off_evbase = take(2 * F);
    L.off_evtot = take(2 * F);
    L.off_permf = take(2 * F);
    L.off_fatt = take(16 * F);
    L.off_fjob = take(F);
    L.off_ffirst = take(F);
    L.off_fcnt = take(F);
    L.off_frisk = take(F);
    L.off_evlist = take(32 * kEvPerLane);
    L.total = o;
    return L;
}

struct MatchParams {
    StepParams sp;
    MatchLayout lay;     // computed on the host: offsets come from the constant bank
    uint8_t* scr_pnh;    // [E][P]    hires per person (0..2)         -> update_kernel
    uint8_t* scr_pnb;    // [E][G][P] purchases per person and good   -> update_kernel
    volatile uint32_t* dev_err;   // device error words of the env
};

__device__ __forceinline__ uint32_t& rec_u32(unsigned char* rec, int off) { return *reinterpret_cast<uint32_t*>(rec + off); }
__device__ __forceinline__ int32_t& rec_i32(unsigned char* rec, int off) { return *reinterpret_cast<int32_t*>(rec + off); }
__device__ __forceinline__ double& rec_value(unsigned char* rec) { return *reinterpret_cast<double*>(rec + kRecValue); }

// inclusive prefix sum over the lanes of a warp
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// byte k of a list kept 4 entries per register
template <int N>
__device__ __forceinline__ uint32_t list_byte(const uint32_t (&w)[N], int k) {
    uint32_t v = w[0];
#pragma unroll
    for (int q = 1; q < N; q++) if ((k >> 2) == q) v = w[q];
    return (v >> (8 * (k & 3))) & 0xFFu;
}
template <int N>
__device__ __forceinline__ void list_set_byte(uint32_t (&w)[N], int k, uint32_t b) {
    const uint32_t sh = 8u * (uint32_t)(k & 3);
#pragma unroll
    for (int q = 0; q < N; q++) if ((k >> 2) == q) w[q] = (w[q] & ~(0xFFu << sh)) | (b << sh);
}

// SMAX: compile-time bound of the stack size S (12 or 16).  Request lists are RIGHT-ALIGNED in SMAX/4
// registers per list (slot i at position i + SMAX - S) so that the fully unrolled evaluation is entered at
// position SMAX - S through one jump and runs without per-slot bound checks.
template <int G, int SMAX>
__global__ void __launch_bounds__(32, 28) match_kernel(const MatchParams mp) {
    FASTACE_DYN_SMEM(smem);
    const StepParams& p = mp.sp;
    const int e = blockIdx.x;
    const int lane = threadIdx.x;
    const int P = p.P, F = p.F, S = p.S;
    const int cap = F * G;
    const MatchLayout& L = mp.lay;
    constexpr int NW = SMAX / 4;
    const int skip = SMAX - S;          // first occupied position of a request list

    unsigned char* matJ = smem + L.off_mat;
    unsigned char* matM = matJ + (F + 1) * kMatBytes;
    unsigned char* recJ = smem + L.off_rec;
    unsigned char* recM = recJ + (F + 1) * kRecBytes;
    double* s_fmoney = reinterpret_cast<double*>(smem + L.off_fmoney);
    double* s_finv = reinterpret_cast<double*>(smem + L.off_finv);
    double* s_flast = reinterpret_cast<double*>(smem + L.off_flast);
    uint32_t* s_fnh = reinterpret_cast<uint32_t*>(smem + L.off_fnh);
    uint32_t* s_fok = reinterpret_cast<uint32_t*>(smem + L.off_fok);
    uint32_t* s_flive = reinterpret_cast<uint32_t*>(smem + L.off_flive);
    uint8_t* s_evcnt = matJ;                                              // [F][32], commit only
    uint16_t* s_evpos = reinterpret_cast<uint16_t*>(matJ + 32 * F);       // [F][32], commit only
    uint16_t* s_evbase = reinterpret_cast<uint16_t*>(smem + L.off_evbase);
    uint16_t* s_evtot = reinterpret_cast<uint16_t*>(smem + L.off_evtot);
    uint16_t* s_permf = reinterpret_cast<uint16_t*>(smem + L.off_permf);
    uint8_t* s_fatt = smem + L.off_fatt;
    uint8_t* s_fjob = smem + L.off_fjob;
    uint8_t* s_ffirst = smem + L.off_ffirst;
    uint8_t* s_fcnt = smem + L.off_fcnt;
    uint8_t* s_frisk = smem + L.off_frisk;
    uint8_t* s_evlist = smem + L.off_evlist;

    const size_t eP = (size_t)e * P, eF = (size_t)e * F, eCap = (size_t)e * cap;
    // a step may be taken in two calls (FASTACE_STEP_PERSONS, then FASTACE_STEP_FIRMS): the state in HBM between
    // them is the economy as it stands when the last person has acted and no firm has (economy.cpp:118-123)
    const bool do_persons = !(p.flags & FASTACE_STEP_FIRMS);
    const bool do_firms = !(p.flags & (FASTACE_STEP_PERSONS | FASTACE_STEP_PERSONS_TRADE));
    const int NM = p.st.m_count[e];
    const int NJ = p.st.j_count[e];
    const int NR = NJ + NM;
    // combined row index R: job offers first, then goods offers
    auto mat_of = [&](int R) { return R < NJ ? matJ + R * kMatBytes : matM + (R - NJ) * kMatBytes; };
    auto rec_of = [&](int R) { return R < NJ ? recJ + R * kRecBytes : recM + (R - NJ) * kRecBytes; };

    // ------------------------------ stage: books and firms ---------------------------------
    const IndexMap mapJ(NJ, p.flags), mapM(NM, p.flags);
    for (int n = lane; n <= NJ; n += 32) {
        unsigned char* rec = recJ + n * kRecBytes;
        const bool real = n < NJ;   // row NJ = "no request": never has room, costs and pays nothing
        rec_value(rec) = real ? p.st.j_wage[eF + n] : 0.0;
        rec_u32(rec, kRecLeft) = real ? p.st.j_left[eF + n] : 0u;
        rec_u32(rec, kRecTaken) = real ? p.st.j_taken[eF + n] : 0u;
        rec_u32(rec, kRecMeta) = real ? (uint32_t)(p.st.j_owner[eF + n] & 0xFF) : 0u;
    }
    for (int n = lane; n <= NM; n += 32) {
        unsigned char* rec = recM + n * kRecBytes;
        const bool real = n < NM;
        rec_value(rec) = real ? p.st.m_price[eCap + n] : 0.0;
        rec_u32(rec, kRecLeft) = real ? p.st.m_left[eCap + n] : 0u;
        rec_u32(rec, kRecTaken) = real ? p.st.m_taken[eCap + n] : 0u;
        rec_u32(rec, kRecMeta) = real ? (uint32_t)(p.st.m_owner[eCap + n] & 0xFF) | ((uint32_t)(p.st.m_good[eCap + n] & 0xFF) << 8) : 0u;
    }
    for (int f = lane; f < F; f += 32) {
        s_fmoney[f] = p.st.f_money[eF + f];
        s_permf[f] = do_firms ? (uint16_t)perm_firm_at(p, eF + f) : (uint16_t)f;
        s_fnh[f] = 0;
        s_fok[f] = 0;
        s_fcnt[f] = 0;
        s_ffirst[f] = 0;
        s_frisk[f] = 0;
        s_fjob[f] = (uint8_t)kNone;
#pragma unroll
        for (int g = 0; g < G; g++) s_finv[g * F + f] = p.st.f_inv[((size_t)e * G + g) * F + f];
        if (do_firms) {
            s_flast[f] = p.st.f_last_money[eF + f];
            const size_t k0 = (size_t)e * S * F + f;
            uint32_t w[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
            if (NM > 0) {   // empty book: no requests at all (decisionNetHandler.cpp:398-403)
                const uint32_t tg = p.compact ? p.cz.f_good_take[eF + f] : 0u;
#pragma unroll
                for (int i = 0; i < kMaxStack; i++) {
                    if (i < S) {
                        const bool take = p.compact ? ((tg >> i) & 1u) != 0 : p.ac.f_good_take[k0 + (size_t)i * F] != 0;
                        const int raw = p.compact ? (int)p.cz.f_good_idx[k0 + (size_t)i * F] : p.ac.f_good_idx[k0 + (size_t)i * F];
                        list_set_byte<4>(w, i, take ? (uint32_t)mapM(raw) : (uint32_t)kNone);
                    }
                }
            }
            *reinterpret_cast<uint4*>(s_fatt + f * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncwarp();
    for (int n = lane; n < NJ; n += 32) s_fjob[rec_u32(recJ + n * kRecBytes, kRecMeta) & 0xFFu] = (uint8_t)n;
    for (int n = lane; n < NM; n += 32) {
        // a firm's entries are contiguous in market order (it posts all its goods in one turn)
        const uint32_t owner = rec_u32(recM + n * kRecBytes, kRecMeta) & 0xFFu;
        const int prev = (n > 0) ? (int)(rec_u32(recM + (n - 1) * kRecBytes, kRecMeta) & 0xFFu) : -1;
        if ((int)owner != prev) s_ffirst[owner] = (uint8_t)n;
        if (!(rec_value(recM + n * kRecBytes) >= 0.0)) s_frisk[owner] = 1;   // a sale would not raise the seller's money
    }
    __syncwarp();
    for (int n = lane; n < NM; n += 32) {
        const uint32_t owner = rec_u32(recM + n * kRecBytes, kRecMeta) & 0xFFu;
        const int next = (n + 1 < NM) ? (int)(rec_u32(recM + (n + 1) * kRecBytes, kRecMeta) & 0xFFu) : -1;
        if ((int)owner != next) s_fcnt[owner] = (uint8_t)(n + 1 - s_ffirst[owner]);
    }
    __syncwarp();

    // ------------------------------ persons: windows of 32 visiting ranks -------------------
    const size_t row0 = (size_t)e * S * P;
    for (int base = 0; do_persons && base < P; base += 32) {
        // ---- (1) rows: death ordinals and initial rooms of the window
        bool lj = false, lm = false;
        for (int R = lane; R < NR; R += 32) {
            unsigned char* rec = rec_of(R);
            const uint32_t left = rec_u32(rec, kRecLeft);
            uint32_t d = left;
            if (R < NJ) {
                lj |= left > 0;
            } else {
                const uint32_t meta = rec_u32(rec, kRecMeta);
                const int sel = (int)(meta & 0xFFu), good = (int)((meta >> 8) & 0xFFu);
                d = min(left, unit_sales_possible(s_finv[good * F + sel]));
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (g != good && s_finv[g * F + sel] < 0.0) d = 0;   // agent.cpp:140 on a zero quantity
                lm |= left > 0;
            }
            const int di = (int)min(d, 0x7FFFFFFFu);
            rec_i32(rec, kRecD) = di;
            rec_u32(rec, kRecMeta) &= 0xFFFFu;   // not re-scanned yet
            const uint32_t rm = (uint32_t)min(di, kRoomMax) * 0x01010101u;
            uint4* q = reinterpret_cast<uint4*>(mat_of(R));
            q[0] = make_uint4(0u, 0u, 0u, 0u);
            q[1] = make_uint4(0u, 0u, 0u, 0u);
            q[2] = make_uint4(rm, rm, rm, rm);
            q[3] = make_uint4(rm, rm, rm, rm);
        }
        if (lane >= 30) {   // the two "no request" rows never have room (their bytes double as sort scratch)
            uint4* q = reinterpret_cast<uint4*>(lane == 30 ? matJ + NJ * kMatBytes : matM + NM * kMatBytes);
            q[2] = make_uint4(0u, 0u, 0u, 0u);
            q[3] = make_uint4(0u, 0u, 0u, 0u);
        }
        const bool liveJ = __any_sync(0xffffffffu, lj), liveM = __any_sync(0xffffffffu, lm);
        if (!liveJ && !liveM) {
            // both books are sold out: no person from here on can trade
            if (lane == 0) FASTACE_STAT(kStatDeadExits, 1);
            for (int r = base + lane; r < P; r += 32) {
                const int pid = p.compact ? (int)p.cz.perm_person[eP + r] : p.ac.perm_person[eP + r];
                mp.scr_pnh[eP + pid] = 0;
#pragma unroll
                for (int g = 0; g < G; g++) mp.scr_pnb[((size_t)e * G + g) * P + pid] = 0;
                write_person_ok(p, e, pid, 0u);
            }
            break;
        }
        // ---- (2) the lane's person: money and request lists (positions skip..SMAX-1; "no request" = row NJ / NM)
        const int r = base + lane;
        const bool active = r < P;
        const int pid = active ? (p.compact ? (int)p.cz.perm_person[eP + r] : p.ac.perm_person[eP + r]) : 0;
        const double money0 = active ? p.st.p_money[eP + pid] : 0.0;
        uint32_t aj[NW], ag[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) { aj[k] = (uint32_t)NJ * 0x01010101u; ag[k] = (uint32_t)NM * 0x01010101u; }
        if (active) {
            const uint32_t tj = (p.compact && liveJ) ? p.cz.p_job_take[eP + pid] : 0u;
            const uint32_t tg = (p.compact && liveM) ? p.cz.p_good_take[eP + pid] : 0u;
#pragma unroll
            for (int k = 0; k < SMAX; k++) {
                const int i = k - skip;
                if (i >= 0) {
                    const size_t a = row0 + (size_t)i * P + pid;
                    if (liveJ) {
                        const bool take = p.compact ? ((tj >> i) & 1u) != 0 : p.ac.p_job_take[a] != 0;
                        const int raw = p.compact ? (int)p.cz.p_job_idx[a] : p.ac.p_job_idx[a];
                        uint32_t n = take ? (uint32_t)mapJ(raw) : (uint32_t)kNone;
                        if (n == (uint32_t)kNone) n = (uint32_t)NJ;
                        aj[k >> 2] = (aj[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (n << (8 * (k & 3)));
                    }
                    if (liveM) {
                        const bool take = p.compact ? ((tg >> i) & 1u) != 0 : p.ac.p_good_take[a] != 0;
                        const int raw = p.compact ? (int)p.cz.p_good_idx[a] : p.ac.p_good_idx[a];
                        uint32_t n = take ? (uint32_t)mapM(raw) : (uint32_t)kNone;
                        if (n == (uint32_t)kNone) n = (uint32_t)NM;
                        ag[k >> 2] = (ag[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (n << (8 * (k & 3)));
                    }
                }
            }
        }
        __syncwarp();

        // ---- (3) fixed-point rounds
        double money = money0;
        int nh = 0;
        uint32_t okm = 0;      // successes: job positions in bits 0..15, goods positions in bits 16..31
        unsigned char* const matJl = matJ + lane;
        unsigned char* const matMl = matM + lane;
        if (lane == 0) FASTACE_STAT(kStatWindows, 1);
        for (int round = 0;; round++) {
            if (lane == 0) { FASTACE_STAT(kStatRounds, 1); FASTACE_STAT(kStatRoundsW0 + min(base >> 5, 3), 1); }
            money = money0; nh = 0; okm = 0;
            if (liveJ) {
                // utilMaxer.cpp:76-85; eligible = person.cpp:39 (0.5 * nh + 0.5 <= 1)
#define FASTACE_JOB_SLOT(K)                                                                   \
    case K: if (K < SMAX) {                                                                   \
        const uint32_t n = (aj[(K) >> 2] >> (8 * ((K) & 3))) & 0xFFu;                         \
        unsigned char* cell = matJl + n * kMatBytes;                                          \
        const uint32_t c = cell[kMatCnt];                                                     \
        const bool el = nh < 2;                                                               \
        if (el) cell[kMatCnt] = (uint8_t)(c + 1u);                                            \
        if (el && c < cell[kMatRoom]) {                                                       \
            nh++;                                                                             \
            money += rec_value(recJ + n * kRecBytes);   /* person.cpp:49 */                   \
            okm |= 1u << (K);                                                                 \
        }                                                                                     \
    }
                switch (skip) {
                    FASTACE_JOB_SLOT(0) FASTACE_JOB_SLOT(1) FASTACE_JOB_SLOT(2) FASTACE_JOB_SLOT(3)
                    FASTACE_JOB_SLOT(4) FASTACE_JOB_SLOT(5) FASTACE_JOB_SLOT(6) FASTACE_JOB_SLOT(7)
                    FASTACE_JOB_SLOT(8) FASTACE_JOB_SLOT(9) FASTACE_JOB_SLOT(10) FASTACE_JOB_SLOT(11)
                    FASTACE_JOB_SLOT(12) FASTACE_JOB_SLOT(13) FASTACE_JOB_SLOT(14) FASTACE_JOB_SLOT(15)
                    default: break;
                }
#undef FASTACE_JOB_SLOT
            }
            if (liveM) {
                // utilMaxer.cpp:64-73; eligible = agent.cpp:102
#define FASTACE_GOOD_SLOT(K)                                                                  \
    case K: if (K < SMAX) {                                                                   \
        const uint32_t n = (ag[(K) >> 2] >> (8 * ((K) & 3))) & 0xFFu;                         \
        unsigned char* cell = matMl + n * kMatBytes;                                          \
        const double price = rec_value(recM + n * kRecBytes);                                 \
        const uint32_t c = cell[kMatCnt];                                                     \
        const bool el = money >= price;                                                       \
        if (el) cell[kMatCnt] = (uint8_t)(c + 1u);                                            \
        if (el && c < cell[kMatRoom]) {                                                       \
            money -= price;                       /* agent.cpp:108 */                         \
            okm |= 1u << (16 + (K));                                                          \
        }                                                                                     \
    }
                switch (skip) {
                    FASTACE_GOOD_SLOT(0) FASTACE_GOOD_SLOT(1) FASTACE_GOOD_SLOT(2) FASTACE_GOOD_SLOT(3)
                    FASTACE_GOOD_SLOT(4) FASTACE_GOOD_SLOT(5) FASTACE_GOOD_SLOT(6) FASTACE_GOOD_SLOT(7)
                    FASTACE_GOOD_SLOT(8) FASTACE_GOOD_SLOT(9) FASTACE_GOOD_SLOT(10) FASTACE_GOOD_SLOT(11)
                    FASTACE_GOOD_SLOT(12) FASTACE_GOOD_SLOT(13) FASTACE_GOOD_SLOT(14) FASTACE_GOOD_SLOT(15)
                    default: break;
                }
#undef FASTACE_GOOD_SLOT
            }
            __syncwarp();

            // ---- rows: demand of the window; over-subscribed rows whose counts moved get their rooms re-scanned
            bool changed = false;
            for (int cb = 0; cb < NR; cb += 32) {
                const int R = cb + lane;
                bool needs = false;
                if (R < NR) {
                    unsigned char* rec = rec_of(R);
                    const uint4* q = reinterpret_cast<const uint4*>(mat_of(R));
                    const uint4 a = q[0], b = q[1];
                    const uint32_t s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;          // per byte <= 8 * 16
                    const uint32_t t = (s & 0x00FF00FFu) + ((s >> 8) & 0x00FF00FFu);
                    const int tot = (int)((t + (t >> 16)) & 0xFFFFu);
                    rec_i32(rec, kRecTot) = tot;
                    const int d = rec_i32(rec, kRecD);
                    if (d > 0 && tot > 0) {
                        if ((rec_u32(rec, kRecMeta) >> 16) != 0) {
                            // scanned before: its rooms stand as long as the counts they were computed for do
                            const uint4 c = q[4], dd = q[5];
                            needs = ((a.x ^ c.x) | (a.y ^ c.y) | (a.z ^ c.z) | (a.w ^ c.w) |
                                     (b.x ^ dd.x) | (b.y ^ dd.y) | (b.z ^ dd.z) | (b.w ^ dd.w)) != 0u;
                        } else {
                            needs = tot > d;   // initial rooms (= D in every lane) cannot bind while the demand fits
                        }
                    }
                }
                unsigned slow = __ballot_sync(0xffffffffu, needs);
                while (slow) {
                    const int R0 = cb + __ffs((int)slow) - 1;
                    slow &= slow - 1;
                    unsigned char* cell = mat_of(R0) + lane;
                    unsigned char* rec = rec_of(R0);
                    if (lane == 0) { FASTACE_STAT(kStatRescans, 1); if (base == 0) FASTACE_STAT(kStatRescansW0, 1); }
                    const int c = cell[kMatCnt];
                    const int excl = warp_inclusive_scan(c, lane) - c;
                    const int rm = max(0, min(kRoomMax, rec_i32(rec, kRecD) - excl));
                    const int old = cell[kMatRoom];
                    changed |= min(c, rm) != min(c, old);
                    cell[kMatRoom] = (uint8_t)rm;
                    cell[kMatPrev] = (uint8_t)c;
                    if (lane == 0) rec_u32(rec, kRecMeta) |= 0x10000u;
                }
            }
            __syncwarp();

            // ---- job offers whose firm may run out of money (firm.cpp:80-84): exact running money in visiting order
            for (int cb = 0; cb < NJ; cb += 32) {
                const int R = cb + lane;
                bool risky = false;
                uint32_t left = 0;
                int f = 0;
                double w = 0.0, m0 = 0.0;
                unsigned char* rec = recJ + (R < NJ ? R : 0) * kRecBytes;
                unsigned char* mat = matJ + (R < NJ ? R : 0) * kMatBytes;
                if (R < NJ) {
                    left = rec_u32(rec, kRecLeft);
                    f = (int)(rec_u32(rec, kRecMeta) & 0xFFu);
                    w = rec_value(rec);
                    m0 = s_fmoney[f];
                    const int tot = rec_i32(rec, kRecTot);
                    const int most = min((int)min(left, 0x7FFFFFFFu), tot);
                    // sufficient for "every hire of the window can be paid": wages are subtracted one by one (rounding
                    // error << 1e-9 relative for most <= 512) and sales only add money unless a price is negative / NaN
                    const bool safe = w >= 0.0 && s_frisk[f] == 0 && (m0 - w * (double)most >= w * (1.0 + 1e-9));
                    risky = tot > 0 && left > 0 && !safe;
                }
                if (__any_sync(0xffffffffu, risky)) {
                    // does the firm sell anything in this window?  (its goods rows' successes = min(tot, D))
                    bool fsales = false;
                    if (risky) {
                        const int first = s_ffirst[f], cnt = s_fcnt[f];
                        for (int n = first; n < first + cnt; n++) {
                            unsigned char* g = recM + n * kRecBytes;
                            fsales |= min(rec_i32(g, kRecTot), rec_i32(g, kRecD)) > 0;
                        }
                    }
                    if (__any_sync(0xffffffffu, fsales)) {
                        // every lane publishes the goods it currently buys, in request order
                        int ne = 0;
                        for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                            const int k = __ffs((int)m) - 1;
                            s_evlist[lane * kEvPerLane + 1 + ne] = (uint8_t)list_byte<NW>(ag, k);
                            ne++;
                        }
                        s_evlist[lane * kEvPerLane] = (uint8_t)ne;
                        __syncwarp();
                    }
                    if (R < NJ) {
                        int d = (int)min(left, 0x7FFFFFFFu);
                        if (risky && !fsales) {
                            // hires are the firm's only events: the order of the applicants does not matter
                            FASTACE_STAT(kStatRiskyWalks, 1);
                            double m = m0;
                            const int most = min(d, rec_i32(rec, kRecTot));
                            for (int h = 0; h < most; h++) {
                                if (m < w) { d = h; break; }                            // firm.cpp:80-84
                                m -= w;                                                 // firm.cpp:108
                            }
                        } else if (risky) {
                            FASTACE_STAT(kStatRiskyWalks, 1);
                            double m = m0;
                            int h = 0;
                            bool done = false;
                            for (int l = 0; l < 32 && !done; l++) {
                                const int c = mat[kMatCnt + l];
                                for (int k = 0; k < c; k++) {
                                    if (h >= d) { done = true; break; }                 // firm.cpp:64
                                    if (m < w) { d = h; done = true; break; }           // firm.cpp:80-84
                                    m -= w;                                             // firm.cpp:108
                                    h++;
                                }
                                const int nb = s_evlist[l * kEvPerLane];
                                for (int k = 0; k < nb && !done; k++) {
                                    unsigned char* g = recM + s_evlist[l * kEvPerLane + 1 + k] * kRecBytes;
                                    if ((int)(rec_u32(g, kRecMeta) & 0xFFu) == f) m += rec_value(g);   // agent.cpp:158
                                }
                            }
                        }
                        if (d != rec_i32(rec, kRecD)) {
                            // the offer dies earlier / later than assumed: its rooms follow the new ordinal
                            changed = true;
                            rec_i32(rec, kRecD) = d;
                            rec_u32(rec, kRecMeta) |= 0x10000u;
                            int run = 0;
                            for (int l = 0; l < 32; l++) {
                                const int c = mat[kMatCnt + l];
                                mat[kMatRoom + l] = (uint8_t)max(0, min(kRoomMax, d - run));
                                mat[kMatPrev + l] = (uint8_t)c;
                                run += c;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            const bool again = __any_sync(0xffffffffu, changed);
            if (!again) break;
            if (round >= kRoundCap) {
                if (lane == 0) mp.dev_err[kDevErrRounds] = 1u;
                break;
            }
            for (int R = lane; R < NR; R += 32) {
                uint4* q = reinterpret_cast<uint4*>(mat_of(R));
                q[0] = make_uint4(0u, 0u, 0u, 0u);
                q[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
        }

        // ---- (4) commit the window
        const bool sales = __any_sync(0xffffffffu, (okm >> 16) != 0);
        if (active) {
            p.st.p_money[eP + pid] = money;
            mp.scr_pnh[eP + pid] = (uint8_t)nh;
            uint32_t nb[G];
#pragma unroll
            for (int g = 0; g < G; g++) nb[g] = 0;
            for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {          // one purchase per set bit
                const int k = __ffs((int)m) - 1;
                const uint32_t good = (rec_u32(recM + list_byte<NW>(ag, k) * kRecBytes, kRecMeta) >> 8) & 0xFFu;
#pragma unroll
                for (int g = 0; g < G; g++) nb[g] += (good == (uint32_t)g);
            }
#pragma unroll
            for (int g = 0; g < G; g++) mp.scr_pnb[((size_t)e * G + g) * P + pid] = (uint8_t)nb[g];
            write_person_ok(p, e, pid, ((okm & 0xFFFFu) >> skip) | (((okm >> 16) >> skip) << 16));
        }
        for (int R = lane; R < NR; R += 32) {
            unsigned char* rec = rec_of(R);
            const int tot = rec_i32(rec, kRecTot);
            const int n = min(tot, rec_i32(rec, kRecD));
            const uint32_t left = rec_u32(rec, kRecLeft);
            if (R < NJ) {
                rec_u32(rec, kRecLeft) = (tot > n) ? 0u : left - (uint32_t)n;    // exhausted or killed (firm.cpp:83)
            } else {
                uint32_t nl = left - (uint32_t)n;
                if (tot > n && nl > 0) nl = 0;                                    // killed (agent.cpp:143)
                rec_u32(rec, kRecLeft) = nl;
                const uint32_t meta = rec_u32(rec, kRecMeta);
                s_finv[((meta >> 8) & 0xFFu) * F + (meta & 0xFFu)] -= (double)n; // n exact unit subtractions (agent.cpp:156)
            }
            rec_u32(rec, kRecTaken) += (uint32_t)n;
            rec_i32(rec, kRecTot) = n;
        }
        __syncwarp();
        if (!sales) {
            // the window's only events at a firm are hires: firm.cpp:108, one subtraction per hire
            for (int f = lane; f < F; f += 32) {
                const int j = s_fjob[f];
                if (j != kNone) {
                    unsigned char* rec = recJ + j * kRecBytes;
                    const int h = rec_i32(rec, kRecTot);
                    const double w = rec_value(rec);
                    double m = s_fmoney[f];
                    for (int k = 0; k < h; k++) m -= w;
                    s_fmoney[f] = m;
                    s_fnh[f] += (uint32_t)h;
                }
            }
        } else {
            // events sorted by firm, stable in (lane, jobs before goods, request order): counting sort over (firm, lane)
            if (lane == 0) FASTACE_STAT(kStatSalesWindows, 1);
            for (int k = lane; k < F * 8; k += 32) reinterpret_cast<uint32_t*>(s_evcnt)[k] = 0u;
            __syncwarp();
            if (active) {
                for (uint32_t m = okm & 0xFFFFu; m != 0; m &= m - 1) {
                    const int k = __ffs((int)m) - 1;
                    const uint32_t f = rec_u32(recJ + list_byte<NW>(aj, k) * kRecBytes, kRecMeta) & 0xFFu;
                    s_evcnt[f * 32 + lane] += 1;
                }
                for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                    const int k = __ffs((int)m) - 1;
                    const uint32_t f = rec_u32(recM + list_byte<NW>(ag, k) * kRecBytes, kRecMeta) & 0xFFu;
                    s_evcnt[f * 32 + lane] += 1;
                }
            }
            __syncwarp();
            int carry = 0;
            for (int fb = 0; fb < F; fb += 32) {
                const int f = fb + lane;
                int run = 0;
                if (f < F) {
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(s_evcnt + f * 32);
                    uint32_t* pw = reinterpret_cast<uint32_t*>(s_evpos + f * 32);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t x = cw[q];
                        const int b0 = x & 0xFFu, b1 = (x >> 8) & 0xFFu, b2 = (x >> 16) & 0xFFu, b3 = x >> 24;
                        pw[2 * q] = (uint32_t)run | ((uint32_t)(run + b0) << 16);
                        pw[2 * q + 1] = (uint32_t)(run + b0 + b1) | ((uint32_t)(run + b0 + b1 + b2) << 16);
                        run += b0 + b1 + b2 + b3;
                    }
                }
                const int incl = warp_inclusive_scan(run, lane);
                if (f < F) { s_evbase[f] = (uint16_t)(carry + incl - run); s_evtot[f] = (uint16_t)run; }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            // (a hire is list entry 0, a sale of the firm's k-th market entry is 1 + k)
            if (active) {
                for (uint32_t m = okm & 0xFFFFu; m != 0; m &= m - 1) {
                    const int k = __ffs((int)m) - 1;
                    const uint32_t f = rec_u32(recJ + list_byte<NW>(aj, k) * kRecBytes, kRecMeta) & 0xFFu;
                    const uint32_t pos = s_evpos[f * 32 + lane];
                    s_evpos[f * 32 + lane] = (uint16_t)(pos + 1);
                    s_evlist[s_evbase[f] + pos] = 0;
                }
                for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                    const int k = __ffs((int)m) - 1;
                    const uint32_t n = list_byte<NW>(ag, k);
                    const uint32_t f = rec_u32(recM + n * kRecBytes, kRecMeta) & 0xFFu;
                    const uint32_t pos = s_evpos[f * 32 + lane];
                    s_evpos[f * 32 + lane] = (uint16_t)(pos + 1);
                    s_evlist[s_evbase[f] + pos] = (uint8_t)(1 + n - s_ffirst[f]);
                }
            }
            __syncwarp();
            for (int f = lane; f < F; f += 32) {
                const int j = s_fjob[f];
                const double w = (j != kNone) ? rec_value(recJ + j * kRecBytes) : 0.0;
                const int first = s_ffirst[f], b0 = s_evbase[f], n = s_evtot[f];
                double m = s_fmoney[f];
                uint32_t h = 0;
                for (int k = 0; k < n; k++) {
                    const int t = s_evlist[b0 + k];
                    if (t == 0) { m -= w; h++; }                                       // firm.cpp:108
                    else m += rec_value(recM + (first + t - 1) * kRecBytes);           // agent.cpp:158
                }
                s_fmoney[f] = m;
                s_fnh[f] += h;
            }
        }
        __syncwarp();
    }

    // job counters are final after the person phase
    if (do_persons && p.out.old_j_left) for (int n = lane; n < NJ; n += 32) p.out.old_j_left[eF + n] = rec_u32(recJ + n * kRecBytes, kRecLeft);
    if (do_persons && p.out.old_j_taken) for (int n = lane; n < NJ; n += 32) p.out.old_j_taken[eF + n] = rec_u32(recJ + n * kRecBytes, kRecTaken);
    if (!do_firms) {
        // persons-only call: the books' counters go back to HBM for the firms call (a full step never needs them
        // there: update_kernel replaces the books)
        for (int n = lane; n < NM; n += 32) {
            p.st.m_left[eCap + n] = rec_u32(recM + n * kRecBytes, kRecLeft);
            p.st.m_taken[eCap + n] = rec_u32(recM + n * kRecBytes, kRecTaken);
        }
        for (int n = lane; n < NJ; n += 32) {
            p.st.j_left[eF + n] = rec_u32(recJ + n * kRecBytes, kRecLeft);
            p.st.j_taken[eF + n] = rec_u32(recJ + n * kRecBytes, kRecTaken);
        }
    }

    // ------------------------------ firms ---------------------------------------------------
    if (do_firms) {
        // requests on entries that are sold out can only fail (amountLeft never grows within a step)
        bool lv = false;
        for (int f = lane; f < F; f += 32) {
            const uint4 slots = *reinterpret_cast<const uint4*>(s_fatt + f * 16);
            const uint32_t w[4] = {slots.x, slots.y, slots.z, slots.w};
            uint32_t live = 0;
#pragma unroll
            for (int i = 0; i < kMaxStack; i++) {
                const uint32_t n = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;      // kNone beyond S
                if (n != (uint32_t)kNone && rec_u32(recM + n * kRecBytes, kRecLeft) > 0) live |= 1u << i;
            }
            s_flive[f] = live;
            lv |= live != 0;
        }
        const bool anylive = __any_sync(0xffffffffu, lv);
        __syncwarp();
        // one firm's turn up to sell_goods (firm.cpp:23-37); `buys`: walk its live requests
        auto firm_turn = [&](int f, bool buys) {
            const int first = s_ffirst[f], cnt = s_fcnt[f];
            double money = s_fmoney[f];
            // Agent::check_my_offers (base/agent.cpp:54-97): running inventoryLeft over own entries
            {
                double invLeft[G];
#pragma unroll
                for (int g = 0; g < G; g++) invLeft[g] = s_finv[g * F + f];
                for (int n = first; n < first + cnt; n++) {
                    unsigned char* rec = recM + n * kRecBytes;
                    const int good = (int)((rec_u32(rec, kRecMeta) >> 8) & 0xFFu);
                    uint32_t left = rec_u32(rec, kRecLeft);
                    double delta = kAmountPerOffer * (double)left;     // agent.cpp:73 (other goods: 0*left = 0)
                    for (;;) {
                        bool okk = true;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            const double dg = (g == good) ? delta : 0.0;
                            if (dg > invLeft[g]) okk = false;
                        }
                        if (okk || left == 0) break;                   // left==0 guard: see SURVEY.md B.2
                        delta -= kAmountPerOffer;                      // agent.cpp:79-80
                        left--;
                    }
                    rec_u32(rec, kRecLeft) = left;
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == good) invLeft[g] -= delta;  // agent.cpp:83
                }
            }
            // first decision: profit of the previous step (neuralFirmDecisionMaker.cpp:65-74)
            {
                const double last = s_flast[f];
                s_flast[f] = (p.time_before > 0) ? (money - last) : 0.0;   // profit, written out below
                p.st.f_last_money[eF + f] = money;
            }
            // ProfitMaxer::buy_goods (firms/profitMaxer.cpp:102-111); the buyer's money stays in a register
            uint32_t ok = 0;
            if (buys) {
                for (uint32_t lm = s_flive[f]; lm != 0; lm &= lm - 1) {
                    const int i = __ffs((int)lm) - 1;
                    const int n = s_fatt[f * 16 + i];
                    unsigned char* rec = recM + n * kRecBytes;
                    const double price = rec_value(rec);
                    if (money >= price) {                                  // agent.cpp:102
                        const uint32_t left = rec_u32(rec, kRecLeft);
                        if (left > 0) {                                    // agent.cpp:124
                            const uint32_t meta = rec_u32(rec, kRecMeta);
                            const int s = (int)(meta & 0xFFu), good = (int)((meta >> 8) & 0xFFu);
                            bool short_ = false;                           // agent.cpp:140
#pragma unroll
                            for (int g = 0; g < G; g++) {
                                const double q = (g == good) ? kAmountPerOffer : 0.0;
                                if (s_finv[g * F + s] < q) short_ = true;
                            }
                            if (short_) {
                                rec_u32(rec, kRecLeft) = 0;                // agent.cpp:143
                            } else {
                                // seller first (agent.cpp:155-160), then buyer (agent.cpp:108-109)
                                if (s == f) money += price; else s_fmoney[s] += price;
                                s_finv[good * F + s] -= kAmountPerOffer;
                                rec_u32(rec, kRecLeft) = left - 1;
                                rec_u32(rec, kRecTaken) += 1;
                                money -= price;
                                s_finv[good * F + f] += kAmountPerOffer;
                                ok |= 1u << i;
                            }
                        }
                    }
                }
            }
            s_fmoney[f] = money;
            s_fok[f] = ok;
            // ProfitMaxer::sell_goods withdraws last step's offers (firms/profitMaxer.cpp:79-81);
            // nothing between buy_goods and that point touches another agent.
            for (int n = first; n < first + cnt; n++) {
                unsigned char* rec = recM + n * kRecBytes;
                rec_i32(rec, kRecD) = (int)rec_u32(rec, kRecLeft);   // final counter of the withdrawn entry
                rec_u32(rec, kRecLeft) = 0;
            }
        };
        if (!anylive) {
            // no purchase can happen: the firms' turns do not interact
            for (int f = lane; f < F; f += 32) firm_turn(f, false);
        } else if (lane == 0) {
            FASTACE_STAT(kStatFirmSerial, 1);
            for (int r = 0; r < F; r++) firm_turn((int)s_permf[r], true);   // visiting order (economy.cpp:121-123)
        }
        __syncwarp();
    }
    // ------------------------------ firms: results to HBM -----------------------------------
    if (do_firms && p.out.old_m_left) for (int n = lane; n < NM; n += 32) p.out.old_m_left[eCap + n] = (uint32_t)rec_i32(recM + n * kRecBytes, kRecD);
    if (do_firms && p.out.old_m_taken) for (int n = lane; n < NM; n += 32) p.out.old_m_taken[eCap + n] = rec_u32(recM + n * kRecBytes, kRecTaken);
    for (int f = lane; f < F; f += 32) {
        p.st.f_money[eF + f] = s_fmoney[f];
        if (do_firms) p.out.f_profit[eF + f] = s_flast[f];
#pragma unroll
        for (int g = 0; g < G; g++) p.st.f_inv[((size_t)e * G + g) * F + f] = s_finv[g * F + f];
        double labor = p.st.f_labor[eF + f];
        const uint32_t nhf = s_fnh[f];
        for (uint32_t k = 0; k < nhf; k++) labor += kLaborPerOffer;  // firm.cpp:109, one add per hire
        p.st.f_labor[eF + f] = labor;
        if (do_firms && p.out.f_good_ok) {
            const uint32_t ok = s_fok[f];
            for (int i = 0; i < S; i++) p.out.f_good_ok[((size_t)e * S + i) * F + f] = (ok >> i) & 1u;
        }
    }
}

}  // namespace fastace
