// match_kernel<G> — all matching of one Economy::time_step, one warp per economy (kernel v4).
//
// What it computes (reference, paths under /root/reference/src): the person phase and the firm phase of
// Economy::time_step (base/economy.cpp:113-123) as far as they touch OTHER agents — labour acceptance
// (persons/utilMaxer.cpp:76-85, base/person.cpp:36-54, base/firm.cpp:56-113), goods purchases of persons and
// firms (utilMaxer.cpp:64-73, firms/profitMaxer.cpp:102-111, base/agent.cpp:99-161), the firms' stale-offer check
// (agent.cpp:54-97), profit record (neural/neuralFirmDecisionMaker.cpp:65-74) and withdrawal of last step's offers
// (profitMaxer.cpp:79-81, 93-95).  Everything that touches only the agent itself (consumption, utility,
// production, the new offers) is update_kernel's.
//
// Shared memory holds the economy's two books (a 32-byte record and three 32-byte lane vectors per offer), the
// firms' money / inventories and, per window, the 32 persons' request lists as offer numbers (36 bytes per lane).
// Person state is never staged: a person's money and request slots go straight from HBM (prefetched into L2 by one
// bulk prefetch per array at kernel start) to the lane that owns it — three 32-bit loads per list in the compact
// encoding, whose index bytes are agent-major.
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
// Every offer has a MATRIX row (three 32-byte lane vectors) and a RECORD row.  Rows are numbered
//   job offer n -> n,   "no job request" -> NJ,   goods offer n -> NJ + 1 + n,   "no goods request" -> NJ + NM + 1,
// the two "no request" rows being ordinary rows with no lots: a request on them is eligible and never succeeds.
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
constexpr int kRecMeta = 24;    // u32     owner | good << 8
constexpr int kRecScanned = 28; // u8      the rooms of this window have been re-scanned (they follow `prev`)
constexpr int kRecEval = 29;    // u8      job rows: the death ordinal has been derived from the firm's money in this window
constexpr int kRecSafe = 30;    // u16     job rows: hires of this window the firm can certainly pay (see the prologue)
constexpr int kRoomMax = 127;   // rooms are clamped here (a lane has at most FASTACE_MAX_STACK requests)
constexpr int kRoundCap = 200;  // > 2 * 32 + 2, the proven bound: reaching it raises kDevErrRounds
constexpr int kEvPerLane = 2 + FASTACE_MAX_STACK;   // list of one lane's purchases: count byte + <= S rows
constexpr int kMaxHiresPerWindow = 64;              // 32 persons, at most two jobs each (person.cpp:39)
constexpr int kReqStride = 2 * FASTACE_MAX_STACK + 4;   // a lane's request lists: jobs at +0, goods at +16 (9 words: conflict-free)
constexpr int kReqGoods = FASTACE_MAX_STACK;

// device error words (fastace_env_t::dev_err, host-mapped memory), surfaced as FASTACE_ERR_NOT_CONVERGED by
// fastace_env_sync, fastace_env_get_state and the next step call
constexpr int kDevErrRounds = 0;        // the window iteration hit kRoundCap (would break bit-exactness)
constexpr int kDevErrLargeRounds = 1;   // the large-economy iteration hit its round cap
constexpr int kDevErrQueue = 2;         // update_kernel gave up waiting on the completion queue

// Completion queue (full steps): a warp of match_kernel that has finished its economy takes a ticket and writes
// (tag << 24 | economy) into slot `ticket - base`; block k of update_kernel, which may already be resident (programmatic
// dependent launch), processes the k-th economy to finish.  So the element-wise update of the early economies runs in
// the SM time the slow ones leave idle.  Tickets count up for ever (`base` = queue launches so far x E, mod 2^32).
constexpr int kQueueTagShift = 24;
constexpr uint32_t kQueueEconMask = (1u << kQueueTagShift) - 1u;

struct MatchLayout {
    int off_mat, off_rec;                                 // matrix rows, record rows (F + F*G + 2 each)
    int off_fmoney, off_finv, off_flast;                  // double
    int off_fnh, off_fok, off_flive;                      // u32
    int off_permf;                                        // u16
    int off_fatt;                                         // u8 [F][16]
    int off_fjob, off_ffirst, off_fcnt, off_frisk;        // u8 [F]
    int off_flabor;                                       // f64 [F] laborHired at the start of the step
    int off_evlist;                                       // u8 [32][kEvPerLane]
    int off_req;                                          // u8 [32][kReqStride]
    int total;
};

__host__ __device__ inline MatchLayout make_match_layout(int P, int F, int G, int S) {
    (void)P; (void)S;
    MatchLayout L;
    const int rows = F + F * G + 2;
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
    L.off_permf = take(2 * F);
    L.off_fatt = take(16 * F);
    L.off_fjob = take(F);
    L.off_ffirst = take(F);
    L.off_fcnt = take(F);
    L.off_frisk = take(F);
    L.off_flabor = take(8 * F);
    L.off_evlist = take(32 * kEvPerLane);
    L.off_req = take(32 * kReqStride);
    L.total = o;
    return L;
}

struct MatchParams {
    StepParams sp;
    MatchLayout lay;     // computed on the host: offsets come from the constant bank
    uint8_t* scr_pnh;    // [E][P]    hires per person (0..2)         -> update_kernel
    uint8_t* scr_pnb;    // [E][G][P] purchases per person and good   -> update_kernel
    volatile uint32_t* dev_err;   // device error words of the env
    uint32_t* done_list;          // completion queue [E] (null: update_kernel waits for the whole grid instead)
    uint32_t* done_count;         // its ticket counter
    uint32_t ticket_base, done_tag;
};

// inclusive prefix sum over the lanes of a warp
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// L2 prefetch of a contiguous array slab with one instruction (sm_90+ bulk prefetch; 16-byte granules)
__device__ __forceinline__ void prefetch_slab_l2(const void* ptr, size_t bytes) {
#ifndef FASTACE_HAVE_SMEM_OPS
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)15;
    const uint32_t n = (uint32_t)((reinterpret_cast<uintptr_t>(ptr) + bytes - a + 15) & ~(size_t)15);
    if (n != 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a), "r"(n) : "memory");
#else
    (void)ptr; (void)bytes;
#endif
}

// Four request bytes of the compact encoding -> four offer numbers.  `take4`: the four slots' take bits; a slot
// without a request becomes `dummy` (the "no request" row).
template <bool MODULO>
__device__ __forceinline__ uint32_t map_request_word(uint32_t x, uint32_t take4, const IndexMap& map, uint32_t dummy) {
    if (MODULO) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t n = map.mod24((x >> (8 * k)) & 0xFFu);
            r |= (((take4 >> k) & 1u) ? n : dummy) << (8 * k);
        }
        return r;
    }
    // absolute indices: keep byte b iff its take bit is set and b < count (count <= 254).  Byte-parallel: the even
    // and the odd bytes in 16-bit lanes, bit 8 of (b + 256 - count) says b >= count
    const uint32_t k16 = (256u - map.count) * 0x00010001u;
    const uint32_t ge_e = ((x & 0x00FF00FFu) + k16) & 0x01000100u;
    const uint32_t ge_o = (((x >> 8) & 0x00FF00FFu) + k16) & 0x01000100u;
    const uint32_t ge = (ge_e >> 8) | ge_o;                                  // 0x01 in every byte that is out of range
    const uint32_t tk = ((take4 & 0xFu) * 0x00204081u) & 0x01010101u;        // take bit k -> byte k
    const uint32_t keep = (tk & ~ge) * 0xFFu;                                // 0xFF in every byte that is a request
    return (x & keep) | ((dummy * 0x01010101u) & ~keep);
}

// All shared-memory traffic goes through explicit 32-bit shared addresses (common.cuh: lds_* / sts_*); the loops
// over request slots are rolled (the lists live in shared memory), so the hot code of a window is a few hundred
// instructions.
#ifdef FASTACE_CTA_TIMING
// profiling build only (tools/cta_timing.py): start / end time and SM of every economy's warp
__device__ unsigned long long g_cta_times[3 * 65536];
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

static_assert(kMaxStack == 16, "request lists are staged as four words");

// MODE specialises the kernel for an input encoding, so that the code of the other encodings is not interleaved with
// the hot path (a quarter of the generic kernel's instructions belong to paths a given launch never takes, and the
// instruction fetch stall is the one issue stall that is not inherent to the algorithm):
//   kModeGeneric  everything decided at run time (any encoding, any output set)
//   kModeCompact  compact encoding (full steps only), none of the optional outputs requested; bit kModeModulo:
//                 FASTACE_IDX_MODULO; bit kModeSmall: all rows of both books and all firms fit one pass of the warp
constexpr int kModeGeneric = 0, kModeCompact = 1, kModeModulo = 2, kModeSmall = 4;

template <int G, int MODE = kModeGeneric>
__global__ void __launch_bounds__(32, 28) match_kernel(const MatchParams mp) {
    FASTACE_DYN_SMEM(smem);
#ifdef FASTACE_CTA_TIMING
    const unsigned long long cta_t0 = global_ns();
#endif
    const StepParams& p = mp.sp;
    const bool compact = (MODE & kModeCompact) ? true : p.compact != 0;
    const bool modulo = (MODE & kModeCompact) ? (MODE & kModeModulo) != 0 : (p.flags & FASTACE_IDX_MODULO) != 0;
    const bool person_flags = (MODE & kModeCompact) ? false : (p.out.p_job_ok != nullptr || p.out.p_good_ok != nullptr);
    const int e = blockIdx.x;
    const int lane = threadIdx.x;
    const int P = p.P, F = p.F, S = p.S;
    const int cap = F * G;
    const MatchLayout& L = mp.lay;

    const uint32_t sb = smem_addr(smem);
    const uint32_t aMat = sb + L.off_mat, aRec = sb + L.off_rec;
    const uint32_t aFmoney = sb + L.off_fmoney, aFinv = sb + L.off_finv, aFlast = sb + L.off_flast;
    const uint32_t aFnh = sb + L.off_fnh, aFok = sb + L.off_fok, aFlive = sb + L.off_flive, aPermf = sb + L.off_permf;
    const uint32_t aFatt = sb + L.off_fatt, aFjob = sb + L.off_fjob, aFfirst = sb + L.off_ffirst, aFcnt = sb + L.off_fcnt;
    const uint32_t aFrisk = sb + L.off_frisk, aEv = sb + L.off_evlist, aFlabor = sb + L.off_flabor;
    const uint32_t aReqL = keep_u32(sb + L.off_req + (uint32_t)lane * kReqStride);   // this lane's request lists

    const size_t eP = (size_t)e * P, eF = (size_t)e * F, eCap = (size_t)e * cap;
    // a step may be taken in two calls (FASTACE_STEP_PERSONS, then FASTACE_STEP_FIRMS): the state in HBM between
    // them is the economy as it stands when the last person has acted and no firm has (economy.cpp:118-123)
    // (the compact encoding is only accepted for full steps; the specialised kernel is launched only when none of the
    // optional outputs is requested)
    const bool do_persons = (MODE & kModeCompact) ? true : !(p.flags & FASTACE_STEP_FIRMS);
    const bool do_firms = (MODE & kModeCompact) ? true : !(p.flags & (FASTACE_STEP_PERSONS | FASTACE_STEP_PERSONS_TRADE));
    const bool optional_out = !(MODE & kModeCompact);
    // kModeSmall: every book and the firm list fit one pass of the warp (F * (G + 1) + 2 <= 32): the strided loops over
    // rows / offers / firms run exactly once
    constexpr int kMorePasses = (MODE & kModeSmall) ? 0 : 1;
    if (do_persons && compact) {
        // the economy's person-side inputs are five contiguous slabs: ask L2 for them now, use them window by window
        if (lane == 0) prefetch_slab_l2(p.cz.p_job_idx + (size_t)e * P * S, (size_t)P * S);
        if (lane == 1) prefetch_slab_l2(p.cz.p_good_idx + (size_t)e * P * S, (size_t)P * S);
        if (lane == 2) prefetch_slab_l2(p.cz.p_job_take + (size_t)e * P, (size_t)P * 2);
        if (lane == 3) prefetch_slab_l2(p.cz.p_good_take + (size_t)e * P, (size_t)P * 2);
        if (lane == 4) prefetch_slab_l2(p.cz.perm_person + (size_t)e * P, (size_t)P * 2);
        if (lane == 5) prefetch_slab_l2(p.st.p_money + (size_t)e * P, (size_t)P * 8);
    }
    // everything above touches only this step's inputs; the books and agent state come from the previous step
    grid_dependency_wait();
    grid_launch_dependents();
    const int NM = p.st.m_count[e];
    const int NJ = p.st.j_count[e];
    const int NT = NJ + NM + 2;                                     // all rows, the two "no request" rows included
    const uint32_t aRecM = keep_u32(aRec + (uint32_t)(NJ + 1) * kRecBytes);   // record / matrix row of goods offer 0
    const uint32_t aMatM = keep_u32(aMat + (uint32_t)(NJ + 1) * kMatBytes);

    // ------------------------------ stage: books and firms ---------------------------------
    const IndexMap mapJ(NJ, modulo), mapM(NM, modulo);
    for (int R = lane, more = 1; R < NT && more; R += 32, more = kMorePasses) {
        const bool isJ = R <= NJ;
        const int n = isJ ? R : R - NJ - 1;
        const bool real = isJ ? n < NJ : n < NM;
        double value = 0.0;
        uint32_t left = 0, taken = 0, meta = 0;
        if (real && isJ) {
            value = p.st.j_wage[eF + n]; left = p.st.j_left[eF + n]; taken = p.st.j_taken[eF + n];
            meta = (uint32_t)(p.st.j_owner[eF + n] & 0xFF);
        } else if (real) {
            value = p.st.m_price[eCap + n]; left = p.st.m_left[eCap + n]; taken = p.st.m_taken[eCap + n];
            meta = (uint32_t)(p.st.m_owner[eCap + n] & 0xFF) | ((uint32_t)(p.st.m_good[eCap + n] & 0xFF) << 8);
        }
        const uint32_t rec = aRec + (uint32_t)R * kRecBytes;
        sts_f64<kRecValue>(rec, value);
        sts_u32<kRecLeft>(rec, left);
        sts_u32<kRecTaken>(rec, taken);
        sts_u32<kRecMeta>(rec, meta);
    }
    for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) {
        sts_f64(aFmoney + 8u * f, p.st.f_money[eF + f]);
        sts_u16(aPermf + 2u * f, do_firms ? (uint32_t)(compact ? (int)p.cz.perm_firm[eF + f] : p.ac.perm_firm[eF + f]) : (uint32_t)f);
        sts_u32(aFnh + 4u * f, 0u);
        sts_u32(aFok + 4u * f, 0u);
        sts_u8(aFcnt + f, 0u);
        sts_u8(aFfirst + f, 0u);
        sts_u8(aFrisk + f, 0u);
        sts_u8(aFjob + f, (uint32_t)kNone);
#pragma unroll
        for (int g = 0; g < G; g++) sts_f64(aFinv + 8u * (g * F + f), p.st.f_inv[((size_t)e * G + g) * F + f]);
        if (do_firms) sts_f64(aFlast + 8u * f, p.st.f_last_money[eF + f]);
        sts_f64(aFlabor + 8u * f, p.st.f_labor[eF + f]);   // used at the very end (laborHired += 0.5 per hire)
    }
    if (do_firms) {
        // the firms' own requests: byte i of firm f = goods offer number or kNone (empty book: no requests at all,
        // decisionNetHandler.cpp:398-403)
        for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) {
            const uint32_t none4 = (uint32_t)kNone * 0x01010101u;
            uint32_t w[kMaxStack / 4];
#pragma unroll
            for (int q = 0; q < kMaxStack / 4; q++) w[q] = none4;
            if (NM > 0 && compact) {
                // agent-major bytes: the aligned words that cover the firm's S request bytes, four slots at a time
                const uint8_t* lst = p.cz.f_good_idx + (eF + f) * (size_t)S;
                const uint32_t take = p.cz.f_good_take[eF + f] & ((1u << S) - 1u);
                const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(lst) & 3u);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(lst - sh);
                uint32_t raw[kMaxStack / 4 + 1];
#pragma unroll
                for (int q = 0; q <= kMaxStack / 4; q++) raw[q] = (4u * q < sh + (uint32_t)S) ? wp[q] : 0u;
#pragma unroll
                for (int q = 0; q < kMaxStack / 4; q++) {
                    if (4 * q < S) {
                        const uint32_t x = __funnelshift_r(raw[q], raw[q + 1], 8u * sh);
                        const uint32_t t4 = take >> (4 * q);
                        w[q] = modulo ? map_request_word<true>(x, t4, mapM, (uint32_t)kNone & 0xFFu) : map_request_word<false>(x, t4, mapM, (uint32_t)kNone & 0xFFu);
                    }
                }
            }
            sts_v4(aFatt + 16u * f, make_uint4(w[0], w[1], w[2], w[3]));
        }
        if (NM > 0 && !compact) {
            // int32 encoding ([E][S][F]): one (firm, slot) pair per lane and pass
            __syncwarp();
            for (int k = lane; k < F * S; k += 32) {
                const int i = k / F, f = k - i * F;
                const size_t a = (size_t)e * S * F + (size_t)k;
                if (p.ac.f_good_take[a] != 0) sts_u8(aFatt + 16u * f + (uint32_t)i, (uint32_t)mapM(p.ac.f_good_idx[a]));
            }
        }
    }
    __syncwarp();
    for (int n = lane, more = 1; n < NJ && more; n += 32, more = kMorePasses) sts_u8(aFjob + (lds_u32<kRecMeta>(aRec + (uint32_t)n * kRecBytes) & 0xFFu), (uint32_t)n);
    for (int n = lane, more = 1; n < NM && more; n += 32, more = kMorePasses) {
        // a firm's entries are contiguous in market order (it posts all its goods in one turn)
        const uint32_t rec = aRecM + (uint32_t)n * kRecBytes;
        const uint32_t owner = lds_u32<kRecMeta>(rec) & 0xFFu;
        const int prev = (n > 0) ? (int)(lds_u32<kRecMeta - kRecBytes>(rec) & 0xFFu) : -1;
        if ((int)owner != prev) sts_u8(aFfirst + owner, (uint32_t)n);
        if (!(lds_f64<kRecValue>(rec) >= 0.0)) sts_u8(aFrisk + owner, 1u);   // a sale would not raise the seller's money
    }
    __syncwarp();
    for (int n = lane, more = 1; n < NM && more; n += 32, more = kMorePasses) {
        const uint32_t rec = aRecM + (uint32_t)n * kRecBytes;
        const uint32_t owner = lds_u32<kRecMeta>(rec) & 0xFFu;
        const int next = (n + 1 < NM) ? (int)(lds_u32<kRecMeta + kRecBytes>(rec) & 0xFFu) : -1;
        if ((int)owner != next) sts_u8(aFcnt + owner, (uint32_t)(n + 1) - lds_u8(aFfirst + owner));
    }
    __syncwarp();

    // ------------------------------ persons: windows of 32 visiting ranks -------------------
    const size_t row0 = (size_t)e * S * P;
    // the visiting order is read one window ahead: the load of the next window's person number is in flight during this
    // window's rounds, and its money / request lines are asked for (L1 prefetch) before the window's firm-money fold
    int pid_next = (do_persons && lane < P) ? (compact ? (int)p.cz.perm_person[eP + lane] : p.ac.perm_person[eP + lane]) : 0;
    for (int base = 0; do_persons && base < P; base += 32) {
        // ---- (1) rows: death ordinals and initial rooms of the window
        bool lj = false, lm = false, rk = false;
        for (int R = lane, more = 1; R < NT && more; R += 32, more = kMorePasses) {
            const uint32_t rec = aRec + (uint32_t)R * kRecBytes;
            const uint32_t left = lds_u32<kRecLeft>(rec);
            const uint32_t meta = lds_u32<kRecMeta>(rec);
            const int owner = (int)(meta & 0xFFu);
            uint32_t d = left;
            if (R <= NJ) {
                lj |= left > 0;
                if (left > 0) {
                    // How many hires of this window can the firm certainly pay?  k is safe when
                    // m0 - w*k >= w*(1 + 1e-9): wages are subtracted one by one (rounding error << 1e-9 relative for
                    // k <= 64) and sales only add money unless a price is negative / NaN (frisk).  Largest safe k
                    // (<= 64, the most a window can hire), found by stepping down from the quotient.
                    const double w = lds_f64<kRecValue>(rec), m0 = lds_f64(aFmoney + 8u * owner);
                    int k = -1;
                    if (w >= 0.0 && lds_u8(aFrisk + owner) == 0u && m0 - w * (double)kMaxHiresPerWindow >= w * (1.0 + 1e-9)) {
                        k = kMaxHiresPerWindow;                            // the firm can pay whatever the window brings
                    } else if (w >= 0.0 && lds_u8(aFrisk + owner) == 0u) {
                        const double q = (m0 - w * (1.0 + 1e-9)) / w;
                        k = q >= (double)kMaxHiresPerWindow ? kMaxHiresPerWindow : (q >= 0.0 ? (int)q : -1);   // NaN -> -1
                        while (k >= 0 && !(m0 - w * (double)k >= w * (1.0 + 1e-9))) k--;
                    }
                    sts_u16<kRecSafe>(rec, (uint32_t)(k + 1));   // 0: not even the first applicant is certain
                    rk |= k < (int)min(left, (uint32_t)kMaxHiresPerWindow);
                }
            } else {
                const int good = (int)((meta >> 8) & 0xFFu);
                d = min(left, unit_sales_possible(lds_f64(aFinv + 8u * (good * F + owner))));
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (g != good && lds_f64(aFinv + 8u * (g * F + owner)) < 0.0) d = 0;   // agent.cpp:140 on a zero quantity
                lm |= left > 0;
            }
            const uint32_t di = min(d, 0x7FFFFFFFu);
            sts_u32<kRecD>(rec, di);
            sts_u16<kRecScanned>(rec, 0u);                               // kRecScanned and kRecEval
            const uint32_t rm = min(di, (uint32_t)kRoomMax) * 0x01010101u;
            const uint32_t mat = aMat + (uint32_t)R * kMatBytes;
            sts_v4<kMatCnt>(mat, make_uint4(0u, 0u, 0u, 0u));
            sts_v4<kMatCnt + 16>(mat, make_uint4(0u, 0u, 0u, 0u));
            sts_v4<kMatRoom>(mat, make_uint4(rm, rm, rm, rm));
            sts_v4<kMatRoom + 16>(mat, make_uint4(rm, rm, rm, rm));
        }
        const bool liveJ = __any_sync(0xffffffffu, lj), liveM = __any_sync(0xffffffffu, lm);
        const bool risk_possible = __any_sync(0xffffffffu, rk);
        if (!liveJ && !liveM) {
            // both books are sold out: no person from here on can trade
            if (lane == 0) FASTACE_STAT(kStatDeadExits, 1);
            for (int r = base + lane; r < P; r += 32) {
                const int pid = compact ? (int)p.cz.perm_person[eP + r] : p.ac.perm_person[eP + r];
                mp.scr_pnh[eP + pid] = 0;
#pragma unroll
                for (int g = 0; g < G; g++) mp.scr_pnb[((size_t)e * G + g) * P + pid] = 0;
                if (person_flags) write_person_ok(p, e, pid, 0u);
            }
            break;
        }
        // ---- (2) the lane's person: money, and its request lists as offer numbers in shared memory
        //      (slot i of the job list at aReqL + i, of the goods list at aReqL + 16 + i; "no request" = offer NJ / NM)
        const int r = base + lane;
        const bool active = r < P;
        const int pid = pid_next;
        if (r + 32 < P) pid_next = compact ? (int)p.cz.perm_person[eP + r + 32] : p.ac.perm_person[eP + r + 32];
        const double money0 = active ? p.st.p_money[eP + pid] : 0.0;
        // keep: the slots that are worth evaluating — requested, on an existing offer that still has lots (a request on a
        // sold-out offer fails without any effect: firm.cpp:64, agent.cpp:124); jobs in bits 0..15, goods in 16..31
        uint32_t keep = 0;
        {
    #pragma unroll
            for (int ph = 0; ph < 2; ph++) {
                const bool live = ph == 0 ? liveJ : liveM;
                if (!live || !active) continue;                        // a sold-out book is not evaluated at all
                const uint32_t dst = aReqL + (ph == 0 ? 0u : (uint32_t)kReqGoods);
                const uint32_t dummy = (uint32_t)(ph == 0 ? NJ : NM);
                const uint32_t recs = ph == 0 ? aRec : aRecM;
                const IndexMap& map = ph == 0 ? mapJ : mapM;
                uint32_t kp = 0;
                if (compact) {
                    // agent-major bytes: the aligned words that cover [pid*S, pid*S + S), shifted into place
                    const uint8_t* lst = (ph == 0 ? p.cz.p_job_idx : p.cz.p_good_idx) + (eP + pid) * (size_t)S;
                    const uint32_t take = (ph == 0 ? p.cz.p_job_take : p.cz.p_good_take)[eP + pid] & ((1u << S) - 1u);
                    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(lst) & 3u);
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(lst - sh);
                    uint32_t raw[kMaxStack / 4 + 1];
#pragma unroll
                    for (int q = 0; q <= kMaxStack / 4; q++) raw[q] = (4u * q < sh + (uint32_t)S) ? wp[q] : 0u;
#pragma unroll
                    for (int q = 0; q < kMaxStack / 4; q++) {
                        if (4 * q < S) {
                            const uint32_t x = __funnelshift_r(raw[q], raw[q + 1], 8u * sh);
                            const uint32_t t4 = take >> (4 * q);
                            const uint32_t w = modulo ? map_request_word<true>(x, t4, map, dummy) : map_request_word<false>(x, t4, map, dummy);
                            sts_u32(dst + 4u * q, w);
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const uint32_t n = __byte_perm(w, 0u, 0x4440u + k);
                                if (lds_u32<kRecLeft>(recs + n * kRecBytes) > 0u) kp |= 1u << (4 * q + k);   // the "no request" row has none
                            }
                        }
                    }
                } else {
                    const size_t a0 = row0 + (size_t)pid;
                    const int32_t* idx = (ph == 0 ? p.ac.p_job_idx : p.ac.p_good_idx) + a0;
                    const uint8_t* tk = (ph == 0 ? p.ac.p_job_take : p.ac.p_good_take) + a0;
                    for (int i = 0; i < S; i++) {
                        const uint32_t raw = (uint32_t)idx[(size_t)i * P];
                        uint32_t n = modulo ? map.mod24(map.mod24(raw >> 16) * map.k16 + (raw & 0xFFFFu)) : (raw < map.count ? raw : dummy);
                        if (tk[(size_t)i * P] == 0) n = dummy;
                        sts_u8(dst + (uint32_t)i, n);
                        if (lds_u32<kRecLeft>(recs + n * kRecBytes) > 0u) kp |= 1u << i;
                    }
                }
                keep |= ph == 0 ? kp : kp << 16;
            }
        }
        __syncwarp();

        // ---- (3) fixed-point rounds
        double money = money0;
        int nh = 0;
        uint32_t okm = 0;      // successes: job slots in bits 0..15, goods slots in bits 16..31
        uint32_t last_goods = 0xFFFFFFFFu;   // the goods successes of the previous round
        const uint32_t aCellJ = keep_u32(aMat + (uint32_t)lane), aCellM = keep_u32(aMatM + (uint32_t)lane);
        if (lane == 0) FASTACE_STAT(kStatWindows, 1);
        for (int round = 0;; round++) {
            if (lane == 0) { FASTACE_STAT(kStatRounds, 1); FASTACE_STAT(kStatRoundsW0 + min(base >> 5, 3), 1); }
            money = money0; nh = 0; okm = 0;
            // ---- evaluate: job requests (utilMaxer.cpp:76-85; eligible = person.cpp:39: 0.5 * nh + 0.5 <= 1).  Every
            //      lane walks ITS OWN kept slots in request order; the loop ends when no lane can apply any more
            if (liveJ) {
                uint32_t rem = keep & 0xFFFFu;
                for (;;) {
                    const bool go = rem != 0u && nh < 2;
                    if (!__any_sync(0xffffffffu, go)) break;
                    if (go) {
                        const uint32_t low = rem & (0u - rem);               // the lane's next kept slot, as a mask
                        rem ^= low;
                        const int i = 31 - __clz((int)low);
                        const uint32_t cell = aCellJ + lds_u8(aReqL + (uint32_t)i) * kMatBytes;
                        const uint32_t c = lds_u8<kMatCnt>(cell);
                        const uint32_t rm = lds_u8<kMatRoom>(cell);
                        sts_u8<kMatCnt>(cell, c + 1u);
                        if (c < rm) { nh++; okm |= low; }
                    }
                }
                // the wages of the (at most two) jobs, in request order (person.cpp:49)
                uint32_t mj = okm;
                if (mj) { money += lds_f64<kRecValue>(aRec + lds_u8(aReqL + (uint32_t)(__ffs((int)mj) - 1)) * kRecBytes); mj &= mj - 1; }
                if (mj) money += lds_f64<kRecValue>(aRec + lds_u8(aReqL + (uint32_t)(__ffs((int)mj) - 1)) * kRecBytes);
            }
            // ---- evaluate: goods requests (utilMaxer.cpp:64-73; eligible = agent.cpp:102)
            if (liveM) {
                uint32_t rem = keep >> 16;
                for (;;) {
                    const bool go = rem != 0u;
                    if (!__any_sync(0xffffffffu, go)) break;
                    if (go) {
                        const uint32_t low = rem & (0u - rem);
                        rem ^= low;
                        const int i = 31 - __clz((int)low);
                        const uint32_t n = lds_u8<kReqGoods>(aReqL + (uint32_t)i);
                        const double price = lds_f64<kRecValue>(aRecM + n * kRecBytes);
                        if (money >= price) {
                            const uint32_t cell = aCellM + n * kMatBytes;
                            const uint32_t c = lds_u8<kMatCnt>(cell);
                            const uint32_t rm = lds_u8<kMatRoom>(cell);
                            sts_u8<kMatCnt>(cell, c + 1u);
                            if (c < rm) { money -= price; okm |= low << 16; }   // agent.cpp:108
                        }
                    }
                }
            }
            __syncwarp();

            // ---- rows: demand of the window; over-subscribed rows whose counts moved get their rooms re-scanned
            bool changed = false;
            // did any lane's purchases move since the last round?  (a death ordinal derived from the firm's money stands
            // while the applications to the row and everybody's purchases do)
            const bool goods_moved = risk_possible && __any_sync(0xffffffffu, (okm >> 16) != last_goods);
            last_goods = okm >> 16;
            for (int cb = 0, more = 1; cb < NT && more; cb += 32, more = kMorePasses) {
                const int R = cb + lane;
                bool needs = false;
                if (R < NT) {
                    const uint32_t rec = aRec + (uint32_t)R * kRecBytes, mat = aMat + (uint32_t)R * kMatBytes;
                    const uint4 a = lds_v4<kMatCnt>(mat), b = lds_v4<kMatCnt + 16>(mat);
                    const uint32_t s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;          // per byte <= 8 * 16
                    const uint32_t t = (s & 0x00FF00FFu) + ((s >> 8) & 0x00FF00FFu);
                    const uint32_t tot = (t + (t >> 16)) & 0xFFFFu;
                    sts_u32<kRecTot>(rec, tot);
                    const uint32_t d = lds_u32<kRecD>(rec);
                    if (d > 0 && tot > 0) {
                        if (lds_u8<kRecScanned>(rec) != 0u) {
                            // scanned before: its rooms stand as long as the counts they were computed for do
                            const uint4 c = lds_v4<kMatPrev>(mat), dd = lds_v4<kMatPrev + 16>(mat);
                            needs = ((a.x ^ c.x) | (a.y ^ c.y) | (a.z ^ c.z) | (a.w ^ c.w) |
                                     (b.x ^ dd.x) | (b.y ^ dd.y) | (b.z ^ dd.z) | (b.w ^ dd.w)) != 0u;
                        } else {
                            needs = tot > d;   // initial rooms (= D in every lane) cannot bind while the demand fits
                        }
                    }
                }
                // ---- job offers whose firm may run out of money (firm.cpp:80-84): the death ordinal of such a row
                //      follows the firm's exact running money in visiting order; a row whose ordinal moves is re-scanned
                if (risk_possible && cb < NJ) {
                    bool risky = false;
                    uint32_t left = 0;
                    int f = 0, tot = 0;
                    double w = 0.0, m0 = 0.0;
                    const uint32_t rec = aRec + (uint32_t)(R < NJ ? R : 0) * kRecBytes;
                    if (R < NJ) {
                        left = lds_u32<kRecLeft>(rec);
                        tot = (int)lds_u32<kRecTot>(rec);
                        // the applicants that could be hired exceed what the firm can certainly pay
                        risky = tot > 0 && left > 0 && min(min((int)min(left, 0x7FFFFFFFu), tot), kMaxHiresPerWindow) >= (int)lds_u16<kRecSafe>(rec);
                        if (risky && !goods_moved && lds_u16<kRecScanned>(rec) == 0x0101u) {
                            // evaluated before in this window and re-scanned since: if `prev` (the applications the
                            // rooms were computed for) is still what the lanes apply, the ordinal in the record stands
                            const uint32_t mat = aMat + (uint32_t)R * kMatBytes;
                            const uint4 a = lds_v4<kMatCnt>(mat), b = lds_v4<kMatCnt + 16>(mat);
                            const uint4 c = lds_v4<kMatPrev>(mat), dd = lds_v4<kMatPrev + 16>(mat);
                            risky = ((a.x ^ c.x) | (a.y ^ c.y) | (a.z ^ c.z) | (a.w ^ c.w) | (b.x ^ dd.x) | (b.y ^ dd.y) | (b.z ^ dd.z) | (b.w ^ dd.w)) != 0u;
                            if (!risky) FASTACE_STAT(kStatRiskyKept, 1);
                        }
                        if (risky) {
                            f = (int)(lds_u32<kRecMeta>(rec) & 0xFFu);
                            w = lds_f64<kRecValue>(rec);
                            m0 = lds_f64(aFmoney + 8u * f);
                        }
                    }
                    if (__any_sync(0xffffffffu, risky)) {
                        // does the firm sell anything in this window?  (its goods rows' successes = min(demand, D);
                        // the demand is summed here: those rows may belong to a later chunk)
                        bool fsales = false;
                        if (risky) {
                            const int first = (int)lds_u8(aFfirst + f), cnt = (int)lds_u8(aFcnt + f);
                            for (int n = first; n < first + cnt; n++) {
                                const uint32_t gm = aMatM + (uint32_t)n * kMatBytes;
                                const uint4 a = lds_v4<kMatCnt>(gm), b = lds_v4<kMatCnt + 16>(gm);
                                fsales |= (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0u && lds_u32<kRecD>(aRecM + (uint32_t)n * kRecBytes) > 0u;
                            }
                        }
                        int d = (int)min(left, 0x7FFFFFFFu);
                        if (risky && !fsales) {
                            // hires are the firm's only events: the order of the applicants does not matter
                            FASTACE_STAT(kStatRiskyWalks, 1);
                            double m = m0;
                            const int most = min(d, tot);
                            for (int h = 0; h < most; h++) {
                                if (m < w) { d = h; break; }                            // firm.cpp:80-84
                                m -= w;                                                 // firm.cpp:108
                            }
                        }
                        // rows whose firm also sells in this window, one at a time with all lanes: lane l holds person l's
                        // applications to the row and its purchases from the firm
                        for (unsigned sm = __ballot_sync(0xffffffffu, risky && fsales); sm != 0; sm &= sm - 1) {
                            const int src = __ffs((int)sm) - 1;
                            const int fb = __shfl_sync(0xffffffffu, f, src);
                            const double wb = __shfl_sync(0xffffffffu, w, src), mb = __shfl_sync(0xffffffffu, m0, src);
                            const int leftb = __shfl_sync(0xffffffffu, d, src);
                            const uint32_t matb = aMat + (uint32_t)(cb + src) * kMatBytes;
                            const int c = (int)lds_u8<kMatCnt>(matb + (uint32_t)lane);
                            double s = 0.0;
                            for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                                const uint32_t g = aRecM + lds_u8<kReqGoods>(aReqL + (uint32_t)(__ffs((int)m) - 1)) * kRecBytes;
                                if ((int)(lds_u32<kRecMeta>(g) & 0xFFu) == fb) s += lds_f64<kRecValue>(g);
                            }
                            // prefixes over the visiting order: applications and sales before this lane's events
                            int ac = c;
                            double sx = s;
#pragma unroll
                            for (int dd = 1; dd < 32; dd <<= 1) {
                                const int uc = __shfl_up_sync(0xffffffffu, ac, dd);
                                const double us = __shfl_up_sync(0xffffffffu, sx, dd);
                                if (lane >= dd) { ac += uc; sx += us; }
                            }
                            const int atot = __shfl_sync(0xffffffffu, ac, 31);
                            const double stot = __shfl_sync(0xffffffffu, sx, 31);
                            ac -= c; sx -= s;
                            // Application number a (0-based, in visiting order) meets m0 - a*w + (sales before it) as long as
                            // every earlier one was a hire (jobs precede purchases, person.cpp:26-29).  These sums differ from
                            // the reference's event-by-event money by a few hundred roundings at most: a comparison that
                            // clears `tol` is the reference's; anything closer is walked exactly below.
                            const double tol = 1e-9 * (fabs(mb) + stot + (double)atot * wb);
                            int failk = -1;
                            bool unclear = !(wb >= 0.0) || lds_u8(aFrisk + (uint32_t)fb) != 0u;
                            for (int k = 0; k < c && failk < 0; k++) {
                                const int a = ac + k;
                                if (a >= leftb) break;                                  // firm.cpp:64: exhausted
                                const double gap = (mb - (double)a * wb + sx) - wb;
                                if (!(fabs(gap) > tol)) unclear = true;
                                if (gap < 0.0) failk = k;                               // firm.cpp:80-84
                            }
                            int db = leftb;
                            const unsigned failed = __ballot_sync(0xffffffffu, failk >= 0);
                            if (failed) db = __shfl_sync(0xffffffffu, ac + failk, __ffs((int)failed) - 1);
                            if (__any_sync(0xffffffffu, unclear)) {
                                // every lane publishes the goods it currently buys, in request order; the row's lane walks
                                uint32_t ne = 0;
                                for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                                    sts_u8(aEv + (uint32_t)lane * kEvPerLane + 1u + ne, lds_u8<kReqGoods>(aReqL + (uint32_t)(__ffs((int)m) - 1)));
                                    ne++;
                                }
                                sts_u8(aEv + (uint32_t)lane * kEvPerLane, ne);
                                __syncwarp();
                                if (lane == src) {
                                    FASTACE_STAT(kStatRiskyWalks, 1);
                                    double m = mb;
                                    int h = 0;
                                    bool done = false;
                                    db = leftb;
                                    for (int l = 0; l < 32 && !done; l++) {
                                        const int cl = (int)lds_u8<kMatCnt>(matb + (uint32_t)l);
                                        for (int k = 0; k < cl; k++) {
                                            if (h >= db) { done = true; break; }               // firm.cpp:64
                                            if (m < wb) { db = h; done = true; break; }        // firm.cpp:80-84
                                            m -= wb;                                           // firm.cpp:108
                                            h++;
                                        }
                                        const int nb = (int)lds_u8(aEv + (uint32_t)l * kEvPerLane);
                                        for (int k = 0; k < nb && !done; k++) {
                                            const uint32_t g = aRecM + lds_u8(aEv + (uint32_t)l * kEvPerLane + 1u + (uint32_t)k) * kRecBytes;
                                            if ((int)(lds_u32<kRecMeta>(g) & 0xFFu) == fb) m += lds_f64<kRecValue>(g);   // agent.cpp:158
                                        }
                                    }
                                }
                                __syncwarp();
                                db = __shfl_sync(0xffffffffu, db, src);
                            } else if (lane == src) {
                                FASTACE_STAT(kStatRiskyCoop, 1);
                            }
                            // every lane holds the applications before it (ac): the row's rooms follow the ordinal right
                            // here, and `prev` records what they were computed for — no separate re-scan of this row
                            {
                                const int rmn = max(0, min(kRoomMax, db - ac));
                                const int old = (int)lds_u8<kMatRoom>(matb + (uint32_t)lane);
                                changed |= min(c, rmn) != min(c, old);
                                sts_u8<kMatRoom>(matb + (uint32_t)lane, (uint32_t)rmn);
                                sts_u8<kMatPrev>(matb + (uint32_t)lane, (uint32_t)c);
                            }
                            if (lane == src) {
                                d = db;
                                needs = false;
                                sts_u32<kRecD>(rec, (uint32_t)db);
                                sts_u8<kRecScanned>(rec, 1u);
                            }
                        }
                        if (risky) sts_u8<kRecEval>(rec, 1u);
                        if (risky && d != (int)lds_u32<kRecD>(rec)) {
                            // the offer dies earlier / later than assumed: its rooms follow the new ordinal (the re-scan
                            // below tells whether that changes anybody's outcome)
                            needs = true;
                            sts_u32<kRecD>(rec, (uint32_t)d);
                        }
                        __syncwarp();
                    }
                }
                unsigned slow = __ballot_sync(0xffffffffu, needs);
                while (slow) {
                    // two rows per shuffle scan (16-bit halves: a prefix is at most 32 * 16)
                    const int R0 = cb + __ffs((int)slow) - 1;
                    slow &= slow - 1;
                    const int R1 = slow ? cb + __ffs((int)slow) - 1 : R0;
                    slow &= slow - 1;
                    const uint32_t cell0 = aMat + (uint32_t)R0 * kMatBytes + (uint32_t)lane, cell1 = aMat + (uint32_t)R1 * kMatBytes + (uint32_t)lane;
                    const uint32_t rec0 = aRec + (uint32_t)R0 * kRecBytes, rec1 = aRec + (uint32_t)R1 * kRecBytes;
                    if (lane == 0) { FASTACE_STAT(kStatRescans, R1 != R0 ? 2 : 1); if (base == 0) FASTACE_STAT(kStatRescansW0, R1 != R0 ? 2 : 1); }
                    const uint32_t c0 = lds_u8<kMatCnt>(cell0), c1 = lds_u8<kMatCnt>(cell1);
                    const uint32_t both = c0 | (c1 << 16);
                    const uint32_t excl = warp_inclusive_scan(both, lane) - both;
                    const int rm0 = max(0, min(kRoomMax, (int)lds_u32<kRecD>(rec0) - (int)(excl & 0xFFFFu)));
                    const int rm1 = max(0, min(kRoomMax, (int)lds_u32<kRecD>(rec1) - (int)(excl >> 16)));
                    const int old0 = (int)lds_u8<kMatRoom>(cell0), old1 = (int)lds_u8<kMatRoom>(cell1);
                    changed |= (min((int)c0, rm0) != min((int)c0, old0)) | (min((int)c1, rm1) != min((int)c1, old1));
                    sts_u8<kMatRoom>(cell0, (uint32_t)rm0);
                    sts_u8<kMatPrev>(cell0, c0);
                    sts_u8<kMatRoom>(cell1, (uint32_t)rm1);
                    sts_u8<kMatPrev>(cell1, c1);
                    sts_u8<kRecScanned>(rec0, 1u);
                    sts_u8<kRecScanned>(rec1, 1u);
                }
            }
            __syncwarp();

            // ---- round end: converged? reset the counts
            const bool again = __any_sync(0xffffffffu, changed);
            if (!again) break;
            if (round >= kRoundCap) {
                if (lane == 0) mp.dev_err[kDevErrRounds] = 1u;
                break;
            }
            for (int R = lane, more = 1; R < NT && more; R += 32, more = kMorePasses) {
                const uint32_t mat = aMat + (uint32_t)R * kMatBytes;
                sts_v4<kMatCnt>(mat, make_uint4(0u, 0u, 0u, 0u));
                sts_v4<kMatCnt + 16>(mat, make_uint4(0u, 0u, 0u, 0u));
            }
            __syncwarp();
        }

        // ---- (4) commit the window
        uint32_t hire1 = (uint32_t)kNone, hire2 = (uint32_t)kNone;    // the job offers this person was hired on
        if (active) {
            p.st.p_money[eP + pid] = money;
            mp.scr_pnh[eP + pid] = (uint8_t)nh;
            uint32_t mj = okm & 0xFFFFu;
            if (mj) { hire1 = lds_u8(aReqL + (uint32_t)(__ffs((int)mj) - 1)); mj &= mj - 1; }
            if (mj) hire2 = lds_u8(aReqL + (uint32_t)(__ffs((int)mj) - 1));
        }
        uint32_t nbuy = 0;
        const unsigned buyers = __ballot_sync(0xffffffffu, (okm >> 16) != 0u);
        if (buyers != 0u) {
            // aFlive (free until the firm phase) collects, per firm, the lanes that bought from it in this window
            for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) sts_u32(aFlive + 4u * f, 0u);
            __syncwarp();
        }
        {
            // purchases: per good for update_kernel, and as an ordered list for the firms' money
            uint32_t nb[G];
#pragma unroll
            for (int g = 0; g < G; g++) nb[g] = 0;
            for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {
                const uint32_t n = lds_u8<kReqGoods>(aReqL + (uint32_t)(__ffs((int)m) - 1));
                const uint32_t meta = lds_u32<kRecMeta>(aRecM + n * kRecBytes);
                const uint32_t good = (meta >> 8) & 0xFFu;
#pragma unroll
                for (int g = 0; g < G; g++) nb[g] += (good == (uint32_t)g);
                sts_u8(aEv + (uint32_t)lane * kEvPerLane + 1u + nbuy, n);
                reds_or_u32(aFlive + 4u * (meta & 0xFFu), 1u << lane);
                nbuy++;
            }
            sts_u8(aEv + (uint32_t)lane * kEvPerLane, nbuy);
            if (active) {
#pragma unroll
                for (int g = 0; g < G; g++) mp.scr_pnb[((size_t)e * G + g) * P + pid] = (uint8_t)nb[g];
                if (person_flags) write_person_ok(p, e, pid, okm);
            }
        }
        for (int R = lane, more = 1; R < NT && more; R += 32, more = kMorePasses) {
            const uint32_t rec = aRec + (uint32_t)R * kRecBytes;
            const uint32_t tot = lds_u32<kRecTot>(rec);
            const uint32_t n = min(tot, lds_u32<kRecD>(rec));
            const uint32_t left = lds_u32<kRecLeft>(rec);
            if (R <= NJ) {
                sts_u32<kRecLeft>(rec, (tot > n) ? 0u : left - n);                 // exhausted or killed (firm.cpp:83)
            } else {
                uint32_t nl = left - n;
                if (tot > n && nl > 0) nl = 0;                                      // killed (agent.cpp:143)
                sts_u32<kRecLeft>(rec, nl);
                if (n > 0) {
                    const uint32_t meta = lds_u32<kRecMeta>(rec);
                    const uint32_t a = aFinv + 8u * (((meta >> 8) & 0xFFu) * F + (meta & 0xFFu));
                    sts_f64(a, lds_f64(a) - (double)n);                             // n exact unit subtractions (agent.cpp:156)
                }
            }
            sts_u32<kRecTaken>(rec, lds_u32<kRecTaken>(rec) + n);
            sts_u32<kRecTot>(rec, n);
        }
        if (r + 32 < P) {
            prefetch_l1(p.st.p_money + eP + pid_next);
            if (compact) {
                prefetch_l1(p.cz.p_job_idx + (eP + pid_next) * (size_t)S);
                prefetch_l1(p.cz.p_good_idx + (eP + pid_next) * (size_t)S);
            }
        }
        __syncwarp();
        // ---- firm money of the window, event by event in the visiting order
        for (int fb = 0, more = 1; fb < F && more; fb += 32, more = kMorePasses) {
            const int f = fb + lane;
            const uint32_t j = f < F ? lds_u8(aFjob + f) : (uint32_t)kNone;
            const bool has = j != (uint32_t)kNone;
            const uint32_t rec = aRec + (has ? j : 0u) * kRecBytes;
            const double w = lds_f64<kRecValue>(rec);
            double m = f < F ? lds_f64(aFmoney + 8u * f) : 0.0;
            uint32_t hires = has ? lds_u32<kRecTot>(rec) : 0u;
            if (buyers == 0u) {
                // the window's only events at a firm are hires: firm.cpp:108, one subtraction per hire
                for (uint32_t k = 0; k < hires; k++) m -= w;
            } else {
                // sales interleave with hires: which lanes were hired on this firm's offer (once / twice) ...
                if (lane == 0) FASTACE_STAT(kStatSalesWindows, 1);
                uint32_t b1 = 0, b2 = 0;
                for (int R = 0; R < NJ; R++) {
                    const unsigned x = __ballot_sync(0xffffffffu, hire1 == (uint32_t)R || hire2 == (uint32_t)R);
                    const unsigned y = __ballot_sync(0xffffffffu, hire1 == (uint32_t)R && hire2 == (uint32_t)R);
                    if (j == (uint32_t)R) { b1 = x; b2 = y; }
                }
                // ... then, for the lanes that bought from THIS firm, buyer by buyer: the hires up to and including that
                // lane (jobs precede purchases, person.cpp:26-29), then its purchases from this firm in request order
                uint32_t before = 0;
                for (unsigned bm = f < F ? lds_u32(aFlive + 4u * f) : 0u; bm != 0; bm &= bm - 1) {
                    const int l = __ffs((int)bm) - 1;
                    const uint32_t upto = 0xFFFFFFFFu >> (31 - l);
                    const uint32_t h = (uint32_t)(__popc(b1 & upto & ~before) + __popc(b2 & upto & ~before));
                    before = upto;
                    for (uint32_t k = 0; k < h; k++) m -= w;                                   // firm.cpp:108
                    const uint32_t ev = aEv + (uint32_t)l * kEvPerLane;
                    const uint32_t ne = lds_u8(ev);
                    for (uint32_t k = 0; k < ne; k++) {
                        const uint32_t g = aRecM + lds_u8(ev + 1u + k) * kRecBytes;
                        if ((int)(lds_u32<kRecMeta>(g) & 0xFFu) == f) m += lds_f64<kRecValue>(g);   // agent.cpp:158
                    }
                }
                const uint32_t h = (uint32_t)(__popc(b1 & ~before) + __popc(b2 & ~before));
                for (uint32_t k = 0; k < h; k++) m -= w;
            }
            if (f < F) {
                sts_f64(aFmoney + 8u * f, m);
                sts_u32(aFnh + 4u * f, lds_u32(aFnh + 4u * f) + hires);
            }
        }
        __syncwarp();
    }

    // ---- person phase: counters to HBM
    // job counters are final after the person phase
    if (do_persons && optional_out && p.out.old_j_left) for (int n = lane, more = 1; n < NJ && more; n += 32, more = kMorePasses) p.out.old_j_left[eF + n] = lds_u32<kRecLeft>(aRec + (uint32_t)n * kRecBytes);
    if (do_persons && optional_out && p.out.old_j_taken) for (int n = lane, more = 1; n < NJ && more; n += 32, more = kMorePasses) p.out.old_j_taken[eF + n] = lds_u32<kRecTaken>(aRec + (uint32_t)n * kRecBytes);
    if (!do_firms) {
        // persons-only call: the books' counters go back to HBM for the firms call (a full step never needs them
        // there: update_kernel replaces the books)
        for (int n = lane, more = 1; n < NM && more; n += 32, more = kMorePasses) {
            p.st.m_left[eCap + n] = lds_u32<kRecLeft>(aRecM + (uint32_t)n * kRecBytes);
            p.st.m_taken[eCap + n] = lds_u32<kRecTaken>(aRecM + (uint32_t)n * kRecBytes);
        }
        for (int n = lane, more = 1; n < NJ && more; n += 32, more = kMorePasses) {
            p.st.j_left[eF + n] = lds_u32<kRecLeft>(aRec + (uint32_t)n * kRecBytes);
            p.st.j_taken[eF + n] = lds_u32<kRecTaken>(aRec + (uint32_t)n * kRecBytes);
        }
    }

    // ------------------------------ firms ---------------------------------------------------
    if (do_firms) {
        // requests on entries that are sold out can only fail (amountLeft never grows within a step)
        bool lv = false;
        for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) {
            const uint4 slots = lds_v4(aFatt + 16u * f);
            const uint32_t w[4] = {slots.x, slots.y, slots.z, slots.w};
            uint32_t live = 0;
#pragma unroll
            for (int i = 0; i < kMaxStack; i++) {
                const uint32_t n = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;      // kNone beyond S
                if (n != (uint32_t)kNone && lds_u32<kRecLeft>(aRecM + n * kRecBytes) > 0u) live |= 1u << i;
            }
            sts_u32(aFlive + 4u * f, live);
            lv |= live != 0;
        }
        const bool anylive = __any_sync(0xffffffffu, lv);
        __syncwarp();
        // one firm's turn up to sell_goods (firm.cpp:23-37); `buys`: walk its live requests
        auto firm_turn = [&](int f, bool buys) {
            const int first = (int)lds_u8(aFfirst + f), cnt = (int)lds_u8(aFcnt + f);
            double money = lds_f64(aFmoney + 8u * f);
            // Agent::check_my_offers (base/agent.cpp:54-97): running inventoryLeft over own entries
            {
                double invLeft[G];
#pragma unroll
                for (int g = 0; g < G; g++) invLeft[g] = lds_f64(aFinv + 8u * (g * F + f));
                for (int n = first; n < first + cnt; n++) {
                    const uint32_t rec = aRecM + (uint32_t)n * kRecBytes;
                    const int good = (int)((lds_u32<kRecMeta>(rec) >> 8) & 0xFFu);
                    uint32_t left = lds_u32<kRecLeft>(rec);
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
                    sts_u32<kRecLeft>(rec, left);
#pragma unroll
                    for (int g = 0; g < G; g++) if (g == good) invLeft[g] -= delta;  // agent.cpp:83
                }
            }
            // first decision: profit of the previous step (neuralFirmDecisionMaker.cpp:65-74)
            {
                const double last = lds_f64(aFlast + 8u * f);
                sts_f64(aFlast + 8u * f, (p.time_before > 0) ? (money - last) : 0.0);   // profit, written out below
                p.st.f_last_money[eF + f] = money;
            }
            // ProfitMaxer::buy_goods (firms/profitMaxer.cpp:102-111); the buyer's money stays in a register
            uint32_t ok = 0;
            if (buys) {
                for (uint32_t lm = lds_u32(aFlive + 4u * f); lm != 0; lm &= lm - 1) {
                    const int i = __ffs((int)lm) - 1;
                    const uint32_t rec = aRecM + lds_u8(aFatt + 16u * f + (uint32_t)i) * kRecBytes;
                    const double price = lds_f64<kRecValue>(rec);
                    if (money >= price) {                                  // agent.cpp:102
                        const uint32_t left = lds_u32<kRecLeft>(rec);
                        if (left > 0) {                                    // agent.cpp:124
                            const uint32_t meta = lds_u32<kRecMeta>(rec);
                            const int s = (int)(meta & 0xFFu), good = (int)((meta >> 8) & 0xFFu);
                            bool short_ = false;                           // agent.cpp:140
#pragma unroll
                            for (int g = 0; g < G; g++) {
                                const double q = (g == good) ? kAmountPerOffer : 0.0;
                                if (lds_f64(aFinv + 8u * (g * F + s)) < q) short_ = true;
                            }
                            if (short_) {
                                sts_u32<kRecLeft>(rec, 0u);                // agent.cpp:143
                            } else {
                                // seller first (agent.cpp:155-160), then buyer (agent.cpp:108-109)
                                if (s == f) money += price; else sts_f64(aFmoney + 8u * s, lds_f64(aFmoney + 8u * s) + price);
                                sts_f64(aFinv + 8u * (good * F + s), lds_f64(aFinv + 8u * (good * F + s)) - kAmountPerOffer);
                                sts_u32<kRecLeft>(rec, left - 1);
                                sts_u32<kRecTaken>(rec, lds_u32<kRecTaken>(rec) + 1u);
                                money -= price;
                                sts_f64(aFinv + 8u * (good * F + f), lds_f64(aFinv + 8u * (good * F + f)) + kAmountPerOffer);
                                ok |= 1u << i;
                            }
                        }
                    }
                }
            }
            sts_f64(aFmoney + 8u * f, money);
            sts_u32(aFok + 4u * f, ok);
            // ProfitMaxer::sell_goods withdraws last step's offers (firms/profitMaxer.cpp:79-81);
            // nothing between buy_goods and that point touches another agent.
            for (int n = first; n < first + cnt; n++) {
                const uint32_t rec = aRecM + (uint32_t)n * kRecBytes;
                sts_u32<kRecD>(rec, lds_u32<kRecLeft>(rec));     // final counter of the withdrawn entry
                sts_u32<kRecLeft>(rec, 0u);
            }
        };
        if (!anylive) {
            // no purchase can happen: the firms' turns do not interact
            for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) firm_turn(f, false);
        } else if (lane == 0) {
            FASTACE_STAT(kStatFirmSerial, 1);
            for (int r = 0; r < F; r++) firm_turn((int)lds_u16(aPermf + 2u * r), true);   // visiting order (economy.cpp:121-123)
        }
        __syncwarp();
    }
    // ------------------------------ firms: results to HBM -----------------------------------
    // the queue ticket is taken now: its round trip overlaps the write-back (only the flag store must follow the fence)
    uint32_t ticket = 0;
    if (mp.done_list && lane == 0) ticket = ticket_add(mp.done_count);
    if (do_firms && optional_out && p.out.old_m_left) for (int n = lane, more = 1; n < NM && more; n += 32, more = kMorePasses) p.out.old_m_left[eCap + n] = lds_u32<kRecD>(aRecM + (uint32_t)n * kRecBytes);
    if (do_firms && optional_out && p.out.old_m_taken) for (int n = lane, more = 1; n < NM && more; n += 32, more = kMorePasses) p.out.old_m_taken[eCap + n] = lds_u32<kRecTaken>(aRecM + (uint32_t)n * kRecBytes);
    for (int f = lane, more = 1; f < F && more; f += 32, more = kMorePasses) {
        p.st.f_money[eF + f] = lds_f64(aFmoney + 8u * f);
        if (do_firms) p.out.f_profit[eF + f] = lds_f64(aFlast + 8u * f);
#pragma unroll
        for (int g = 0; g < G; g++) p.st.f_inv[((size_t)e * G + g) * F + f] = lds_f64(aFinv + 8u * (g * F + f));
        double labor = lds_f64(aFlabor + 8u * f);
        const uint32_t nhf = lds_u32(aFnh + 4u * f);
        for (uint32_t k = 0; k < nhf; k++) labor += kLaborPerOffer;  // firm.cpp:109, one add per hire
        p.st.f_labor[eF + f] = labor;
        if (do_firms && optional_out && p.out.f_good_ok) {
            const uint32_t ok = lds_u32(aFok + 4u * f);
            for (int i = 0; i < S; i++) p.out.f_good_ok[((size_t)e * S + i) * F + f] = (ok >> i) & 1u;
        }
    }
    if (mp.done_list) {
        // this economy's matching results are complete: hand it to update_kernel (every lane fences its own writes,
        // then lane 0 publishes)
        fence_gpu();
        __syncwarp();
        if (lane == 0) store_relaxed_u32(mp.done_list + (ticket - mp.ticket_base), (mp.done_tag << kQueueTagShift) | (uint32_t)e);
    }
#ifdef FASTACE_CTA_TIMING
    if (lane == 0 && e < 65536) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_cta_times[3 * e] = cta_t0; g_cta_times[3 * e + 1] = global_ns(); g_cta_times[3 * e + 2] = smid;
    }
#endif
}

}  // namespace fastace
