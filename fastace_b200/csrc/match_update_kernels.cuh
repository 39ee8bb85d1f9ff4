// Economy::time_step for E independent economies — default path: two sm_100a kernels.
//
//   match_kernel<G>   one warp per economy.  Stages the economy's offer books, firm state and
//                     the agents' request lists in shared memory and resolves ALL matching of
//                     the step (person phase, then firm phase) first-come-first-served in the
//                     visiting order.  No transcendental math, no inventories of persons.
//   update_kernel<G>  fully parallel element-wise work that depends on the matching result:
//                     one thread per person (apply purchases, consume, CES utility = reward);
//                     one lane per (firm, output good) (CES production, decode of the new goods /
//                     job offers, warp scan -> new books in market order).
//
// Person matching is lane-parallel and exact: lanes = persons.  A window of 32 persons with
// consecutive visiting ranks evaluates its whole request chains in parallel.  For every offer R
// the ordinal number of an ELIGIBLE request (person-side check passed: person.cpp:39 /
// agent.cpp:102) is
//     ord = (# eligible requests on R by lower lanes of the window) + (own earlier ones)
// and the request succeeds iff ord < D[R], the offer's death ordinal:
//     job offer   : lots left (firm.cpp:64), lowered to the first request its firm cannot pay
//                   (firm.cpp:80-84: that request kills the offer);
//     goods offer : min(lots left, floor(seller inventory)) (agent.cpp:124,140-143).
// Offers only ever lose availability, so the serial first-come-first-served outcome is the unique
// fixed point of (prefix counts, D).  It is reached by iterating
//     evaluate (lanes = persons)  ->  prefix scan (lanes = offers)  ->  evaluate ...
// Per (offer, lane) two bytes live in shared memory: cnt = eligible requests of that lane in the
// current evaluation, room = clamp(D - prefix, 0, 31) = how many of them can succeed.  The
// iteration stops as soon as min(cnt, room) is unchanged for every cell: the next evaluation would
// reproduce this one, so this one already is the fixed point.  Lane k is exact after k+2 rounds at
// the latest; in practice 2-3 rounds per window.
// Person money is accumulated in the person's own request order (bit-exact).  A firm's money is
// updated once per window as M - wage*hires + sum price*sales: the same value up to fp64 rounding
// order; FASTACE_STEP_SERIAL selects the strictly ordered serial kernel instead.
//
// Reference code restated (paths under /root/reference/src): see step_kernel.cuh header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fastace {

constexpr int kRowStride = 36;   // bytes per cell row: 9 words -> consecutive offers start in distinct banks
constexpr int kMaxStack = FASTACE_MAX_STACK;

struct MatchLayout {
    int off_mrec;                                                                // uint4 [F*G]: price (f64), owner | good<<8
    int off_pmoney, off_fmoney, off_finv, off_jwage;                             // double
    int off_mleft, off_mtaken, off_jleft, off_jtaken, off_fnh, off_fok;          // u32
    int off_dord, off_tot;                                                       // i32 [F*(G+1)]
    int off_permp, off_permf;                                                    // u16
    int off_att;                                                                 // u8 [P][AS]: jobs at [0,S), goods at [S,2S), AS = 2S rounded up to 4
    int off_fatt;                                                                // u8 [F][16]
    int off_jowner, off_pnh, off_pnb, off_ffirst, off_fcnt, off_fjob;
    int off_cnt, off_room;                                                       // u8 [F*(G+1)][kRowStride]
    int Pp;                                                                      // P rounded up to 4
    int AS;                                                                      // bytes per person in the request table
    int total;
};

__host__ __device__ inline MatchLayout make_match_layout(int P, int F, int G, int S) {
    MatchLayout L;
    const int cap = F * G, nr = F * (G + 1);
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 7) & ~7; return r; };
    L.Pp = (P + 3) & ~3;
    L.AS = (2 * S + 3) & ~3;
    L.off_mrec = take(16 * cap);          // first: 16-byte aligned
    L.off_fatt = take(16 * F);
    L.off_pmoney = take(8 * P);
    L.off_fmoney = take(8 * F);
    L.off_finv = take(8 * G * F);
    L.off_jwage = take(8 * F);
    L.off_mleft = take(4 * cap);
    L.off_mtaken = take(4 * cap);
    L.off_jleft = take(4 * F);
    L.off_jtaken = take(4 * F);
    L.off_fnh = take(4 * F);
    L.off_fok = take(4 * F);
    L.off_dord = take(4 * nr);
    L.off_tot = take(4 * nr);
    L.off_permp = take(2 * P);
    L.off_permf = take(2 * F);
    L.off_att = take(P * L.AS + 8);   // +8: the gather may read one word past the last row
    L.off_jowner = take(F);
    L.off_pnh = take(L.Pp);
    L.off_pnb = take(G * L.Pp);
    L.off_ffirst = take(F);
    L.off_fcnt = take(F);
    L.off_fjob = take(F);
    L.off_cnt = take(kRowStride * nr);
    L.off_room = take(kRowStride * nr);
    L.total = o;
    return L;
}

struct MatchParams {
    StepParams sp;
    MatchLayout lay;   // computed on the host: offsets come from the constant bank instead of being
                       // re-derived in registers all over the kernel
    uint8_t* scr_pnh;  // [E][P]    hires per person (0..2)         -> update_kernel
    uint8_t* scr_pnb;  // [E][G][P] purchases per person and good   -> update_kernel
};

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24);
}

// SMAX: compile-time bound of the stack size S (12 or 16): request lists live in SMAX/4 registers
// per list and the evaluation is fully unrolled over SMAX slots.
template <int G, int SMAX>
__global__ void __launch_bounds__(32, 28) match_kernel(const MatchParams mp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const StepParams& p = mp.sp;
    const int e = blockIdx.x;
    const int lane = threadIdx.x;
    const int P = p.P, F = p.F, S = p.S;
    const int cap = F * G;
    const MatchLayout& L = mp.lay;
    const int Pp = L.Pp, AS = L.AS;

    double* s_pmoney = reinterpret_cast<double*>(smem + L.off_pmoney);
    double* s_fmoney = reinterpret_cast<double*>(smem + L.off_fmoney);
    double* s_finv = reinterpret_cast<double*>(smem + L.off_finv);
    uint4* s_mrec = reinterpret_cast<uint4*>(smem + L.off_mrec);
    double* s_jwage = reinterpret_cast<double*>(smem + L.off_jwage);
    uint32_t* s_mleft = reinterpret_cast<uint32_t*>(smem + L.off_mleft);
    uint32_t* s_mtaken = reinterpret_cast<uint32_t*>(smem + L.off_mtaken);
    uint32_t* s_jleft = reinterpret_cast<uint32_t*>(smem + L.off_jleft);
    uint32_t* s_jtaken = reinterpret_cast<uint32_t*>(smem + L.off_jtaken);
    uint32_t* s_fnh = reinterpret_cast<uint32_t*>(smem + L.off_fnh);
    uint32_t* s_fok = reinterpret_cast<uint32_t*>(smem + L.off_fok);
    int32_t* s_dord = reinterpret_cast<int32_t*>(smem + L.off_dord);
    int32_t* s_tot = reinterpret_cast<int32_t*>(smem + L.off_tot);
    uint16_t* s_permp = reinterpret_cast<uint16_t*>(smem + L.off_permp);
    uint16_t* s_permf = reinterpret_cast<uint16_t*>(smem + L.off_permf);
    uint8_t* s_att = smem + L.off_att;
    uint8_t* s_fatt = smem + L.off_fatt;
    uint8_t* s_jowner = smem + L.off_jowner;
    uint8_t* s_pnh = smem + L.off_pnh;
    uint8_t* s_pnb = smem + L.off_pnb;
    uint8_t* s_ffirst = smem + L.off_ffirst;
    uint8_t* s_fcnt = smem + L.off_fcnt;
    uint8_t* s_fjob = smem + L.off_fjob;
    uint8_t* s_cnt = smem + L.off_cnt;
    uint8_t* s_room = smem + L.off_room;
    double* s_flast = reinterpret_cast<double*>(s_room);   // firm phase only (cells are free then): last_money in, profit out
    auto rec_price = [&](int n) { return *reinterpret_cast<const double*>(&s_mrec[n]); };
    auto rec_owner = [&](int n) { return (int)(s_mrec[n].z & 0xFFu); };
    auto rec_good = [&](int n) { return (int)((s_mrec[n].z >> 8) & 0xFFu); };

    const size_t eP = (size_t)e * P, eF = (size_t)e * F, eCap = (size_t)e * cap;
    // a step may be taken in two calls (FASTACE_STEP_PERSONS, then FASTACE_STEP_FIRMS): the state in HBM between
    // them is the economy as it stands when the last person has acted and no firm has (economy.cpp:118-123)
    const bool do_persons = !(p.flags & FASTACE_STEP_FIRMS);
    const bool do_firms = !(p.flags & (FASTACE_STEP_PERSONS | FASTACE_STEP_PERSONS_TRADE));
    const int NM = p.st.m_count[e];
    const int NJ = p.st.j_count[e];
    const int NR = NJ + NM;

    // ------------------------------ stage --------------------------------------------------
    {
        const IndexMap mapJ(NJ, p.flags), mapM(NM, p.flags);
        const bool hasJ = NJ > 0, hasM = NM > 0;   // empty book: no requests at all (decisionNetHandler.cpp:398-403, 476-480)
        for (int pid = lane; do_persons && pid < P; pid += 32) {
            s_pmoney[pid] = p.st.p_money[eP + pid];
            s_permp[pid] = p.compact ? p.cz.perm_person[eP + pid] : (uint16_t)p.ac.perm_person[eP + pid];
        }
        // request lists -> person-major rows; 4 consecutive persons per lane when rows are 16B-aligned
        const size_t row0 = (size_t)e * S * P;
        if (!do_persons) {
            // firms-only call: no person requests are read
        } else if (p.compact) {
            for (int pid = lane; pid < P; pid += 32) {
                const uint32_t tj = hasJ ? p.cz.p_job_take[eP + pid] : 0u, tg = hasM ? p.cz.p_good_take[eP + pid] : 0u;
                uint8_t* row = s_att + pid * AS;
                for (int i = 0; i < S; i++) {
                    const size_t k = row0 + (size_t)i * P + pid;
                    row[i] = (uint8_t)(((tj >> i) & 1u) ? mapJ((int)p.cz.p_job_idx[k]) : kNone);
                    row[S + i] = (uint8_t)(((tg >> i) & 1u) ? mapM((int)p.cz.p_good_idx[k]) : kNone);
                }
            }
        } else if ((P & 3) == 0) {
            const int Q = P >> 2;
#pragma unroll 2
            for (int i = 0; i < S; i++) {
                const int4* ji = reinterpret_cast<const int4*>(p.ac.p_job_idx + row0 + (size_t)i * P);
                const int4* gi = reinterpret_cast<const int4*>(p.ac.p_good_idx + row0 + (size_t)i * P);
                const uchar4* jt = reinterpret_cast<const uchar4*>(p.ac.p_job_take + row0 + (size_t)i * P);
                const uchar4* gt = reinterpret_cast<const uchar4*>(p.ac.p_good_take + row0 + (size_t)i * P);
                for (int q = lane; q < Q; q += 32) {
                    const int4 a = ji[q], b = gi[q];
                    const uchar4 ta = jt[q], tb = gt[q];
                    uint8_t* row = s_att + (4 * q) * AS + i;
                    row[0] = (uint8_t)((ta.x && hasJ) ? mapJ(a.x) : kNone);
                    row[AS] = (uint8_t)((ta.y && hasJ) ? mapJ(a.y) : kNone);
                    row[2 * AS] = (uint8_t)((ta.z && hasJ) ? mapJ(a.z) : kNone);
                    row[3 * AS] = (uint8_t)((ta.w && hasJ) ? mapJ(a.w) : kNone);
                    row += S;
                    row[0] = (uint8_t)((tb.x && hasM) ? mapM(b.x) : kNone);
                    row[AS] = (uint8_t)((tb.y && hasM) ? mapM(b.y) : kNone);
                    row[2 * AS] = (uint8_t)((tb.z && hasM) ? mapM(b.z) : kNone);
                    row[3 * AS] = (uint8_t)((tb.w && hasM) ? mapM(b.w) : kNone);
                }
            }
        } else {
            for (int i = 0; i < S; i++) {
                const size_t rb = row0 + (size_t)i * P;
                for (int pid = lane; pid < P; pid += 32) {
                    s_att[pid * AS + i] = (uint8_t)((p.ac.p_job_take[rb + pid] && hasJ) ? mapJ(p.ac.p_job_idx[rb + pid]) : kNone);
                    s_att[pid * AS + S + i] = (uint8_t)((p.ac.p_good_take[rb + pid] && hasM) ? mapM(p.ac.p_good_idx[rb + pid]) : kNone);
                }
            }
        }
        {
            uint32_t* z = reinterpret_cast<uint32_t*>(s_pnh);
            for (int k = lane; k < (Pp >> 2); k += 32) z[k] = 0u;
            z = reinterpret_cast<uint32_t*>(s_pnb);
            for (int k = lane; k < (G * Pp >> 2); k += 32) z[k] = 0u;
        }
        for (int f = lane; f < F; f += 32) {
            s_fmoney[f] = p.st.f_money[eF + f];
            s_permf[f] = do_firms ? (uint16_t)perm_firm_at(p, eF + f) : (uint16_t)f;
            s_fnh[f] = 0;
            s_fok[f] = 0;
            s_fcnt[f] = 0;
            s_ffirst[f] = 0;
            s_fjob[f] = (uint8_t)kNone;
#pragma unroll
            for (int g = 0; g < G; g++) s_finv[g * F + f] = p.st.f_inv[((size_t)e * G + g) * F + f];
            const size_t k0 = (size_t)e * S * F + f;
            for (int i = do_firms ? S : 0; i < 16; i++) s_fatt[f * 16 + i] = (uint8_t)kNone;
            if (!do_firms) {
                // persons-only call: the firms' requests are not read
            } else if (p.compact) {
                const uint32_t tg = hasM ? p.cz.f_good_take[eF + f] : 0u;
                for (int i = 0; i < S; i++)
                    s_fatt[f * 16 + i] = (uint8_t)(((tg >> i) & 1u) ? mapM((int)p.cz.f_good_idx[k0 + (size_t)i * F]) : kNone);
            } else {
                for (int i = 0; i < S; i++)
                    s_fatt[f * 16 + i] = (uint8_t)((p.ac.f_good_take[k0 + (size_t)i * F] && hasM) ? mapM(p.ac.f_good_idx[k0 + (size_t)i * F]) : kNone);
            }
        }
    }
    for (int n = lane; n < NJ; n += 32) {
        s_jowner[n] = (uint8_t)p.st.j_owner[eF + n];
        s_jleft[n] = p.st.j_left[eF + n];
        s_jtaken[n] = p.st.j_taken[eF + n];
        s_jwage[n] = p.st.j_wage[eF + n];
    }
    {
        uint32_t* cz = reinterpret_cast<uint32_t*>(s_cnt);
        const int words = (kRowStride / 4) * NR;
        for (int k = lane; k < words; k += 32) cz[k] = 0u;
    }
    __syncwarp();
    for (int n = lane; n < NJ; n += 32) s_fjob[s_jowner[n]] = (uint8_t)n;
    for (int n = lane; n < NM; n += 32) {
        const int owner = p.st.m_owner[eCap + n];
        const double price = p.st.m_price[eCap + n];
        s_mrec[n] = make_uint4((uint32_t)__double2loint(price), (uint32_t)__double2hiint(price),
                               (uint32_t)(owner & 0xFF) | ((uint32_t)(p.st.m_good[eCap + n] & 0xFF) << 8), 0u);
        s_mleft[n] = p.st.m_left[eCap + n];
        s_mtaken[n] = p.st.m_taken[eCap + n];
        // a firm's entries are contiguous in market order (it posts all goods in one turn)
        const int prev = (n > 0) ? p.st.m_owner[eCap + n - 1] : -1;
        if (owner != prev) s_ffirst[owner] = (uint8_t)n;
    }
    __syncwarp();
    for (int n = lane; n < NM; n += 32) {
        const int owner = rec_owner(n);
        const int next = (n + 1 < NM) ? rec_owner(n + 1) : -1;
        if (owner != next) s_fcnt[owner] = (uint8_t)(n + 1 - s_ffirst[owner]);
    }
    __syncwarp();

    // ------------------------------ persons: windows of 32 visiting ranks -------------------
    for (int base = 0; do_persons && base < P; base += 32) {
        const int r = base + lane;
        const bool active = r < P;
        const int pid = active ? (int)s_permp[r] : 0;
        const double money0 = active ? s_pmoney[pid] : 0.0;
        // the lane's request lists, 4 slots per register (kNone beyond S); goods start at byte S
        uint32_t aj[SMAX / 4], ag[SMAX / 4];
        {
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(s_att + pid * AS);
            const int gw = S >> 2, gs = 8 * (S & 3);
#pragma unroll
            for (int k = 0; k < SMAX / 4; k++) {
                const int valid = min(max(S - 4 * k, 0), 4);                       // slots of this word that exist
                const uint32_t none = valid >= 4 ? 0u : (0xFFFFFFFFu << (8 * valid));
                uint32_t j = 0xFFFFFFFFu, g = 0xFFFFFFFFu;
                if (active && valid > 0) {
                    j = aw[k] | none;
                    g = __funnelshift_r(aw[gw + k], aw[gw + k + 1], gs) | none;
                }
                aj[k] = j;
                ag[k] = g;
            }
        }
        for (int R = lane; R < NR; R += 32) {
            uint32_t d;
            if (R < NJ) {
                d = s_jleft[R];
            } else {
                const int o = R - NJ, sel = rec_owner(o), good = rec_good(o);
                d = min(s_mleft[o], unit_sales_possible(s_finv[good * F + sel]));
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (g != good && s_finv[g * F + sel] < 0.0) d = 0;  // agent.cpp:140 on a zero quantity
            }
            const int di = (int)min(d, 0x7FFFFFFFu);
            s_dord[R] = di;
            // initial guess for the window: no lower lane is eligible for anything
            uint32_t* rr = reinterpret_cast<uint32_t*>(s_room + R * kRowStride);
            const uint32_t rm = (uint32_t)min(di, 31) * 0x01010101u;
#pragma unroll
            for (int k = 0; k < 8; k++) rr[k] = rm;
        }
        __syncwarp();
        double money = money0;
        int nh = 0;
        uint32_t okm = 0;
        for (int round = 0; round < 80; round++) {
            // ---- evaluate the window's request chains against room[][]
            money = money0; nh = 0; okm = 0;
#pragma unroll
            for (int i = 0; i < SMAX; i++) {                       // utilMaxer.cpp:76-85
                if (i < S) {
                    const int n = (int)((aj[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                    if (n != kNone && nh < 2) {                         // person.cpp:39 (0.5*nh + 0.5 <= 1)
                        const int a = n * kRowStride + lane;
                        const uint32_t c = s_cnt[a];
                        s_cnt[a] = (uint8_t)(c + 1u);
                        if (c < s_room[a]) {
                            nh++;
                            money += s_jwage[n];                        // person.cpp:49
                            okm |= 1u << i;
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < SMAX; i++) {                       // utilMaxer.cpp:64-73
                if (i < S) {
                    const int n = (int)((ag[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                    if (n != kNone) {
                        const double price = rec_price(n);
                        if (money >= price) {                           // agent.cpp:102
                            const int a = (NJ + n) * kRowStride + lane;
                            const uint32_t c = s_cnt[a];
                            s_cnt[a] = (uint8_t)(c + 1u);
                            if (c < s_room[a]) {
                                money -= price;                         // agent.cpp:108
                                okm |= 1u << (16 + i);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // ---- lanes = offers: prefix of the eligible counts over the window's lanes -> room
            bool changed = false;
            for (int R = lane; R < NR; R += 32) {
                const uint32_t* cw = reinterpret_cast<const uint32_t*>(s_cnt + R * kRowStride);
                uint32_t* rw = reinterpret_cast<uint32_t*>(s_room + R * kRowStride);
                const int d = s_dord[R];
                int run = 0;
                uint32_t cvs[8], ros[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { cvs[k] = cw[k]; ros[k] = rw[k]; }   // all loads in flight at once
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t cv = cvs[k], ro = ros[k];
                    uint32_t rn = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int c = (int)((cv >> (8 * b)) & 0xFFu);
                        const int oldroom = (int)((ro >> (8 * b)) & 0xFFu);
                        const int room = max(0, min(31, d - run));
                        changed |= min(c, room) != min(c, oldroom);
                        rn |= (uint32_t)room << (8 * b);
                        run += c;
                    }
                    rw[k] = rn;
                }
                s_tot[R] = run;
            }
            __syncwarp();
            // ---- job offers whose firm may run out of money: first request it cannot pay
            for (int R = lane; R < NJ; R += 32) {
                const uint32_t left = s_jleft[R];
                const int f = s_jowner[R];
                const double w = s_jwage[R], m0 = s_fmoney[f];
                const int tot = s_tot[R];
                int d = (int)min(left, 0x7FFFFFFFu);
                const int most = min(d, tot);
                if (!(m0 - w * (double)most >= w * (1.0 + 1e-9))) {
                    const uint8_t* row = s_cnt + R * kRowStride;
                    const int first = s_ffirst[f], cnt = s_fcnt[f];
                    int pre[G];          // eligible requests on the firm's own goods offers by lower lanes
#pragma unroll
                    for (int g = 0; g < G; g++) pre[g] = 0;
                    int h = 0;
                    bool done = false;
                    for (int l = 0; l < 32 && !done; l++) {
                        const int c = row[l];
                        if (c != 0) {
                            double sales = 0.0;   // income from goods sold to lower lanes of this window
#pragma unroll
                            for (int g = 0; g < G; g++)
                                if (g < cnt) {
                                    const int sold = min(pre[g], s_dord[NJ + first + g]);
                                    if (sold > 0) sales += rec_price(first + g) * (double)sold;
                                }
                            for (int k = 0; k < c; k++) {
                                if (h >= d) { done = true; break; }                                  // firm.cpp:64
                                if ((m0 + sales) - w * (double)h < w) { d = h; done = true; break; } // firm.cpp:80
                                h++;
                            }
                        }
#pragma unroll
                        for (int g = 0; g < G; g++)
                            if (g < cnt) pre[g] += s_cnt[(NJ + first + g) * kRowStride + l];
                    }
                }
                if (d != s_dord[R]) {
                    // the offer dies earlier/later than assumed: redo its row with the new ordinal
                    changed = true;
                    s_dord[R] = d;
                    const uint8_t* row = s_cnt + R * kRowStride;
                    uint8_t* rrow = s_room + R * kRowStride;
                    int run = 0;
                    for (int l = 0; l < 32; l++) { rrow[l] = (uint8_t)max(0, min(31, d - run)); run += row[l]; }
                }
            }
            const bool again = __any_sync(0xffffffffu, changed);
            __syncwarp();
            {
                uint32_t* cz = reinterpret_cast<uint32_t*>(s_cnt);
                const int words = (kRowStride / 4) * NR;
                for (int k = lane; k < words; k += 32) cz[k] = 0u;
            }
            __syncwarp();
            if (!again) break;
        }
        // ---- commit the window
        if (active) {
            s_pmoney[pid] = money;
            s_pnh[pid] = (uint8_t)nh;
            write_person_ok(p, e, pid, okm);
            for (uint32_t m = okm >> 16; m != 0; m &= m - 1) {      // one purchase per set bit
                const int i = __ffs(m) - 1;
                uint32_t w = ag[0];
#pragma unroll
                for (int k = 1; k < SMAX / 4; k++) if ((i >> 2) == k) w = ag[k];
                const int n = (int)((w >> (8 * (i & 3))) & 0xFFu);
                s_pnb[rec_good(n) * Pp + pid] += 1;
            }
        }
        for (int R = lane; R < NR; R += 32) {
            const int tot = s_tot[R];
            const int n = min(tot, s_dord[R]);
            if (R < NJ) {
                const uint32_t left = s_jleft[R];
                s_jleft[R] = (tot > n) ? 0u : left - (uint32_t)n;   // exhausted or killed (firm.cpp:83)
                s_jtaken[R] += (uint32_t)n;
            } else {
                const int o = R - NJ;
                uint32_t left = s_mleft[o] - (uint32_t)n;
                if (tot > n && left > 0) left = 0;                  // killed (agent.cpp:143)
                s_mleft[o] = left;
                s_mtaken[o] += (uint32_t)n;
                s_finv[rec_good(o) * F + rec_owner(o)] -= (double)n;   // n exact unit subtractions
            }
            s_tot[R] = n;
        }
        __syncwarp();
        for (int f = lane; f < F; f += 32) {
            double m = s_fmoney[f];
            const int j = s_fjob[f];
            if (j != kNone) {
                const int h = s_tot[j];
                if (h > 0) m = m - s_jwage[j] * (double)h;
                s_fnh[f] += (uint32_t)h;
            }
            const int first = s_ffirst[f], cnt = s_fcnt[f];
            for (int o = first; o < first + cnt; o++) {
                const int sold = s_tot[NJ + o];
                if (sold > 0) m = m + rec_price(o) * (double)sold;   // (an unsold offer may carry an inf/NaN price)
            }
            s_fmoney[f] = m;
        }
        __syncwarp();
    }

    // ------------------------------ persons: results to HBM ---------------------------------
    for (int pid = lane; do_persons && pid < P; pid += 32) {
        p.st.p_money[eP + pid] = s_pmoney[pid];
        mp.scr_pnh[eP + pid] = s_pnh[pid];
#pragma unroll
        for (int g = 0; g < G; g++) mp.scr_pnb[((size_t)e * G + g) * P + pid] = s_pnb[g * Pp + pid];
    }
    // job counters are final after the person phase
    if (do_persons && p.out.old_j_left) for (int n = lane; n < NJ; n += 32) p.out.old_j_left[eF + n] = s_jleft[n];
    if (do_persons && p.out.old_j_taken) for (int n = lane; n < NJ; n += 32) p.out.old_j_taken[eF + n] = s_jtaken[n];
    if (!do_firms) {
        // persons-only call: the books' counters go back to HBM for the firms call (a full step never needs them
        // there: update_kernel replaces the books)
        for (int n = lane; n < NM; n += 32) { p.st.m_left[eCap + n] = s_mleft[n]; p.st.m_taken[eCap + n] = s_mtaken[n]; }
        for (int n = lane; n < NJ; n += 32) { p.st.j_left[eF + n] = s_jleft[n]; p.st.j_taken[eF + n] = s_jtaken[n]; }
    }

    // ------------------------------ firms: serial walk in visiting order ---------------------
    for (int f = lane; do_firms && f < F; f += 32) s_flast[f] = p.st.f_last_money[eF + f];
    __syncwarp();
    if (lane == 0 && do_firms) {
        for (int r = 0; r < F; r++) {
            const int f = s_permf[r];
            const int first = s_ffirst[f], cnt = s_fcnt[f];
            const uint4 slots = *reinterpret_cast<const uint4*>(s_fatt + f * 16);   // the firm's request bytes
            double money = s_fmoney[f];
            // Agent::check_my_offers (base/agent.cpp:54-97): running inventoryLeft over own entries
            {
                double invLeft[G];
#pragma unroll
                for (int g = 0; g < G; g++) invLeft[g] = s_finv[g * F + f];
                for (int n = first; n < first + cnt; n++) {
                    const int good = rec_good(n);
                    uint32_t left = s_mleft[n];
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
                    s_mleft[n] = left;
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
#pragma unroll 1
            for (int i = 0; i < S; i++) {
                const uint32_t word = (i < 4) ? slots.x : (i < 8) ? slots.y : (i < 12) ? slots.z : slots.w;
                const int n = (int)((word >> (8 * (i & 3))) & 0xFFu);
                if (n != kNone) {
                    const uint4 rec = s_mrec[n];
                    const double price = __hiloint2double((int)rec.y, (int)rec.x);
                    if (money >= price) {                                  // agent.cpp:102
                        const uint32_t left = s_mleft[n];
                        if (left > 0) {                                    // agent.cpp:124
                            const int s = (int)(rec.z & 0xFFu), good = (int)((rec.z >> 8) & 0xFFu);
                            bool short_ = false;                           // agent.cpp:140
#pragma unroll
                            for (int g = 0; g < G; g++) {
                                const double q = (g == good) ? kAmountPerOffer : 0.0;
                                if (s_finv[g * F + s] < q) short_ = true;
                            }
                            if (short_) {
                                s_mleft[n] = 0;                            // agent.cpp:143
                            } else {
                                // seller first (agent.cpp:155-160), then buyer (agent.cpp:108-109)
                                if (s == f) money += price; else s_fmoney[s] += price;
                                s_finv[good * F + s] -= kAmountPerOffer;
                                s_mleft[n] = left - 1;
                                s_mtaken[n] += 1;
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
                s_dord[n] = (int)s_mleft[n];   // final counters of the withdrawn entry (s_dord is free now)
                s_mleft[n] = 0;
            }
        }
    }
    __syncwarp();
    // ------------------------------ firms: results to HBM -----------------------------------
    if (do_firms && p.out.old_m_left) for (int n = lane; n < NM; n += 32) p.out.old_m_left[eCap + n] = (uint32_t)s_dord[n];
    if (do_firms && p.out.old_m_taken) for (int n = lane; n < NM; n += 32) p.out.old_m_taken[eCap + n] = s_mtaken[n];
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

// -------------------------------------------------------------------------------------------------
// update_kernel: blocks [0, firm_blocks) handle firms (one warp per economy, lanes = (visiting rank,
// output good) pairs; scheduled first because their fp64 pow chains are the longest), the remaining
// blocks handle persons (one thread each).
struct UpdateParams {
    StepParams sp;
    const uint8_t* scr_pnh;
    const uint8_t* scr_pnb;
    int firm_blocks;
};

constexpr int kUpdateThreads = 128;

template <int G>
__global__ void __launch_bounds__(kUpdateThreads, 8) update_kernel(const UpdateParams up) {
    const StepParams& p = up.sp;
    const int P = p.P, F = p.F;
    if ((int)blockIdx.x >= up.firm_blocks) {
        // ---- UtilMaxer::consume_goods (utilMaxer.cpp:88-92) + choose_goods_to_consume
        //      (neuralPersonDecisionMaker.cpp:93-111) + UtilMaxer::u (utilMaxer.cpp:54-62)
        const size_t t = (size_t)((int)blockIdx.x - up.firm_blocks) * kUpdateThreads + threadIdx.x;
        if (t >= (size_t)p.E * P) return;
        const int e = (int)(t / P), pid = (int)(t % P);
        const double labor = kLaborPerOffer * (double)up.scr_pnh[t];       // exact: 0, 0.5 or 1.0
        if (p.flags & FASTACE_STEP_PERSONS_TRADE) {
            // trade-only call: purchases and labour go into the state, consumption follows in its own call
            // (it touches nobody but the person, so deferring it changes nothing: utilMaxer.cpp:88-92)
#pragma unroll
            for (int g = 0; g < G; g++) {
                const size_t k = ((size_t)e * G + g) * P + pid;
                double v = p.st.p_inv[k];
                const int nb = up.scr_pnb[k];
                for (int q = 0; q < nb; q++) v += kAmountPerOffer;         // agent.cpp:109, one add per purchase
                p.st.p_inv[k] = v;
            }
            p.st.p_labor[t] = labor;
            return;
        }
        const bool applied = (p.flags & FASTACE_STEP_PERSONS_CONSUME) != 0;   // purchases are already in p_inv
        double x[G + 1], inv[G];
        x[0] = 1 - labor;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const size_t k = ((size_t)e * G + g) * P + pid;
            double v = p.st.p_inv[k];
            const int nb = applied ? 0 : up.scr_pnb[k];
            for (int q = 0; q < nb; q++) v += kAmountPerOffer;             // agent.cpp:109, one add per purchase
            const double c = v * (double)p.ac.p_consume[k];                 // neuralPersonDecisionMaker.cpp:99
            x[g + 1] = c;
            inv[g] = v - c;                                                 // utilMaxer.cpp:91
        }
        double share[G + 1], theta[G + 1];
#pragma unroll
        for (int i = 0; i <= G; i++) {
            share[i] = p.st.p_util_share[((size_t)e * (G + 1) + i) * P + pid];
            theta[i] = (p.util_kind == FASTACE_FN_STONE_GEARY) ? p.st.p_util_theta[((size_t)e * (G + 1) + i) * P + pid] : 0.0;
        }
        p.out.p_reward[t] = eval_function<G + 1, true>(p.util_kind, p.st.p_util_tfp[t], share, theta, p.st.p_util_rho[t], x);
        p.st.p_labor[t] = labor;
#pragma unroll
        for (int g = 0; g < G; g++) p.st.p_inv[((size_t)e * G + g) * P + pid] = inv[g];
        return;
    }
    // ---- firms: produce (profitMaxer.cpp:68-72), sell_goods / search_for_laborers decode
    //      (neuralFirmDecisionMaker.cpp:111-180), new books in market order (economy.cpp:52-59, 125-126)
    const int warps_per_block = kUpdateThreads / 32;
    const int e = (int)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (e >= p.E) return;
    const int lane = threadIdx.x & 31;
    const int cap = F * G;
    const size_t eF = (size_t)e * F, eCap = (size_t)e * cap;
    int base_m = 0, base_j = 0;
    // whole firms per pass (FPC firms x G goods <= 32 lanes), so that all lanes reading a firm's
    // inventories / labour do so before any lane of the same pass overwrites them
    const int FPC = 32 / G;
    for (int r0 = 0; r0 < F; r0 += FPC) {
        const int r = r0 + lane / G, g = lane % G;
        const bool active = lane < FPC * G && r < F;
        int lots = 0, jlots = 0, f = 0;
        double price = 0.0, jwage = 0.0, newinv = 0.0;
        if (active) {
            f = perm_firm_at(p, eF + r);
            double in[G + 1];
            in[0] = p.st.f_labor[eF + f];
            double xg = 0.0, invg = 0.0;
#pragma unroll
            for (int k = 0; k < G; k++) {
                const size_t a = ((size_t)e * G + k) * F + f;
                const double iv = p.st.f_inv[a];
                const double xk = iv * (double)p.ac.f_prod[a];              // neuralFirmDecisionMaker.cpp:101
                in[k + 1] = xk;
                if (k == g) { xg = xk; invg = iv; }
            }
            const size_t ag = ((size_t)e * G + g) * F + f;
            double share[G + 1], theta[G + 1];
#pragma unroll
            for (int i = 0; i <= G; i++) {
                const size_t k = (((size_t)e * G + g) * (G + 1) + i) * F + f;
                share[i] = p.st.f_prod_share[k];
                theta[i] = (p.prod_kind == FASTACE_FN_STONE_GEARY) ? p.st.f_prod_theta[k] : 0.0;
            }
            const double outg = eval_function<G + 1, false>(p.prod_kind, p.st.f_prod_tfp[ag], share, theta, p.st.f_prod_rho[ag], in);
            newinv = invg + (outg - xg);                                    // profitMaxer.cpp:71
            // decisionNetHandler.cpp:591 amounts = proportion * inventory; neuralFirmDecisionMaker.cpp:129
            const double amount = (double)p.ac.f_offer_amt[ag] * newinv;
            lots = x86_double_to_int(amount / kAmountPerOffer);
            price = (double)p.ac.f_offer_price[ag] / kAmountPerOffer;
            if (g == 0) {
                // neuralFirmDecisionMaker.cpp:164-176, decisionNetHandler.cpp:631-635
                double wage = (double)p.ac.f_job_wage[eF + f];
                if (wage > kLargeNumber) wage = kLargeNumber;
                jlots = x86_double_to_int((double)p.ac.f_job_labor[eF + f] / kLaborPerOffer);
                jwage = wage / kLaborPerOffer;
            }
        }
        __syncwarp();   // every lane has read its firm's inventories and labour before any is overwritten
        const unsigned mmask = __ballot_sync(0xffffffffu, active && lots > 0);
        const unsigned jmask = __ballot_sync(0xffffffffu, active && g == 0 && jlots > 0);
        if (active) {
            const size_t ag = ((size_t)e * G + g) * F + f;
            p.st.f_inv[ag] = newinv;
            if (g == 0) p.st.f_labor[eF + f] = 0.0;                          // firm.cpp:41
            if (lots > 0) {
                const int slot = base_m + __popc(mmask & ((1u << lane) - 1u));
                p.st.m_owner[eCap + slot] = f;
                p.st.m_good[eCap + slot] = g;
                p.st.m_left[eCap + slot] = (uint32_t)lots;
                p.st.m_taken[eCap + slot] = 0;
                p.st.m_price[eCap + slot] = price;
            }
            if (g == 0 && jlots > 0) {
                const int slot = base_j + __popc(jmask & ((1u << lane) - 1u));
                p.st.j_owner[eF + slot] = f;
                p.st.j_left[eF + slot] = (uint32_t)jlots;
                p.st.j_taken[eF + slot] = 0;
                p.st.j_wage[eF + slot] = jwage;
            }
        }
        base_m += __popc(mmask);
        base_j += __popc(jmask);
    }
    if (lane == 0) {
        p.st.m_count[e] = base_m;
        p.st.j_count[e] = base_j;
    }
}

}  // namespace fastace
