"""Prototype (numpy) of the lane-parallel exact matching used by kernel v2 — person phase.

Window of 32 persons (consecutive visiting ranks) evaluates its request chains in parallel.
Per offer R: ordinal of an eligible request = (# eligible requests on R by earlier lanes) + (own
earlier eligible requests on R); ok = eligible and ordinal < D[R], where D[R] is the offer's
death ordinal: remaining lots (jobs) / min(lots, floor(inventory)) (goods), lowered for a job offer
when its firm cannot pay (money-kill, firm.cpp:80-84).  (P, D) are iterated to the fixed point,
which equals the serial first-come-first-served outcome.
Prints iteration statistics and checks the outcome against the oracle's success flags.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastace_b200 import _abi, scenario
from oracle.loader import Oracle

W = 32
INIT_REQ = True   # initial guess: every taken request is eligible

def person_phase(st, act, e, dims, flags, stats, walks):
    E, P, F, G, S = dims
    NJ, NM = int(st['j_count'][e]), int(st['m_count'][e])
    def mapidx(raw, n):
        if n <= 0: return -1
        if flags & 1: return int(np.uint32(raw) % np.uint32(n))
        return int(raw) if 0 <= raw < n else -1
    j_left = st['j_left'][e].copy(); j_wage = st['j_wage'][e]; j_owner = st['j_owner'][e]
    m_left = st['m_left'][e].copy(); m_price = st['m_price'][e]; m_owner = st['m_owner'][e]; m_good = st['m_good'][e]
    f_money = st['f_money'][e].copy(); f_inv = st['f_inv'][e].copy()
    ok_job = np.zeros((S, P), np.uint8); ok_good = np.zeros((S, P), np.uint8)
    perm = act['perm_person'][e]
    offers_of = {f: [o for o in range(NM) if m_owner[o] == f] for f in range(F)}
    for base in range(0, P, W):
        lanes = perm[base:base + W]
        nl = len(lanes)
        remJ = j_left[:NJ].astype(np.int64)
        capM = np.array([min(int(m_left[o]), int(np.floor(f_inv[m_good[o], m_owner[o]]))) for o in range(NM)], dtype=np.int64)
        jslot = np.full((nl, S), -1); gslot = np.full((nl, S), -1)
        for l, p in enumerate(lanes):
            for i in range(S):
                if act['p_job_take'][e, i, p]: jslot[l, i] = mapidx(act['p_job_idx'][e, i, p], NJ)
                if act['p_good_take'][e, i, p]: gslot[l, i] = mapidx(act['p_good_idx'][e, i, p], NM)
        if INIT_REQ:
            rJ = np.array([[(jslot[l] == j).sum() for l in range(nl)] for j in range(NJ)]).reshape(NJ, nl)
            rM = np.array([[(gslot[l] == o).sum() for l in range(nl)] for o in range(NM)]).reshape(NM, nl)
            PJ = np.cumsum(rJ, axis=1) - rJ; PM = np.cumsum(rM, axis=1) - rM
        else:
            PJ = np.zeros((NJ, nl), np.int64); PM = np.zeros((NM, nl), np.int64)
        DJ = remJ.copy()
        it = 0
        while True:
            it += 1
            cJ = np.zeros((NJ, nl), np.int64); cM = np.zeros((NM, nl), np.int64)
            okJ = np.zeros((nl, S), bool); okG = np.zeros((nl, S), bool)
            elJ = np.zeros((nl, S), bool); elG = np.zeros((nl, S), bool)
            for l, p in enumerate(lanes):
                money = st['p_money'][e, p]; nh = 0
                for i in range(S):
                    s = jslot[l, i]
                    if s < 0: continue
                    if nh < 2:
                        ordn = PJ[s, l] + cJ[s, l]; cJ[s, l] += 1
                        elJ[l, i] = True
                        if ordn < DJ[s]:
                            okJ[l, i] = True; nh += 1; money += j_wage[s]
                for i in range(S):
                    s = gslot[l, i]
                    if s < 0: continue
                    if money >= m_price[s]:
                        ordn = PM[s, l] + cM[s, l]; cM[s, l] += 1
                        elG[l, i] = True
                        if ordn < capM[s]:
                            okG[l, i] = True; money -= m_price[s]
            newPJ = np.cumsum(cJ, axis=1) - cJ; newPM = np.cumsum(cM, axis=1) - cM
            # death ordinals of job offers: lots, or the first eligible request the firm cannot pay
            newDJ = remJ.copy()
            for j in range(NJ):
                tot = cJ[j].sum()
                f = j_owner[j]; w = j_wage[j]
                if f_money[f] - w * min(remJ[j], tot) >= w * (1 + 1e-9):
                    continue                       # cheap sufficient condition: can always pay
                walks.append(1)
                h = 0; dead = False
                for l in range(nl):
                    if cJ[j, l] == 0: continue
                    sales = sum(m_price[o] * min(newPM[o, l], capM[o]) for o in offers_of[f])
                    for k in range(cJ[j, l]):
                        if h >= remJ[j]: dead = True; break
                        if f_money[f] + sales - w * h < w:
                            newDJ[j] = h; dead = True; break
                        h += 1
                    if dead: break
            if np.array_equal(newPJ, PJ) and np.array_equal(newPM, PM) and np.array_equal(newDJ, DJ):
                break
            PJ, PM, DJ = newPJ, newPM, newDJ
        stats.append(it)
        for l, p in enumerate(lanes):
            ok_job[:, p] = okJ[l]; ok_good[:, p] = okG[l]
        for j in range(NJ):
            h = okJ[jslot == j].sum(); tot = (elJ & (jslot == j)).sum()
            j_left[j] -= h; f_money[j_owner[j]] -= j_wage[j] * h
            if tot > h: j_left[j] = 0          # exhausted, or killed by a request the firm could not pay
        for o in range(NM):
            n = okG[gslot == o].sum(); tot = (elG & (gslot == o)).sum()
            m_left[o] -= n; f_inv[m_good[o], m_owner[o]] -= n; f_money[m_owner[o]] += m_price[o] * n
            if tot > n and m_left[o] > 0: m_left[o] = 0
    return ok_job, ok_good, j_left, m_left

def main():
    dims = (8, 100, 10, 2, 10)
    E, P, F, G, S = dims
    orc = Oracle()
    for preset_name, preset in (("bench", scenario.BENCH_PRESET), ("untuned (bankrupt firms)", dict(labor_mu=1.0))):
        state, _ = scenario.custom_initial_state(dims, 5)
        orders = scenario.OrderStream(dims, 6)
        allstats = []; walks = []; nwin = 0
        for t in range(40):
            act = scenario.synthetic_actions(dims, seed=7, step=t, perms=orders.next(), **preset)
            before = {k: v.copy() for k, v in state.items()}
            out = _abi.alloc_host("out", dims)
            orc.step(dims, state, act, out, flags=1, time_before=t)
            if t in (1, 2, 3, 5, 8, 10, 15, 20, 30, 39):
                for e in range(E):
                    stats = []
                    okj, okg, jl, ml = person_phase(before, act, e, dims, 1, stats, walks)
                    nwin += len(stats)
                    assert np.array_equal(okj, out['p_job_ok'][e]), (t, e, 'job flags differ')
                    assert np.array_equal(okg, out['p_good_ok'][e]), (t, e, 'goods flags differ')
                    nj = before['j_count'][e]
                    assert np.array_equal(jl[:nj], out['old_j_left'][e][:nj]), (t, e, 'job left differ', jl[:nj], out['old_j_left'][e][:nj])
                    allstats.append(stats)
        a = np.array(allstats)
        print("%s: exact on all checked steps; evals/window mean %s ; per step mean %.2f max %d ; money-walks %d over %d windows" %
              (preset_name, a.mean(axis=0).round(2), a.sum(axis=1).mean(), a.sum(axis=1).max(), len(walks), nwin))

if __name__ == '__main__':
    for init in (False, True):
        INIT_REQ = init
        print("initial guess: every request eligible" if init else "initial guess: nobody else eligible")
        main()
