"""Constructed near-ties of a firm's money: economies in which the reference's event-by-event
fp64 running money lands EXACTLY on the side of `money < wage` (base/firm.cpp:80) opposite to where the same real number
summed in another order (money - wage*hires + price*sales) lands, so that a third applicant is hired in one and
kills the offer in the other.  A step implementation agrees with the reference here only if it keeps the reference's
operation order for firm money."""
import numpy as np

from fastace_b200 import _abi


def find_ties(count, seed=0):
    """(wage_action f32, price f32, firm money f64, hired) tuples.  Person A is hired and buys one unit, B is hired,
    then C applies: the reference's money ((M - w) + p) - w is >= w (C is hired) while the reordered sum (M - 2w) + p
    is < w (C would be refused and the offer killed).  Every second tuple is the control one ulp-step below, where the
    reference itself refuses C."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        wa = np.float32(rng.uniform(0.3, 3.0))          # f_job_wage action; the offer's wage is wa / 0.5
        w = float(wa) / 0.5
        pr = np.float32(rng.uniform(0.2, 1.5) * w)      # p < 2w: the firm can pay A and B
        p = float(pr)
        m = 3 * w - p
        for _ in range(8):
            m = float(np.nextafter(m, -np.inf))
        for _ in range(16):
            seq = ((m - w) + p) - w
            alt = (m - 2 * w) + p
            if seq >= w and alt < w:
                out.append((wa, pr, m, True))
                lo = m
                while ((lo - w) + p) - w >= w:
                    lo = float(np.nextafter(lo, -np.inf))
                out.append((wa, pr, lo, False))
                break
            m = float(np.nextafter(m, np.inf))
    return out[:count]


def build(E, seed=0):
    """dims, initial state and the two steps' actions: step 0 posts the offers (empty markets), step 1 trades."""
    dims = (E, 3, 1, 1, 2)
    ties = find_ties(E, seed)
    st = _abi.alloc_host("state", dims)
    st["p_money"][:] = 100.0
    st["p_inv"][:] = 1.0
    st["p_util_tfp"][:] = 1.0
    st["p_util_share"][:] = 0.5
    st["p_util_rho"][:] = -0.1
    st["f_inv"][:] = 50.0
    st["f_prod_tfp"][:] = 1.0
    st["f_prod_share"][:] = 0.5
    st["f_prod_rho"][:] = -0.1
    for e, (wa, pr, m, hired) in enumerate(ties):
        st["f_money"][e, 0] = m

    def actions(step):
        a = {n: np.zeros(shp, dtype=dt) for n, (dt, shp) in _abi.shapes("actions", dims).items()}
        a["perm_person"][:] = np.arange(3, dtype=np.int32)
        a["perm_firm"][:] = 0
        a["f_offer_amt"][:] = 0.5          # lots = (int)(0.5 * inventory) > 0
        a["f_job_labor"][:] = 5.0          # 10 lots
        for e, (wa, pr, m, hired) in enumerate(ties):
            a["f_job_wage"][e, 0] = wa
            a["f_offer_price"][e, 0, 0] = pr
        if step == 1:
            a["p_job_take"][:, 0, :] = 1   # every person applies once (slot 0, entry 0) ...
            a["p_good_take"][:, 0, 0] = 1  # ... and the first one also buys one unit
        return a

    return dims, st, [actions(0), actions(1)], ties
