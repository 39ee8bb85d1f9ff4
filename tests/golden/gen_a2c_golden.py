"""Generates tests/golden/a2c_episode.npz from the reference's UNMODIFIED trainer and decision-net handler
(oracle/_ref/libfastace_refa2c.so, built by `make -C oracle refa2c`): initial parameters of the 11 nets, the
injected agent states and rewards of a few short episodes, the random draws the reference consumed (replayed
from the shared torch CPU generator), and what the reference produced — recorded log-probabilities and values,
episode losses, gradients, parameters after the Adam steps and the learning rates after every episode."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(ROOT, "oracle", "_ref", "libfastace_refa2c.so")

dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
CFG = dict(numGoods=2, stackSize=5, encodingSize=4, hiddenSize=16, nHidden=3, nHiddenSmall=2)
NETS = ("purchaseNet", "firmPurchaseNet", "laborSearchNet", "consumptionNet", "productionNet", "offerNet",
        "jobOfferNet", "valueNet", "firmValueNet")
KINDS = ("logp_purchase", "logp_firmPurchase", "logp_laborSearch", "logp_consumption", "logp_production",
         "logp_offer", "logp_jobOffer", "value", "reward")


def d(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)


def load():
    L = C.CDLL(LIB)
    L.refa2c_create.restype = C.c_void_p
    L.refa2c_create.argtypes = [C.c_int] * 8 + [C.c_double, dp, dp, dp, C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_uint64]
    L.refa2c_destroy.argtypes = [C.c_void_p]
    L.refa2c_post_market.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), dp, C.c_int, dp]
    L.refa2c_reset.argtypes = [C.c_void_p]
    L.refa2c_step.argtypes = [C.c_void_p] + [dp] * 8
    L.refa2c_record.restype = C.c_double
    L.refa2c_record.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.refa2c_train_on_episode.restype = C.c_double
    L.refa2c_train_on_episode.argtypes = [C.c_void_p]
    L.refa2c_lrs.argtypes = [C.c_void_p, dp]
    L.refa2c_num_params.argtypes = [C.c_void_p]
    L.refa2c_param_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    L.refa2c_param_data.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
    return L


def export_params(L, h, grad=0):
    out = {}
    shape = (C.c_int64 * 2)()
    name = C.create_string_buffer(256)
    for i in range(L.refa2c_num_params(h)):
        n = L.refa2c_param_info(h, i, name, 256, shape)
        buf = np.zeros(n, dtype=np.float32)
        L.refa2c_param_data(h, i, grad, buf.ctypes.data_as(fp))
        out[name.value.decode()] = buf.reshape((shape[0], shape[1]) if shape[1] else (shape[0],))
    return out


def replay_draws(T, P, F, G, S, nM, nJ):
    """The reference's torch RNG call sequence for T steps (refa2c_step order), from the current generator state."""
    z = lambda *s: torch.zeros(*s)
    out = []
    for _ in range(T):
        dr = dict(pidxM=z(P, S).long(), pidxJ=z(P, S).long(), fidxM=z(F, S).long(), fidxJ=z(F, S).long(),
                  u_job=z(P, S), u_good=z(P, S), u_fgood=z(F, S), n_cons=z(P, G), n_prod=z(F, G), n_amt=z(F, G),
                  n_price=z(F, G), n_lab=z(F), n_wage=z(F))
        for p in range(P):
            if nM: dr["pidxM"][p] = torch.randint(0, nM, (S,))      # generate_offerIndices
            if nJ: dr["pidxJ"][p] = torch.randint(0, nJ, (S,))      # generate_jobOfferIndices
            if nJ: dr["u_job"][p] = torch.rand(S)                   # create_joboffer_requests
            if nM: dr["u_good"][p] = torch.rand(S)                  # create_offer_requests
            dr["n_cons"][p] = torch.randn(G)                        # sample_logitNormal(consumptionNet)
        for f in range(F):
            if nM: dr["fidxM"][f] = torch.randint(0, nM, (S,))
            if nJ: dr["fidxJ"][f] = torch.randint(0, nJ, (S,))
            if nM: dr["u_fgood"][f] = torch.rand(S)
            dr["n_prod"][f] = torch.randn(G)
            dr["n_amt"][f] = torch.randn(G)
            dr["n_price"][f] = torch.randn(G)
            dr["n_lab"][f] = torch.randn(1)[0]
            dr["n_wage"][f] = torch.randn(1)[0]
        out.append({k: v.unsqueeze(0).numpy() for k, v in dr.items()})   # leading economy axis
    return out


def run_reference(L, cfg, P, F, T, episodes, seed, market=True, lr=1e-3, sched=(1, 2, 0.5, 2), discount=0.9):
    """-> dict of numpy arrays (everything the parity test needs)"""
    G, S = cfg["numGoods"], cfg["stackSize"]
    rng = np.random.default_rng(seed)
    U = G + 3
    up = rng.uniform(0.2, 1.0, (P, U)); pp = rng.uniform(0.2, 1.0, (F, G, U))
    lrs = np.full(9, lr) * rng.uniform(0.5, 2.0, 9)
    h = L.refa2c_create(G, S, cfg["encodingSize"], cfg["hiddenSize"], cfg["nHidden"], cfg["nHiddenSmall"], P, F, discount,
                        d(up), d(pp), d(lrs), sched[0], sched[1], sched[2], sched[3], seed)
    nM, nJ = (min(3, F * G), min(2, F)) if market else (0, 0)
    good = rng.integers(0, G, nM).astype(np.int32); price = rng.uniform(0.5, 3.0, nM); wage = rng.uniform(0.5, 3.0, nJ)
    if market:
        L.refa2c_post_market(h, nM, good.ctypes.data_as(C.POINTER(C.c_int)), d(price), nJ, d(wage))
    res = dict(P=P, F=F, T=T, episodes=episodes, nM=nM, nJ=nJ, good=good, price=price, wage=wage, util_params=up,
               prod_params=pp, lrs=lrs, sched=np.array(sched, dtype=np.float64), discount=discount)
    for k, v in export_params(L, h).items():
        res["init/" + k] = v
    for ep in range(episodes):
        st = dict(p_money=rng.uniform(1, 9, (T, P)), p_labor=rng.uniform(0, 1, (T, P)), p_inv=rng.uniform(0, 5, (T, P, G)),
                  p_reward=rng.uniform(0, 3, (T, P)), f_money=rng.uniform(1, 9, (T, F)), f_labor=rng.uniform(0, 2, (T, F)),
                  f_inv=rng.uniform(0, 5, (T, F, G)), f_profit=rng.uniform(-1, 1, (T, F)))
        torch.manual_seed(seed * 100 + ep)
        L.refa2c_reset(h)
        for t in range(T):
            L.refa2c_step(h, *(d(st[k][t]) for k in ("p_money", "p_labor", "p_inv", "p_reward", "f_money", "f_labor", "f_inv", "f_profit")))
        torch.manual_seed(seed * 100 + ep)
        draws = replay_draws(T, P, F, G, S, nM, nJ)
        for k, v in st.items():
            res[f"ep{ep}/state/{k}"] = v
        for t in range(T):
            for k, v in draws[t].items():
                res[f"ep{ep}/draws/{t}/{k}"] = v
        rec = np.array([[[L.refa2c_record(h, kind, t, a) for a in range(P + F)] for t in range(T)] for kind in range(9)])
        res[f"ep{ep}/records"] = rec      # [kind][t][agent]; -12345 = not recorded
        res[f"ep{ep}/loss"] = L.refa2c_train_on_episode(h)
        for k, v in export_params(L, h, grad=1).items():
            res[f"ep{ep}/grad/{k}"] = v
        for k, v in export_params(L, h).items():
            res[f"ep{ep}/param/{k}"] = v
        out = np.zeros(9); L.refa2c_lrs(h, d(out)); res[f"ep{ep}/lrs"] = out
    L.refa2c_destroy(h)
    return res


if __name__ == "__main__":
    L = load()
    a = run_reference(L, CFG, P=3, F=2, T=4, episodes=3, seed=11)
    b = run_reference(L, CFG, P=2, F=2, T=3, episodes=1, seed=12, market=False)
    out = {("market/" + k): v for k, v in a.items()}
    out.update({("empty/" + k): v for k, v in b.items()})
    for k, v in CFG.items():
        out["cfg/" + k] = np.array([v])
    np.savez_compressed(os.path.join(HERE, "a2c_episode.npz"), **out)
    print("wrote a2c_episode.npz", sum(np.asarray(v).nbytes for v in out.values()), "bytes")
