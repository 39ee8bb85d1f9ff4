"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container, where /root/reference exists:
    python tests/golden/gen_golden.py
Each fixture holds the initial state, every step's injected actions together with the
visiting orders the reference's own std::shuffle produced, and the reference's state and
outputs after every step.  The fixtures pin the C oracle and the CUDA path on machines
where the reference itself is not available (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fastace_b200 import _abi, scenario  # noqa: E402
from oracle.loader import Reference, build  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_OUT = [n for n, _, _, _ in _abi.OUT_FIELDS if n not in ("p_job_ok", "p_good_ok", "f_good_ok")]


def simple_scenario_state():
    """SimpleScenario::setup (src/neural/neuralScenarios.cpp:37-74): 2 persons + 1 firm."""
    from oracle.loader import Oracle
    orc = Oracle()
    dims = (1, 2, 1, 2, 10)
    st = _abi.alloc_host("state", dims)
    st["p_money"][0] = [20.0, 20.0]
    st["p_inv"][0, :, 0] = [10.0, 10.0]
    st["p_inv"][0, :, 1] = [10.0, 10.0]
    st["p_util_tfp"][0] = 1.0
    for p, (share, el) in enumerate([((0.5, 0.5, 0.5), 1.3), ((0.2, 0.6, 0.4), 1.3)]):
        s, rho = orc.ces_params(share, el)
        st["p_util_share"][0, :, p] = s
        st["p_util_rho"][0, p] = rho
    st["f_money"][0, 0] = 50.0
    st["f_inv"][0, :, 0] = [10.0, 20.0]
    for g, (tfp, share, el) in enumerate([(0.5, (1.0, 0.0, 1.0), 3.0), (1.0, (1.0, 0.0, 1.0), 5.0)]):
        s, rho = orc.ces_params(share, el)
        st["f_prod_tfp"][0, g, 0] = tfp
        st["f_prod_share"][0, g, :, 0] = s
        st["f_prod_rho"][0, g, 0] = rho
    return dims, st


def record(name, dims, state, steps, seed, flags, preset, mutate=None):
    ref = Reference(dims, state, seed=seed)
    data = {"dims": np.array(dims, dtype=np.int32), "flags": np.array([flags], dtype=np.uint32),
            "steps": np.array([steps], dtype=np.int32)}
    for k, v in state.items():
        data[f"init/{k}"] = v
    for t in range(steps):
        act = scenario.synthetic_actions(dims, seed=seed + 1, step=t, **preset)
        if mutate:
            mutate(t, act, dims)
        out = _abi.alloc_host("out", dims, names=ALL_OUT)
        pp, pf = ref.step(act, out, flags=flags)
        act["perm_person"], act["perm_firm"] = pp, pf
        st, tt = ref.get_state()
        assert tt == t + 1
        for k, v in act.items():
            data[f"act{t}/{k}"] = v
        for k, v in out.items():
            data[f"out{t}/{k}"] = v
        for k, v in st.items():
            if not k.startswith(("p_util", "f_prod")):
                data[f"state{t}/{k}"] = v
    ref.close()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    print(name, os.path.getsize(path) // 1024, "KiB")


def absolute_small_indices(t, act, dims):
    E, P, F, G, S = dims
    rng = np.random.default_rng(1000 + t)
    for k, hi in (("p_job_idx", F + 2), ("p_good_idx", F * G + 2), ("f_good_idx", F * G + 2)):
        act[k] = rng.integers(-1, hi, act[k].shape, dtype=np.int32)


def stress(t, act, dims):
    if t % 4 == 1:
        act["f_job_wage"][:, 0] = 3e9      # 1e8 clip
        act["f_job_labor"][:, 1] = 5e9     # (int) overflow -> no offer
    if t % 4 == 2:
        act["p_job_take"][:] = 1
        act["p_good_take"][:] = 1
        act["f_good_take"][:] = 1
    if t % 4 == 3:
        act["f_offer_price"][:] *= 0.05    # cheap goods: inventories, not money, bind


def main():
    build()
    dims, st = simple_scenario_state()
    record("simple_scenario", dims, st, steps=30, seed=5, flags=_abi.IDX_MODULO, preset=dict(labor_mu=0.5))
    dims = (2, 48, 12, 2, 10)     # config A: py/train.py defaults
    record("config_a", dims, scenario.custom_initial_state(dims, 101)[0], steps=40, seed=7,
           flags=_abi.IDX_MODULO, preset={})
    dims = (2, 100, 10, 2, 10)    # config B shape
    record("config_b", dims, scenario.custom_initial_state(dims, 202)[0], steps=40, seed=9,
           flags=_abi.IDX_MODULO, preset=scenario.BENCH_PRESET)
    dims = (2, 30, 5, 2, 10)
    record("absolute_idx", dims, scenario.custom_initial_state(dims, 303)[0], steps=16, seed=11,
           flags=_abi.IDX_ABSOLUTE, preset=scenario.BENCH_PRESET, mutate=absolute_small_indices)
    dims = (2, 40, 6, 2, 10)
    record("stress", dims, scenario.custom_initial_state(dims, 404)[0], steps=24, seed=13,
           flags=_abi.IDX_MODULO, preset=scenario.BENCH_PRESET, mutate=stress)
    dims = (1, 37, 7, 3, 6)       # 3 goods, ragged sizes
    record("three_goods", dims, scenario.generic_initial_state(dims, 505), steps=20, seed=15,
           flags=_abi.IDX_MODULO, preset=scenario.BENCH_PRESET)
    dims = (1, 64, 31, 8, 4)      # config D's 8 goods in miniature (F*G = 248 book entries)
    record("eight_goods", dims, scenario.generic_initial_state(dims, 606), steps=10, seed=17,
           flags=_abi.IDX_MODULO, preset=scenario.BENCH_PRESET)
    dims = (1, 50, 40, 3, 6)      # more firms than lanes in a warp
    record("many_firms", dims, scenario.generic_initial_state(dims, 707), steps=10, seed=19,
           flags=_abi.IDX_MODULO, preset=scenario.BENCH_PRESET)


if __name__ == "__main__":
    main()
