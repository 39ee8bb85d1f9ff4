"""Function plugins beyond CES: Linear, CobbDouglas(CRS), StoneGeary, Leontief as
selectable utility / production families.  The oracle is compared bit-for-bit with the reference's own
classes (live, when oracle/_ref is present, and through committed known answers); the CUDA path is compared
with the oracle."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests import golden_util as GU
from tests import helpers as H

KINDS = {"ces": _abi.FN_CES, "cobb_douglas": _abi.FN_COBB_DOUGLAS, "stone_geary": _abi.FN_STONE_GEARY,
         "leontief": _abi.FN_LEONTIEF, "linear": _abi.FN_LINEAR}

# known answers produced by the reference classes (vecToScalar.cpp) through oracle/_ref:
# (kind, tfp, share, theta, x, float.hex of VecToScalar::f)
KAT = [
    ("linear", 1.0, (0.5, 1.5, 2.0), None, (1.0, 2.0, 3.0), (0.5 * 1.0 + 1.5 * 2.0 + 2.0 * 3.0).hex()),
    ("leontief", 1.0, (0.5, 1.5, 2.0), None, (4.0, 1.0, 3.0), (1.5).hex()),
    ("cobb_douglas", 1.0, (0.5, 0.5), None, (4.0, 9.0), "0x1.8000000000000p+2"),
    ("cobb_douglas", 2.0, (0.3, 0.3, 0.4), None, (1.5, 2.5, 3.5), "0x1.3a15a4c88c403p+2"),
]


@pytest.mark.parametrize("kind,tfp,share,theta,x,want", KAT)
def test_known_answers(oracle, kind, tfp, share, theta, x, want):
    assert oracle.function_f(KINDS[kind], tfp, share, theta, 0.0, x).hex() == want


def _state(dims, seed, with_theta):
    st = scenario.generic_initial_state(dims, seed)
    E, P, F, G, S = dims
    if with_theta:
        rng = np.random.default_rng(seed)
        st["p_util_theta"][:] = rng.uniform(0.0, 0.05, st["p_util_theta"].shape)
        st["f_prod_theta"][:] = rng.uniform(0.0, 0.05, st["f_prod_theta"].shape)
    return st


@pytest.mark.parametrize("util,prod", [("cobb_douglas", "cobb_douglas"), ("leontief", "linear"), ("linear", "leontief"),
                                        ("stone_geary", "stone_geary"), ("ces", "cobb_douglas")])
def test_oracle_matches_reference_classes(oracle, util, prod):
    loader = pytest.importorskip("oracle.loader")
    if not loader.have_reference():
        pytest.skip("oracle/_ref not built here")
    dims = (2, 24, 5, 2, 6)
    st = _state(dims, 11, "stone" in util or "stone" in prod)
    ref = loader.Reference(dims, st, seed=3, util_kind=KINDS[util], prod_kind=KINDS[prod])
    oracle.set_function_kinds(KINDS[util], KINDS[prod])
    try:
        ost = {k: v.copy() for k, v in st.items()}
        for t in range(12):
            act = scenario.synthetic_actions(dims, seed=5, step=t, **scenario.BENCH_PRESET)
            rout = _abi.alloc_host("out", dims, names=("p_reward", "f_profit"))
            oout = _abi.alloc_host("out", dims, names=("p_reward", "f_profit"))
            pp, pf = ref.step(act, rout, flags=_abi.IDX_MODULO)
            act["perm_person"], act["perm_firm"] = pp, pf
            oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
            rst, _ = ref.get_state()
            for k in ("p_money", "p_inv", "f_money", "f_inv", "m_count", "j_count"):
                assert GU.bits_equal(rst[k], ost[k]), (t, k)
            # NaN rewards (negative base of a fractional power) must agree as NaN
            assert np.array_equal(rout["p_reward"], oout["p_reward"], equal_nan=True), t
    finally:
        oracle.set_function_kinds(_abi.FN_CES, _abi.FN_CES)
        ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("util,prod", [("cobb_douglas", "cobb_douglas"), ("leontief", "linear"), ("linear", "leontief"),
                                        ("stone_geary", "stone_geary"), ("ces", "ces")])
@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL])
def test_cuda_matches_oracle(oracle, util, prod, mode):
    from fastace_b200.env import BatchedEconomy
    dims = (12, 50, 8, 2, 8)
    st = _state(dims, 21, "stone" in util or "stone" in prod)
    env = BatchedEconomy(dims)
    env.set_function_kinds(KINDS[util], KINDS[prod])
    env.set_state(st)
    oracle.set_function_kinds(KINDS[util], KINDS[prod])
    try:
        ost = H.copy_state(st)
        orders = scenario.OrderStream(dims, 9)
        for t in range(12):
            act = scenario.synthetic_actions(dims, seed=8, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
            before = H.copy_state(ost)
            oout = _abi.alloc_host("out", dims)
            oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
            gout = _abi.alloc_host("out", dims)
            env.time_step_host(act, gout, flags=_abi.IDX_MODULO | mode)
            H.compare_outputs(gout, oout, dims, before)
            H.compare_states(env.get_state(), ost, dims)
    finally:
        oracle.set_function_kinds(_abi.FN_CES, _abi.FN_CES)
        env.close()
