"""The C oracle against the golden vectors generated from the UNMODIFIED reference
(tests/golden/gen_golden.py).  Every double must be bit-identical: same compiler family,
same libm pow, same operation order."""
import numpy as np
import pytest

from fastace_b200 import _abi
from tests import golden_util as GU


@pytest.mark.parametrize("name", GU.NAMES)
def test_oracle_reproduces_reference_bits(oracle, name):
    g = GU.Golden(name)
    st = g.initial_state()
    for t in range(g.steps):
        before = {k: v.copy() for k, v in st.items()}
        out = _abi.alloc_host("out", g.dims)
        oracle.step(g.dims, st, g.actions(t), out, flags=g.flags, time_before=t)
        GU.assert_out_bits(out, g.outputs(t), g.dims, before, where=f"{name} step {t}")
        GU.assert_state_bits(st, g.state(t), g.dims, where=f"{name} step {t}")


def test_golden_fixtures_exercise_the_market():
    """the fixtures are not trivial: trades, hires, sold-out and killed offers all occur"""
    g = GU.Golden("config_b")
    taken_m = sum(int(g.outputs(t)["old_m_taken"].sum()) for t in range(g.steps))
    taken_j = sum(int(g.outputs(t)["old_j_taken"].sum()) for t in range(g.steps))
    assert taken_m > 500 and taken_j > 5000
    s = GU.Golden("stress")
    dead = 0
    for t in range(1, s.steps):
        n = s.state(t - 1)["j_count"]
        for e in range(s.dims[0]):
            dead += int((s.outputs(t)["old_j_left"][e, : n[e]] == 0).sum())
    assert dead > 0


def test_money_is_conserved_in_golden_runs():
    """invariant of the single-threaded reference: total money is constant"""
    for name in ("config_a", "config_b"):
        g = GU.Golden(name)
        init = g.initial_state()
        total0 = init["p_money"].sum(axis=1) + init["f_money"].sum(axis=1)
        for t in range(g.steps):
            s = g.state(t)
            total = s["p_money"].sum(axis=1) + s["f_money"].sum(axis=1)
            assert np.allclose(total, total0, rtol=1e-12, atol=0)
