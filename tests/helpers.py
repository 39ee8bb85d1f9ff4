"""Shared comparison helpers for the parity tests."""
import numpy as np

from fastace_b200 import _abi

# north_star tolerance: integer / index / counter work bit-exact, floating point 1e-5 relative
RTOL = 1e-5

EXACT_STATE = ["m_count", "j_count"]
EXACT_BOOK_M = ["m_owner", "m_good", "m_left", "m_taken"]
EXACT_BOOK_J = ["j_owner", "j_left", "j_taken"]
FLOAT_STATE = ["p_money", "p_inv", "p_labor", "f_money", "f_inv", "f_labor", "f_last_money"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    denom = np.maximum(np.abs(b), 1e-300)
    err = np.abs(a - b) / np.maximum(denom, 1e-12)
    err = np.where(both_nan | same_inf, 0.0, err)
    # absolute floor for values that are ~0 in both
    err = np.where((np.abs(a) < 1e-12) & (np.abs(b) < 1e-12), 0.0, err)
    return err


def assert_close(name, a, b, rtol=RTOL):
    err = rel_err(a, b)
    assert not np.isnan(err).any(), f"{name}: NaN mismatch"
    m = float(err.max()) if err.size else 0.0
    assert m <= rtol, f"{name}: max relative error {m:.3e} > {rtol}"
    return m


def compare_states(got, want, dims, rtol=RTOL):
    """got / want: host state dicts.  Books are compared over their live prefix."""
    E, P, F, G, S = dims
    worst = 0.0
    for k in EXACT_STATE:
        assert np.array_equal(got[k], want[k]), f"{k} differs: {got[k][:8]} vs {want[k][:8]}"
    for e in range(E):
        n = int(want["m_count"][e])
        for k in EXACT_BOOK_M:
            assert np.array_equal(got[k][e, :n], want[k][e, :n]), f"{k}[{e}] differs: {got[k][e,:n]} vs {want[k][e,:n]}"
        worst = max(worst, assert_close(f"m_price[{e}]", got["m_price"][e, :n], want["m_price"][e, :n], rtol))
        n = int(want["j_count"][e])
        for k in EXACT_BOOK_J:
            assert np.array_equal(got[k][e, :n], want[k][e, :n]), f"{k}[{e}] differs"
        worst = max(worst, assert_close(f"j_wage[{e}]", got["j_wage"][e, :n], want["j_wage"][e, :n], rtol))
    for k in FLOAT_STATE:
        worst = max(worst, assert_close(k, got[k], want[k], rtol))
    return worst


def compare_outputs(got, want, dims, state_before, rtol=RTOL):
    """Step outputs: success flags and old-book counters exact, rewards/profits 1e-5."""
    E, P, F, G, S = dims
    worst = 0.0
    for k in ("p_job_ok", "p_good_ok", "f_good_ok"):
        if k in got and k in want:
            assert np.array_equal(got[k], want[k]), f"{k} differs ({(got[k] != want[k]).sum()} flags)"
    for e in range(E):
        nm = int(state_before["m_count"][e])
        nj = int(state_before["j_count"][e])
        for k, n in (("old_m_left", nm), ("old_m_taken", nm), ("old_j_left", nj), ("old_j_taken", nj)):
            if k in got and k in want:
                assert np.array_equal(got[k][e, :n], want[k][e, :n]), f"{k}[{e}] differs: {got[k][e,:n]} vs {want[k][e,:n]}"
    worst = max(worst, assert_close("p_reward", got["p_reward"], want["p_reward"], rtol))
    worst = max(worst, assert_close("f_profit", got["f_profit"], want["f_profit"], rtol))
    return worst


def compare_all_fast(got_state, want_state, got_out, want_out, dims, state_before, rtol=RTOL):
    """compare_states + compare_outputs without per-economy Python loops (full-size runs): books over their live
    prefix through masks; integers exact, floating point `rtol`; money / labour bit-exact."""
    E, P, F, G, S = dims
    for k in EXACT_STATE:
        assert np.array_equal(got_state[k], want_state[k]), k
    live_m = np.arange(F * G)[None, :] < want_state["m_count"][:, None]
    live_j = np.arange(F)[None, :] < want_state["j_count"][:, None]
    for k in EXACT_BOOK_M:
        assert np.array_equal(got_state[k][live_m], want_state[k][live_m]), k
    for k in EXACT_BOOK_J:
        assert np.array_equal(got_state[k][live_j], want_state[k][live_j]), k
    assert_close("m_price", got_state["m_price"][live_m], want_state["m_price"][live_m], rtol)
    assert_close("j_wage", got_state["j_wage"][live_j], want_state["j_wage"][live_j], rtol)
    for k in FLOAT_STATE:
        assert_close(k, got_state[k], want_state[k], rtol)
    for k in ("p_money", "f_money", "p_labor", "f_last_money"):
        assert np.array_equal(got_state[k], want_state[k], equal_nan=True), k + " is not bit-identical"
    for k in ("p_job_ok", "p_good_ok", "f_good_ok"):
        assert np.array_equal(got_out[k], want_out[k]), k
    old_m = np.arange(F * G)[None, :] < state_before["m_count"][:, None]
    old_j = np.arange(F)[None, :] < state_before["j_count"][:, None]
    for k, m in (("old_m_left", old_m), ("old_m_taken", old_m), ("old_j_left", old_j), ("old_j_taken", old_j)):
        assert np.array_equal(got_out[k][m], want_out[k][m]), k
    assert_close("p_reward", got_out["p_reward"], want_out["p_reward"], rtol)
    assert_close("f_profit", got_out["f_profit"], want_out["f_profit"], rtol)


def copy_state(state):
    return {k: v.copy() for k, v in state.items()}


ALL_OUT = [n for n, _, _, _ in _abi.OUT_FIELDS]
