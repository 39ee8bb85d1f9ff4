"""Large-economy path (csrc/large_economy.cuh, BASELINE config D) against the CPU oracle.
Same bar as the warp-per-economy kernels — matching, counters, market order bit-exact; fp within 1e-5 —
and, because every agent's fp64 state is updated inside one thread in the reference's order, money and
inventories are additionally checked for bit equality here."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests import helpers as H
from tests.test_gpu_parity import _run_episode

pytestmark = pytest.mark.gpu
LARGE = _abi.STEP_LARGE


@pytest.mark.parametrize("dims", [
    (3, 1, 1, 1, 1), (4, 33, 1, 1, 3), (2, 100, 33, 2, 10), (3, 64, 40, 3, 16), (2, 31, 9, 8, 10),
    (2, 0, 3, 2, 4), (3, 7, 5, 5, 0), (2, 257, 12, 4, 7), (5, 100, 10, 2, 10),
])
def test_forced_large_path_on_small_shapes(oracle, dims):
    """the same shape sweep as the warp kernels, through FASTACE_STEP_LARGE"""
    _run_episode(oracle, dims, 10, seed=dims[1] + 7 * dims[2], preset=scenario.BENCH_PRESET, flags=_abi.IDX_MODULO | LARGE)


def test_large_path_bankrupt_firms_and_host_api(oracle):
    # untuned recipe: firms run out of money, job offers are killed by the first applicant that finds money < wage
    _run_episode(oracle, (6, 100, 10, 2, 10), 30, seed=5, preset=dict(labor_mu=1.0), flags=_abi.IDX_MODULO | LARGE)
    _run_episode(oracle, (3, 37, 7, 2, 10), 12, seed=5, preset=scenario.BENCH_PRESET, flags=_abi.IDX_MODULO | LARGE, host_api=True)


def _episode_bitwise(oracle, dims, steps, seed, preset, flags=_abi.IDX_MODULO):
    """large-only dims: state after every step equals the oracle's bit for bit (books, counters, money, persons' inventories)"""
    from fastace_b200.env import BatchedEconomy
    E, P, F, G, S = dims
    state = scenario.generic_initial_state(dims, seed)
    env = BatchedEconomy(dims)
    env.set_state(state, time=0)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, seed + 17)
    rounds, trades = [], 0
    for t in range(steps):
        act = scenario.synthetic_actions(dims, seed=seed + 1, step=t, perms=orders.next(), **preset)
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        gout = _abi.alloc_host("out", dims)
        env.time_step_host(act, gout, flags=flags)
        rounds.append(env.large_stats())
        H.compare_outputs(gout, oout, dims, before)
        got = env.get_state()
        H.compare_states(got, ost, dims)
        # f_inv goes through pow() (production): 1e-5 like everywhere; everything else is +,-,* in the reference's order
        for k in ("p_money", "f_money", "p_inv", "p_labor", "f_last_money"):
            assert np.array_equal(got[k], ost[k], equal_nan=True), f"{k}: not bit-identical at step {t}"
        trades += int(oout["p_good_ok"].sum()) + int(oout["p_job_ok"].sum()) + int(oout["f_good_ok"].sum())
    env.close()
    return rounds, trades


def test_beyond_warp_kernel_limits(oracle):
    # F*G = 400 > 254: only the large path can take this env
    rounds, trades = _episode_bitwise(oracle, (2, 3000, 100, 4, 10), 8, seed=9, preset=scenario.BENCH_PRESET)
    assert trades > 1000
    print("rounds (person, firm) per step:", rounds)


def test_config_d_full_size(oracle):
    """BASELINE config D: one economy of 100 000 persons + 5 000 firms, 8 goods, stack 10"""
    rounds, trades = _episode_bitwise(oracle, (1, 100000, 5000, 8, 10), 4, seed=21, preset=scenario.BENCH_PRESET)
    assert trades > 100000
    print("rounds (person, firm) per step:", rounds, "transactions:", trades)


def test_large_path_extreme_actions(oracle):
    """inf / NaN prices and wages, zero and huge amounts: same comparisons as the reference's"""
    from fastace_b200.env import BatchedEconomy
    dims = (4, 60, 8, 3, 8)
    state = scenario.generic_initial_state(dims, 4)
    env = BatchedEconomy(dims)
    env.set_state(state, time=0)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 6)
    rng = np.random.default_rng(0)
    for t in range(8):
        act = scenario.synthetic_actions(dims, seed=3, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        pr, wg = act["f_offer_price"], act["f_job_wage"]
        pr[rng.random(pr.shape) < 0.15] = np.inf
        pr[rng.random(pr.shape) < 0.1] = np.nan
        pr[rng.random(pr.shape) < 0.1] = 0.0
        wg[rng.random(wg.shape) < 0.15] = np.inf
        wg[rng.random(wg.shape) < 0.1] = np.nan
        act["f_job_labor"][rng.random(wg.shape) < 0.1] = 1e30
        act["f_offer_amt"][rng.random(pr.shape) < 0.1] = 1e30
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
        gout = _abi.alloc_host("out", dims)
        env.time_step_host(act, gout, flags=_abi.IDX_MODULO | LARGE)
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)
    env.close()


@pytest.mark.parametrize("dims", [(1, 100, 10, 2, 10), (1, 3000, 100, 4, 10)])
def test_large_path_phase_wise_equals_one_call(oracle, dims):
    """TRADE + CONSUME + FIRMS (and PERSONS + FIRMS) on the large-economy path == one call, bit for bit"""
    from fastace_b200.env import BatchedEconomy
    state = scenario.generic_initial_state(dims, 7) if dims[3] != 2 else scenario.custom_initial_state(dims, 7)[0]
    one, two = BatchedEconomy(dims), BatchedEconomy(dims)
    one.set_state(state); two.set_state(state)
    orders = scenario.OrderStream(dims, 8)
    base = _abi.IDX_MODULO | LARGE
    for t in range(8):
        act = scenario.synthetic_actions(dims, seed=9, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        out1, out2 = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        one.time_step_host(act, out1, flags=base)
        person_flags = ("p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")
        if t % 2 == 0:
            two.time_step_host(act, {k: out2[k] for k in person_flags}, flags=base | _abi.STEP_PERSONS_TRADE)
            mid = two.get_state()
            assert np.array_equal(mid["p_labor"], 0.5 * out2["p_job_ok"].sum(axis=1))
            two.time_step_host(act, {"p_reward": out2["p_reward"]}, flags=base | _abi.STEP_PERSONS_CONSUME)
        else:
            two.time_step_host(act, {k: out2[k] for k in person_flags + ("p_reward",)}, flags=base | _abi.STEP_PERSONS)
        mid = two.get_state()
        assert np.array_equal(mid["p_money"], one.get_state()["p_money"]) and two.get_time() == t
        two.time_step_host(act, {k: out2[k] for k in ("f_profit", "f_good_ok", "old_m_left", "old_m_taken")}, flags=base | _abi.STEP_FIRMS)
        s1, s2 = one.get_state(), two.get_state()
        for k in s1:
            assert np.array_equal(s1[k], s2[k], equal_nan=s1[k].dtype.kind == "f"), (k, t)
        for k in ("p_reward", "f_profit", "p_job_ok", "p_good_ok", "f_good_ok"):
            assert np.array_equal(out1[k], out2[k], equal_nan=out1[k].dtype.kind == "f"), (k, t)
    one.close(); two.close()
