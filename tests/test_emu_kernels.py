"""The step kernels' CUDA SOURCE against the oracle — without a GPU.

tests/emu compiles fastace_b200/csrc/match_kernel.cuh and update_kernel for the CPU under a SIMT emulator (one fiber
per CUDA thread, warp collectives with their real semantics) and drives them as launch_step does.  This is the check
of the kernel LOGIC that can run in the GPU-less container; the `-m gpu` tests check the compiled sm_100a code the
same way on the device.  Matching, counters, market order: bit-exact.  Money, labour: bit-exact (the kernel keeps the
reference's fp64 operation order).  Inventories / rewards / profits: 1e-5 (pow)."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests import helpers as H

EXACT_FLOAT = ("p_money", "f_money", "p_labor", "f_last_money")


@pytest.fixture(scope="module")
def emu():
    from tests.emu.loader import EmuKernels
    k = EmuKernels()
    yield k
    k.close()


def _episode(emu, oracle, dims, steps, seed, preset, flags=_abi.IDX_MODULO, compact=False, tweak=None):
    E, P, F, G, S = dims
    state = scenario.custom_initial_state(dims, seed)[0] if G == 2 else scenario.generic_initial_state(_abi.make_dims(*dims), seed)
    ost, est = H.copy_state(state), H.copy_state(state)
    orders = scenario.OrderStream(dims, seed + 17)
    emu.stats()
    for t in range(steps):
        act = scenario.synthetic_actions(dims, seed=seed + 1, step=t, perms=orders.next(), **preset)
        if tweak:
            tweak(t, act)
        before = H.copy_state(ost)
        oout, eout = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        if compact:
            cz = _abi.compact_actions_for_counts(act, before["j_count"], before["m_count"], bool(flags & _abi.IDX_MODULO))
            emu.step(dims, est, None, eout, flags=flags, time_before=t, compact=cz)
        else:
            emu.step(dims, est, act, eout, flags=flags, time_before=t)
        H.compare_outputs(eout, oout, dims, before)
        H.compare_states(est, ost, dims)
        for k in EXACT_FLOAT:
            assert np.array_equal(est[k], ost[k], equal_nan=True), (k, t)
    return emu.stats()


def test_config_b_shape(emu, oracle):
    st = _episode(emu, oracle, (3, 100, 10, 2, 10), 24, 11, scenario.BENCH_PRESET)
    assert st["sales_windows"] > 0 and st["rescans"] > 0      # the contended paths were exercised


def test_bankrupt_firms_kill_job_offers(emu, oracle):
    st = _episode(emu, oracle, (4, 100, 10, 2, 10), 20, 5, dict(labor_mu=1.0))
    assert st["risky_walks"] > 0


@pytest.mark.parametrize("dims", [(3, 1, 1, 1, 1), (4, 33, 1, 1, 3), (2, 100, 33, 2, 10), (3, 64, 40, 3, 16), (2, 31, 9, 8, 10),
                                  (2, 0, 3, 2, 4), (3, 7, 5, 5, 0), (2, 257, 12, 4, 7)])
def test_shapes(emu, oracle, dims):
    _episode(emu, oracle, dims, 8, dims[1] + 7 * dims[2], scenario.BENCH_PRESET)


@pytest.mark.parametrize("modulo", [True, False])
def test_compact_encoding(emu, oracle, modulo):
    dims = (3, 70, 8, 2, 10)
    rng = np.random.default_rng(5)

    def tweak(t, act):
        if not modulo:
            for k, hi in (("p_job_idx", 8 + 2), ("p_good_idx", 16 + 2), ("f_good_idx", 16 + 2)):
                act[k][...] = rng.integers(-1, hi, act[k].shape, dtype=np.int32)
    _episode(emu, oracle, dims, 10, 77, scenario.BENCH_PRESET, flags=_abi.IDX_MODULO if modulo else _abi.IDX_ABSOLUTE,
             compact=True, tweak=tweak)


def test_goods_rich_market_firms_buy(emu, oracle):
    """plenty of goods and cheap prices: persons and FIRMS buy all episode long (serial firm walk, sale folds)"""
    st = _episode(emu, oracle, (3, 60, 8, 2, 10), 12, 31, dict(take_prob=0.6, prod_scale=0.2, wage_scale=0.3, price_scale=0.05, labor_mu=1.5))
    assert st["firm_serial"] > 0 and st["sales_windows"] > 6


def test_extreme_actions(emu, oracle):
    dims = (4, 40, 5, 2, 10)

    def tweak(t, act):
        act["f_job_wage"][1] = 3e9
        act["f_job_labor"][2] = 3e9
        act["f_job_labor"][3, :2] = np.inf
        act["f_offer_amt"][0] = 1.0
        act["f_offer_price"][0] = 1e-3
        act["f_offer_price"][1, 0, :2] = np.inf
        act["f_offer_price"][1, 1, 2:4] = np.nan
        act["f_offer_price"][2, 0, :3] = -0.75          # a sale LOWERS the seller's money
        act["f_job_wage"][3, 2:4] = np.inf
        act["f_job_wage"][2, 1] = -0.25                 # a hire RAISES the firm's money
        act["p_consume"][1] = 0.0
        act["p_consume"][2] = 1.0
        if t % 3 == 0:
            act["p_job_take"][:] = 1; act["p_good_take"][:] = 1; act["f_good_take"][:] = 1
        if t % 3 == 1:
            act["p_job_take"][:] = 0; act["p_good_take"][:] = 0; act["f_good_take"][:] = 0
    _episode(emu, oracle, dims, 9, 33, scenario.BENCH_PRESET, tweak=tweak)


def test_phase_wise_calls_equal_one_call(emu, oracle):
    """PERSONS_TRADE + PERSONS_CONSUME + FIRMS == one call"""
    dims = (3, 50, 6, 2, 10)
    state = scenario.custom_initial_state(dims, 51)[0]
    one, two = H.copy_state(state), H.copy_state(state)
    orders = scenario.OrderStream(dims, 52)
    for t in range(6):
        act = scenario.synthetic_actions(dims, seed=53, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        o1, o2 = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        emu.step(dims, one, act, o1, flags=_abi.IDX_MODULO, time_before=t)
        emu.step(dims, two, act, {k: o2[k] for k in ("p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")},
                 flags=_abi.IDX_MODULO | _abi.STEP_PERSONS_TRADE, time_before=t)
        emu.step(dims, two, act, {"p_reward": o2["p_reward"]}, flags=_abi.IDX_MODULO | _abi.STEP_PERSONS_CONSUME, time_before=t)
        emu.step(dims, two, act, {k: o2[k] for k in ("f_profit", "f_good_ok", "old_m_left", "old_m_taken")},
                 flags=_abi.IDX_MODULO | _abi.STEP_FIRMS, time_before=t)
        for k in ("p_money", "p_inv", "p_labor", "f_money", "f_inv", "f_labor", "f_last_money", "m_count", "j_count"):
            assert np.array_equal(one[k], two[k], equal_nan=True), (k, t)
        for k in ("p_reward", "f_profit", "p_job_ok", "p_good_ok", "f_good_ok"):
            assert np.array_equal(o1[k], o2[k], equal_nan=True), (k, t)


def test_firm_money_near_tie(emu, oracle):
    """the third applicant's fate hangs on the last bit of the firm's running money (tests/near_tie.py)"""
    from tests import near_tie
    dims, state, acts, ties = near_tie.build(24, seed=3)
    ost, est = H.copy_state(state), H.copy_state(state)
    for t, act in enumerate(acts):
        before = H.copy_state(ost)
        oout, eout = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_ABSOLUTE, time_before=t)
        emu.step(dims, est, act, eout, flags=_abi.IDX_ABSOLUTE, time_before=t)
        H.compare_outputs(eout, oout, dims, before)
        H.compare_states(est, ost, dims)
        for k in EXACT_FLOAT:
            assert np.array_equal(est[k], ost[k], equal_nan=True), (k, t)
    hired = np.array([h for (_, _, _, h) in ties])
    assert np.array_equal(oout["p_job_ok"][:, 0, 2].astype(bool), hired)     # the reference decides as constructed ...
    assert hired.any() and (~hired).any()                                   # ... both ways
