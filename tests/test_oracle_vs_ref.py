"""Live comparison of the C oracle with the unmodified reference (oracle/_ref) on fresh
random inputs.  Skipped where the reference library is not present; the committed golden
vectors (test_oracle_golden.py) cover that case."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests import golden_util as GU

loader = pytest.importorskip("oracle.loader")
pytestmark = pytest.mark.skipif(not loader.have_reference(), reason="oracle/_ref not built here")

OUTS = [n for n, _, _, _ in _abi.OUT_FIELDS if n not in ("p_job_ok", "p_good_ok", "f_good_ok")]


def _compare(oracle, dims, steps, seed, flags, preset, mutate=None, generic=False, outs=OUTS):
    state = scenario.generic_initial_state(dims, seed) if generic else scenario.custom_initial_state(dims, seed)[0]
    ref = loader.Reference(dims, state, seed=seed + 5)
    ost = {k: v.copy() for k, v in state.items()}
    for t in range(steps):
        act = scenario.synthetic_actions(dims, seed=seed + 1, step=t, **preset)
        if mutate:
            mutate(t, act)
        rout = _abi.alloc_host("out", dims, names=outs)
        oout = _abi.alloc_host("out", dims, names=outs)
        pp, pf = ref.step(act, rout, flags=flags)
        act["perm_person"], act["perm_firm"] = pp, pf
        before = {k: v.copy() for k, v in ost.items()}
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        rst, rt = ref.get_state()
        assert rt == t + 1
        GU.assert_out_bits(oout, rout, dims, before, where=f"step {t}")
        GU.assert_state_bits(ost, rst, dims, where=f"step {t}")
    ref.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_default_scenario(oracle, seed):
    _compare(oracle, (3, 48, 12, 2, 10), 40, seed, _abi.IDX_MODULO, {})


def test_bench_workload(oracle):
    _compare(oracle, (4, 100, 10, 2, 10), 40, 77, _abi.IDX_MODULO, scenario.BENCH_PRESET)


def test_bankrupt_firms_and_dead_markets(oracle):
    # untuned recipe: firms run out of money -> job offers are killed, the goods market empties
    _compare(oracle, (3, 100, 10, 2, 10), 30, 5, _abi.IDX_MODULO, dict(labor_mu=1.0))


@pytest.mark.parametrize("dims", [(2, 1, 1, 2, 3), (2, 17, 3, 1, 5), (1, 40, 35, 4, 16), (1, 25, 6, 8, 10), (2, 9, 2, 2, 0)])
def test_shapes(oracle, dims):
    _compare(oracle, dims, 12, dims[1], _abi.IDX_MODULO, scenario.BENCH_PRESET, generic=True)


def test_absolute_indices(oracle):
    rng = np.random.default_rng(3)

    def mutate(t, act):
        for k, hi in (("p_job_idx", 9), ("p_good_idx", 16), ("f_good_idx", 16)):
            act[k] = rng.integers(-2, hi, act[k].shape, dtype=np.int32)

    _compare(oracle, (3, 30, 6, 2, 10), 15, 9, _abi.IDX_ABSOLUTE, scenario.BENCH_PRESET, mutate)


def test_extreme_actions(oracle):
    def mutate(t, act):
        act["f_job_wage"][0] = 3e9
        act["f_job_labor"][1] = 3e9
        act["f_job_labor"][2, :1] = np.inf
        act["f_offer_price"][0] = 1e-3
        act["p_consume"][1] = 0.0
        act["p_consume"][2] = 1.0
        act["f_offer_amt"][1] = 1.0
        if t % 3 == 0:
            act["p_job_take"][:] = 1; act["p_good_take"][:] = 1; act["f_good_take"][:] = 1

    _compare(oracle, (3, 40, 5, 2, 10), 15, 4, _abi.IDX_MODULO, scenario.BENCH_PRESET, mutate)


def test_config_d_full_size(oracle):
    """BASELINE config D at full size (one economy of 100 000 persons + 5 000 firms, 8 goods): the oracle that the
    large-economy GPU tests are checked against is itself bit-identical to the reference here, 3 steps (the first
    trades against empty books, the next two move ~800 k requests each)"""
    # (the harness samples the old books' counters at every decision, O(agents x book): state, rewards and profits only)
    _compare(oracle, (1, 100000, 5000, 8, 10), 3, 21, _abi.IDX_MODULO, scenario.BENCH_PRESET, generic=True,
             outs=("p_reward", "f_profit"))


def test_negative_prices_and_wages(oracle):
    """a sale that lowers the seller's money, a hire that raises the firm's: the cases the kernels' "can the firm pay"
    shortcut must not take"""
    def mutate(t, act):
        act["f_offer_price"][0, 0, :3] = -0.75
        act["f_job_wage"][1, 1] = -0.25
        act["f_offer_price"][2, 1, 2:4] = np.nan
    _compare(oracle, (3, 40, 5, 2, 10), 12, 33, _abi.IDX_MODULO, scenario.BENCH_PRESET, mutate)
