"""The driver's smoke() entry point is part of the GPU suite, so that it cannot rot unnoticed."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.gpu
def test_graft_entry_smoke(capsys):
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


@pytest.mark.gpu
def test_market_stats_device_reduction(native_lib):
    """fastace_env_market_stats == the reference's print_info sums (src/pybindings.cpp:20-75) computed from the books"""
    import numpy as np
    from fastace_b200 import _abi, scenario
    from fastace_b200.env import BatchedEconomy
    dims = (7, 40, 6, 3, 10)
    E, P, F, G, S = dims
    env = BatchedEconomy(dims)
    env.set_state(scenario.generic_initial_state(_abi.make_dims(*dims), 3))
    for t in range(3):
        act = scenario.synthetic_actions(dims, seed=5, step=t, **scenario.BENCH_PRESET)
        env.time_step(env.alloc_actions(act), env.alloc_outputs(), flags=_abi.IDX_MODULO)
        st = env.get_state()
        ms = {k: v.cpu().numpy() for k, v in env.market_stats().items()}
        for e in range(E):
            n, nj = int(st["m_count"][e]), int(st["j_count"][e])
            for g in range(G):
                sel = st["m_good"][e, :n] == g
                want = 0.0
                for pr in st["m_price"][e, :n][sel]:
                    want += 1.0 / pr                      # same order as the reference's loop
                assert ms["sum_quantity_per_price"][e, g] == want and ms["offers"][e, g] == sel.sum()
                assert ms["lots"][e, g] == st["m_left"][e, :n][sel].sum()
            want = 0.0
            for w in st["j_wage"][e, :nj]:
                want += w / 0.5
            assert ms["sum_wage_per_labor"][e] == want and ms["job_offers"][e] == nj and ms["job_lots"][e] == st["j_left"][e, :nj].sum()
    env.close()
