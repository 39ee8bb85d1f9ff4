"""Host-side logic of the C ABI that needs no GPU: library loads and exports every
declared symbol, legacy struct layouts, scenario draws, visiting orders."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fastace_b200 import _abi, lib, scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "fastace_b200.h")).read()
    # every function prototype in the header: `name(`
    declared = set(re.findall(r"^\s*(?:int|void|const char\*|fastace_custom_scenario_params_t|fastace_training_params_t)\s+\**(\w+)\s*\(",
                              header, flags=re.M))
    assert declared, "no prototypes found"
    assert declared == set(lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(native_lib, name), name
    assert native_lib.fastace_abi_version() == _abi.ABI_VERSION


def test_legacy_struct_layouts_match_reference():
    # sizeof(CustomScenarioParams)=344, sizeof(TrainingParams)=136 and the offsets checked in SURVEY.md §5
    assert C.sizeof(_abi.CustomScenarioParams) == 344
    assert C.sizeof(_abi.TrainingParams) == 136
    assert _abi.TrainingParams.firmValueNetLR.offset == 104
    assert _abi.TrainingParams.episodeBatchSizeForLRDecay.offset == 112
    assert _abi.TrainingParams.multiplierForLRDecay.offset == 120
    assert _abi.TrainingParams.reverseAnnealingPeriod.offset == 128


def test_create_params_defaults(native_lib):
    p = scenario.scenario_params(48, 12)
    assert (p.numPeople, p.numFirms) == (48, 12)
    # src/neural/neuralScenarios.h:61-114
    assert (p.money_mu, p.money_sigma, p.good2_mu, p.good2_sigma) == (10.0, 2.0, 1.0, 0.2)
    assert (p.firm_money_mu, p.firm_good1_sigma, p.firm_good2_mu) == (50.0, 4.0, 30.0)
    assert (p.firm_good2_share1_sigma, p.firm_good2_share2_sigma) == (0.02, 0.05)
    assert (p.firm_elasticity2_mu, p.firm_elasticity2_sigma) == (10.0, 2.5)
    t = scenario.training_params()
    # src/neural/neuralConstants.h:13-29
    assert (t.numEpisodes, t.episodeLength, t.stackSize, t.encodingSize, t.hiddenSize, t.nHidden, t.nHiddenSmall) == \
        (100, 20, 10, 10, 100, 12, 6)
    assert t.purchaseNetLR == 1e-5 and t.firmValueNetLR == 1e-5 and t.multiplierForLRDecay == 0.5
    assert (t.episodeBatchSizeForLRDecay, t.patienceForLRDecay, t.reverseAnnealingPeriod) == (10, 5, 3)


def test_shuffle_matches_libstdcxx_kat(native_lib):
    # SURVEY.md Appendix C KAT: minstd_rand0(1234), std::shuffle of 0..9, twice (g++ 13, glibc 2.39)
    dims = (1, 10, 10, 2, 1)
    s = scenario.OrderStream(dims, 1234)
    pp, pf = s.next()
    assert pp[0].tolist() == [5, 0, 4, 8, 1, 2, 7, 6, 3, 9]
    # persons then firms use the SAME engine (economy.cpp:110-111): the firm vector (identity
    # here) receives the engine's NEXT shuffle.  std::shuffle acts on positions, so applying
    # that position permutation to the first result must give the KAT's "shuffled again" line.
    assert [int(pp[0][i]) for i in pf[0]] == [8, 7, 0, 3, 4, 1, 9, 5, 2, 6]
    # cumulative: the second step permutes the previous order
    pp2, _ = s.next()
    assert sorted(pp2[0].tolist()) == list(range(10)) and pp2[0].tolist() != pp[0].tolist()


def test_shuffle_against_reference_engine_when_present(native_lib):
    loader = pytest.importorskip("oracle.loader")
    if not loader.have_reference():
        pytest.skip("oracle/_ref not built here")
    L = loader.ref_lib()
    for seed, n in ((1, 100), (77, 48), (123456, 7), (2**31 - 1, 33), (0, 5)):
        out = (C.c_int32 * (2 * n))()
        L.fastace_ref_shuffle_kat(seed, n, 2, out)
        s = scenario.OrderStream((1, n, n, 2, 1), seed)
        pp, pf = s.next()
        # ref: two cumulative shuffles of one vector; ours: persons then firms from identity ->
        # first person order must equal the reference's first round
        assert pp[0].tolist() == list(out[:n])


def test_orders_follow_the_reference_economy(native_lib, oracle):
    """the orders produced by fastace_shuffle_orders are exactly those the reference's own
    Economy::time_step uses when its rng is seeded the same way"""
    loader = pytest.importorskip("oracle.loader")
    if not loader.have_reference():
        pytest.skip("oracle/_ref not built here")
    dims = (3, 20, 6, 2, 4)
    state = scenario.custom_initial_state(dims, 5)[0]
    ref = loader.Reference(dims, state, seed=4242)
    stream = scenario.OrderStream(dims, 4242)
    for t in range(6):
        act = scenario.synthetic_actions(dims, seed=1, step=t)
        out = _abi.alloc_host("out", dims, names=("p_reward", "f_profit"))
        pp, pf = ref.step(act, out, flags=_abi.IDX_MODULO)
        qp, qf = stream.next()
        assert np.array_equal(pp, qp) and np.array_equal(pf, qf)


def test_custom_initial_state_distributions(native_lib):
    dims = (64, 100, 10, 2, 10)
    st, disc = scenario.custom_initial_state(dims, 9)
    assert abs(st["p_money"].mean() - 10.0) < 0.1 and abs(st["p_money"].std() - 2.0) < 0.1
    assert abs(st["p_inv"][:, 0].mean() - 10.0) < 0.1 and abs(st["p_inv"][:, 1].mean() - 1.0) < 0.02
    assert abs(st["f_money"].mean() - 50.0) < 1.5
    # good 1 of firms is drawn with firm_good2_sigma = 5 (sic, neuralScenarios.cpp:130), not 4
    assert abs(st["f_inv"][:, 0].std() - 5.0) < 0.5
    assert np.allclose(st["p_util_share"].sum(axis=1), 1.0) and np.allclose(st["f_prod_share"].sum(axis=2), 1.0)
    assert (st["p_inv"] >= 0).all() and (st["f_inv"] >= 0).all() and (st["p_money"] >= 0).all()
    assert (disc > 0).all() and (disc < 1).all()
    assert (st["m_count"] == 0).all() and (st["j_count"] == 0).all()
    # different economies draw different states; the same seed reproduces
    assert not np.array_equal(st["p_money"][0], st["p_money"][1])
    st2, _ = scenario.custom_initial_state(dims, 9)
    assert all(np.array_equal(st[k], st2[k]) for k in st)
    # economy e of seed s equals economy 0 of seed s+e (seed_e = base + e)
    st3, _ = scenario.custom_initial_state((1,) + dims[1:], 9 + 5)
    assert np.array_equal(st3["p_money"][0], st["p_money"][5])


def test_ces_normalisation_in_scenario_equals_reference_rule(native_lib, oracle):
    st, _ = scenario.custom_initial_state((2, 5, 2, 2, 10), 3)
    # rho = 1/(1-sigma) < 0 for sigma > 1 ; shares normalised
    assert ((st["p_util_rho"] < 0) | (st["p_util_rho"] > 1)).all()
    assert np.allclose(st["p_util_share"].sum(axis=1), 1.0, rtol=0, atol=1e-15)


def test_env_create_fails_loudly_without_gpu(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fastace_b200.env import BatchedEconomy
    with pytest.raises(lib.FastaceError) as e:
        BatchedEconomy((4, 10, 2, 2, 10))
    assert "no CPU path" in str(e.value) or "status -3" in str(e.value)


def test_env_create_rejects_bad_dims(native_lib):
    h = C.c_void_p()
    for bad in ((0, 10, 2, 2, 10), (4, 10, 2, 9, 10), (4, 10, 2, 2, 17), (4, 10, 70000, 2, 10), (1, 2000000, 10, 2, 10)):
        d = _abi.make_dims(*bad)
        assert native_lib.fastace_env_create(C.byref(d), 0, C.byref(h)) == -1
    # beyond the warp-per-economy kernels (F*G > 254) is NOT an error: such envs take the large-economy path;
    # without a GPU the call gets as far as the device check
    d = _abi.make_dims(1, 1000, 200, 2, 10)
    assert native_lib.fastace_env_create(C.byref(d), 0, C.byref(h)) in (0, -3)


def test_host_block_layout_matches_staging_layout():
    """alloc_host_block: fields in struct order, 256 B aligned and padded — the layout step_host copies in one piece"""
    dims = (3, 11, 4, 2, 5)
    for kind in ("actions", "compact", "out"):
        blk, holder = _abi.alloc_host_block(kind, dims)
        shp = _abi.shapes(kind, dims)
        assert list(blk) == list(shp)
        at = None
        for n, (dt, shape) in shp.items():
            a = blk[n]
            assert a.dtype == dt and tuple(a.shape) == tuple(shape) and a.flags["C_CONTIGUOUS"]
            assert a.ctypes.data % 256 == 0
            if at is not None:
                assert a.ctypes.data == at
            at = a.ctypes.data + (a.nbytes + 255) // 256 * 256
        _abi.struct_from_numpy(kind, blk, dims)      # accepted as is
