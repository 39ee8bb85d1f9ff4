"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical
injected actions and visiting orders.  Bit-exact for matching / counters / market order,
1e-5 relative for floating point (north_star)."""
import numpy as np
import pytest

from fastace_b200 import _abi, scenario
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run_episode(oracle, dims, steps, seed, flags=_abi.IDX_MODULO, preset=None, state=None, host_api=False,
                 shuffled=True):
    from fastace_b200.env import BatchedEconomy
    E, P, F, G, S = dims
    if state is None:
        state = scenario.custom_initial_state(dims, seed)[0] if G == 2 else scenario.generic_initial_state(dims, seed)
    env = BatchedEconomy(dims)
    env.set_state(state, time=0)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, seed + 17) if shuffled else None
    worst = 0.0
    for t in range(steps):
        perms = orders.next() if shuffled else None
        act = scenario.synthetic_actions(dims, seed=seed + 1, step=t, perms=perms, **(preset or {}))
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        if host_api:
            gout = _abi.alloc_host("out", dims)
            env.time_step_host(act, gout, flags=flags)
        else:
            dact = env.alloc_actions(act)
            dout = env.alloc_outputs(names=None)
            env.time_step(dact, dout, flags=flags)
            gout = {k: v.cpu().numpy().view(_abi.shapes("out", dims)[k][0]) for k, v in dout.items()}
        worst = max(worst, H.compare_outputs(gout, oout, dims, before))
        got = env.get_state()
        worst = max(worst, H.compare_states(got, ost, dims))
        # the kernels keep the reference's fp64 operation order: money and labour are bit-identical
        for k in ("p_money", "f_money", "p_labor", "f_last_money"):
            assert np.array_equal(got[k], ost[k], equal_nan=True), (k, t)
        assert env.get_time() == t + 1
    env.close()
    return worst


@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL])
def test_config_b_small_batch(oracle, mode):
    # config B shape (100 persons + 10 firms, 2 goods, stack 10), 64 economies, 40-step episode;
    # both matching kernels (lane-parallel v2 = default, serial v1)
    worst = _run_episode(oracle, (64, 100, 10, 2, 10), 40, seed=11, preset=scenario.BENCH_PRESET,
                         flags=_abi.IDX_MODULO | mode)
    print("max rel err", worst)


@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL])
def test_bankrupt_firms(oracle, mode):
    # untuned recipe: firms run out of money, job offers are killed by the first applicant
    _run_episode(oracle, (32, 100, 10, 2, 10), 30, seed=5, preset=dict(labor_mu=1.0), flags=_abi.IDX_MODULO | mode)


def test_config_a_default_scenario(oracle):
    # config A: py/train.py defaults, 48 persons + 12 firms, 40 steps
    _run_episode(oracle, (8, 48, 12, 2, 10), 40, seed=3)


def test_host_api_matches(oracle):
    _run_episode(oracle, (5, 37, 7, 2, 10), 12, seed=5, preset=scenario.BENCH_PRESET, host_api=True)


@pytest.mark.parametrize("dims", [
    (3, 1, 1, 1, 1), (4, 33, 1, 1, 3), (2, 100, 33, 2, 10), (3, 64, 40, 3, 16), (2, 31, 9, 8, 10),
    (2, 0, 3, 2, 4), (3, 7, 5, 5, 0), (2, 257, 12, 4, 7),
])
@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL])
def test_shapes(oracle, dims, mode):
    _run_episode(oracle, dims, 10, seed=dims[1] + 7 * dims[2], preset=scenario.BENCH_PRESET, flags=_abi.IDX_MODULO | mode)


def test_absolute_indices_with_out_of_range(oracle):
    from fastace_b200.env import BatchedEconomy
    dims = (6, 50, 6, 2, 10)
    E, P, F, G, S = dims
    state = scenario.custom_initial_state(dims, 21)[0]
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    rng = np.random.default_rng(0)
    for t in range(15):
        act = scenario.synthetic_actions(dims, seed=9, step=t, **scenario.BENCH_PRESET)
        # small absolute indices, some negative, some beyond any possible book size
        for k, hi in (("p_job_idx", F + 3), ("p_good_idx", F * G + 3), ("f_good_idx", F * G + 3)):
            act[k] = rng.integers(-2, hi, act[k].shape, dtype=np.int32)
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_ABSOLUTE, time_before=t)
        gout = _abi.alloc_host("out", dims)
        env.time_step_host(act, gout, flags=_abi.IDX_ABSOLUTE)
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)


def test_extreme_actions(oracle):
    """wage clip at 1e8, labour / amounts that overflow (int) conversion, zero proportions,
    everyone requests everything, nobody requests anything."""
    from fastace_b200.env import BatchedEconomy
    dims = (4, 40, 5, 2, 10)
    E, P, F, G, S = dims
    state = scenario.custom_initial_state(dims, 33)[0]
    state["f_inv"][0] *= 1e6  # huge inventories -> many lots
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    for t in range(12):
        act = scenario.synthetic_actions(dims, seed=4, step=t, **scenario.BENCH_PRESET)
        act["f_job_wage"][1] = 3e9           # clipped to 1e8 then /0.5
        act["f_job_labor"][2] = 3e9          # (int) overflow -> no offer
        act["f_job_labor"][3, :2] = np.inf
        act["f_offer_amt"][0] = 1.0
        act["f_offer_price"][0] = 1e-3
        act["f_offer_price"][1, 0, :2] = np.inf     # exp() overflow of a policy net: nobody can afford it
        act["f_offer_price"][1, 1, 2:4] = np.nan    # never compares affordable
        act["f_job_wage"][3, 2:4] = np.inf          # clipped to 1e8
        act["p_consume"][1] = 0.0
        act["p_consume"][2] = 1.0
        if t % 3 == 0:
            act["p_job_take"][:] = 1; act["p_good_take"][:] = 1; act["f_good_take"][:] = 1
        if t % 3 == 1:
            act["p_job_take"][:] = 0; act["p_good_take"][:] = 0; act["f_good_take"][:] = 0
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
        gout = _abi.alloc_host("out", dims)
        env.time_step_host(act, gout, flags=_abi.IDX_MODULO)
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)


@pytest.mark.parametrize("modulo", [True, False])
@pytest.mark.parametrize("device_path", [True, False])
def test_compact_encoding_matches(oracle, modulo, device_path):
    """fastace_actions_compact_t carries the same decisions: same results as the oracle on the
    standard encoding (host-pointer path with ASYNC pipelining, and device-pointer path)."""
    from fastace_b200.env import BatchedEconomy
    dims = (9, 100, 10, 2, 10)
    E, P, F, G, S = dims
    state = scenario.custom_initial_state(dims, 77)[0]
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 78)
    flags = _abi.IDX_MODULO if modulo else _abi.IDX_ABSOLUTE
    rng = np.random.default_rng(5)
    keep = []
    for t in range(14):
        act = scenario.synthetic_actions(dims, seed=79, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        if not modulo:
            for k, hi in (("p_job_idx", F + 2), ("p_good_idx", F * G + 2), ("f_good_idx", F * G + 2)):
                act[k] = rng.integers(-1, hi, act[k].shape, dtype=np.int32)
        cz = _abi.compact_actions_for_counts(act, ost["j_count"], ost["m_count"], modulo)
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        gout = _abi.alloc_host("out", dims)
        if device_path:
            dcz = env.alloc_compact_actions(cz)
            dout = env.alloc_outputs(names=None)
            env.time_step(env.pack_device("compact", dcz), dout, flags=flags)
            gout = {k: v.cpu().numpy().view(_abi.shapes("out", dims)[k][0]) for k, v in dout.items()}
        else:
            czs = _abi.struct_from_numpy("compact", cz, env.dims)
            keep.append((cz, czs))
            env.time_step_host(czs, gout, flags=flags | _abi.STEP_ASYNC)
            env.sync()
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)
    env.close()


def test_async_host_pipeline_equals_sync(oracle):
    """several asynchronous host-pointer steps in flight (double-buffered staging) give the same
    trajectory as synchronous stepping; per-step outputs land in per-step host buffers"""
    from fastace_b200.env import BatchedEconomy
    dims = (16, 100, 10, 2, 10)
    state = scenario.custom_initial_state(dims, 5)[0]
    acts = [scenario.synthetic_actions(dims, seed=6, step=t, **scenario.BENCH_PRESET) for t in range(12)]
    envs, outs = [], []
    for mode in (0, _abi.STEP_ASYNC):
        env = BatchedEconomy(dims)
        env.set_state(state)
        o = [_abi.alloc_host("out", dims, names=("p_reward", "f_profit")) for _ in acts]
        for t, a in enumerate(acts):
            env.time_step_host(a, o[t], flags=_abi.IDX_MODULO | mode)
        env.sync()
        envs.append(env.get_state())
        outs.append(o)
        env.close()
    for k in envs[0]:
        assert np.array_equal(envs[0][k], envs[1][k]), k
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a["p_reward"], b["p_reward"]) and np.array_equal(a["f_profit"], b["f_profit"])


@pytest.mark.parametrize("kind", ["actions", "compact"])
def test_host_block_single_copy_path(oracle, kind):
    """host arrays carved from one block (alloc_host_block) take the single-copy staging path; results unchanged"""
    from fastace_b200.env import BatchedEconomy
    dims = (6, 100, 10, 2, 10)
    state = scenario.custom_initial_state(dims, 41)[0]
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 42)
    for t in range(8):
        act = scenario.synthetic_actions(dims, seed=43, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
        src = act if kind == "actions" else _abi.compact_actions_for_counts(act, before["j_count"], before["m_count"], True)
        blk, _keep = _abi.alloc_host_block(kind, dims)
        for k, v in src.items():
            blk[k][...] = v
        gout, _keep2 = _abi.alloc_host_block("out", dims)
        env.time_step_host(_abi.struct_from_numpy(kind, blk, dims), gout, flags=_abi.IDX_MODULO)
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)
    env.close()


def test_two_call_step_equals_one_call(oracle):
    """FASTACE_STEP_PERSONS + FASTACE_STEP_FIRMS == one full step, bit for bit (state, outputs), and the oracle"""
    from fastace_b200 import lib
    from fastace_b200.env import BatchedEconomy
    dims = (12, 100, 10, 2, 10)
    state = scenario.custom_initial_state(dims, 51)[0]
    one, two = BatchedEconomy(dims), BatchedEconomy(dims)
    one.set_state(state); two.set_state(state)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 52)
    for t in range(10):
        act = scenario.synthetic_actions(dims, seed=53, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        before = H.copy_state(ost)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
        out1, out2 = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        one.time_step_host(act, out1, flags=_abi.IDX_MODULO)
        person_out = {k: out2[k] for k in ("p_reward", "p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")}
        firm_out = {k: out2[k] for k in ("f_profit", "f_good_ok", "old_m_left", "old_m_taken")}
        two.time_step_host(act, person_out, flags=_abi.IDX_MODULO | _abi.STEP_PERSONS)
        assert two.get_time() == t                                  # the step is not complete yet
        mid = two.get_state()
        assert np.array_equal(mid["p_money"], one.get_state()["p_money"])        # persons are done ...
        assert np.array_equal(mid["m_count"], before["m_count"])                 # ... the books are still last step's
        with pytest.raises(lib.FastaceError):
            two.time_step_host(act, out2, flags=_abi.IDX_MODULO)                 # a full step cannot start mid-step
        two.time_step_host(act, firm_out, flags=_abi.IDX_MODULO | _abi.STEP_FIRMS)
        assert two.get_time() == t + 1
        s1, s2 = one.get_state(), two.get_state()
        for k in s1:
            if k.startswith("m_") and k != "m_count" or k.startswith("j_") and k != "j_count":
                cnt = s1["m_count"] if k.startswith("m_") else s1["j_count"]     # books: live prefix (the rest is not market)
                for e in range(dims[0]):
                    assert np.array_equal(s1[k][e, :cnt[e]], s2[k][e, :cnt[e]], equal_nan=s1[k].dtype.kind == "f"), (k, e)
            else:
                assert np.array_equal(s1[k], s2[k], equal_nan=s1[k].dtype.kind == "f"), k
        for k in ("p_reward", "f_profit", "p_job_ok", "p_good_ok", "f_good_ok"):
            assert np.array_equal(out1[k], out2[k], equal_nan=out1[k].dtype.kind == "f"), k
        H.compare_outputs(out2, out1, dims, before)      # old_* counters: live prefix of the previous books
        H.compare_outputs(out2, oout, dims, before)
        H.compare_states(s2, ost, dims)
    with pytest.raises(lib.FastaceError):
        two.time_step_host(act, out2, flags=_abi.IDX_MODULO | _abi.STEP_FIRMS)   # firms need the person phase first
    # the person phase itself in two calls: trades, then consumption (which touches nobody but the person)
    for t in range(10, 14):
        act = scenario.synthetic_actions(dims, seed=53, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        out1, out2 = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        before = one.get_state()
        one.time_step_host(act, out1, flags=_abi.IDX_MODULO)
        two.time_step_host(act, {k: out2[k] for k in ("p_job_ok", "p_good_ok", "old_j_left", "old_j_taken")},
                           flags=_abi.IDX_MODULO | _abi.STEP_PERSONS_TRADE)
        mid = two.get_state()
        bought = out2["p_good_ok"].sum()
        assert np.isclose((mid["p_inv"] - before["p_inv"]).sum(), bought)       # purchases are in, nothing consumed yet
        assert np.array_equal(mid["p_labor"], 0.5 * out2["p_job_ok"].sum(axis=1))
        with pytest.raises(lib.FastaceError):
            two.time_step_host(act, {"f_profit": out2["f_profit"]}, flags=_abi.IDX_MODULO | _abi.STEP_FIRMS)   # consumption is due
        two.time_step_host(act, {"p_reward": out2["p_reward"]}, flags=_abi.IDX_MODULO | _abi.STEP_PERSONS_CONSUME)
        two.time_step_host(act, {k: out2[k] for k in ("f_profit", "f_good_ok", "old_m_left", "old_m_taken")},
                           flags=_abi.IDX_MODULO | _abi.STEP_FIRMS)
        s1, s2 = one.get_state(), two.get_state()
        for k in ("p_money", "p_inv", "p_labor", "f_money", "f_inv", "f_labor", "f_last_money", "m_count", "j_count"):
            assert np.array_equal(s1[k], s2[k], equal_nan=s1[k].dtype.kind == "f"), k
        for k in ("p_reward", "f_profit", "p_job_ok", "p_good_ok", "f_good_ok"):
            assert np.array_equal(out1[k], out2[k], equal_nan=out1[k].dtype.kind == "f"), k
        H.compare_states(s2, s1, dims)
    assert one.get_time() == two.get_time() == 14
    one.close(); two.close()


@pytest.mark.parametrize("mode", [0, _abi.STEP_SERIAL, _abi.STEP_LARGE])
def test_firm_money_near_tie(oracle, mode):
    """constructed near-ties (tests/near_tie.py): the third applicant is hired / refused by the LAST BIT of the firm's
    running money, which comes out right only in the reference's operation order (hire, sale, hire)"""
    from fastace_b200.env import BatchedEconomy
    from tests import near_tie
    dims, state, acts, ties = near_tie.build(48, seed=5)
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    for t, act in enumerate(acts):
        before = H.copy_state(ost)
        oout, gout = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_ABSOLUTE, time_before=t)
        env.time_step_host(act, gout, flags=_abi.IDX_ABSOLUTE | mode)
        H.compare_outputs(gout, oout, dims, before)
        got = env.get_state()
        H.compare_states(got, ost, dims)
        for k in ("p_money", "f_money", "p_labor"):
            assert np.array_equal(got[k], ost[k]), (k, t)
    hired = np.array([h for (_, _, _, h) in ties])
    assert np.array_equal(gout["p_job_ok"][:, 0, 2].astype(bool), hired) and hired.any() and (~hired).any()
    env.close()


@pytest.mark.parametrize("own_orders", [False, True])
def test_packed_host_encoding_matches(oracle, own_orders):
    """fastace_env_step_host_packed: bit-field indices without take masks, optionally without visiting orders (the env
    shuffles on the device) — same trajectory as the oracle on the int32 encoding with the host's std::shuffle orders"""
    from fastace_b200.env import BatchedEconomy
    dims = (9, 100, 10, 2, 10)
    state = scenario.custom_initial_state(dims, 77)[0]
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 78)
    if own_orders:
        env.restart_orders(78)
    keep = []
    for t in range(14):
        act = scenario.synthetic_actions(dims, seed=79, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        pk = _abi.packed_actions_for_counts(act, ost["j_count"], ost["m_count"], True, with_orders=not own_orders)
        before = H.copy_state(ost)
        oout, gout = _abi.alloc_host("out", dims), _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_MODULO, time_before=t)
        pks = _abi.struct_from_numpy("packed", pk, env.dims)
        keep.append((pk, pks))
        env.time_step_host(pks, gout, flags=_abi.STEP_ASYNC if t % 2 else 0)
        env.sync()
        H.compare_outputs(gout, oout, dims, before)
        H.compare_states(env.get_state(), ost, dims)
    env.close()


def test_completion_queue_equals_whole_grid_wait(native_lib, monkeypatch):
    """update_kernel fed by match_kernel's completion queue (the default for full steps) and update_kernel waiting
    for the whole matching grid (FASTACE_NO_QUEUE=1, read when the env is created) give bit-identical states, also
    when the economy count is not a multiple of the queue's group size and over many back-to-back launches."""
    from fastace_b200.env import BatchedEconomy
    dims = (333, 100, 10, 2, 10)
    state0 = scenario.custom_initial_state(dims, 77)[0]
    runs = []
    for no_queue in ("0", "1"):
        monkeypatch.setenv("FASTACE_NO_QUEUE", no_queue)
        env = BatchedEconomy(dims)
        env.set_state(state0)
        orders = scenario.OrderStream(dims, 5)
        dout = env.alloc_outputs(names=None)
        for t in range(30):
            act = scenario.synthetic_actions(dims, seed=11, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
            env.time_step(env.alloc_actions(act), dout, flags=_abi.IDX_MODULO)
        runs.append((env.get_state(), {k: v.cpu().numpy().copy() for k, v in dout.items()}))
        env.close()
    (sa, oa), (sb, ob) = runs
    for k in sa:
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    for k in oa:
        assert np.array_equal(oa[k].view(np.uint8), ob[k].view(np.uint8)), k


@pytest.mark.parametrize("modulo", [True, False])
def test_specialised_compact_kernels_match(oracle, modulo):
    """compact encoding on the device path with only the mandatory outputs: launch_step picks
    match_kernel<G, kModeCompact [| kModeModulo]> (the other encodings compiled out); states against the oracle"""
    from fastace_b200.env import BatchedEconomy
    dims = (37, 100, 10, 2, 10)
    E, P, F, G, S = dims
    state = scenario.custom_initial_state(dims, 177)[0]
    env = BatchedEconomy(dims)
    env.set_state(state)
    ost = H.copy_state(state)
    orders = scenario.OrderStream(dims, 178)
    flags = _abi.IDX_MODULO if modulo else _abi.IDX_ABSOLUTE
    rng = np.random.default_rng(6)
    for t in range(30):
        act = scenario.synthetic_actions(dims, seed=179, step=t, perms=orders.next(), **scenario.BENCH_PRESET)
        if not modulo:
            for k, hi in (("p_job_idx", F + 2), ("p_good_idx", F * G + 2), ("f_good_idx", F * G + 2)):
                act[k] = rng.integers(-1, hi, act[k].shape, dtype=np.int32)
        cz = _abi.compact_actions_for_counts(act, ost["j_count"], ost["m_count"], modulo)
        oout = _abi.alloc_host("out", dims)
        oracle.step(dims, ost, act, oout, flags=flags, time_before=t)
        dout = env.alloc_outputs()                                   # p_reward and f_profit only
        env.time_step(env.pack_device("compact", env.alloc_compact_actions(cz)), dout, flags=flags)
        got = env.get_state()
        H.compare_states(got, ost, dims)
        for k in ("p_money", "f_money", "p_labor", "f_last_money"):
            assert np.array_equal(got[k], ost[k], equal_nan=True), (k, t)
        assert np.allclose(dout["p_reward"].cpu().numpy(), oout["p_reward"], rtol=1e-5, equal_nan=True)
        assert np.allclose(dout["f_profit"].cpu().numpy(), oout["f_profit"], rtol=1e-5, atol=1e-9, equal_nan=True)
    env.close()
