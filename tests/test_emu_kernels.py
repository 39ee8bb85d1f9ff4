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
OPTIONAL_OUT = ("p_job_ok", "p_good_ok", "f_good_ok", "old_j_left", "old_j_taken", "old_m_left", "old_m_taken")


@pytest.fixture(scope="module")
def emu():
    from tests.emu.loader import EmuKernels
    k = EmuKernels()
    yield k
    k.close()


def _episode(emu, oracle, dims, steps, seed, preset, flags=_abi.IDX_MODULO, compact=False, tweak=None, drop_out=()):
    """drop_out: outputs the emulated kernels are not asked for (compact encoding without the persons' success flags =
    the call the specialised match_kernel<G, kModeCompact...> is compiled for)"""
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
        for k in drop_out:
            del eout[k]
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


@pytest.mark.parametrize("modulo", [True, False])
def test_specialised_compact_kernels(emu, oracle, modulo):
    """compact encoding, persons' success flags not requested: launch_step (and the emulator harness) pick
    match_kernel<G, kModeCompact [| kModeModulo]> — the same source with the other encodings compiled out"""
    rng = np.random.default_rng(8)

    def tweak(t, act):                       # absolute indices: in range, with a few out-of-range ones
        if not modulo:
            for k in ("p_job_idx", "p_good_idx", "f_good_idx"):
                act[k][...] = rng.integers(-1, 14, act[k].shape, dtype=np.int32)
    st = _episode(emu, oracle, (5, 100, 10, 2, 10), 24, 31, scenario.BENCH_PRESET, tweak=tweak,
                  flags=_abi.IDX_MODULO if modulo else _abi.IDX_ABSOLUTE, compact=True, drop_out=OPTIONAL_OUT)
    assert st["sales_windows"] > 0 and st["rescans"] > 0
    _episode(emu, oracle, (3, 64, 40, 3, 16), 8, 9, dict(take_prob=0.7, prod_scale=0.8, wage_scale=3.0, price_scale=0.3, labor_mu=1.5),
             flags=_abi.IDX_MODULO if modulo else _abi.IDX_ABSOLUTE, compact=True, tweak=tweak, drop_out=OPTIONAL_OUT)


@pytest.mark.parametrize("order", [1, 2])
def test_completion_queue_in_any_finishing_order(emu, oracle, order):
    """match_kernel's warps hand their economies to update_kernel through the completion queue in the order they
    finish: with the emulated blocks run in descending / scrambled order the queue is a permutation of the economies,
    and the step must not notice (also with an economy count that is not a multiple of the queue's group size)"""
    emu.block_order(order)
    try:
        _episode(emu, oracle, (37, 40, 6, 2, 10), 6, 21, scenario.BENCH_PRESET)
        _episode(emu, oracle, (5, 100, 10, 2, 10), 4, 22, scenario.BENCH_PRESET, compact=True, drop_out=OPTIONAL_OUT)
    finally:
        emu.block_order(0)


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


def test_firm_money_near_tie_specialised_kernel(emu, oracle):
    """the same last-bit ties through the compact encoding without optional outputs (specialised match_kernel)"""
    from tests import near_tie
    dims, state, acts, ties = near_tie.build(24, seed=3)
    ost, est = H.copy_state(state), H.copy_state(state)
    for t, act in enumerate(acts):
        oout = _abi.alloc_host("out", dims)
        eout = {k: v for k, v in _abi.alloc_host("out", dims).items() if k not in OPTIONAL_OUT}
        cz = _abi.compact_actions_for_counts(act, ost["j_count"], ost["m_count"], False)
        oracle.step(dims, ost, act, oout, flags=_abi.IDX_ABSOLUTE, time_before=t)
        emu.step(dims, est, None, eout, flags=_abi.IDX_ABSOLUTE, time_before=t, compact=cz)
        H.compare_states(est, ost, dims)
        for k in EXACT_FLOAT:
            assert np.array_equal(est[k], ost[k], equal_nan=True), (k, t)


def test_device_shuffle_equals_libstdcxx(emu):
    """shuffle_orders_kernel (shared-memory and global-memory variants, pairwise and per-element branches of
    std::shuffle) against the host's libstdc++ through fastace_shuffle_orders, cumulatively"""
    import ctypes as C
    L = emu.lib
    pt = lambda a: a.ctypes.data_as(C.c_void_p)
    for (E, P, F, smem, calls) in [(70, 100, 10, 1, (3, 1, 4)), (3, 48, 12, 1, (1, 1)), (5, 1, 1, 1, (2,)), (3, 2, 3, 1, (2, 2)),
                                   (2, 1000, 37, 0, (1, 1)), (1, 50000, 3, 0, (1,))]:
        host = scenario.OrderStream((E, P, F, 2, 10), 1234)
        rng = np.zeros(E, np.uint64)
        sp_, sf_ = np.zeros((E, P), np.int32), np.zeros((E, F), np.int32)
        for call, steps in enumerate(calls):
            op, of = np.zeros((steps, E, P), np.int32), np.zeros((steps, E, F), np.int32)
            op16, of16 = np.zeros((steps, E, P), np.uint16), np.zeros((steps, E, F), np.uint16)
            L.fastace_emu_shuffle(E, P, F, C.c_uint32(1234), 1 if call == 0 else 0, steps, pt(rng), pt(sp_), pt(sf_),
                                  pt(op) if smem else None, pt(of) if smem else None, pt(op16) if smem else None,
                                  pt(of16) if smem else None, smem)
            for k in range(steps):
                hp, hf = host.next()
                if smem:
                    assert np.array_equal(op[k], hp) and np.array_equal(of[k], hf) and np.array_equal(op16[k], hp) and np.array_equal(of16[k], hf)
            assert np.array_equal(sp_, hp) and np.array_equal(sf_, hf) and np.array_equal(rng, host.rng_state)
    # KAT of SURVEY.md App. C: minstd_rand0(1234), 0..9
    rng, sp_, sf_ = np.zeros(1, np.uint64), np.zeros((1, 10), np.int32), np.zeros((1, 1), np.int32)
    op, of = np.zeros((2, 1, 10), np.int32), np.zeros((2, 1, 1), np.int32)
    L.fastace_emu_shuffle(1, 10, 1, C.c_uint32(1234), 1, 2, pt(rng), pt(sp_), pt(sf_), pt(op), pt(of), None, None, 1)
    assert op[0, 0].tolist() == [5, 0, 4, 8, 1, 2, 7, 6, 3, 9]


def test_segmented_sort_is_a_stable_sort_by_firm(emu):
    """segmented_sort.cuh (histogram / scan / ranked scatter) == numpy's stable argsort, "no request" keys last"""
    import ctypes as C
    L = emu.lib
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(1)
    for (n, F, chunk) in [(5000, 37, 256), (1, 1, 32), (0, 5, 64), (777, 1000, 96), (20000, 300, 4096), (4097, 2, 4096)]:
        key = rng.integers(0, F, n).astype(np.uint16)
        key[rng.random(n) < 0.3] = 0xFFFF
        val = rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
        bins = np.where(key == 0xFFFF, F, key).astype(np.int64)
        seg = np.zeros(F + 1, np.uint32)
        seg[1:] = np.cumsum(np.bincount(bins, minlength=F + 1)[:F])
        hist = np.zeros((max((n + chunk - 1) // chunk, 1), F + 1), np.uint32)
        ko, vo = np.zeros(n, np.uint16), np.zeros(n, np.uint32)
        L.fastace_emu_segmented_sort(p(key), p(val), p(ko), p(vo), p(seg), p(hist), n, F, chunk)
        order = np.argsort(bins, kind="stable")
        assert np.array_equal(ko, key[order]) and np.array_equal(vo, val[order]), (n, F, chunk)


def test_packed_encoding_expands_to_the_compact_one(emu):
    """_abi.packed_actions_for_counts (host packer) -> expand_packed_kernel == _abi.compact_actions_for_counts"""
    import ctypes as C
    L = emu.lib
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for dims, modulo in (((3, 7, 5, 3, 10), True), ((2, 33, 10, 2, 10), False), ((2, 5, 40, 6, 16), True), ((1, 4, 1, 1, 1), True)):
        E, P, F, G, S = dims
        act = scenario.synthetic_actions(dims, seed=1, step=0, **scenario.BENCH_PRESET)
        rng = np.random.default_rng(E)
        jc, mc = rng.integers(0, F + 1, E), rng.integers(0, F * G + 1, E)
        if not modulo:
            for k, hi in (("p_job_idx", F + 2), ("p_good_idx", F * G + 2), ("f_good_idx", F * G + 2)):
                act[k] = rng.integers(-1, hi, act[k].shape, dtype=np.int32)
        pk = _abi.packed_actions_for_counts(act, jc, mc, modulo)
        cz = _abi.compact_actions_for_counts(act, jc, mc, modulo)
        bj, nbj, bg, nbg = _abi.packed_layout(F, G, S)
        for name, tname, bits, nb, cnt in (("p_job_idx", "p_job_take", bj, nbj, jc), ("p_good_idx", "p_good_take", bg, nbg, mc),
                                           ("f_good_idx", "f_good_take", bg, nbg, mc)):
            n = pk[name].shape[0] * pk[name].shape[1]
            idx, take = np.zeros((n, S), np.uint8), np.zeros(n, np.uint16)
            L.fastace_emu_expand_packed(n, S, bits, nb, p(pk[name]), p(idx), p(take))
            idx, take = idx.reshape(cz[name].shape), take.reshape(cz[tname].shape)
            # a request = take bit set AND index inside the book; both encodings must name the same requests
            want_idx = cz[name].astype(np.int64)
            want = ((cz[tname][:, :, None] >> np.arange(S)) & 1).astype(bool) & (want_idx < cnt[:, None, None]) & (cnt[:, None, None] > 0)
            got = ((take[:, :, None] >> np.arange(S)) & 1).astype(bool)
            assert np.array_equal(got, want), name
            assert np.array_equal(np.where(want, idx, 0), np.where(want, want_idx, 0)), name


def test_reward_pow_accuracy(emu):
    """pow_reward (common.cuh: the short-polynomial x^y of the utility path) against libm pow: far inside the 1e-5
    reward tolerance over the whole range it serves, and the library fallbacks for everything else"""
    import ctypes as C
    import math
    f = emu.lib.fastace_emu_pow_reward
    f.restype, f.argtypes = C.c_double, [C.c_double, C.c_double]
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(20000):
        x, y = 10 ** rng.uniform(-9, 9), rng.uniform(-30, 30)
        if abs(y * math.log(x)) > 690:
            continue
        worst = max(worst, abs(f(x, y) - math.pow(x, y)) / math.pow(x, y))
    assert worst < 5e-11, worst
    assert f(2.0, -9.0) == 2.0 ** -9 and f(1.0, 5.0) == 1.0
    assert f(float("inf"), -1.0) == 0.0 and math.isinf(f(0.0, -0.1)) and math.isnan(f(float("nan"), 1.0))
    assert f(1e-8, 80.0) == 0.0 and abs(f(1e-320, 0.5) / 1e-160 - 1) < 1e-5      # underflow / denormal input: library path
