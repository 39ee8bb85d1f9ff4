"""Host-side scenario data: initial states, visiting orders and synthetic injected actions.

Thin numpy layer over the host entry points of the C ABI (they need no GPU):
  * initial state     — fastace_scenario_custom_init (CustomScenario::setup distributions,
                        /root/reference/src/neural/neuralScenarios.cpp:93-161)
  * visiting orders   — fastace_shuffle_orders (Economy::time_step's std::shuffle calls,
                        /root/reference/src/base/economy.cpp:110-111)
  * synthetic actions — the "fixed injected actions" recipe of SURVEY.md §8(d) config B,
                        drawn with numpy Philox so that every consumer (GPU env, oracle,
                        reference harness) is fed identical tensors.
"""
import ctypes as C

import numpy as np

from . import _abi, lib


def scenario_params(num_people, num_firms):
    """create_scenario_params (src/pybindings.cpp:8-13): the reference's defaults."""
    return lib.load().create_scenario_params(num_people, num_firms)


def training_params():
    """create_training_params (src/pybindings.cpp:15-18)."""
    return lib.load().create_training_params()


def custom_initial_state(dims, seed, params=None):
    """Host state dict for E economies drawn as CustomScenario::setup does (G must be 2).
    Returns (state, discount) with discount [E][P]."""
    d = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
    if params is None:
        params = scenario_params(d.num_persons, d.num_firms)
    state = _abi.alloc_host("state", d)
    st = _abi.struct_from_numpy("state", state, d)
    disc = np.zeros((d.num_econ, d.num_persons), dtype=np.float64)
    lib.check(lib.load().fastace_scenario_custom_init(
        C.byref(d), C.byref(params), int(seed) & 0xFFFFFFFF, C.byref(st), disc.ctypes.data_as(C.POINTER(C.c_double))))
    return state, disc


def generic_initial_state(dims, seed):
    """Initial state for any number of goods G (config D uses 8): the same families of
    distributions as CustomScenario, cycled over goods.  numpy Philox draws."""
    E, P, F, G, S = dims.tuple if isinstance(dims, _abi.Dims) else tuple(dims)
    rng = np.random.Generator(np.random.Philox(seed))
    st = _abi.alloc_host("state", (E, P, F, G, S))
    nonneg = lambda a: np.maximum(a, 0.0)
    st["p_money"][:] = nonneg(rng.normal(10.0, 2.0, (E, P)))
    for g in range(G):
        mu, sg = ((10.0, 2.0), (1.0, 0.2))[g % 2]
        st["p_inv"][:, g, :] = nonneg(rng.normal(mu, sg, (E, P)))
    st["p_util_tfp"][:] = 1.0
    share = np.empty((E, G + 1, P))
    share[:, 0, :] = rng.normal(0.4, 0.1, (E, P))
    for g in range(G):
        mu, sg = ((0.4, 0.1), (0.1, 0.02))[g % 2]
        share[:, g + 1, :] = rng.normal(mu, sg, (E, P))
    ssum = np.zeros((E, P))
    for i in range(G + 1):
        ssum = ssum + share[:, i, :]
    st["p_util_share"][:] = share / ssum[:, None, :]
    el = rng.normal(10.0, 2.5, (E, P))
    el = np.where(el <= 0, 1e-8, el)
    st["p_util_rho"][:] = 1 / (1 - el)
    st["f_money"][:] = nonneg(rng.normal(50.0, 10.0, (E, F)))
    for g in range(G):
        mu, sg = ((10.0, 5.0), (30.0, 5.0))[g % 2]
        st["f_inv"][:, g, :] = nonneg(rng.normal(mu, sg, (E, F)))
    st["f_prod_tfp"][:] = nonneg(rng.normal(1.0, 0.2, (E, G, F)))
    fs = np.empty((E, G, G + 1, F))
    fs[:, :, 0, :] = rng.normal(0.4, 0.05, (E, G, F))
    for g in range(G):
        mu, sg = ((0.1, 0.02), (0.4, 0.02))[g % 2]
        fs[:, :, g + 1, :] = rng.normal(mu, sg, (E, G, F))
    fsum = np.zeros((E, G, F))
    for i in range(G + 1):
        fsum = fsum + fs[:, :, i, :]
    st["f_prod_share"][:] = fs / fsum[:, :, None, :]
    fel = rng.normal(10.0, 2.5, (E, G, F))
    fel = np.where(fel <= 0, 1e-8, fel)
    st["f_prod_rho"][:] = 1 / (1 - fel)
    return st


class OrderStream:
    """Per-economy cumulative std::shuffle orders, one call per step (host libstdc++)."""

    def __init__(self, dims, seed):
        self.dims = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
        E, P, F, G, S = self.dims.tuple
        self.seed = int(seed) & 0xFFFFFFFF
        self.rng_state = np.zeros(E, dtype=np.uint64)
        self.perm_person = np.zeros((E, P), dtype=np.int32)
        self.perm_firm = np.zeros((E, F), dtype=np.int32)
        self.first = True

    def next(self):
        L = lib.load()
        lib.check(L.fastace_shuffle_orders(
            C.byref(self.dims), self.seed, self.rng_state.ctypes.data_as(C.POINTER(C.c_uint64)),
            self.perm_person.ctypes.data_as(C.POINTER(C.c_int32)),
            self.perm_firm.ctypes.data_as(C.POINTER(C.c_int32)), 1 if self.first else 0))
        self.first = False
        return self.perm_person.copy(), self.perm_firm.copy()


class DeviceOrderStream:
    """The same stream of visiting orders generated ON THE DEVICE by the env (fastace_env_shuffle_orders: one thread
    per economy replays minstd_rand0 and libstdc++'s std::shuffle, bit for bit) — nothing crosses the bus.  next()
    returns (perm_person, perm_firm) as int32 CUDA tensors [E][P] / [E][F]; orders are produced `chunk` steps ahead."""

    def __init__(self, env, seed, chunk=32):
        self.env, self.seed, self.chunk = env, int(seed) & 0xFFFFFFFF, int(chunk)
        self.started = False
        self.buf = None
        self.pos = 0

    def next(self):
        if self.buf is None or self.pos == self.chunk:
            self.buf = self.env.shuffle_orders(seed=self.seed, restart=not self.started, steps=self.chunk)
            self.started = True
            self.pos = 0
        k = self.pos
        self.pos += 1
        return self.buf[0][k], self.buf[1][k]


# Injected-action distributions that keep both books non-empty over a 40-step episode of
# the default scenario (with SURVEY.md's untuned recipe firms go bankrupt and the goods
# market is empty after ~12 steps, so a benchmark would time an idle market).  Measured
# with the CPU oracle: 9-16 goods offers, 10 job offers, ~170 hires and 13-140 purchases
# per economy-step throughout the episode.
BENCH_PRESET = dict(take_prob=0.5, prod_scale=0.5, wage_scale=0.1, price_scale=1.5, labor_mu=2.3)


def synthetic_actions(dims, seed, step, take_prob=0.5, wage_scale=1.0, price_scale=1.0,
                      labor_mu=1.0, prod_scale=1.0, perms=None):
    """One step of injected decisions:
    indices = raw uniform draws (to be used with IDX_MODULO), take ~ Bernoulli(take_prob),
    proportions ~ U(0,1) float32, price ~ LogNormal(0,0.5), labour ~ LogNormal(labor_mu,0.5),
    wage ~ LogNormal(0,0.5).  Philox keyed by (seed, step) so any step can be regenerated.
    `perms` = (perm_person, perm_firm) or None for identity orders."""
    E, P, F, G, S = dims.tuple if isinstance(dims, _abi.Dims) else tuple(dims)
    rng = np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(step)]))
    a = {}
    if perms is None:
        a["perm_person"] = np.ascontiguousarray(np.broadcast_to(np.arange(P, dtype=np.int32), (E, P)))
        a["perm_firm"] = np.ascontiguousarray(np.broadcast_to(np.arange(F, dtype=np.int32), (E, F)))
    else:
        a["perm_person"] = np.ascontiguousarray(perms[0], dtype=np.int32)
        a["perm_firm"] = np.ascontiguousarray(perms[1], dtype=np.int32)
    a["p_job_idx"] = rng.integers(0, 2**31 - 1, (E, S, P), dtype=np.int32)
    a["p_job_take"] = (rng.random((E, S, P)) < take_prob).astype(np.uint8)
    a["p_good_idx"] = rng.integers(0, 2**31 - 1, (E, S, P), dtype=np.int32)
    a["p_good_take"] = (rng.random((E, S, P)) < take_prob).astype(np.uint8)
    a["p_consume"] = rng.random((E, G, P), dtype=np.float32)
    a["f_good_idx"] = rng.integers(0, 2**31 - 1, (E, S, F), dtype=np.int32)
    a["f_good_take"] = (rng.random((E, S, F)) < take_prob).astype(np.uint8)
    a["f_prod"] = rng.random((E, G, F), dtype=np.float32) * np.float32(prod_scale)
    a["f_offer_amt"] = rng.random((E, G, F), dtype=np.float32)
    a["f_offer_price"] = (price_scale * np.exp(rng.normal(0.0, 0.5, (E, G, F)))).astype(np.float32)
    a["f_job_labor"] = np.exp(rng.normal(labor_mu, 0.5, (E, F))).astype(np.float32)
    a["f_job_wage"] = (wage_scale * np.exp(rng.normal(0.0, 0.5, (E, F)))).astype(np.float32)
    return a
