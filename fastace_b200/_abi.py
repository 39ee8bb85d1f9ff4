"""ctypes mirror of include/fastace_b200.h.

The reference binds its native library with plain ``ctypes`` over ``extern "C"``
(/root/reference/py/main.py:10-126); this module does the same for the B200 library.
It holds only struct layouts and array shape tables — no compute.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 3
FASTACE_OK = 0
IDX_ABSOLUTE = 0
IDX_MODULO = 1
STEP_SERIAL = 2
STEP_PROFILE = 4
STEP_ASYNC = 8
STEP_LARGE = 16
STEP_PERSONS = 32     # person phase only; STEP_FIRMS must follow and completes the step
STEP_FIRMS = 64
STEP_PERSONS_TRADE = 128      # job search + purchases; STEP_PERSONS_CONSUME must follow
STEP_PERSONS_CONSUME = 256
MAX_GOODS = 8
FN_CES, FN_COBB_DOUGLAS, FN_STONE_GEARY, FN_LEONTIEF, FN_LINEAR = range(5)
MAX_STACK = 16

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(C.c_uint8)
_hp = C.POINTER(C.c_uint16)


class Dims(C.Structure):
    _fields_ = [
        ("num_econ", C.c_int32),
        ("num_persons", C.c_int32),
        ("num_firms", C.c_int32),
        ("num_goods", C.c_int32),
        ("stack_size", C.c_int32),
    ]

    @property
    def tuple(self):
        return (self.num_econ, self.num_persons, self.num_firms, self.num_goods, self.stack_size)


def make_dims(E, P, F, G, S):
    return Dims(int(E), int(P), int(F), int(G), int(S))


# name -> (ctypes pointer type, numpy dtype, shape function of (E,P,F,G,S))
STATE_FIELDS = [
    ("p_money", _dp, np.float64, lambda E, P, F, G, S: (E, P)),
    ("p_inv", _dp, np.float64, lambda E, P, F, G, S: (E, G, P)),
    ("p_labor", _dp, np.float64, lambda E, P, F, G, S: (E, P)),
    ("p_util_tfp", _dp, np.float64, lambda E, P, F, G, S: (E, P)),
    ("p_util_share", _dp, np.float64, lambda E, P, F, G, S: (E, G + 1, P)),
    ("p_util_rho", _dp, np.float64, lambda E, P, F, G, S: (E, P)),
    ("f_money", _dp, np.float64, lambda E, P, F, G, S: (E, F)),
    ("f_inv", _dp, np.float64, lambda E, P, F, G, S: (E, G, F)),
    ("f_labor", _dp, np.float64, lambda E, P, F, G, S: (E, F)),
    ("f_last_money", _dp, np.float64, lambda E, P, F, G, S: (E, F)),
    ("f_prod_tfp", _dp, np.float64, lambda E, P, F, G, S: (E, G, F)),
    ("f_prod_share", _dp, np.float64, lambda E, P, F, G, S: (E, G, G + 1, F)),
    ("f_prod_rho", _dp, np.float64, lambda E, P, F, G, S: (E, G, F)),
    ("m_count", _ip, np.int32, lambda E, P, F, G, S: (E,)),
    ("m_owner", _ip, np.int32, lambda E, P, F, G, S: (E, F * G)),
    ("m_good", _ip, np.int32, lambda E, P, F, G, S: (E, F * G)),
    ("m_left", _up, np.uint32, lambda E, P, F, G, S: (E, F * G)),
    ("m_taken", _up, np.uint32, lambda E, P, F, G, S: (E, F * G)),
    ("m_price", _dp, np.float64, lambda E, P, F, G, S: (E, F * G)),
    ("j_count", _ip, np.int32, lambda E, P, F, G, S: (E,)),
    ("j_owner", _ip, np.int32, lambda E, P, F, G, S: (E, F)),
    ("j_left", _up, np.uint32, lambda E, P, F, G, S: (E, F)),
    ("j_taken", _up, np.uint32, lambda E, P, F, G, S: (E, F)),
    ("j_wage", _dp, np.float64, lambda E, P, F, G, S: (E, F)),
    ("p_util_theta", _dp, np.float64, lambda E, P, F, G, S: (E, G + 1, P)),
    ("f_prod_theta", _dp, np.float64, lambda E, P, F, G, S: (E, G, G + 1, F)),
]

ACTION_FIELDS = [
    ("perm_person", _ip, np.int32, lambda E, P, F, G, S: (E, P)),
    ("perm_firm", _ip, np.int32, lambda E, P, F, G, S: (E, F)),
    ("p_job_idx", _ip, np.int32, lambda E, P, F, G, S: (E, S, P)),
    ("p_job_take", _bp, np.uint8, lambda E, P, F, G, S: (E, S, P)),
    ("p_good_idx", _ip, np.int32, lambda E, P, F, G, S: (E, S, P)),
    ("p_good_take", _bp, np.uint8, lambda E, P, F, G, S: (E, S, P)),
    ("p_consume", _fp, np.float32, lambda E, P, F, G, S: (E, G, P)),
    ("f_good_idx", _ip, np.int32, lambda E, P, F, G, S: (E, S, F)),
    ("f_good_take", _bp, np.uint8, lambda E, P, F, G, S: (E, S, F)),
    ("f_prod", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_amt", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_price", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_job_labor", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
    ("f_job_wage", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
]

# fastace_actions_compact_t: one-byte indices, bit-mask takes, 16-bit orders
COMPACT_FIELDS = [
    ("perm_person", _hp, np.uint16, lambda E, P, F, G, S: (E, P)),
    ("perm_firm", _hp, np.uint16, lambda E, P, F, G, S: (E, F)),
    ("p_job_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, P, S)),      # agent-major: one agent's S bytes are contiguous
    ("p_job_take", _hp, np.uint16, lambda E, P, F, G, S: (E, P)),
    ("p_good_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, P, S)),
    ("p_good_take", _hp, np.uint16, lambda E, P, F, G, S: (E, P)),
    ("p_consume", _fp, np.float32, lambda E, P, F, G, S: (E, G, P)),
    ("f_good_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, F, S)),
    ("f_good_take", _hp, np.uint16, lambda E, P, F, G, S: (E, F)),
    ("f_prod", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_amt", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_price", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_job_labor", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
    ("f_job_wage", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
]

def packed_layout(F, G, S):
    """(bits_job, bytes_job, bits_good, bytes_good) of the packed host encoding (fastace_packed_layout): the smallest
    field width whose all-ones value is not a valid index."""
    def bits_for(n):
        b = 1
        while (1 << b) - 1 < n:
            b += 1
        return b
    bj, bg = bits_for(F), bits_for(F * G)
    return bj, (S * bj + 7) // 8, bg, (S * bg + 7) // 8


# fastace_actions_packed_t: S bit fields per agent, all-ones = no request, no take masks; orders optional
PACKED_FIELDS = [
    ("perm_person", _hp, np.uint16, lambda E, P, F, G, S: (E, P)),
    ("perm_firm", _hp, np.uint16, lambda E, P, F, G, S: (E, F)),
    ("p_job_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, P, packed_layout(F, G, S)[1])),
    ("p_good_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, P, packed_layout(F, G, S)[3])),
    ("p_consume", _fp, np.float32, lambda E, P, F, G, S: (E, G, P)),
    ("f_good_idx", _bp, np.uint8, lambda E, P, F, G, S: (E, F, packed_layout(F, G, S)[3])),
    ("f_prod", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_amt", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_offer_price", _fp, np.float32, lambda E, P, F, G, S: (E, G, F)),
    ("f_job_labor", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
    ("f_job_wage", _fp, np.float32, lambda E, P, F, G, S: (E, F)),
]

OUT_FIELDS = [
    ("p_reward", _dp, np.float64, lambda E, P, F, G, S: (E, P)),
    ("f_profit", _dp, np.float64, lambda E, P, F, G, S: (E, F)),
    ("p_job_ok", _bp, np.uint8, lambda E, P, F, G, S: (E, S, P)),
    ("p_good_ok", _bp, np.uint8, lambda E, P, F, G, S: (E, S, P)),
    ("f_good_ok", _bp, np.uint8, lambda E, P, F, G, S: (E, S, F)),
    ("old_m_left", _up, np.uint32, lambda E, P, F, G, S: (E, F * G)),
    ("old_m_taken", _up, np.uint32, lambda E, P, F, G, S: (E, F * G)),
    ("old_j_left", _up, np.uint32, lambda E, P, F, G, S: (E, F)),
    ("old_j_taken", _up, np.uint32, lambda E, P, F, G, S: (E, F)),
]
OUT_MANDATORY = ("p_reward", "f_profit")


class State(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in STATE_FIELDS]


class Actions(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in ACTION_FIELDS]


class ActionsCompact(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in COMPACT_FIELDS]


class ActionsPacked(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in PACKED_FIELDS]


class StepOut(C.Structure):
    _fields_ = [(n, t) for n, t, _, _ in OUT_FIELDS]


class MarketStats(C.Structure):
    """fastace_market_stats_t"""
    _fields_ = [("sum_quantity_per_price", _dp), ("offers", _up), ("lots", _up),
                ("sum_wage_per_labor", _dp), ("job_offers", _up), ("job_lots", _up)]


class CustomScenarioParams(C.Structure):
    """neural::CustomScenarioParams, /root/reference/src/neural/neuralScenarios.h:49-115
    (ctypes mirror as in /root/reference/py/main.py:12-60)."""

    _fields_ = [("numPeople", C.c_uint), ("numFirms", C.c_uint)] + [
        (n, C.c_double)
        for n in (
            "money_mu money_sigma good1_mu good1_sigma good2_mu good2_sigma "
            "labor_share_mu labor_share_sigma good1_share_mu good1_share_sigma good2_share_mu good2_share_sigma "
            "discount_mu discount_sigma elasticity_mu elasticity_sigma "
            "firm_money_mu firm_money_sigma firm_good1_mu firm_good1_sigma firm_good2_mu firm_good2_sigma "
            "firm_tfp1_mu firm_tfp1_sigma firm_tfp2_mu firm_tfp2_sigma "
            "firm_labor_share1_mu firm_labor_share1_sigma firm_good1_share1_mu firm_good1_share1_sigma "
            "firm_good2_share1_mu firm_good2_share1_sigma "
            "firm_labor_share2_mu firm_labor_share2_sigma firm_good1_share2_mu firm_good1_share2_sigma "
            "firm_good2_share2_mu firm_good2_share2_sigma "
            "firm_elasticity1_mu firm_elasticity1_sigma firm_elasticity2_mu firm_elasticity2_sigma"
        ).split()
    ]


class TrainingParams(C.Structure):
    """neural::TrainingParams, /root/reference/src/neural/neuralScenarios.h:148-186
    (ctypes mirror as in /root/reference/py/main.py:63-85)."""

    _fields_ = (
        [(n, C.c_uint) for n in "numEpisodes episodeLength updateEveryNEpisodes checkpointEveryNEpisodes "
                                "stackSize encodingSize hiddenSize nHidden nHiddenSmall".split()]
        + [(n, C.c_double) for n in "purchaseNetLR firmPurchaseNetLR laborSearchNetLR consumptionNetLR "
                                    "productionNetLR offerNetLR jobOfferNetLR valueNetLR firmValueNetLR".split()]
        + [("episodeBatchSizeForLRDecay", C.c_uint), ("patienceForLRDecay", C.c_uint),
           ("multiplierForLRDecay", C.c_double), ("reverseAnnealingPeriod", C.c_uint)]
    )


def field_table(kind):
    return {"state": STATE_FIELDS, "actions": ACTION_FIELDS, "compact": COMPACT_FIELDS, "packed": PACKED_FIELDS, "out": OUT_FIELDS}[kind]


def shapes(kind, dims):
    """{name: (numpy dtype, shape)} for a Dims/tuple (E,P,F,G,S)."""
    t = dims.tuple if isinstance(dims, Dims) else tuple(dims)
    return {n: (dt, fn(*t)) for n, _, dt, fn in field_table(kind)}


def alloc_host(kind, dims, names=None):
    """dict of zero-filled C-contiguous numpy arrays for the given struct."""
    out = {}
    for n, (dt, shp) in shapes(kind, dims).items():
        if names is None or n in names:
            out[n] = np.zeros(shp, dtype=dt)
    return out


def alloc_host_block(kind, dims, names=None, pinned=False):
    """Like alloc_host, but all arrays are views into ONE buffer laid out as the library's device staging block
    (fields in struct order, each padded to 256 B): fastace_env_step_host* then moves the whole struct with a single
    host<->device copy.  pinned=True page-locks the buffer (through torch).  Returns (dict of arrays, buffer)."""
    table = [(n, dt, shp) for n, (dt, shp) in shapes(kind, dims).items() if names is None or n in names]
    offs, total = [], 0
    for n, dt, shp in table:
        offs.append(total)
        total += (int(np.prod(shp)) * np.dtype(dt).itemsize + 255) // 256 * 256
    if pinned:
        import torch
        holder = torch.zeros(max(total, 1) + 256, dtype=torch.uint8).pin_memory()
        buf = holder.numpy()
    else:
        holder = buf = np.zeros(max(total, 1) + 256, dtype=np.uint8)
    base = (-buf.ctypes.data) % 256          # 256 B aligned start
    out = {}
    for (n, dt, shp), o in zip(table, offs):
        nbytes = int(np.prod(shp)) * np.dtype(dt).itemsize
        out[n] = buf[base + o: base + o + nbytes].view(dt).reshape(shp)
    return out, holder


def struct_from_numpy(kind, arrays, dims=None):
    """Build the ctypes struct from a dict of numpy arrays (missing names -> NULL).
    Arrays must be C-contiguous with the exact dtype; shapes are checked when dims given."""
    cls = {"state": State, "actions": Actions, "compact": ActionsCompact, "packed": ActionsPacked, "out": StepOut}[kind]
    s = cls()
    shp = shapes(kind, dims) if dims is not None else None
    for n, ptr_t, dt, _ in field_table(kind):
        a = arrays.get(n)
        if a is None:
            continue
        if not isinstance(a, np.ndarray) or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{n}: need C-contiguous numpy array of dtype {np.dtype(dt)}")
        if shp is not None and tuple(a.shape) != tuple(shp[n][1]):
            raise ValueError(f"{n}: shape {a.shape} != expected {shp[n][1]}")
        setattr(s, n, a.ctypes.data_as(ptr_t))
    s._keepalive = arrays  # keep the buffers alive as long as the struct
    return s


def struct_from_pointers(kind, ptrs):
    """Build the ctypes struct from a dict name -> integer address (device pointers)."""
    cls = {"state": State, "actions": Actions, "compact": ActionsCompact, "packed": ActionsPacked, "out": StepOut}[kind]
    s = cls()
    for n, ptr_t, _, _ in field_table(kind):
        p = ptrs.get(n)
        if p:
            setattr(s, n, C.cast(C.c_void_p(int(p)), ptr_t))
    return s


def compact_actions_for_counts(actions, j_count, m_count, modulo):
    """Same decisions in the compact encoding, given the current book sizes ([E] arrays).
    modulo: the byte stored is (raw % count) so that byte % count == raw % count.
    absolute: out-of-range / negative indices become 255 (no request)."""
    a = actions
    E = a["perm_person"].shape[0]
    jc = np.asarray(j_count, dtype=np.int64).reshape(E, 1, 1)
    mc = np.asarray(m_count, dtype=np.int64).reshape(E, 1, 1)

    def idx8(raw, cnt):   # [E][S][N] int32 -> [E][N][S] u8 (agent-major)
        raw = raw.astype(np.int64)
        if modulo:
            u = raw & 0xFFFFFFFF
            b = np.where(cnt > 0, u % np.maximum(cnt, 1), 0).astype(np.uint8)
        else:
            ok = (raw >= 0) & (raw < cnt)
            b = np.where(ok, raw, 255).astype(np.uint8)
        return b.transpose(0, 2, 1)

    def mask16(take):  # [E][S][N] u8 -> [E][N] u16
        S = take.shape[1]
        w = (1 << np.arange(S, dtype=np.uint32)).reshape(1, S, 1)
        return ((take != 0).astype(np.uint32) * w).sum(axis=1).astype(np.uint16)

    out = {
        "perm_person": a["perm_person"].astype(np.uint16), "perm_firm": a["perm_firm"].astype(np.uint16),
        "p_job_idx": idx8(a["p_job_idx"], jc), "p_job_take": mask16(a["p_job_take"]),
        "p_good_idx": idx8(a["p_good_idx"], mc), "p_good_take": mask16(a["p_good_take"]),
        "f_good_idx": idx8(a["f_good_idx"], mc), "f_good_take": mask16(a["f_good_take"]),
    }
    for k in ("p_consume", "f_prod", "f_offer_amt", "f_offer_price", "f_job_labor", "f_job_wage"):
        out[k] = a[k]
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


def packed_actions_for_counts(actions, j_count, m_count, modulo, with_orders=True):
    """Same decisions in the packed host encoding (fastace_actions_packed_t), given the current book sizes: every request
    slot becomes one bit field holding the offer index (draw % count with `modulo`), all-ones where nothing is
    requested (not taken, or out of range).  with_orders=False leaves the visiting orders out (the env generates them)."""
    a = actions
    E, S, P = a["p_job_idx"].shape
    F = a["f_good_idx"].shape[2]
    G = a["p_consume"].shape[1]
    bj, nbj, bg, nbg = packed_layout(F, G, S)
    jc = np.asarray(j_count, dtype=np.int64).reshape(E, 1, 1)
    mc = np.asarray(m_count, dtype=np.int64).reshape(E, 1, 1)

    def fields(raw, take, cnt, bits, nbytes):     # [E][S][N] -> [E][N][nbytes]
        raw = raw.astype(np.int64)
        if modulo:
            v = np.where(cnt > 0, (raw & 0xFFFFFFFF) % np.maximum(cnt, 1), (1 << bits) - 1)
            ok = (take != 0) & (cnt > 0)
        else:
            v = raw
            ok = (take != 0) & (raw >= 0) & (raw < cnt)
        v = np.where(ok, v, (1 << bits) - 1).astype(np.uint64).transpose(0, 2, 1)          # [E][N][S]
        word = np.zeros(v.shape[:2], dtype=np.uint64)       # S * bits <= 128: two 64-bit halves
        word_hi = np.zeros(v.shape[:2], dtype=np.uint64)
        for i in range(S):
            sh = i * bits
            if sh < 64:
                word |= v[:, :, i] << np.uint64(sh)
                if sh + bits > 64:
                    word_hi |= v[:, :, i] >> np.uint64(64 - sh)
            else:
                word_hi |= v[:, :, i] << np.uint64(sh - 64)
        out = np.zeros(v.shape[:2] + (nbytes,), dtype=np.uint8)
        for k in range(nbytes):
            src = word if k < 8 else word_hi
            out[:, :, k] = ((src >> np.uint64(8 * (k % 8))) & np.uint64(0xFF)).astype(np.uint8)
        return out

    out = {
        "p_job_idx": fields(a["p_job_idx"], a["p_job_take"], jc, bj, nbj),
        "p_good_idx": fields(a["p_good_idx"], a["p_good_take"], mc, bg, nbg),
        "f_good_idx": fields(a["f_good_idx"], a["f_good_take"], mc, bg, nbg),
    }
    if with_orders:
        out["perm_person"] = a["perm_person"].astype(np.uint16)
        out["perm_firm"] = a["perm_firm"].astype(np.uint16)
    for k in ("p_consume", "f_prod", "f_offer_amt", "f_offer_price", "f_job_labor", "f_job_wage"):
        out[k] = a[k]
    return {k: np.ascontiguousarray(v) for k, v in out.items()}
