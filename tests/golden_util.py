"""Loader for tests/golden/*.npz (vectors generated from the unmodified reference by
tests/golden/gen_golden.py)."""
import glob
import os

import numpy as np

from fastace_b200 import _abi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if os.path.basename(p) not in ("policy_nets.npz", "a2c_episode.npz"))   # step fixtures only


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.dims = tuple(int(x) for x in z["dims"])
        self.flags = int(z["flags"][0])
        self.steps = int(z["steps"][0])

    def initial_state(self):
        st = _abi.alloc_host("state", self.dims)
        for k in st:
            if f"init/{k}" in self.z.files:   # fields added later (StoneGeary thresholds) stay zero
                st[k][...] = self.z[f"init/{k}"]
        return st

    def actions(self, t):
        return {k: np.ascontiguousarray(self.z[f"act{t}/{k}"]) for k, _, _, _ in _abi.ACTION_FIELDS}

    def outputs(self, t):
        return {k: self.z[f"out{t}/{k}"] for k, _, _, _ in _abi.OUT_FIELDS if f"out{t}/{k}" in self.z.files}

    def state(self, t):
        return {k: self.z[f"state{t}/{k}"] for k, _, _, _ in _abi.STATE_FIELDS if f"state{t}/{k}" in self.z.files}


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_state_bits(got, want, dims, where=""):
    """Bit-exact comparison of every state field (books over their live prefix)."""
    E, P, F, G, S = dims
    for k, w in want.items():
        g = got[k]
        if k in ("m_owner", "m_good", "m_left", "m_taken", "m_price"):
            for e in range(E):
                n = int(want["m_count"][e])
                assert bits_equal(g[e, :n], w[e, :n]), f"{where} {k}[{e}] {g[e,:n]} != {w[e,:n]}"
        elif k in ("j_owner", "j_left", "j_taken", "j_wage"):
            for e in range(E):
                n = int(want["j_count"][e])
                assert bits_equal(g[e, :n], w[e, :n]), f"{where} {k}[{e}] {g[e,:n]} != {w[e,:n]}"
        else:
            assert bits_equal(g, w), f"{where} {k}: {np.abs(np.asarray(g, float) - np.asarray(w, float)).max()}"


def assert_out_bits(got, want, dims, before, where=""):
    E, P, F, G, S = dims
    for k, w in want.items():
        g = got[k]
        if k.startswith("old_m"):
            for e in range(E):
                n = int(before["m_count"][e])
                assert bits_equal(g[e, :n], w[e, :n]), f"{where} {k}[{e}] {g[e,:n]} != {w[e,:n]}"
        elif k.startswith("old_j"):
            for e in range(E):
                n = int(before["j_count"][e])
                assert bits_equal(g[e, :n], w[e, :n]), f"{where} {k}[{e}] {g[e,:n]} != {w[e,:n]}"
        else:
            assert bits_equal(g, w), f"{where} {k} differs"
