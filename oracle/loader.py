"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU checkers.

  * ``Oracle``    : oracle/liboracle.so, our C restatement (fastace_oracle.c)
  * ``Reference`` : oracle/_ref/libfastace_ref.so, the unmodified reference env sources +
                    replay harness (ref_harness.cpp); exists only where it was built
                    (this container) or travelled as a prebuilt file (the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package fastace_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fastace_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfastace_ref.so")


def build(quiet=True):
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have_reference():
    return os.path.exists(REF_SO)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.fastace_oracle_step.restype = C.c_int
        L.fastace_oracle_step.argtypes = [C.POINTER(_abi.Dims), C.POINTER(_abi.State), C.POINTER(_abi.Actions),
                                          C.POINTER(_abi.StepOut), C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.fastace_oracle_step_mt.restype = C.c_int
        L.fastace_oracle_step_mt.argtypes = [C.POINTER(_abi.Dims), C.POINTER(_abi.State), C.POINTER(_abi.Actions),
                                             C.POINTER(_abi.StepOut), C.c_uint32, C.c_uint32, C.c_int]
        L.fastace_oracle_ces_f.restype = C.c_double
        L.fastace_oracle_ces_f.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.fastace_oracle_ces_params.restype = None
        L.fastace_oracle_ces_params.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.fastace_oracle_cobb_douglas_f.restype = C.c_double
        L.fastace_oracle_cobb_douglas_f.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.fastace_oracle_set_function_kinds.restype = None
        L.fastace_oracle_set_function_kinds.argtypes = [C.c_int, C.c_int]
        L.fastace_oracle_function_f.restype = C.c_double
        L.fastace_oracle_function_f.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                                C.POINTER(C.c_double), C.c_int, C.c_int]
        L.fastace_oracle_double_to_int.restype = C.c_int32
        L.fastace_oracle_double_to_int.argtypes = [C.c_double]

    def step(self, dims, state, actions, out, flags=0, time_before=0, nthreads=1):
        """In-place step of host numpy dicts `state`; fills `out` dict."""
        d = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
        st = _abi.struct_from_numpy("state", state, d)
        ac = _abi.struct_from_numpy("actions", actions, d)
        ou = _abi.struct_from_numpy("out", out, d)
        if nthreads > 1:
            rc = self.lib.fastace_oracle_step_mt(C.byref(d), C.byref(st), C.byref(ac), C.byref(ou), flags, time_before, nthreads)
        else:
            rc = self.lib.fastace_oracle_step(C.byref(d), C.byref(st), C.byref(ac), C.byref(ou), flags, time_before, 0, d.num_econ)
        if rc != 0:
            raise RuntimeError("fastace_oracle_step failed")

    def set_function_kinds(self, util_kind=0, prod_kind=0):
        self.lib.fastace_oracle_set_function_kinds(int(util_kind), int(prod_kind))

    def function_f(self, kind, tfp, share, theta, rho, x):
        s = np.ascontiguousarray(share, dtype=np.float64)
        th = np.ascontiguousarray(theta if theta is not None else np.zeros_like(s), dtype=np.float64)
        q = np.ascontiguousarray(x, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        return self.lib.fastace_oracle_function_f(int(kind), tfp, s.ctypes.data_as(dp), th.ctypes.data_as(dp), rho,
                                                  q.ctypes.data_as(dp), len(q), 1)

    def ces_f(self, tfp, share_norm, rho, x):
        s = np.ascontiguousarray(share_norm, dtype=np.float64)
        q = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.fastace_oracle_ces_f(tfp, s.ctypes.data_as(C.POINTER(C.c_double)), rho,
                                             q.ctypes.data_as(C.POINTER(C.c_double)), len(q), 1)

    def ces_params(self, share_raw, elasticity):
        s = np.ascontiguousarray(share_raw, dtype=np.float64)
        o = np.zeros_like(s)
        rho = C.c_double()
        self.lib.fastace_oracle_ces_params(s.ctypes.data_as(C.POINTER(C.c_double)), elasticity, len(s),
                                           o.ctypes.data_as(C.POINTER(C.c_double)), C.byref(rho))
        return o, rho.value

    def cobb_douglas_f(self, tfp, elast, x):
        e = np.ascontiguousarray(elast, dtype=np.float64)
        q = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.fastace_oracle_cobb_douglas_f(tfp, e.ctypes.data_as(C.POINTER(C.c_double)),
                                                      q.ctypes.data_as(C.POINTER(C.c_double)), len(q))


class Reference:
    """The reference's own Economy objects behind the same array interface."""

    def __init__(self, dims, state, seed, util_kind=0, prod_kind=0):
        if not have_reference():
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.fastace_ref_set_function_kinds.argtypes = [C.c_int, C.c_int]
        L.fastace_ref_set_function_kinds(int(util_kind), int(prod_kind))
        L.fastace_ref_create.restype = C.c_void_p
        L.fastace_ref_create.argtypes = [C.POINTER(_abi.Dims), C.POINTER(_abi.State), C.c_uint32]
        L.fastace_ref_destroy.argtypes = [C.c_void_p]
        L.fastace_ref_step.restype = C.c_int
        L.fastace_ref_step.argtypes = [C.c_void_p, C.POINTER(_abi.Actions), C.POINTER(_abi.StepOut), C.c_uint32,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
        L.fastace_ref_get_state.restype = C.c_int
        L.fastace_ref_get_state.argtypes = [C.c_void_p, C.POINTER(_abi.State), C.POINTER(C.c_uint32)]
        self.dims = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
        st = _abi.struct_from_numpy("state", state, self.dims)
        self.h = L.fastace_ref_create(C.byref(self.dims), C.byref(st), seed)

    def close(self):
        if self.h:
            self.lib.fastace_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, actions, out, flags=0, nthreads=1, want_perms=True):
        """Steps every economy once.  Returns (perm_person, perm_firm) the reference used
        (None, None when want_perms is False: skips the harness's order export)."""
        E, P, F, G, S = self.dims.tuple
        ac = _abi.struct_from_numpy("actions", {k: v for k, v in actions.items() if k not in ("perm_person", "perm_firm")}, self.dims)
        ou = _abi.struct_from_numpy("out", out, self.dims)
        if not want_perms:
            rc = self.lib.fastace_ref_step(self.h, C.byref(ac), C.byref(ou), flags, None, None, nthreads)
            if rc != 0:
                raise RuntimeError("fastace_ref_step failed")
            return None, None
        pp = np.zeros((E, P), dtype=np.int32)
        pf = np.zeros((E, F), dtype=np.int32)
        rc = self.lib.fastace_ref_step(self.h, C.byref(ac), C.byref(ou), flags,
                                       pp.ctypes.data_as(C.POINTER(C.c_int32)), pf.ctypes.data_as(C.POINTER(C.c_int32)), nthreads)
        if rc != 0:
            raise RuntimeError("fastace_ref_step failed")
        return pp, pf

    def get_state(self):
        state = _abi.alloc_host("state", self.dims)
        st = _abi.struct_from_numpy("state", state, self.dims)
        t = C.c_uint32()
        self.lib.fastace_ref_get_state(self.h, C.byref(st), C.byref(t))
        return state, t.value


def ref_lib():
    """Raw handle for the KAT helpers (ces_f, ces_params, cobb_douglas_f, shuffle_kat)."""
    L = C.CDLL(REF_SO)
    L.fastace_ref_ces_f.restype = C.c_double
    L.fastace_ref_ces_f.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_int]
    L.fastace_ref_ces_params.restype = None
    L.fastace_ref_ces_params.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fastace_ref_cobb_douglas_f.restype = C.c_double
    L.fastace_ref_cobb_douglas_f.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    L.fastace_ref_shuffle_kat.restype = None
    L.fastace_ref_shuffle_kat.argtypes = [C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    return L
