"""BatchedEconomy — host-side mirror of the reference's ``Economy`` for E economies.

Same vocabulary as the reference (``time_step``, ``get_time``, ``get_market``,
``get_jobMarket``, ``get_persons`` / ``get_firms`` state; /root/reference/src/base/base.h:83-132)
but every call acts on all economies of the env.  All compute happens inside
libfastace_b200.so (hand-written sm_100a kernels); this file only moves pointers.
PyTorch is used for device memory of action/output tensors and for streams.

There is no CPU path: constructing a BatchedEconomy without a CUDA device raises.
"""
import ctypes as C

import numpy as np

from . import _abi, lib

_TORCH_DTYPES = None


def _torch():
    import torch
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32,
                         np.uint32: torch.int32, np.uint8: torch.uint8}
    return torch


class BatchedEconomy:
    """E independent economies resident on one GPU.

    dims = (E, P, F, G, S).  State lives in buffers owned by the native env; actions and
    outputs are caller-provided device tensors (``time_step``) or host numpy arrays
    (``time_step_host``).
    """

    def __init__(self, dims, device=0):
        self.dims = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
        self._lib = lib.load()
        h = C.c_void_p()
        lib.check(self._lib.fastace_env_create(C.byref(self.dims), int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    # ---- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fastace_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Economy read API --------------------------------------------------------------
    def get_time(self):
        t = C.c_uint32()
        lib.check(self._lib.fastace_env_time(self._h, C.byref(t)))
        return t.value

    def get_numGoods(self):
        return self.dims.num_goods

    def set_function_kinds(self, util_kind=_abi.FN_CES, prod_kind=_abi.FN_CES):
        """VecToScalar family of the persons' utility / the firms' per-good production functions."""
        lib.check(self._lib.fastace_env_set_function_kinds(self._h, int(util_kind), int(prod_kind)))

    def launch_count(self):
        n = C.c_uint64()
        lib.check(self._lib.fastace_env_launch_count(self._h, C.byref(n)))
        return n.value

    def kernel_times(self):
        """(match_ms, update_ms, steps) accumulated by steps run with STEP_PROFILE; resets."""
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        lib.check(self._lib.fastace_env_kernel_times(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def set_state(self, state, time=0):
        """Load a host state dict (numpy arrays in the layout of include/fastace_b200.h)."""
        st = _abi.struct_from_numpy("state", state, self.dims)
        lib.check(self._lib.fastace_env_set_state(self._h, C.byref(st), int(time)))

    def set_time(self, time):
        """Set the economy clock only (a state struct with every member NULL)."""
        st = _abi.State()
        lib.check(self._lib.fastace_env_set_state(self._h, C.byref(st), int(time)))

    def get_state(self, names=None):
        """Copy the state back to host numpy arrays (all fields, or the named ones)."""
        state = _abi.alloc_host("state", self.dims, names)
        st = _abi.struct_from_numpy("state", state, self.dims)
        lib.check(self._lib.fastace_env_get_state(self._h, C.byref(st)))
        return state

    def device_state_pointers(self):
        """{name: device address} of the env's own buffers (zero-copy sharing)."""
        st = _abi.State()
        lib.check(self._lib.fastace_env_device_state(self._h, C.byref(st)))
        return {n: C.cast(getattr(st, n), C.c_void_p).value for n, _, _, _ in _abi.STATE_FIELDS}

    def device_state_tensors(self):
        """Zero-copy torch views of the env's state buffers (for a policy network that
        reads agent state in place).  uint32 fields are exposed as int32."""
        torch = _torch()
        ptrs = self.device_state_pointers()
        out = {}
        for n, (dt, shp) in _abi.shapes("state", self.dims).items():
            out[n] = _CudaView(ptrs[n], shp, dt, self.device).tensor()
        return out

    # ---- the hot path --------------------------------------------------------------------
    def alloc_actions(self, host_actions=None):
        """Device tensors for one step's actions (optionally filled from a host dict)."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        out = {}
        for n, (dt, shp) in _abi.shapes("actions", self.dims).items():
            if host_actions is not None:
                out[n] = torch.from_numpy(np.ascontiguousarray(host_actions[n])).to(dev)
            else:
                out[n] = torch.zeros(shp, dtype=_TORCH_DTYPES[dt], device=dev)
        return out

    def alloc_compact_actions(self, host_compact):
        """Device tensors (uint16 viewed as int16) for a compact action dict."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        out = {}
        for n, (dt, shp) in _abi.shapes("compact", self.dims).items():
            a = np.ascontiguousarray(host_compact[n])
            if a.dtype == np.uint16:
                a = a.view(np.int16)
            out[n] = torch.from_numpy(a).to(dev)
        return out

    def alloc_outputs(self, names=_abi.OUT_MANDATORY):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        out = {}
        for n, (dt, shp) in _abi.shapes("out", self.dims).items():
            if names is None or n in names:
                out[n] = torch.zeros(shp, dtype=_TORCH_DTYPES[dt], device=dev)
        return out

    def pack_device(self, kind, tensors):
        """ctypes struct of raw device pointers for a dict of torch CUDA tensors."""
        shp = _abi.shapes(kind, self.dims)
        ptrs = {}
        for n, t in tensors.items():
            if t is None:
                continue
            if not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != tuple(shp[n][1]):
                raise ValueError(f"{n}: need a contiguous CUDA tensor of shape {shp[n][1]}")
            if t.element_size() != np.dtype(shp[n][0]).itemsize:
                raise ValueError(f"{n}: wrong element size")
            ptrs[n] = t.data_ptr()
        s = _abi.struct_from_pointers(kind, ptrs)
        s._keepalive = tensors
        return s

    def time_step(self, actions, out, flags=_abi.IDX_ABSOLUTE, stream=None):
        """Economy::time_step for all economies.  `actions` / `out` are dicts of torch CUDA
        tensors or pre-packed structs (pack_device).  Asynchronous on `stream`
        (default: torch's current stream)."""
        ou = out if isinstance(out, _abi.StepOut) else self.pack_device("out", out)
        if stream is None:
            stream = _torch().cuda.current_stream(self.device).cuda_stream
        if isinstance(actions, _abi.ActionsCompact):
            lib.check(self._lib.fastace_env_step_device_compact(self._h, C.byref(actions), C.byref(ou), int(flags), C.c_void_p(stream)))
            return
        ac = actions if isinstance(actions, _abi.Actions) else self.pack_device("actions", actions)
        lib.check(self._lib.fastace_env_step_device(self._h, C.byref(ac), C.byref(ou), int(flags), C.c_void_p(stream)))

    def time_step_host(self, actions, out, flags=_abi.IDX_ABSOLUTE):
        """Same with host numpy arrays (or pre-built structs of host pointers): copies in,
        steps, copies out, synchronises."""
        ou = out if isinstance(out, _abi.StepOut) else _abi.struct_from_numpy("out", out, self.dims)
        if isinstance(actions, _abi.ActionsCompact):
            lib.check(self._lib.fastace_env_step_host_compact(self._h, C.byref(actions), C.byref(ou), int(flags)))
            return
        if isinstance(actions, _abi.ActionsPacked):
            lib.check(self._lib.fastace_env_step_host_packed(self._h, C.byref(actions), C.byref(ou), int(flags)))
            return
        ac = actions if isinstance(actions, _abi.Actions) else _abi.struct_from_numpy("actions", actions, self.dims)
        lib.check(self._lib.fastace_env_step_host(self._h, C.byref(ac), C.byref(ou), int(flags)))

    def restart_orders(self, seed):
        """(Re)start the env's own visiting orders: economy e seeds minstd_rand0(seed + e), identity order.  The packed
        host calls without orders then consume one std::shuffle step each."""
        lib.check(self._lib.fastace_env_shuffle_orders(self._h, int(seed) & 0xFFFFFFFF, 1, 0, None, None, None, None, None))

    def shuffle_orders(self, seed=0, restart=False, steps=1, perm_person=None, perm_firm=None, stream=None):
        """Economy::time_step's visiting orders for the next `steps` steps, generated ON THE DEVICE (bit-identical to
        std::shuffle with each economy's minstd_rand0; economy.cpp:110-111).  `perm_person` / `perm_firm`: torch CUDA
        tensors [steps][E][P] / [steps][E][F], int32 or int16 (= the compact encoding's 16-bit orders); allocated
        (int32) when None.  Returns (perm_person, perm_firm)."""
        torch = _torch()
        E, P, F = self.dims.num_econ, self.dims.num_persons, self.dims.num_firms
        dev = torch.device("cuda", self.device)
        if perm_person is None:
            perm_person = torch.empty((steps, E, P), dtype=torch.int32, device=dev)
        if perm_firm is None:
            perm_firm = torch.empty((steps, E, F), dtype=torch.int32, device=dev)
        for t, n in ((perm_person, steps * E * P), (perm_firm, steps * E * F)):
            if not t.is_cuda or not t.is_contiguous() or t.numel() != n or t.dtype not in (torch.int32, torch.int16):
                raise ValueError("orders: need contiguous CUDA int32 / int16 tensors of [steps][E][agents]")
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        p32 = lambda t: C.c_void_p(t.data_ptr()) if t.dtype == torch.int32 else None
        p16 = lambda t: C.c_void_p(t.data_ptr()) if t.dtype == torch.int16 else None
        lib.check(self._lib.fastace_env_shuffle_orders(self._h, int(seed) & 0xFFFFFFFF, 1 if restart else 0, int(steps),
                                                       p32(perm_person), p32(perm_firm), p16(perm_person), p16(perm_firm),
                                                       C.c_void_p(stream)))
        return perm_person, perm_firm

    def market_stats(self, stream=None):
        """Per-economy statistics of the current books, reduced on the device (print_info of src/pybindings.cpp:20-75):
        dict of torch CUDA tensors — sum_quantity_per_price [E][G], offers [E][G], lots [E][G], sum_wage_per_labor [E],
        job_offers [E], job_lots [E]; average price of good g = sum_quantity_per_price / offers."""
        torch = _torch()
        E, G = self.dims.num_econ, self.dims.num_goods
        dev = torch.device("cuda", self.device)
        t = {"sum_quantity_per_price": torch.empty((E, G), dtype=torch.float64, device=dev),
             "offers": torch.empty((E, G), dtype=torch.int32, device=dev), "lots": torch.empty((E, G), dtype=torch.int32, device=dev),
             "sum_wage_per_labor": torch.empty((E,), dtype=torch.float64, device=dev),
             "job_offers": torch.empty((E,), dtype=torch.int32, device=dev), "job_lots": torch.empty((E,), dtype=torch.int32, device=dev)}
        ms = _abi.MarketStats()
        for k, v in t.items():
            setattr(ms, k, C.cast(C.c_void_p(v.data_ptr()), dict(_abi.MarketStats._fields_)[k]))
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        lib.check(self._lib.fastace_env_market_stats(self._h, C.byref(ms), C.c_void_p(stream)))
        return t

    def large_stats(self):
        """(person-phase rounds, firm-phase rounds) of the last large-economy step"""
        a, b = C.c_uint32(0), C.c_uint32(0)
        lib.check(self._lib.fastace_env_large_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sync(self):
        """Wait for every step enqueued by time_step_host(..., flags | STEP_ASYNC)."""
        lib.check(self._lib.fastace_env_sync(self._h))


class _CudaView:
    """Exposes externally owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr, shape, dtype, device):
        dt = np.dtype(dtype)
        if dt == np.uint32:
            dt = np.dtype(np.int32)
        self.__cuda_array_interface__ = {
            "shape": tuple(int(x) for x in shape), "typestr": dt.str, "data": (int(ptr), False),
            "version": 2, "strides": None,
        }
        self._device = device

    def tensor(self):
        torch = _torch()
        return torch.as_tensor(self, device=torch.device("cuda", self._device))
