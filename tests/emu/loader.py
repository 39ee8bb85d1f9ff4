"""TEST INFRASTRUCTURE ONLY — ctypes face of tests/emu/libfastace_emu.so (the CUDA step kernels' source run on the
CPU under the SIMT emulator).  Same calling shape as oracle.loader.Oracle.step: host numpy dicts, stepped in place."""
import ctypes as C

from fastace_b200 import _abi

from . import build_emu


class EmuKernels:
    def __init__(self):
        self.lib = C.CDLL(build_emu.build())
        L = self.lib
        L.fastace_emu_create.restype = C.c_void_p
        L.fastace_emu_create.argtypes = [C.POINTER(_abi.Dims)]
        L.fastace_emu_destroy.argtypes = [C.c_void_p]
        L.fastace_emu_step.restype = C.c_int
        L.fastace_emu_step.argtypes = [C.c_void_p, C.POINTER(_abi.Dims), C.POINTER(_abi.State), C.c_void_p, C.c_void_p,
                                       C.POINTER(_abi.StepOut), C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        self.handles = {}

    def _handle(self, d):
        key = d.tuple
        if key not in self.handles:
            self.handles[key] = self.lib.fastace_emu_create(C.byref(d))
        return self.handles[key]

    def step(self, dims, state, actions, out, flags=0, time_before=0, compact=None, util_kind=0, prod_kind=0):
        """In-place step of the host numpy dict `state`; `actions` (int32 encoding) or `compact`."""
        d = _abi.make_dims(*dims) if not isinstance(dims, _abi.Dims) else dims
        st = _abi.struct_from_numpy("state", state, d)
        ou = _abi.struct_from_numpy("out", out, d)
        if compact is not None:
            cz = _abi.struct_from_numpy("compact", compact, d)
            rc = self.lib.fastace_emu_step(self._handle(d), C.byref(d), C.byref(st), None, C.byref(cz), C.byref(ou), flags, time_before,
                                           util_kind, prod_kind)
        else:
            ac = _abi.struct_from_numpy("actions", actions, d)
            rc = self.lib.fastace_emu_step(self._handle(d), C.byref(d), C.byref(st), C.byref(ac), None, C.byref(ou), flags, time_before,
                                           util_kind, prod_kind)
        if rc != 0:
            raise RuntimeError(f"emulated kernels raised error words {rc}")

    STAT_NAMES = ("windows", "rounds", "rescans", "risky_walks", "sales_windows", "dead_exits", "firm_serial",
                  "rounds_w0", "rounds_w1", "rounds_w2", "rounds_w3", "rescans_w0", "risky_coop", "risky_kept")

    def stats(self):
        """event counters of the kernels since the last call"""
        buf = (C.c_ulonglong * 16)()
        self.lib.fastace_emu_stats(buf)
        return dict(zip(self.STAT_NAMES, list(buf)))

    def block_order(self, order):
        """0 ascending (default), 1 descending, 2 scrambled: the order in which the blocks of every launch run"""
        self.lib.fastace_emu_block_order(int(order))

    def close(self):
        for h in self.handles.values():
            self.lib.fastace_emu_destroy(h)
        self.handles = {}
