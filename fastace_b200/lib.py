"""ctypes binding of libfastace_b200.so (the C ABI in include/fastace_b200.h).

Mirrors how the reference binds its own library (``ctypes.CDLL("../bin/libpybindings.so")``,
/root/reference/py/main.py:10, prototypes :88-109).  There is no fallback: if the CUDA
library is missing this raises, and env creation raises without a GPU.
"""
import ctypes as C
import os

from . import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
# FASTACE_B200_LIB: load a variant build (profiling defines, other launch bounds) instead of the product library
LIB_PATH = os.environ.get("FASTACE_B200_LIB") or os.path.join(HERE, "libfastace_b200.so")

# every symbol include/fastace_b200.h declares
EXPORTED_SYMBOLS = [
    "fastace_abi_version", "fastace_last_error", "fastace_device_count",
    "fastace_env_create", "fastace_env_destroy", "fastace_env_dims", "fastace_env_time", "fastace_env_set_function_kinds",
    "fastace_env_set_state", "fastace_env_get_state", "fastace_env_device_state",
    "fastace_env_step_device", "fastace_env_step_host", "fastace_env_step_device_compact",
    "fastace_env_step_host_compact", "fastace_env_step_host_packed", "fastace_packed_layout", "fastace_env_sync", "fastace_env_launch_count", "fastace_env_kernel_times", "fastace_env_large_stats", "fastace_mlp_stack_layout", "fastace_mlp_residual_tanh_stack", "fastace_mlp_forward", "fastace_layer_forward", "fastace_layer_backward", "fastace_policy_bernoulli", "fastace_policy_normal",
    "create_scenario_params", "create_training_params", "run", "train",
    "fastace_scenario_custom_init", "fastace_shuffle_orders", "fastace_env_shuffle_orders", "fastace_env_market_stats",
]


class FastaceError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library (never builds, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FastaceError(
            f"{LIB_PATH} is missing: build it with `python -m fastace_b200.build` "
            "(or __graft_entry__.build()).  fastace_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.fastace_abi_version.restype = C.c_int
    L.fastace_last_error.restype = C.c_char_p
    L.fastace_device_count.restype = C.c_int
    L.fastace_env_create.restype = C.c_int
    L.fastace_env_create.argtypes = [C.POINTER(_abi.Dims), C.c_int, C.POINTER(vp)]
    L.fastace_env_destroy.restype = C.c_int
    L.fastace_env_destroy.argtypes = [vp]
    L.fastace_env_dims.restype = C.c_int
    L.fastace_env_dims.argtypes = [vp, C.POINTER(_abi.Dims)]
    L.fastace_env_set_function_kinds.restype = C.c_int
    L.fastace_env_set_function_kinds.argtypes = [vp, C.c_int, C.c_int]
    L.fastace_env_time.restype = C.c_int
    L.fastace_env_time.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.fastace_env_set_state.restype = C.c_int
    L.fastace_env_set_state.argtypes = [vp, C.POINTER(_abi.State), C.c_uint32]
    L.fastace_env_get_state.restype = C.c_int
    L.fastace_env_get_state.argtypes = [vp, C.POINTER(_abi.State)]
    L.fastace_env_device_state.restype = C.c_int
    L.fastace_env_device_state.argtypes = [vp, C.POINTER(_abi.State)]
    L.fastace_env_step_device.restype = C.c_int
    L.fastace_env_step_device.argtypes = [vp, C.POINTER(_abi.Actions), C.POINTER(_abi.StepOut), C.c_uint32, vp]
    L.fastace_env_step_host.restype = C.c_int
    L.fastace_env_step_host.argtypes = [vp, C.POINTER(_abi.Actions), C.POINTER(_abi.StepOut), C.c_uint32]
    L.fastace_env_step_device_compact.restype = C.c_int
    L.fastace_env_step_device_compact.argtypes = [vp, C.POINTER(_abi.ActionsCompact), C.POINTER(_abi.StepOut), C.c_uint32, vp]
    L.fastace_env_step_host_compact.restype = C.c_int
    L.fastace_env_step_host_compact.argtypes = [vp, C.POINTER(_abi.ActionsCompact), C.POINTER(_abi.StepOut), C.c_uint32]
    L.fastace_env_step_host_packed.restype = C.c_int
    L.fastace_env_step_host_packed.argtypes = [vp, C.POINTER(_abi.ActionsPacked), C.POINTER(_abi.StepOut), C.c_uint32]
    L.fastace_packed_layout.restype = C.c_int
    L.fastace_packed_layout.argtypes = [C.POINTER(_abi.Dims)] + [C.POINTER(C.c_int)] * 4
    L.fastace_env_sync.restype = C.c_int
    L.fastace_env_sync.argtypes = [vp]
    L.fastace_env_launch_count.restype = C.c_int
    L.fastace_env_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.fastace_mlp_stack_layout.restype = C.c_int
    L.fastace_mlp_stack_layout.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fastace_mlp_residual_tanh_stack.restype = C.c_int
    L.fastace_mlp_residual_tanh_stack.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp]
    L.fastace_mlp_forward.restype = C.c_int
    L.fastace_mlp_forward.argtypes = [vp, vp]
    L.fastace_layer_forward.restype = C.c_int
    L.fastace_layer_forward.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.c_int, vp]
    L.fastace_layer_backward.restype = C.c_int
    L.fastace_layer_backward.argtypes = [vp, vp, vp, C.c_int64, vp]
    L.fastace_policy_bernoulli.restype = C.c_int
    L.fastace_policy_bernoulli.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.fastace_policy_normal.restype = C.c_int
    L.fastace_policy_normal.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.fastace_env_large_stats.restype = C.c_int
    L.fastace_env_large_stats.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.fastace_env_kernel_times.restype = C.c_int
    L.fastace_env_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.create_scenario_params.restype = _abi.CustomScenarioParams
    L.create_scenario_params.argtypes = [C.c_uint, C.c_uint]
    L.create_training_params.restype = _abi.TrainingParams
    L.create_training_params.argtypes = []
    # the reference's own prototypes, py/main.py:96-109
    L.run.argtypes = [_abi.CustomScenarioParams, _abi.TrainingParams]
    L.run.restype = None
    L.train.argtypes = [C.POINTER(C.c_double), C.POINTER(_abi.CustomScenarioParams), C.POINTER(_abi.TrainingParams), C.c_bool, C.c_double]
    L.train.restype = None
    L.fastace_scenario_custom_init.restype = C.c_int
    L.fastace_scenario_custom_init.argtypes = [C.POINTER(_abi.Dims), C.POINTER(_abi.CustomScenarioParams), C.c_uint32,
                                               C.POINTER(_abi.State), C.POINTER(C.c_double)]
    L.fastace_shuffle_orders.restype = C.c_int
    L.fastace_shuffle_orders.argtypes = [C.POINTER(_abi.Dims), C.c_uint32, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
    L.fastace_env_market_stats.restype = C.c_int
    L.fastace_env_market_stats.argtypes = [vp, C.POINTER(_abi.MarketStats), vp]
    L.fastace_env_shuffle_orders.restype = C.c_int
    L.fastace_env_shuffle_orders.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    if L.fastace_abi_version() != _abi.ABI_VERSION:
        raise FastaceError("libfastace_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load().fastace_last_error()
        raise FastaceError(f"fastace status {rc}: {msg.decode() if msg else ''}")
