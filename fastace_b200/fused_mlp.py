"""Host side of the fused residual tanh stack (csrc/mlp_stack.cuh) — rollout-only fast path of the decision nets.
Library boundary: fastace_mlp_residual_tanh_stack in include/fastace_b200.h; torch is only the tensor container."""
import ctypes as C

import torch

from . import lib

_cache = {}


def layout(hidden):
    """(padded_out, padded_in) of the packed weights for this hidden size"""
    a, b = C.c_int(0), C.c_int(0)
    lib.check(lib.load().fastace_mlp_stack_layout(int(hidden), C.byref(a), C.byref(b)))
    return a.value, b.value


def pack_layers(layers):
    """[L] nn.Linear(H, H) -> (bf16 [L][NP][KP] zero-padded weights, fp32 [L][NP] zero-padded biases); cached until a
    parameter is modified in place (tensor version counters)."""
    key = tuple(id(l) for l in layers)
    ver = tuple((l.weight._version, l.bias._version, l.weight.data_ptr()) for l in layers)
    hit = _cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    H = layers[0].in_features
    assert all(l.in_features == H and l.out_features == H for l in layers), "residual stack needs square layers"
    np_, kp = layout(H)
    dev = layers[0].weight.device
    with torch.no_grad():
        w = torch.zeros(len(layers), np_, kp, dtype=torch.bfloat16, device=dev)
        w[:, :H, :H] = torch.stack([l.weight for l in layers]).to(torch.bfloat16)
        b = torch.zeros(len(layers), np_, dtype=torch.float32, device=dev)
        b[:, :H] = torch.stack([l.bias for l in layers]).float()
    _cache[key] = (ver, w, b)
    return w, b


def residual_tanh_stack(x, layers):
    """x [..., H] (CUDA) -> x after  x <- x + tanh(layer(x))  for every layer, one kernel."""
    if not x.is_cuda:
        raise RuntimeError("the fused stack is a CUDA kernel; there is no CPU path")
    H = x.shape[-1]
    w, b = pack_layers(layers)
    xs = x.contiguous().float()
    y = torch.empty_like(xs)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    lib.check(lib.load().fastace_mlp_residual_tanh_stack(
        C.c_void_p(xs.data_ptr()), C.c_void_p(y.data_ptr()), xs.numel() // H, H, len(layers),
        C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(stream)))
    return y
