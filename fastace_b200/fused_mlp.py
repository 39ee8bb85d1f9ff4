"""Host side of the fused residual tanh stack (csrc/mlp_stack.cuh) — rollout-only fast path of the decision nets.
Library boundary: fastace_mlp_residual_tanh_stack in include/fastace_b200.h; torch is only the tensor container."""
import ctypes as C
import weakref

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
    # ids (and device addresses) are reused once a module is freed: an entry counts only for the very same live objects
    if hit is not None and hit[0] == ver and all(r() is l for r, l in zip(hit[3], layers)):
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
    _cache[key] = (ver, w, b, [weakref.ref(l) for l in layers])
    return w, b


class MlpDesc(C.Structure):       # fastace_mlp_desc_t
    _fields_ = [("rows", C.c_int64), ("hidden", C.c_int32), ("layers", C.c_int32),
                ("x", C.c_void_p), ("y", C.c_void_p), ("w_bf16", C.c_void_p), ("bias", C.c_void_p),
                ("x0", C.c_void_p), ("in_features", C.c_int32), ("w0_bf16", C.c_void_p), ("b0", C.c_void_p),
                ("out", C.c_void_p), ("out_features", C.c_int32), ("activation", C.c_int32),
                ("wl_bf16", C.c_void_p), ("bl", C.c_void_p)]


ACT = {None: 0, "sigmoid": 1, "tanh": 2}


def _pack_edge(layer, rows_pad, cols_pad, key):
    """one Linear as zero-padded bf16 [rows_pad][cols_pad] + fp32 bias [rows_pad]; cached like pack_layers"""
    ver = (layer.weight._version, layer.bias._version, layer.weight.data_ptr())
    hit = _cache.get((key, id(layer)))
    if hit is not None and hit[0] == ver and hit[3]() is layer:
        return hit[1], hit[2]
    o, i = layer.weight.shape
    with torch.no_grad():
        w = torch.zeros(rows_pad, cols_pad, dtype=torch.bfloat16, device=layer.weight.device)
        w[:o, :i] = layer.weight.to(torch.bfloat16)
        b = torch.zeros(rows_pad, dtype=torch.float32, device=layer.weight.device)
        b[:o] = layer.bias.float()
    _cache[(key, id(layer))] = (ver, w, b, weakref.ref(layer))
    return w, b


def supports(first, layers, last, x):
    """can this net body run through fastace_mlp_forward?"""
    H = layers[0].in_features if len(layers) else (first.out_features if first is not None else x.shape[-1])
    if H % 2 or H > 128 or H < 2:
        return False
    if any(l.in_features != H or l.out_features != H for l in layers):
        return False
    if first is not None and (first.out_features != H or first.in_features > 128):
        return False
    if last is not None and (last.in_features != H or last.out_features > 16):
        return False
    return True


def net_forward(x, first, layers, last, activation=None, want_stream=False):
    """act(last(stack(tanh(first(x)))))  in ONE kernel.  first / last may be None (x is then the [.., H] stack input /
    the [.., H] stream is returned).  want_stream=True also returns the stream next to the head output."""
    if not x.is_cuda:
        raise RuntimeError("the fused stack is a CUDA kernel; there is no CPU path")
    H = layers[0].in_features if len(layers) else (first.out_features if first is not None else x.shape[-1])
    np_, kp = layout(H)
    xs = x.contiguous().float()
    rows = xs.numel() // xs.shape[-1]
    d = MlpDesc()
    d.rows, d.hidden, d.layers = rows, H, len(layers)
    keep = [xs]
    if len(layers):
        w, b = pack_layers(layers)
        d.w_bf16, d.bias = w.data_ptr(), b.data_ptr()
        keep += [w, b]
    if first is not None:
        k0p = (first.in_features + 15) // 16 * 16 + 8
        w0, b0 = _pack_edge(first, np_, k0p, "first")
        d.x0, d.in_features, d.w0_bf16, d.b0 = xs.data_ptr(), first.in_features, w0.data_ptr(), b0.data_ptr()
        keep += [w0, b0]
    else:
        d.x = xs.data_ptr()
    y = out = None
    if last is None or want_stream:
        y = torch.empty(*xs.shape[:-1], H, dtype=torch.float32, device=xs.device)
        d.y = y.data_ptr()
    if last is not None:
        wl, bl = _pack_edge(last, 16, kp, "last")
        out = torch.empty(*xs.shape[:-1], last.out_features, dtype=torch.float32, device=xs.device)
        d.out, d.out_features, d.activation, d.wl_bf16, d.bl = out.data_ptr(), last.out_features, ACT[activation], wl.data_ptr(), bl.data_ptr()
        keep += [wl, bl]
    stream = torch.cuda.current_stream(xs.device).cuda_stream
    lib.check(lib.load().fastace_mlp_forward(C.byref(d), C.c_void_p(stream)))
    if last is None:
        return y
    return (out, y) if want_stream else out


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
