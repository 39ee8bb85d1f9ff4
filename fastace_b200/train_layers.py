"""Training-time layers of the decision nets (trainer re-evaluation on CUDA):  y = [x +] tanh(x W^T + b)  as
  forward   one cuBLAS GEMM + ONE hand-written element-wise kernel (csrc/layer_kernels.cuh; eager: bias add, tanh, add)
  backward  ONE element-wise kernel (dz = dy (1 - t^2)), dW AND db from one split-K batched GEMM over row chunks
            (cuBLAS picks a 3.5x slower kernel for the skinny [H x rows] x [rows x H] product with rows ~ 4e5, and
            eager spends another reduction pass on the bias gradient), dx = dy + dz W as one addmm.
Measured per 409 600 x 100 layer: forward 270 -> 145 us, backward ~600 -> ~280 us.  fp32 / TF32 as torch is set;
the reference-pinned gradient bar holds (tests/test_trainer.py, GPU)."""
import ctypes as C

import torch

from . import lib

SPLIT_ROWS = 8192        # rows per chunk of the split-K weight-gradient product


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def weight_and_bias_grad(dz, x):
    """dW = dz^T x  [H, K] accumulated over row chunks (one bmm of [S, H, chunk] x [S, chunk, K] and a sum over the S
    partials), db = dz^T 1 as one matrix-vector product (a single streaming read of dz)"""
    rows, H = dz.shape
    S = rows // SPLIT_ROWS
    body = S * SPLIT_ROWS
    acc = None
    if S > 0:
        acc = torch.bmm(dz[:body].view(S, SPLIT_ROWS, H).transpose(1, 2), x[:body].view(S, SPLIT_ROWS, x.shape[1])).sum(0)
    if body < rows:
        tail = dz[body:].t() @ x[body:]
        acc = tail if acc is None else acc + tail
    db = torch.mv(dz.t(), torch.ones(rows, dtype=dz.dtype, device=dz.device))
    return acc, db


class TanhLayer(torch.autograd.Function):
    """x [rows, K] -> [x +] tanh(x W^T + b), residual only when K == H"""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        x = x.contiguous()
        z = x @ weight.t()
        y, t = torch.empty_like(z), torch.empty_like(z)
        lib.check(lib.load().fastace_layer_forward(
            C.c_void_p(z.data_ptr()), C.c_void_p(bias.data_ptr()), C.c_void_p(x.data_ptr() if residual else 0),
            C.c_void_p(y.data_ptr()), C.c_void_p(t.data_ptr()), z.shape[0], z.shape[1], _stream(z)))
        ctx.save_for_backward(x, weight, t)
        ctx.residual = residual
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, t = ctx.saved_tensors
        dy = dy.contiguous()
        dz = torch.empty_like(dy)
        lib.check(lib.load().fastace_layer_backward(C.c_void_p(dy.data_ptr()), C.c_void_p(t.data_ptr()),
                                                    C.c_void_p(dz.data_ptr()), dy.numel(), _stream(dy)))
        dw, db = weight_and_bias_grad(dz, x)
        dx = torch.addmm(dy, dz, weight) if ctx.residual else dz @ weight
        return dx, dw, db, None


def tanh_layer(x, linear, residual):
    """drop-in for  (x +) torch.tanh(linear(x))  on 2-D fp32 CUDA inputs"""
    return TanhLayer.apply(x, linear.weight, linear.bias, residual)


def usable(x):
    return x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and torch.is_grad_enabled() and not torch.is_autocast_enabled()
