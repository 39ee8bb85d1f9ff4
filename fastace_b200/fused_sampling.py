"""Host side of the sampling kernels (csrc/policy_kernels.cuh) — rollout-only fast path of policy.evaluate: the
Bernoulli takes and the normal-family samples of all agents, their log-probabilities, and the write of the decisions
into the env's action layout, one launch per head.  Library boundary: fastace_policy_bernoulli / fastace_policy_normal
in include/fastace_b200.h; torch is only the tensor container.  The differentiable path (policy.sample_*) is the
checker of these kernels in tests/test_policy.py."""
import torch

from . import lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def bernoulli(probas, uniforms, idx, valid, invalid_nan, out_idx, out_take):
    """probas / uniforms [E,A,S] fp32, idx [E,A,S] int64, valid [E] -> logp [E,A]; out_idx int32 / out_take uint8 [E][S][A]"""
    E, A, S = probas.shape
    probas, uniforms = _f32(probas), _f32(uniforms)
    idx = idx.contiguous()
    assert idx.dtype == torch.int64 and out_idx.dtype == torch.int32 and out_take.dtype == torch.uint8
    assert tuple(out_idx.shape) == (E, S, A) and tuple(out_take.shape) == (E, S, A) and out_idx.is_contiguous() and out_take.is_contiguous()
    valid8 = valid.reshape(E).to(torch.uint8)
    logp = torch.empty(E, A, dtype=torch.float32, device=probas.device)
    lib.check(lib.load().fastace_policy_bernoulli(
        probas.data_ptr(), uniforms.data_ptr(), idx.data_ptr(), valid8.data_ptr(), E, A, S, 1 if invalid_nan else 0,
        out_idx.data_ptr(), out_take.data_ptr(), logp.data_ptr(), _stream()))
    return logp


def normal(params, offset, noise, kind, out_x, logp=None):
    """params [E,A,C,K] (or [E,A,K] with one component) fp32 with (mu, log sigma) at [..., offset:offset+2];
    noise [E,A,C] (or [E,A]); kind "logit" / "log"; out_x fp32 [E][C][A] (or [E][A]).  Returns logp [E,A] (summed over
    the components; added to `logp` when given)."""
    params, noise = _f32(params), _f32(noise)
    E, A = params.shape[0], params.shape[1]
    C = params.shape[2] if params.dim() == 4 else 1
    K = params.shape[-1]
    assert noise.numel() == E * A * C and out_x.numel() == E * A * C and out_x.dtype == torch.float32 and out_x.is_contiguous()
    acc = logp is not None
    if logp is None:
        logp = torch.empty(E, A, dtype=torch.float32, device=params.device)
    lib.check(lib.load().fastace_policy_normal(
        params.data_ptr(), K, offset, noise.data_ptr(), E, A, C, 0 if kind == "logit" else 1, 1 if acc else 0,
        out_x.data_ptr(), logp.data_ptr(), _stream()))
    return logp
