"""Fused TCN engine (host side): drives the tensor-core / fused kernels of libfqss_sm100 for the
24 ConvBlocks of the separator.  Plumbing only -- buffers, pointers, launch order."""
import torch

from . import _native as N
from ._native import check, lib, ptr, stream_ptr


def pw_gemm(act_bf16, w_bf16, s1, s0, M, addend=None, out_dtype=torch.float32):
    """out[b,o,m] = s1[o] * sum_k act[b,k,m] w[o,k] + s0[o] (+ addend) on tcgen05; act [B,K,ld], w [N,K] bf16."""
    N.require_cuda(act_bf16, w_bf16, s1, s0, addend)
    B, K, ld = act_bf16.shape
    Nn = w_bf16.shape[0]
    assert act_bf16.is_contiguous() and w_bf16.is_contiguous() and w_bf16.shape[1] == K
    out = torch.empty((B, Nn, ld), device=act_bf16.device, dtype=out_dtype)
    f32 = out if out_dtype == torch.float32 else None
    b16 = out if out_dtype == torch.bfloat16 else None
    check(lib().fqss_pw_gemm(ptr(act_bf16), ptr(w_bf16), ptr(s1), ptr(s0), ptr(f32) or None, ptr(b16) or None,
                             ptr(addend) or None, B, K, Nn, M, ld, stream_ptr()))
    return out
