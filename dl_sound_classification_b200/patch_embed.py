"""AST patch embedding fed from the frontend's features.

``PatchEmbed.forward`` (src/models/ast_mini.py:7-15, ast_small.py:7-15) and ``ASTModel.forward``
(src/models/ast.py:30,50-56) run ``Conv2d(1, D, 16, stride=10)`` on the ``(B, 1, 128, T)`` spectrogram
and flatten it to ``(B, 12 * Tp, D)``.  ``patch_embed`` computes exactly that as one im2col GEMM on the
tensor cores (``b200fbank_patch_embed``: tcgen05.mma, fp16 operands, fp32 accumulation), reading the
features the fused frontend just wrote while they are still in L2.  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi as K


def _forward(spec: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int,
             out_dtype: torch.dtype) -> torch.Tensor:
    B, _, F, T = spec.shape
    D, P = int(weight.shape[0]), int(weight.shape[2])
    w16 = weight.detach().to(device=spec.device, dtype=torch.float16).reshape(D, P * P).contiguous()
    b32 = None if bias is None else bias.detach().to(device=spec.device, dtype=torch.float32).contiguous()
    Fp, Tp = (F - P) // stride + 1, (T - P) // stride + 1
    out = torch.empty((B, Fp * Tp, D), dtype=out_dtype, device=spec.device)
    with torch.cuda.device(spec.device):
        K.check(K.lib.b200fbank_patch_embed(spec.data_ptr(), B, F, T, w16.data_ptr(), None if b32 is None else b32.data_ptr(),
                                            D, P, int(stride), out.data_ptr(), int(out_dtype == torch.float16),
                                            torch.cuda.current_stream().cuda_stream))
    return out


class _PatchEmbedFn(torch.autograd.Function):
    """Forward on the tcgen05 kernel; backward = the convolution's own gradients from torch (cuDNN, on the GPU), so that
    swapping ``PatchEmbed`` into a model that trains it does not silently freeze ``proj.weight`` / ``proj.bias``."""

    @staticmethod
    def forward(ctx, spec, weight, bias, stride, out_dtype):
        ctx.save_for_backward(spec, weight)
        ctx.stride, ctx.has_bias = int(stride), bias is not None
        return _forward(spec, weight, bias, int(stride), out_dtype)

    @staticmethod
    def backward(ctx, g):
        spec, weight = ctx.saved_tensors
        B, _, F, T = spec.shape
        D, P = int(weight.shape[0]), int(weight.shape[2])
        Fp, Tp = (F - P) // ctx.stride + 1, (T - P) // ctx.stride + 1
        g4 = g.to(torch.float32).transpose(1, 2).reshape(B, D, Fp, Tp)
        w32 = weight.detach().to(device=spec.device, dtype=torch.float32)
        gs = torch.nn.grad.conv2d_input(spec.shape, w32, g4, stride=ctx.stride) if ctx.needs_input_grad[0] else None
        gw = torch.nn.grad.conv2d_weight(spec, weight.shape, g4, stride=ctx.stride).to(weight.dtype) if ctx.needs_input_grad[1] else None
        gb = g4.sum((0, 2, 3)) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gs, gw, gb, None, None


def patch_embed(spec: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 10,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``spec`` (B, 1, F, T) or (B, F, T) float32 CUDA; ``weight`` (D, 1, 16, 16); ``bias`` (D) or None.
    Returns ``(B, Fp * Tp, D)`` in ``out_dtype``: None = float16 under CUDA autocast (what the reference's "16-mixed"
    trainer hands to the transformer), float32 otherwise (what the reference module returns without autocast)."""
    if not spec.is_cuda:
        raise RuntimeError("patch_embed needs CUDA tensors: dl_sound_classification_b200 has no CPU fallback")
    if spec.dim() == 3:
        spec = spec.unsqueeze(1)                                   # src/models/ast.py:52-53
    if spec.dim() != 4 or spec.shape[1] != 1 or spec.dtype != torch.float32:
        raise ValueError("spec must be (B, 1, F, T) float32")
    if weight.dim() != 4 or weight.shape[1] != 1 or weight.shape[2] != weight.shape[3]:
        raise ValueError("weight must be (D, 1, P, P)")
    if out_dtype is None:
        out_dtype = torch.float16 if torch.is_autocast_enabled() else torch.float32
    if out_dtype not in (torch.float16, torch.float32):
        raise TypeError("out_dtype must be float16 or float32")
    spec = spec.contiguous()
    P = int(weight.shape[2])
    if spec.shape[2] < P or spec.shape[3] < P:
        raise ValueError(f"spectrogram {spec.shape[2]} x {spec.shape[3]} is smaller than one patch")
    needs_grad = torch.is_grad_enabled() and (spec.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad))
    if needs_grad:
        return _PatchEmbedFn.apply(spec, weight, bias, int(stride), out_dtype)
    return _forward(spec, weight, bias, int(stride), out_dtype)


class PatchEmbed(torch.nn.Module):
    """Mirror of the reference module (src/models/ast_mini.py:7-15): same constructor, same parameters
    (``proj.weight`` / ``proj.bias``, so a reference state_dict loads), forward on the tensor cores."""

    def __init__(self, in_chans: int = 1, emb_dim: int = 192, patch_size: int = 16, stride: int = 10):
        super().__init__()
        if in_chans != 1:
            raise NotImplementedError("AST spectrograms have one channel")
        self.proj = torch.nn.Conv2d(in_chans, emb_dim, kernel_size=patch_size, stride=stride)
        self.stride = stride

    def forward(self, x: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        return patch_embed(x, self.proj.weight, self.proj.bias, self.stride, out_dtype)
