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


def patch_embed(spec: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 10,
                out_dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """``spec`` (B, 1, F, T) or (B, F, T) float32 CUDA; ``weight`` (D, 1, 16, 16); ``bias`` (D) or None.
    Returns ``(B, Fp * Tp, D)`` in ``out_dtype`` (float16, what fp16 autocast hands to the transformer, or float32)."""
    if not spec.is_cuda:
        raise RuntimeError("patch_embed needs CUDA tensors: dl_sound_classification_b200 has no CPU fallback")
    if spec.dim() == 3:
        spec = spec.unsqueeze(1)                                   # src/models/ast.py:52-53
    if spec.dim() != 4 or spec.shape[1] != 1 or spec.dtype != torch.float32:
        raise ValueError("spec must be (B, 1, F, T) float32")
    if weight.dim() != 4 or weight.shape[1] != 1 or weight.shape[2] != weight.shape[3]:
        raise ValueError("weight must be (D, 1, P, P)")
    if out_dtype not in (torch.float16, torch.float32):
        raise TypeError("out_dtype must be float16 or float32")
    spec = spec.contiguous()
    B, _, F, T = spec.shape
    D, P = int(weight.shape[0]), int(weight.shape[2])
    w16 = weight.detach().to(device=spec.device, dtype=torch.float16).reshape(D, P * P).contiguous()
    b32 = None if bias is None else bias.detach().to(device=spec.device, dtype=torch.float32).contiguous()
    if F < P or T < P:
        raise ValueError(f"spectrogram {F} x {T} is smaller than one patch")
    Fp, Tp = (F - P) // stride + 1, (T - P) // stride + 1
    out = torch.empty((B, Fp * Tp, D), dtype=out_dtype, device=spec.device)
    with torch.cuda.device(spec.device):
        K.check(K.lib.b200fbank_patch_embed(spec.data_ptr(), B, F, T, w16.data_ptr(), None if b32 is None else b32.data_ptr(),
                                            D, P, int(stride), out.data_ptr(), int(out_dtype == torch.float16),
                                            torch.cuda.current_stream().cuda_stream))
    return out


class PatchEmbed(torch.nn.Module):
    """Mirror of the reference module (src/models/ast_mini.py:7-15): same constructor, same parameters
    (``proj.weight`` / ``proj.bias``, so a reference state_dict loads), forward on the tensor cores."""

    def __init__(self, in_chans: int = 1, emb_dim: int = 192, patch_size: int = 16, stride: int = 10):
        super().__init__()
        if in_chans != 1:
            raise NotImplementedError("AST spectrograms have one channel")
        self.proj = torch.nn.Conv2d(in_chans, emb_dim, kernel_size=patch_size, stride=stride)
        self.stride = stride

    def forward(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float16) -> torch.Tensor:
        return patch_embed(x, self.proj.weight, self.proj.bias, self.stride, out_dtype)
