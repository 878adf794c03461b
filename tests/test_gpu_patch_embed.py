"""GPU tests of the patch-embedding row: the tcgen05 im2col GEMM behind
b200fbank_patch_embed against the reference's own PatchEmbed output (golden) and against torch's convolution."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
# fp16 operands (the reference's "16-mixed" AST setting), fp32 accumulation: the float32 convolution of the reference is
# the yardstick (operand rounding 2^-11 relative), the convolution of the fp16-rounded operands the exactness check
ATOL_VS_FP32, ATOL_VS_FP16_OPERANDS = 2e-3, 2e-5


@pytest.fixture(scope="module")
def b2():
    import dl_sound_classification_b200 as m
    assert torch.cuda.is_available()
    return m


def test_matches_the_reference_patch_embed_golden(b2):
    g = np.load(os.path.join(HERE, "golden", "patch_embed.npz"))
    gen = torch.Generator().manual_seed(99)
    x = torch.randn((2, 1, 128, 66), generator=gen) * 0.5
    w, b = torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"])
    got = b2.patch_embed(x.cuda(), w, b, 10, torch.float32).cpu()
    assert tuple(got.shape) == g["out"].shape == (2, 12 * 6, 192)
    assert float((got - torch.from_numpy(g["out"])).abs().max()) < ATOL_VS_FP32
    half = b2.patch_embed(x.cuda(), w, b, 10, torch.float16)
    assert half.dtype == torch.float16 and float((half.float().cpu() - got).abs().max()) < 2e-3
    # module mirror: a reference state_dict loads and the (B, F, T) form of ASTModel.forward is accepted
    mod = b2.PatchEmbed(1, 192, 16, 10)
    mod.load_state_dict({"proj.weight": w, "proj.bias": b})
    assert torch.equal(mod.cuda()(x[:, 0].cuda(), torch.float32).cpu(), got)


@pytest.mark.parametrize("B,T,D", [(3, 512, 768), (5, 200, 384), (1, 16, 192), (130, 40, 192), (2, 101, 192)])
def test_against_torch_convolution(b2, B, T, D):
    torch.manual_seed(B * 1000 + T + D)
    x = (torch.randn(B, 1, 128, T) * 0.5).cuda()
    conv = torch.nn.Conv2d(1, D, 16, stride=10).cuda()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = conv(x).flatten(2).transpose(1, 2)
            ref16 = torch.nn.functional.conv2d(x.half().float(), conv.weight.half().float(), conv.bias,
                                               stride=10).flatten(2).transpose(1, 2)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    got = b2.patch_embed(x, conv.weight, conv.bias, 10, torch.float32)
    assert tuple(got.shape) == tuple(ref.shape)
    assert float((got - ref16).abs().max()) < ATOL_VS_FP16_OPERANDS        # same operands, fp32 accumulation: summation order only
    assert float((got - ref).abs().max()) < ATOL_VS_FP32
    nob = b2.patch_embed(x, conv.weight, None, 10, torch.float32)
    assert float((nob + conv.bias - got).abs().max()) < 1e-6


def test_full_size_feeds_from_the_frontend(b2):
    """BASELINE.json configs[4] tail: 1024 clips -> fused frontend (AST layout) -> patch embedding."""
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    gen = torch.Generator(device="cuda").manual_seed(3)
    wav = torch.rand((1024, 220500), generator=gen, device="cuda") * 2 - 1
    spec, _ = fe(wav, out_frames=512, mean=-6.6268, std=5.0613, layout="bft")
    assert tuple(spec.shape) == (1024, 1, 128, 512)
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(1, 768, 16, stride=10).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        y = b2.patch_embed(spec, conv.weight, conv.bias)              # autocast: fp16 out, as the reference's 16-mixed trainer
    assert tuple(y.shape) == (1024, 600, 768) and y.dtype == torch.float16 and bool(torch.isfinite(y).all())
    idx = [0, 511, 1023]
    with torch.no_grad():
        ref = conv(spec[idx]).flatten(2).transpose(1, 2)
    assert float((y[idx].float() - ref).abs().max()) < 4e-3
    assert torch.equal(b2.patch_embed(spec[idx], conv.weight.detach(), conv.bias.detach(), 10, torch.float16), y[idx])   # batch invariance
    with pytest.raises(NotImplementedError):
        b2.patch_embed(spec[:1], torch.zeros(100, 1, 16, 16), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.patch_embed(spec[:1].cpu(), conv.weight, conv.bias)


@pytest.mark.parametrize("B,F,T,D,stride", [(2, 128, 1024, 192, 10),      # three segments per patch row
                                            (1, 36, 1380, 384, 10),        # odd number of segments (tail tile half empty), mel-dB length
                                            (3, 128, 512, 192, 16),        # non-overlapping patches
                                            (2, 64, 300, 192, 7),          # odd stride: 32-bit strip reads
                                            (1, 128, 2048, 192, 10)])
def test_tma_pipeline_shapes_against_torch_convolution(b2, B, F, T, D, stride):
    """The TMA + tcgen05 pipeline (patch_embed_pipe.cuh; taken whenever T % 4 == 0) on segmentations the AST shape does
    not reach: several segments per patch row, a half-empty tail tile, other strides and heights, both output types."""
    torch.manual_seed(B * 7 + T + D + stride)
    x = (torch.randn(B, 1, F, T) * 0.5).cuda()
    conv = torch.nn.Conv2d(1, D, 16, stride=stride).cuda()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref16 = torch.nn.functional.conv2d(x.half().float(), conv.weight.half().float(), conv.bias,
                                               stride=stride).flatten(2).transpose(1, 2)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    got = b2.patch_embed(x, conv.weight, conv.bias, stride, torch.float32)
    assert tuple(got.shape) == tuple(ref16.shape)
    assert float((got - ref16).abs().max()) < ATOL_VS_FP16_OPERANDS
    half = b2.patch_embed(x, conv.weight, conv.bias, stride, torch.float16)
    assert float((half.float() - got).abs().max()) < 4e-3
    # repeated launches on one stream (barrier phases start from scratch every launch)
    assert torch.equal(b2.patch_embed(x, conv.weight, conv.bias, stride, torch.float32), got)


def test_patch_embed_trains_through_the_kernel(b2):
    """Swapping the module into a model that trains it must not freeze its parameters: the gradients of the kernel's
    forward are the convolution's own (torch, on the GPU), and the default output type follows autocast."""
    torch.manual_seed(5)
    x = (torch.randn(2, 1, 128, 96) * 0.5).cuda().requires_grad_(True)
    mod = b2.PatchEmbed(1, 192, 16, 10).cuda()
    ref = torch.nn.Conv2d(1, 192, 16, stride=10).cuda()
    ref.load_state_dict({"weight": mod.proj.weight.detach().clone(), "bias": mod.proj.bias.detach().clone()})
    y = mod(x)
    assert y.dtype == torch.float32 and y.requires_grad                   # no autocast: float32 like the reference module
    g = torch.randn_like(y)
    y.backward(g)
    x2 = x.detach().clone().requires_grad_(True)
    ref(x2).flatten(2).transpose(1, 2).backward(g)
    assert float((mod.proj.weight.grad - ref.weight.grad).abs().max()) < 2e-3 * float(ref.weight.grad.abs().max())
    assert float((mod.proj.bias.grad - ref.bias.grad).abs().max()) < 1e-4 * float(ref.bias.grad.abs().max())
    assert float((x.grad - x2.grad).abs().max()) < 1e-5 + 1e-4 * float(x2.grad.abs().max())
    with torch.autocast("cuda", dtype=torch.float16), torch.no_grad():
        assert mod(x).dtype == torch.float16


def test_tma_pipeline_random_shapes(b2):
    """Fuzz of the segmentation logic (segments per patch row, patches per segment, tail tiles, strip alignment) of the
    pipelined kernel: random feature-map shapes and strides against torch's convolution of the fp16-rounded operands."""
    g = torch.Generator().manual_seed(2025)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for _ in range(12):
            B = int(torch.randint(1, 6, (1,), generator=g))
            F = int(torch.randint(16, 140, (1,), generator=g))
            T = 4 * int(torch.randint(4, 420, (1,), generator=g))
            stride = int(torch.randint(4, 17, (1,), generator=g))
            x = (torch.randn(B, 1, F, T, generator=g) * 0.5).cuda()
            w = (torch.randn(192, 1, 16, 16, generator=g) * 0.05).cuda()
            bias = torch.randn(192, generator=g).cuda()
            ref = torch.nn.functional.conv2d(x.half().float(), w.half().float(), bias, stride=stride).flatten(2).transpose(1, 2)
            got = b2.patch_embed(x, w, bias, stride, torch.float32)
            assert tuple(got.shape) == tuple(ref.shape), (B, F, T, stride)
            assert float((got - ref).abs().max()) < ATOL_VS_FP16_OPERANDS, (B, F, T, stride)
    finally:
        torch.backends.cudnn.allow_tf32 = old
