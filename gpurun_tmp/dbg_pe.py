import sys, os
sys.path.insert(0, os.getcwd())
import torch
import dl_sound_classification_b200 as b2
torch.manual_seed(0)
for (B, T, D) in ((2, 128, 192), (3, 512, 768), (1, 64, 384)):
    x = torch.randn(B, 1, 128, T, device="cuda") * 0.5
    conv = torch.nn.Conv2d(1, D, 16, stride=10).cuda()
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        ref = conv(x).flatten(2).transpose(1, 2)
        xr = x.half().float(); wr = conv.weight.half().float()
        ref16 = torch.nn.functional.conv2d(xr, wr, conv.bias, stride=10).flatten(2).transpose(1, 2)
        got = b2.patch_embed(x, conv.weight, conv.bias, 10, torch.float32)
        got16 = b2.patch_embed(x, conv.weight, conv.bias, 10, torch.float16)
    torch.cuda.synchronize()
    print(B, T, D, tuple(got.shape), "max|got-ref16|", float((got - ref16).abs().max()), "max|got-ref|", float((got - ref).abs().max()),
          "ref scale", float(ref.abs().mean()), "f16 out err", float((got16.float() - got).abs().max()))
