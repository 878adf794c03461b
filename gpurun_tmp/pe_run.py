import sys, os
sys.path.insert(0, os.getcwd())
import torch
import dl_sound_classification_b200 as b2
x = torch.randn((1024, 1, 128, 512), device="cuda") * 0.5
conv = torch.nn.Conv2d(1, 768, 16, stride=10).cuda()
w, b = conv.weight.detach(), conv.bias.detach()
for _ in range(3):
    y = b2.patch_embed(x, w, b)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
