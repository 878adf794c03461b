"""Debug: phase timing of the patch-embedding kernel (library built with -DB200_PE_TIMING).
usage (GPU box): python tools/pe_timing.py"""
import ctypes, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shutil.copy(os.path.join(ROOT, "tools", "libb200fbank_pe_timing.so"), os.path.join(ROOT, "dl_sound_classification_b200", "lib", "libb200fbank.so"))
import torch
import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import _capi as K
x = torch.randn((1024, 1, 128, 512), device="cuda") * 0.5
conv = torch.nn.Conv2d(1, 768, 16, stride=10).cuda()
buf = (ctypes.c_ulonglong * 4)()
for it in range(3):
    b2.patch_embed(x, conv.weight.detach(), conv.bias.detach(), 10, torch.float16)
    torch.cuda.synchronize()
    K.lib.b200fbank_debug_pe_timing(buf)
v = list(buf)
print("per tile (cycles): build A %.0f  mma+wait %.0f  epilogue %.0f  [%d tiles]" % (v[0] / v[3], v[1] / v[3], v[2] / v[3], v[3]))
