"""Debug: per-role wait / busy cycles of the pipelined patch-embedding kernel (library built with -DB200_PP_TIMING as
tools/build/pp_timing.so).  usage (GPU box): B200FBANK_LIB=$PWD/tools/build/pp_timing.so python tools/pp_timing.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import _capi as K
x = torch.randn((1024, 1, 128, 512), device="cuda") * 0.5
conv = torch.nn.Conv2d(1, 768, 16, stride=10).cuda()
buf = (ctypes.c_ulonglong * 16)()
for it in range(3):
    b2.patch_embed(x, conv.weight.detach(), conv.bias.detach(), 10, torch.float16)
    torch.cuda.synchronize()
    K.lib.b200fbank_debug_pp_timing(buf)
v = list(buf)
t = v[8] or 1          # tiles (counted once per CTA-tile)
print("per CTA-tile (cycles; cutter / epilogue figures are per warp = sum / 4):")
print("  producer wait s_empty %.0f" % (v[0] / t))
print("  cutter   wait s_full %.0f  wait a_empty %.0f  busy %.0f (strip reads + convert %.0f, tcgen05.st + arrive %.0f)" % (v[1] / t / 4, v[2] / t / 4, v[3] / t / 4, v[9] / t / 4, v[10] / t / 4))
print("  mma      wait a_full %.0f  wait acc_empty %.0f" % (v[4] / t, v[5] / t))
print("  epilogue wait acc_full %.0f  busy %.0f (tcgen05.ld + wait %.0f, staging reads + global stores %.0f)   [%d CTA-tiles]" % (v[6] / t / 4, v[7] / t / 4, v[11] / t / 4, v[12] / t / 4, t))
