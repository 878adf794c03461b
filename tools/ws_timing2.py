"""Pipeline balance of the warp-specialised kernel: run with B200FBANK_LIB pointing at a -DB200_WS_TIMING build.
usage (GPU box): B200FBANK_LIB=$PWD/tools/build/timing.so python tools/ws_timing2.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import _capi as K
fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
wav = torch.rand((1024, 220500), device="cuda") * 2 - 1
out = torch.empty((1024, 512, 128), device="cuda")
buf = (ctypes.c_ulonglong * 8)()
for it in range(3):
    fe(wav, out_frames=512, mean=-6.6, std=5.0, out=out, return_n_frames=False)
    torch.cuda.synchronize()
    K.lib.b200fbank_debug_ws_timing(buf)
v = list(buf)
print("WSTIMING R per half-chunk (cycles, per R warp): load-wait %.0f  empty-wait %.0f  compute %.0f  end-barrier %.0f  [%d warp-halves]" % (v[0] / v[3], v[1] / v[3], v[2] / v[3], v[7] / v[3], v[3]))
print("WSTIMING F per pass  (cycles, per F warp): full-wait %.0f  pass %.0f   [%d passes]" % (v[4] / v[6], v[5] / v[6], v[6]))
