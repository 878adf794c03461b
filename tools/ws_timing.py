"""Debug: pipeline balance of the warp-specialised kernel (library built with -DB200_WS_TIMING).
usage (GPU box): python tools/ws_timing.py"""
import ctypes, os, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# run as: B200FBANK_LIB=$PWD/tools/build/ws_timing.so python tools/ws_timing.py   (nvcc ... -DB200_WS_TIMING -o tools/build/ws_timing.so)
assert "B200FBANK_LIB" in os.environ, "point B200FBANK_LIB at a library built with -DB200_WS_TIMING"
import torch
import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import _capi as K
fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
wav = torch.rand((1024, 220500), device="cuda") * 2 - 1
out = torch.empty((1024, 512, 128), device="cuda")
buf = (ctypes.c_ulonglong * 8)()
for it in range(3):
    fe(wav, out_frames=512, mean=-6.6, std=5.0, out=out, return_n_frames=False)
    K.lib.b200fbank_debug_ws_timing(buf)
v = list(buf)
print("R per chunk (cycles, per R warp): load-wait %.0f  empty-wait %.0f  compute %.0f   [%d warp-chunks]" % (v[0] / v[3], v[1] / v[3], v[2] / v[3], v[3]))
print("F per pass  (cycles, per F warp): full-wait %.0f  pass %.0f   [%d passes]" % (v[4] / v[6], v[5] / v[6], v[6]))
