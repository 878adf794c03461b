"""Generate tests/golden/patch_embed.npz with the REFERENCE's own PatchEmbed module.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_patch_embed.py

``PatchEmbed`` is imported from /root/reference/src/models/ast_mini.py (pure torch); float32 on the CPU.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src/models")
from ast_mini import PatchEmbed  # noqa: E402

torch.manual_seed(4242)
torch.set_num_threads(1)
pe = PatchEmbed(in_chans=1, emb_dim=192, patch_size=16, stride=10)
g = torch.Generator().manual_seed(99)
x = torch.randn((2, 1, 128, 66), generator=g) * 0.5          # AST-normalised log-mel scale; 66 frames -> 6 patch columns
with torch.no_grad():
    y = pe(x)
np.savez_compressed(os.path.join(HERE, "patch_embed.npz"), weight=pe.proj.weight.detach().numpy(), bias=pe.proj.bias.detach().numpy(),
                    out=y.numpy(), meta_torch=np.array(torch.__version__))
print("out", tuple(y.shape), float(y.abs().mean()))
