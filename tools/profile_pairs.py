"""Two forwards of an n_filters = 32 generator (batch 32, 180x320) for an ncu launch list of the pixel-pair path."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_srgan_b200.model import Generator  # noqa: E402

g = Generator(types.SimpleNamespace(n_filters=32, n_layers=8), compute_dtype=torch.float16).cuda().eval()
x = (torch.rand((32, 3, 180, 320)) * 2 - 1).cuda()
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        g(x)
torch.cuda.synchronize()
