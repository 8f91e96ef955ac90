"""Minimal driver for ncu: N forward passes of the bench workload (no CPU baseline, no clocks thread)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200.model import Generator  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
h, w = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (180, 320)
dt = torch.bfloat16 if os.environ.get("FSR_DTYPE", "fp16") == "bf16" else torch.float16
g = Generator(types.SimpleNamespace(n_filters=64, n_layers=8), compute_dtype=dt)
g = g.cuda().eval()
x = (torch.rand((batch, 3, h, w)) * 2 - 1).cuda()
with torch.no_grad():
    for _ in range(iters):
        y = g(x)
torch.cuda.synchronize()
print("done", tuple(y.shape))
