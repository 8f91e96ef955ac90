"""Minimal driver for ncu: two launches each of the small / auxiliary kernels at their bench shapes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200 import _lib as L, data, ops  # noqa: E402
from fast_srgan_b200.metrics import ValidationMetrics  # noqa: E402

g = torch.Generator().manual_seed(0)
alpha = torch.tensor([0.25], device="cuda")
w = (torch.randn((64, 3, 3, 3), generator=g) * 0.2).cuda()
b = (torch.randn((64,), generator=g) * 0.1).cuda()
x32 = (torch.rand((32, 3, 180, 320), generator=g) * 2 - 1).cuda()
x64 = (torch.rand((64, 3, 96, 96), generator=g) * 2 - 1).cuda()
act = torch.randn((64, 96, 96, 64), generator=g).cuda().to(torch.bfloat16)
acc = torch.zeros((64, 3, 3, 3), device="cuda")
m = ValidationMetrics("cuda")
rs = np.random.RandomState(0)
cache = data.DeviceImageCache([rs.randint(0, 256, (3, 600, 800), dtype=np.uint8) for _ in range(4)])
samples = torch.tensor([(i % 4, 7 * i, 11 * i) for i in range(64)], dtype=torch.int32)
for _ in range(2):
    ops.neck_conv3x3(x32, w, b, torch.float16, act=L.ACT_PRELU, alpha=alpha)
    ops.neck_conv3x3(x64, w, b, torch.bfloat16, act=L.ACT_LRELU, slope=0.2)
    ops.wgrad_c3(x64, act, acc, flip=False, layout=2)
    m.update(x64, x64 * 0.9)
    data.crop_resize_batch(cache, samples, 24, 4)
torch.cuda.synchronize()
print("done")
