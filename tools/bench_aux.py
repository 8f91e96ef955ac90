"""Micro-benchmarks of the small / auxiliary kernels (CUDA events, median of N launches after warm-up), one JSON line:
  neck conv 3->64 (b32 180x320 fp16; b64 96x96 bf16), 3-channel weight gradient (b64 96x96 bf16), each with the
  tensor-core (mma.sync) kernel and the CUDA-core kernel it replaces (fsr_set_small_mma), against their HBM bytes;
  fused PSNR+SSIM (b64 96x96 and b8 720x1280); GPU crop + antialiased bicubic batch (b64, 24x24 LR / 96x96 HR) with the
  reference's CPU path (torch interpolate per sample, what dataloader.py:24-38 runs in its workers) timed beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200 import _lib as L, data, ops  # noqa: E402
from fast_srgan_b200.metrics import ValidationMetrics  # noqa: E402


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3      # us


out = {}
lib = L.load()
g = torch.Generator().manual_seed(0)
alpha = torch.tensor([0.25], device="cuda")
for tag, (N, H, W), dt in (("neck_b32_180x320_fp16", (32, 180, 320), torch.float16), ("neck_b64_96x96_bf16", (64, 96, 96), torch.bfloat16)):
    x = (torch.rand((N, 3, H, W), generator=g) * 2 - 1).cuda()
    w = (torch.randn((64, 3, 3, 3), generator=g) * 0.2).cuda()
    b = (torch.randn((64,), generator=g) * 0.1).cuda()
    bytes_ = N * H * W * (64 * 2 + 12)
    for mma in (1, 0):
        lib.fsr_set_small_mma(mma)
        us = timed(lambda: ops.neck_conv3x3(x, w, b, dt, act=L.ACT_PRELU, alpha=alpha))
        out[f"{tag}_{'mma' if mma else 'cuda_core'}"] = {"us": us, "GBps_algorithmic": bytes_ / us / 1e3}
N, H, W = 64, 96, 96
img = torch.randn((N, 3, H, W), generator=g).cuda()
act = torch.randn((N, H, W, 64), generator=g).cuda().to(torch.bfloat16)
acc = torch.zeros((64, 3, 3, 3), device="cuda")
for mma in (1, 0):
    lib.fsr_set_small_mma(mma)
    us = timed(lambda: ops.wgrad_c3(img, act, acc, flip=False, layout=2))
    out[f"wgrad_c3_b64_96x96_bf16_{'mma' if mma else 'cuda_core'}"] = {"us": us, "GBps_algorithmic": N * H * W * (128 + 12) / us / 1e3}
lib.fsr_set_small_mma(-1)

m = ValidationMetrics("cuda")
for tag, shape in (("psnr_ssim_b64_96x96", (64, 3, 96, 96)), ("psnr_ssim_b8_720x1280", (8, 3, 720, 1280))):
    a = (torch.rand(shape, generator=g) * 2 - 1).cuda()
    bb = (torch.rand(shape, generator=g) * 2 - 1).cuda()
    us = timed(lambda: (m.reset(), m.update(a, bb)), n=20)
    out[tag] = {"us": us, "GBps_algorithmic": 2 * a.numel() * 4 / us / 1e3}

rs = np.random.RandomState(0)
imgs = [rs.randint(0, 256, (3, 600 + 8 * i, 800 + 16 * i), dtype=np.uint8) for i in range(16)]
cache = data.DeviceImageCache(imgs)
loader = data.GpuCropLoader(cache, data.ShardedReplacementSampler(len(imgs), 64 * 40, 64, seed=0), 24, 4, seed=0)
samples = loader.draw(torch.arange(64) % len(imgs))
us = timed(lambda: data.crop_resize_batch(cache, samples, 24, 4, loader._taps))
out["crop_resize_b64"] = {"us": us, "samples_per_s": 64 / us * 1e6}
t0 = time.perf_counter()
nb = 0
for lr, hr in loader:          # host draw + H2D of 12 B/sample + one launch per batch
    nb += 1
torch.cuda.synchronize()
out["gpu_crop_loader_b64"] = {"batches": nb, "samples_per_s": nb * 64 / (time.perf_counter() - t0)}
# reference CPU path of one worker: crop + v2.Resize-equivalent interpolate per sample (dataloader.py:24-38)
torch.set_num_threads(1)
t0 = time.perf_counter()
for k in range(256):
    i, cy, cx = (int(v) for v in samples[k % 64])
    hr = torch.tensor(np.ascontiguousarray(imgs[i][:, cy:cy + 96, cx:cx + 96]), dtype=torch.float32)
    lr = torch.nn.functional.interpolate(hr[None], size=(24, 24), mode="bicubic", antialias=True)[0]
    hr, lr = hr / 127.5 - 1, lr / 127.5 - 1
dt_cpu = (time.perf_counter() - t0) / 256
out["cpu_dataloader_worker_reference_path"] = {"us_per_sample": dt_cpu * 1e6, "samples_per_s_per_worker": 1 / dt_cpu,
                                                "samples_per_s_16_workers": 16 / dt_cpu}
print(json.dumps(out), flush=True)
