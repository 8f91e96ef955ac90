"""BASELINE configs[4]: generator sweep n_filters x n_layers x input size, batch 32, FPS + fraction of the conv-FLOP
roofline (whole model, ALGORITHMIC FLOPs of the unpadded network over the sustained bf16 GEMM peak of
MEASURED_PEAKS.json).  n_filters = 32 runs zero-padded to 64-channel rows (4x the ideal FLOPs - reported honestly as a
low fraction); n_filters = 128 runs on the general-channel kernels."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200.model import Generator  # noqa: E402

peak = 1425.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", peak)
B = 32
print("| n_filters | n_layers | input | GFLOP/frame | ms/batch(32) | frames/s | whole-model TFLOP/s (algorithmic) | frac of %.0f |" % peak)
print("|---|---|---|---|---|---|---|---|")
FS = tuple(int(a) for a in sys.argv[1:]) or (32, 64, 128)
for F, L in [(f, l) for f in FS for l in (4, 8, 12, 16)]:
    g = Generator(types.SimpleNamespace(n_filters=F, n_layers=L), compute_dtype=torch.float16)
    g = g.cuda().eval()
    for (h, w) in ((90, 160), (180, 320)):
        x = (torch.rand((B, 3, h, w)) * 2 - 1).cuda()
        with torch.no_grad():
            for _ in range(3):
                g(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = h * w * (2 * 27 * F + 2 * 9 * F * F * (2 * L + 1) + 2 * 9 * F * 4 * F * 5 + 16 * 2 * 9 * F * 3)
        tf = fl * B / (ms * 1e-3) / 1e12
        print(f"| {F} | {L} | {h}x{w} | {fl / 1e9:.3f} | {ms:.3f} | {B / (ms * 1e-3):.0f} | {tf:.0f} | {tf / peak:.3f} |", flush=True)
