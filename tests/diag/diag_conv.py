"""First-contact diagnostic for the tcgen05 conv kernel on a real B200 (run under `timeout`).
Prints per-tap / per-k-slice errors so a wrong descriptor field can be localised from one run."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fast_srgan_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
dt = torch.float16
print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


from fast_srgan_b200 import _lib  # noqa: E402
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.load().fsr_set_halo_mode(MODE)
_lib.load().fsr_set_ws_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
print("halo mode", MODE, flush=True)
g = torch.Generator().manual_seed(0)
N, H, W = 1, 16, 32
x = nhwc(torch.randn((N, 64, H, W), generator=g).cuda())
t0 = time.time()
for r in range(3):
    for s in range(3):
        for kq in range(4):
            w = torch.zeros(64, 64, 3, 3, device="cuda")
            w[:, kq * 16:(kq + 1) * 16, r, s] = (torch.randn((64, 16), generator=g) * 0.1).cuda().to(dt).float()
            wp, _ = ops.pack_conv3x3(w, None, dt)
            raw, st = ops.conv3x3_c64_raw_stats(x, wp)
            torch.cuda.synchronize()
            ref = F.conv2d(nchw(x), w, padding=1)
            err = (nchw(raw) - ref).abs().max().item()
            print(f"tap r={r} s={s} k16={kq}: max-abs err {err:.3e} (ref max {ref.abs().max().item():.3e})", flush=True)
print("delta-weight sweep done in %.1fs" % (time.time() - t0), flush=True)

w = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).cuda().to(dt).float()
wp, _ = ops.pack_conv3x3(w, None, dt)
for (N, H, W) in [(1, 8, 16), (2, 13, 21), (4, 90, 160), (32, 180, 320)]:
    x = nhwc(torch.randn((N, 64, H, W), generator=g).cuda())
    raw, st = ops.conv3x3_c64_raw_stats(x, wp)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x), w, padding=1)
    err = (nchw(raw) - ref).abs().max().item()
    s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], dim=-1)
    serr = ((st - s_ref).abs() / (s_ref.abs() + 1.0)).max().item()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(50):
        ops.conv3x3_c64_raw_stats(x, wp, st)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 50
    fl = 2.0 * N * H * W * 64 * 64 * 9
    print(f"full conv N={N} {H}x{W}: max-abs {err:.3e} stats-rel {serr:.3e}  {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
print("DIAG DONE", flush=True)
