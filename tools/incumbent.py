"""The incumbent on the same box: the reference's network written with stock torch.nn.functional ops,
run by PyTorch eager on the GPU (cuDNN, channels_last, bf16 autocast) and, when it compiles in time, torch.compile.
Not part of the product and not a parity check: it only gives the library number beside ours.

    python tools/incumbent.py [--steps 10] [--no-compile] [--no-train]

Prints one JSON line: generator frames/s at b32 180x320 (fp16/bf16 autocast), GAN train-step ms at b64 24x24.
Same architecture as model.py:28-117 (generator), :122-193 (discriminator), :6-23 (VGG19 features to relu5_3), same
step as trainer.py:168-196; weights random (torch default init), data synthetic."""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-compile", action="store_true")
ap.add_argument("--no-train", action="store_true")
ap.add_argument("--compile-mode", default="max-autotune", help="torch.compile mode; the reference uses max-autotune (trainer.py:23-26)")
ap.add_argument("--compile-train", action="store_true", help="also time the GAN step with torch.compile'd G / D / VGG (trainer.py:23-26)")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True


class Res(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(f, f, 3, padding=1, bias=False), nn.Conv2d(f, f, 3, padding=1, bias=False)
        self.n1, self.n2, self.a = nn.InstanceNorm2d(f), nn.InstanceNorm2d(f), nn.PReLU()

    def forward(self, x):
        return x + self.n2(self.c2(self.a(self.n1(self.c1(x)))))


class Up(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.c, self.a = nn.Conv2d(f, 4 * f, 3, padding=1), nn.PReLU()

    def forward(self, x):
        return self.a(F.pixel_shuffle(self.c(x), 2))


class Gen(nn.Module):
    def __init__(self, f=64, layers=8):
        super().__init__()
        self.neck = nn.Sequential(nn.Conv2d(3, f, 3, padding=1), nn.PReLU())
        self.stem = nn.Sequential(*[Res(f) for _ in range(layers)])
        self.bott = nn.Sequential(nn.Conv2d(f, f, 3, padding=1, bias=False), nn.InstanceNorm2d(f))
        self.up = nn.Sequential(Up(f), Up(f))
        self.head = nn.Conv2d(f, 3, 3, padding=1)

    def forward(self, x):
        r = self.neck(x)
        return torch.tanh(self.head(self.up(self.bott(self.stem(r)) + r)))


def disc(f=64):
    layers = [nn.Conv2d(3, f, 3, padding=1), nn.LeakyReLU(0.2)]
    cin = f
    for i, s in enumerate((2, 1, 2, 1, 2, 1, 2)):
        cout = f * min(8, 2 ** ((i + 1) // 2))
        layers += [nn.Conv2d(cin, cout, 3, stride=s, padding=1, bias=False), nn.InstanceNorm2d(cout), nn.LeakyReLU(0.01)]
        cin = cout
    return nn.Sequential(*layers, nn.Conv2d(cin, 1, 1))


def vgg():
    cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512]
    layers, cin = [], 3
    for v in cfg:
        if v == "M":
            layers.append(nn.MaxPool2d(2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU()]
            cin = v
    return nn.Sequential(*layers)


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {"what": "stock PyTorch %s (cuDNN %s) on the same GPU" % (torch.__version__, torch.backends.cudnn.version())}
torch.manual_seed(1234)
gen = Gen().to(dev).to(memory_format=torch.channels_last).eval()
x = (torch.rand(32, 3, 180, 320, device=dev) * 2 - 1).contiguous(memory_format=torch.channels_last)
for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
    def run():
        with torch.no_grad(), torch.autocast("cuda", dtype=dt):
            return gen(x)
    ms = timed(run, args.steps)
    out["generator_eager_%s" % name] = {"ms_per_step": round(ms, 3), "fps": round(32 / ms * 1e3, 1)}
if not args.no_compile:
    try:
        t0 = time.time()
        cg = torch.compile(gen, mode=args.compile_mode)

        def run_c():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return cg(x)
        ms = timed(run_c, args.steps)
        out["generator_compile_bf16"] = {"ms_per_step": round(ms, 3), "fps": round(32 / ms * 1e3, 1), "compile_s": round(time.time() - t0, 1),
                                         "mode": args.compile_mode}
    except Exception as e:  # noqa: BLE001  (inductor needs a C compiler / triton: report, do not fail)
        out["generator_compile_bf16"] = {"error": repr(e)[:200]}
del x

if not args.no_train:
    B = 64
    G = Gen().to(dev).to(memory_format=torch.channels_last).train()
    D = disc().to(dev).to(memory_format=torch.channels_last).train()
    V = vgg().to(dev).to(memory_format=torch.channels_last).eval().requires_grad_(False)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
    og = torch.optim.AdamW(G.parameters(), lr=1e-4, fused=True)
    od = torch.optim.AdamW(D.parameters(), lr=1e-4, fused=True)
    lr = (torch.rand(B, 3, 24, 24, device=dev) * 2 - 1).contiguous(memory_format=torch.channels_last)
    hr = (torch.rand(B, 3, 96, 96, device=dev) * 2 - 1).contiguous(memory_format=torch.channels_last)
    bce, l1 = nn.BCEWithLogitsLoss(), nn.SmoothL1Loss()

    def feat(img):
        return V(((img + 1) / 2 - mean) / std)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            od.zero_grad(set_to_none=True)
            y_real = D(hr)
            sr = G(lr)
            y_fake = D(sr.detach())
            loss_d = 0.5 * bce(y_real, torch.ones_like(y_real) - 0.2 * torch.rand_like(y_real)) \
                + 0.5 * bce(y_fake, 0.2 * torch.rand_like(y_fake))
        loss_d.backward()
        od.step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            og.zero_grad(set_to_none=True)
            sr = G(lr)
            y_fake = D(sr)
            adv = 1e-1 * bce(y_fake, torch.ones_like(y_fake) - 0.2 * torch.rand_like(y_fake))
            content = l1(feat(sr), feat(hr))
            loss_g = 0.5 * adv + 0.5 * content
        loss_g.backward()
        og.step()
        return loss_d
    ms = timed(step, args.steps)
    out["train_step_eager_bf16_b64"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1)}
    if args.compile_train and not args.no_compile:
        try:
            t0 = time.time()
            G, D, V = (torch.compile(m, mode=args.compile_mode) for m in (G, D, V))    # trainer.py:23-26
            ms = timed(step, args.steps)
            out["train_step_compile_bf16_b64"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
                                                  "compile_s": round(time.time() - t0, 1), "mode": args.compile_mode}
        except Exception as e:  # noqa: BLE001
            out["train_step_compile_bf16_b64"] = {"error": repr(e)[:200]}

print(json.dumps(out), flush=True)
