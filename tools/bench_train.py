"""GAN train-step timing (BASELINE configs[2]/[3]): python tools/bench_train.py [--batch 64] [--steps 10]
Under torchrun: per-rank batch = --batch (weak scaling), NCCL gradient all-reduce.  Prints one JSON line."""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200 import _lib as L  # noqa: E402
from fast_srgan_b200 import distributed as D  # noqa: E402
from fast_srgan_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()
rank, local, world = D.init_from_env("nccl")
torch.cuda.set_device(local)
ns = types.SimpleNamespace
cfg = ns(experiment=ns(name="b", seed=0), generator=ns(n_filters=64, n_layers=8), discriminator=ns(n_filters=64, n_layers=7),
         training=ns(device=f"cuda:{local}", generator_lr=1e-4, discriminator_lr=1e-4))
torch.manual_seed(1234)     # torch default init, identical on every rank
tr = Trainer(cfg, compute_dtype=torch.float16 if args.dtype == "fp16" else torch.bfloat16)
B = args.batch
g = torch.Generator().manual_seed(rank)
lr = (torch.rand((B, 3, 24, 24), generator=g) * 2 - 1).cuda()
hr = (torch.rand((B, 3, 96, 96), generator=g) * 2 - 1).cuda()
noise = {k: torch.rand((B, 1, 6, 6), generator=g).cuda() for k in ("d_real", "d_fake", "g_real")}
for _ in range(args.warmup):
    out = tr.train_step(lr, hr, noise=noise)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
n0 = L.load().fsr_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = tr.train_step(lr, hr, noise=noise)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
    flops = 2636e9 * B / 64.0      # needed FLOPs per step at batch 64
    print(json.dumps({"metric": "GAN train-step ms (trainer.py:168-196), 24x24 LR / 96x96 HR", "ms_per_step": ms.item(),
                      "per_gpu_batch": B, "n_gpus": world, "global_batch": B * world, "dtype": args.dtype,
                      "samples_per_s": B * world / (ms.item() / 1e3), "tflops_per_gpu": flops / (ms.item() * 1e-3) / 1e12,
                      "gpu_launches_per_step": (L.load().fsr_launch_count() - n0) / args.steps,
                      "losses": {k: float(v) for k, v in out.items() if k != "sr"}}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
