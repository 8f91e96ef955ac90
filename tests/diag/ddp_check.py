"""torchrun --nproc-per-node 2 tests/diag/ddp_check.py : N-rank sharded train_step == 1-rank full-batch train_step.
Each rank also runs the FULL batch alone (no collective) and compares parameters after the step."""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import srgan_oracle as O  # noqa: E402
from fast_srgan_b200 import distributed as D  # noqa: E402
from fast_srgan_b200.trainer import Trainer  # noqa: E402

rank, local, world = D.init_from_env("nccl")
torch.cuda.set_device(local)


def ns(**k):
    return types.SimpleNamespace(**k)


def make():
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=8), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device=f"cuda:{local}", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=torch.bfloat16)
    tr.generator.load_state_dict(O.make_generator_state(64, 8, 1234))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    tr.perceptual_network.load_state_dict(O.make_vgg19_state(99))
    return tr


B = 4 * world
g = torch.Generator().manual_seed(5)
lr = torch.rand((B, 3, 24, 24), generator=g) * 2 - 1
hr = torch.rand((B, 3, 96, 96), generator=g) * 2 - 1
noise = {k: torch.rand((B, 1, 6, 6), generator=g) for k in ("d_real", "d_fake", "g_real")}

STEPS = int(os.environ.get("DDP_CHECK_STEPS", "4"))   # 2 eager steps, then the captured CUDA graph (collectives inside)
tr = make()                                   # distributed: this rank's shard
if world > 1 and os.environ.get("FSR_NCCL_CAPI", "1") != "0":
    assert tr.engine.comm is not None and tr.engine.comm.native, "the exchange must go through libfsr_b200's fsr_nccl_*"
for _ in range(STEPS):
    out = tr.train_step(D.shard_batch(lr, rank, world), D.shard_batch(hr, rank, world), noise=D.shard_noise(noise, rank, world))
torch.cuda.synchronize()

ref = make()
ref.engine.world = 1                          # full batch on this GPU alone, no collective
for _ in range(STEPS):
    ref.train_step(lr, hr, noise=noise)
torch.cuda.synchronize()

worst = 0.0
for name in ("gp", "dp"):
    a, b = getattr(tr.engine, name).flat, getattr(ref.engine, name).flat
    # both runs move each weight by ~lr*sign(g); differences come only from fp32 summation order (sharded vs full)
    diff = (a - b).abs().max().item()
    frac = ((a - b).abs() > 1e-6).float().mean().item()
    worst = max(worst, diff)
    print(f"[rank {rank}] {name}: max |param_ddp - param_full| = {diff:.3e}, fraction differing > 1e-6: {frac:.4f}", flush=True)
t = torch.tensor([worst], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # all ranks must hold bit-identical parameters after the step
    chk = tr.engine.gp.flat.double().sum() + tr.engine.dp.flat.double().sum()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("replica checksum spread:", (hi - lo).item(), flush=True)
        assert (hi - lo).item() == 0.0
if rank == 0:
    print("DDP CHECK worst", t.item(), flush=True)
    assert t.item() <= 2.05e-4 * STEPS
if world > 1:
    dist.destroy_process_group()
