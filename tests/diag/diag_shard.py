"""Diagnostic: discriminator forward/backward at batch B vs two half batches (must agree: ops are per-sample)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import srgan_oracle as O  # noqa: E402
from fast_srgan_b200 import ops  # noqa: E402
from fast_srgan_b200.trainer import Trainer  # noqa: E402

ns = types.SimpleNamespace
dt = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=2), discriminator=ns(n_filters=64, n_layers=7),
         training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
tr = Trainer(cfg, compute_dtype=dt)
tr.generator.load_state_dict(O.make_generator_state(64, 2, 1234))
tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
e = tr.engine
B = 16
g = torch.Generator().manual_seed(77)
hr = (torch.rand((B, 3, 96, 96), generator=g) * 2 - 1).cuda()
noise = torch.rand((B, 36), generator=g).cuda()


def run(sl):
    e.dp.zero_grad()
    z, ctx = e.D.forward(hr[sl].contiguous(), save=True)
    dz = torch.empty_like(z)
    loss = torch.zeros(1, device="cuda")
    ops.bce_logits(z, noise[sl].contiguous(), 0.3, 0.8, loss, dz, grad_scale=1.0)
    e.D.backward(ctx, dz, wgrad=True, d_img=None)
    torch.cuda.synchronize()
    return z.clone(), {k: v.clone() for k, v in e.dp.g.items()}, [(l[1].clone(), l[2].clone()) for l in ctx["layers"]]


zf, gf, lf = run(slice(0, B))
za, ga, la = run(slice(0, B // 2))
zb, gb, lb = run(slice(B // 2, B))
zf2, gf2, _ = run(slice(0, B))
print("logits full vs halves max-abs", (zf - torch.cat([za, zb])).abs().max().item(), " run-to-run", (zf - zf2).abs().max().item())
for i in range(7):
    raw_f, st_f = lf[i]
    raw_h = torch.cat([la[i][0], lb[i][0]])
    st_h = torch.cat([la[i][1], lb[i][1]])
    print(f"layer {i}: raw max-abs diff {(raw_f.float() - raw_h.float()).abs().max().item():.3e}  stats rel diff "
          f"{((st_f - st_h).abs().max() / st_f.abs().max()).item():.3e}")
for k in gf:
    h = 0.5 * (ga[k] + gb[k])
    print(f"{k:24s} full-vs-halves rel-L2 {((gf[k] - h).norm() / gf[k].norm()).item():.3e}   run-to-run {((gf[k] - gf2[k]).norm() / gf[k].norm()).item():.3e}")
