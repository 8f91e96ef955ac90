"""Per-tensor gradient parity of one GAN step (engine vs fp64 oracle) - diagnostic print-out."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import srgan_oracle as O  # noqa: E402
from test_train_step_gpu import _run_step  # noqa: E402

dt = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
tr, out, res, og, od = _run_step(dt)
e = tr.engine
for name, fp, ref in (("D", e.dp, res["d_grads"]), ("G", e.gp, res["g_grads"])):
    for k, gref in ref.items():
        got = fp.g[k].double().cpu() / e.S
        rel = ((got - gref).norm() / gref.norm().clamp_min(1e-30)).item()
        cos = (got.flatten() @ gref.flatten() / (got.norm() * gref.norm()).clamp_min(1e-30)).item()
        print(f"{dt} {name} {k:28s} rel-L2 {rel:.3e} cos {cos:.5f} |g| {gref.norm().item():.3e} |got| {got.norm().item():.3e}")
