"""Fixture for the ONLY artefact the reference ships for this path: its pretrained generator `models/model.pt`
(3.7 MB, 36 fp32 tensors saved from a torch.compile'd module: keys carry `_orig_mod.`, inference.py:27-33).

    python oracle/make_ckpt_golden.py        # needs /root/reference; writes tests/golden/checkpoint_golden.npz

Imports the UNMODIFIED /root/reference/model.py, loads the checkpoint the way inference.py:27-35 does, runs
`Generator.forward` (model.py:112-117) in fp32 on the CPU on the survey's anchor input
(`torch.manual_seed(0); x = torch.rand(2,3,90,160)*2-1`), asserts the oracle restatement reproduces it,
and stores: the state dict (original key names, `_orig_mod.` prefix kept - the loader must strip it), frame 0 of
the input, the reference's output for frame 0 and the survey's mean/std/abs-max anchor over both frames.
The GPU box has no /root/reference: tests/test_checkpoint_gpu.py uses this file only.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import srgan_oracle as O  # noqa: E402

REF = "/root/reference"


def main():
    sys.path.insert(0, REF)
    import model  # the reference's model.py, unmodified
    raw = torch.load(f"{REF}/models/model.pt", map_location="cpu")
    weights = {k.replace("_orig_mod.", ""): v for k, v in raw.items()}          # inference.py:30-33
    g = model.Generator(types.SimpleNamespace(n_filters=64, n_layers=8))
    print(g.load_state_dict(weights))
    g.eval()
    torch.manual_seed(0)
    x = torch.rand(2, 3, 90, 160) * 2 - 1
    with torch.no_grad():
        y = g(x)
        yo = O.generator_forward(weights, x)
    err = (y - yo).abs().max().item()
    print(f"oracle vs reference on the checkpoint: max-abs {err:.3e}")
    assert err <= 5e-5
    anchor = np.array([y.mean().item(), y.std().item(), y.abs().max().item()])
    print("anchor mean/std/absmax", anchor, "(SURVEY 8c: 0.048907 0.404062 0.998066)")
    out = {"sd/" + k: v.numpy() for k, v in raw.items()}
    out["x0"] = x[0:1].numpy()
    out["y0"] = y[0:1].numpy()
    out["anchor"] = anchor
    path = os.path.join(ROOT, "tests", "golden", "checkpoint_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
