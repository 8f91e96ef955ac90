"""Mirror of the reference's inference.py (CLI: --image_dir --output_dir) on the B200 generator.

The per-image pixel pipeline of inference.py:48-56 (uint8 HWC -> [-1,1] -> G -> uint8, truncating cast) runs fused
on the GPU (`Generator.super_resolve_u8`).  PIL is used for file I/O only, like the reference."""
from __future__ import annotations

import os
from argparse import ArgumentParser

import numpy as np
import torch

from . import config as cfg_mod
from .model import Generator


def main(argv=None):
    parser = ArgumentParser("Real Time Image Super Resolution (B200)")
    parser.add_argument("--image_dir", required=True, type=str)
    parser.add_argument("--output_dir", required=True, type=str)
    parser.add_argument("--config", default=None, type=str)
    parser.add_argument("--weights", default="models/model.pt", type=str)
    args = parser.parse_args(argv)
    from PIL import Image
    os.makedirs(args.output_dir, exist_ok=True)
    config = cfg_mod.load(args.config)
    model = Generator(config.generator)
    model.load_state_dict(torch.load(args.weights, map_location="cpu"))       # inference.py:29-33
    model.to("cuda").eval()
    names = sorted(x for x in os.listdir(args.image_dir) if x.lower().endswith((".png", ".jpg", "jpeg")))
    print(f"Found {len(names)} to super resolve, starting...")
    for name in names:
        img = np.array(Image.open(os.path.join(args.image_dir, name)).convert("RGB"))
        sr = model.super_resolve_u8(torch.from_numpy(img).unsqueeze(0).cuda())[0].cpu().numpy()
        Image.fromarray(sr).save(os.path.join(args.output_dir, os.path.basename(name)))


if __name__ == "__main__":
    main()
