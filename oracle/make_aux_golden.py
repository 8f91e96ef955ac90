"""Writes tests/golden/aux_golden.npz: outputs of the UNMODIFIED reference data path (dataloader.NumpyImagesDataset,
/root/reference/dataloader.py:9-38, torchvision v2.Resize) on seeded synthetic uint8 images, and of
torch.utils.data.RandomSampler(replacement=True) as train.py:69-80 builds it.  Run in the build container only
(`python oracle/make_aux_golden.py`); /root/reference does not exist on the GPU box, the fixture travels instead.
Asserts the oracle restatement (oracle/aux_oracle.py) reproduces the reference before writing.
There is nothing to generate for f3: torchmetrics (Pipfile: ==1.4.0) is not installed here - parity unpinned."""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import aux_oracle as A  # noqa: E402
import dataloader  # noqa: E402  (the reference module)


def main():
    out = {}
    rs = np.random.RandomState(1234)
    # smooth-ish synthetic images (random low-res field upsampled + noise) so that the resize sees structure
    shapes = [(3, 150, 170), (3, 96, 96), (3, 200, 131)]
    images = []
    for i, (c, h, w) in enumerate(shapes):
        base = rs.rand(c, h // 8 + 2, w // 8 + 2)
        img = np.kron(base, np.ones((8, 8)))[:, :h, :w] * 200 + rs.rand(c, h, w) * 55
        images.append(img.astype(np.uint8))
        out[f"img{i}"] = images[-1]
    d = tempfile.mkdtemp()
    paths = []
    for i, im in enumerate(images):
        paths.append(os.path.join(d, f"{i}.npy"))
        np.save(paths[-1], im)
    ds = dataloader.NumpyImagesDataset(paths, 24, 4)
    samples, lrs, hrs = [], [], []
    random.seed(77)
    state = random.getstate()
    for k in range(12):
        idx = k % len(images)
        lr, hr = ds[idx]
        lrs.append(lr.numpy())
        hrs.append(hr.numpy())
    random.setstate(state)
    for k in range(12):
        idx = k % len(images)
        _, h, w = images[idx].shape
        cy, cx = random.randint(0, h - 96), random.randint(0, w - 96)       # dataloader.py:27-29, same draw order
        samples.append((idx, cy, cx))
        lr_o, hr_o = A.crop_and_downscale(images[idx], cy, cx, 24, 4)
        assert np.array_equal(hr_o.numpy(), hrs[k]) and np.abs(lr_o.numpy() - lrs[k]).max() == 0.0
        lr_t, _ = A.crop_and_downscale(images[idx], cy, cx, 24, 4, use_torch=False)
        assert np.abs(lr_t.numpy() - lrs[k]).max() <= 2e-6, np.abs(lr_t.numpy() - lrs[k]).max()
    out["samples"] = np.array(samples, np.int32)
    out["lr"] = np.stack(lrs)
    out["hr_checksum"] = np.array([float(np.abs(h).sum()) for h in hrs])
    # a second geometry: lr 16, scale 2 (support 4, 9 taps)
    ds2 = dataloader.NumpyImagesDataset(paths, 16, 2)
    random.seed(5)
    lr2, hr2 = ds2[0]
    random.seed(5)
    cy, cx = random.randint(0, 150 - 32), random.randint(0, 170 - 32)
    out["samples_s2"] = np.array([(0, cy, cx)], np.int32)
    out["lr_s2"] = lr2.numpy()[None]
    lr_o, _ = A.crop_and_downscale(images[0], cy, cx, 16, 2, use_torch=False)
    assert np.abs(lr_o.numpy() - lr2.numpy()).max() <= 2e-6
    # sampler stream (train.py:69-80)
    from torch.utils.data import RandomSampler
    g = torch.Generator().manual_seed(1234)
    stream = torch.tensor(list(RandomSampler(range(800), replacement=True, num_samples=24 * 7 + 5, generator=g)))
    assert torch.equal(stream, A.replacement_sample_indices(800, 24 * 7 + 5, 1234))
    out["sampler_800_173_seed1234"] = stream.numpy()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "aux_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
