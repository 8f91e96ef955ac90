"""GPU data path of the reference training loop: `NumpyImagesDataset` (dataloader.py:9-38) and the
replacement sampler / DataLoader wiring of train.py:62-113, re-designed for a B200 box.

The reference decodes nothing at train time either (images are pre-converted uint8 CHW .npy files, train.py:40-58), but
crops and bicubic-downscales every sample on CPU workers: at a few ms per GAN step 16 workers cannot feed 8 GPUs.  Here
the uint8 images live in HBM once (DIV2K is ~3 GB as uint8; 180 GB per GPU), and a batch is ONE kernel launch
(`fsr_crop_resize_aa`): crop + antialiased bicubic (the op v2.Resize runs) + x/127.5 - 1 for LR and HR together.
Only 12 bytes per sample (image index, crop_y, crop_x) cross PCIe per step.  No CPU fallback."""
from __future__ import annotations

import math
import random
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def _cubic_aa(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    far = (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def aa_bicubic_tap_table(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per output index of an antialiased bicubic resize in_size -> out_size (align_corners=False): first input index,
    tap count, normalised float32 taps - the table ATen's `_compute_indices_min_size_weights_aa` builds (cubic a=-0.5,
    support 2*scale, window [int(c - s + .5), int(c + s + .5)) clipped to the image, taps renormalised)."""
    scale = in_size / out_size
    support = 2.0 * max(scale, 1.0)
    inv = 1.0 / max(scale, 1.0)
    K = int(math.ceil(support)) * 2 + 1
    centers = scale * (np.arange(out_size, dtype=np.float64) + 0.5)
    lo = np.maximum(0, (centers - support + 0.5).astype(np.int64))          # astype truncates toward zero like C's (int)
    hi = np.minimum(in_size, (centers + support + 0.5).astype(np.int64))
    size = (hi - lo).astype(np.int32)
    j = np.arange(K)[None, :]
    w = _cubic_aa((j + lo[:, None] - centers[:, None] + 0.5) * inv).astype(np.float32)
    w[j >= size[:, None]] = 0.0
    w = w / w.sum(axis=1, dtype=np.float32, keepdims=True)
    return lo.astype(np.int32), size, np.ascontiguousarray(w, dtype=np.float32)


class DeviceImageCache:
    """All training images as uint8 CHW, back to back in one HBM buffer (+ per-image offset / height / width tables)."""

    def __init__(self, images: Sequence[np.ndarray], device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceImageCache lives in GPU memory - there is no CPU fallback")
        offs, hs, ws, total = [], [], [], 0
        for im in images:
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[0] != 3:
                raise ValueError("images must be uint8 CHW arrays with 3 channels (train.py:40-58 writes exactly that)")
            offs.append(total)
            hs.append(im.shape[1])
            ws.append(im.shape[2])
            total += im.size
        host = torch.empty(total, dtype=torch.uint8).pin_memory()
        for im, o in zip(images, offs):
            host[o:o + im.size] = torch.from_numpy(np.ascontiguousarray(im)).reshape(-1)
        self.data = host.to(self.device, non_blocking=True)
        self.offsets = torch.tensor(offs, dtype=torch.int64, device=self.device)
        self.heights = torch.tensor(hs, dtype=torch.int32, device=self.device)
        self.widths = torch.tensor(ws, dtype=torch.int32, device=self.device)
        self.shapes: List[Tuple[int, int]] = list(zip(hs, ws))
        torch.cuda.current_stream(self.device).synchronize()

    @classmethod
    def from_numpy_paths(cls, paths: Sequence[str], device="cuda") -> "DeviceImageCache":
        return cls([np.load(p, mmap_mode="r") for p in paths], device)     # dataloader.py:25

    def __len__(self) -> int:
        return len(self.shapes)


def crop_resize_batch(cache: DeviceImageCache, samples: torch.Tensor, lr_size: int, scale: int, taps=None):
    """samples int32 [B,3] (image index, crop_y, crop_x), host or device -> (lr [B,3,lr,lr], hr [B,3,lr*scale,lr*scale])
    fp32 NCHW in [-1,1] on the cache's device: dataloader.py:24-38 for a whole batch in one launch."""
    dev = cache.device
    samples = samples.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    B = samples.shape[0]
    hr_size = lr_size * scale
    if taps is None:
        taps = tuple(torch.from_numpy(t).to(dev) for t in aa_bicubic_tap_table(hr_size, lr_size))
    tmin, tsize, tw = taps
    lr = torch.empty((B, 3, lr_size, lr_size), dtype=torch.float32, device=dev)
    hr = torch.empty((B, 3, hr_size, hr_size), dtype=torch.float32, device=dev)
    L.check(L.load().fsr_crop_resize_aa(cache.data.data_ptr(), cache.offsets.data_ptr(), cache.heights.data_ptr(),
                                        cache.widths.data_ptr(), samples.data_ptr(), B, lr_size, scale, tmin.data_ptr(),
                                        tsize.data_ptr(), tw.data_ptr(), tw.shape[1], lr.data_ptr(), hr.data_ptr(),
                                        L.stream_ptr(dev)), "crop + resize")
    return lr, hr


class ShardedReplacementSampler:
    """torch.utils.data.RandomSampler(replacement=True, num_samples, generator) as train.py:69-80 builds it, sharded:
    every rank draws the SAME global index stream (same seed) and keeps samples [rank*B/W, (rank+1)*B/W) of each global
    batch - the batch partition of SURVEY.md 8e, so W ranks at B/W reproduce one process at batch B."""

    def __init__(self, num_images: int, num_samples: int, global_batch: int, seed: int, rank: int = 0, world: int = 1):
        if global_batch % world:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
        self.num_images, self.num_samples, self.global_batch = num_images, num_samples, global_batch
        self.seed, self.rank, self.world = seed, rank, world

    def global_indices(self) -> torch.Tensor:
        g = torch.Generator().manual_seed(self.seed)
        chunks = [torch.randint(high=self.num_images, size=(32,), dtype=torch.int64, generator=g)
                  for _ in range(self.num_samples // 32)]
        chunks.append(torch.randint(high=self.num_images, size=(self.num_samples % 32,), dtype=torch.int64, generator=g))
        return torch.cat(chunks)

    def __len__(self) -> int:
        return self.num_samples // self.global_batch            # drop_last=True (train.py:96,107)

    def __iter__(self) -> Iterator[torch.Tensor]:
        idx = self.global_indices()
        per = self.global_batch // self.world
        for b in range(len(self)):
            s = b * self.global_batch + self.rank * per
            yield idx[s:s + per]


class GpuCropLoader:
    """Drop-in for the train / pretrain DataLoaders of train.py:92-113: iterating yields (lr_images, hr_images) fp32 NCHW
    CUDA batches.  Crop offsets are drawn per sample with random.randint like dataloader.py:27-29, from a per-rank
    `random.Random(seed + rank)`."""

    def __init__(self, cache: DeviceImageCache, sampler: ShardedReplacementSampler, lr_image_size: int, scale_factor: int,
                 seed: int = 0):
        self.cache, self.sampler = cache, sampler
        self.lr_size, self.scale = int(lr_image_size), int(scale_factor)
        self._rng = random.Random(seed + sampler.rank)
        self._taps = tuple(torch.from_numpy(t).to(cache.device) for t in aa_bicubic_tap_table(self.lr_size * self.scale, self.lr_size))
        hr = self.lr_size * self.scale
        for h, w in cache.shapes:
            if h < hr or w < hr:
                raise ValueError(f"image {h}x{w} is smaller than the {hr}x{hr} HR crop")

    def __len__(self) -> int:
        return len(self.sampler)

    def draw(self, indices: torch.Tensor) -> torch.Tensor:
        hr = self.lr_size * self.scale
        rows = []
        for i in indices.tolist():
            h, w = self.cache.shapes[i]
            rows.append((i, self._rng.randint(0, h - hr), self._rng.randint(0, w - hr)))    # dataloader.py:27-29
        return torch.tensor(rows, dtype=torch.int32).pin_memory()

    def __iter__(self):
        for indices in self.sampler:
            yield crop_resize_batch(self.cache, self.draw(indices), self.lr_size, self.scale, self._taps)
