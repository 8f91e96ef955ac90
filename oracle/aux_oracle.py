"""CPU oracle for the rows SURVEY.md section 8 marks "next": validation metrics (f3) and the
crop + bicubic-antialias data path (f4).  TEST INFRASTRUCTURE - never shipped, never timed as product;
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

f3  PSNR / SSIM as the reference computes them over the validation set (reference trainer.py:46-51, 53-69):
    torchmetrics.image.{PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure}(data_range=1.0, reduction="none").
    The arithmetic lives in a third-party dependency that is ABSENT here and under /root/reference:
    torchmetrics == 1.4.0 (reference Pipfile:13).  Restated below from its published algorithm
    (functional/image/psnr.py `_psnr_update/_psnr_compute`, functional/image/ssim.py `_ssim_update`):
      PSNR  = 10*log10(data_range^2 / (SSE_total / numel_total)), SSE accumulated over ALL update() calls (dim=None);
      SSIM  = per image: reflect-pad 5, 11x11 gaussian (sigma 1.5, normalised 1-D taps, outer product) depthwise conv of
              x, y, x*x, y*y, x*y; c1=(0.01*dr)^2, c2=(0.03*dr)^2;
              map = ((2 mu_x mu_y + c1)(2 s_xy + c2)) / ((mu_x^2 + mu_y^2 + c1)(s_x + s_y + c2));
              crop the padded border [5:-5] (only windows fully inside the image remain); mean over C x H' x W';
              reduction="none" keeps one value per image, `compute().mean()` (trainer.py:67) averages the images.
    PARITY UNPINNED for f3: no torchmetrics here to run, the reference holds no expected metric values.  Pinned only by
    known answers (identical images -> SSIM 1; constant offset d -> PSNR -20 log10 d) and scipy's gaussian filter.

f4  NumpyImagesDataset.__getitem__ (reference dataloader.py:24-38): random HR crop of a uint8 CHW image, LR =
    v2.Resize((lr, lr), antialias=True, interpolation=BICUBIC) of the float32 crop, both mapped x/127.5 - 1.
    The resize is torch's `_upsample_bicubic2d_aa` (separable, PIL-style: cubic a = -0.5, support = 2*scale, taps
    renormalised per output pixel): restated as explicit tap tables in `aa_bicubic_taps`.  PINNED against the unmodified
    reference class in tests/test_aux_oracle_cpu.py when /root/reference is present, and against the committed fixture
    tests/golden/aux_golden.npz (written by oracle/make_aux_golden.py) everywhere else.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- f3: PSNR / SSIM (torchmetrics 1.4.0 semantics)
def gaussian_taps(kernel_size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> Tensor:
    """torchmetrics functional/image/helper.py `_gaussian`: exp(-(d/sigma)^2/2) over d = -(k-1)/2..(k-1)/2, sum-normalised."""
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=dtype)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return g / g.sum()


def ssim_per_image(pred: Tensor, target: Tensor, data_range: float = 1.0, kernel_size: int = 11, sigma: float = 1.5,
                   k1: float = 0.01, k2: float = 0.03) -> Tensor:
    """torchmetrics `_ssim_update` (gaussian_kernel=True, reduction='none'): [N,C,H,W] x2 -> [N]."""
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    N, C, H, W = pred.shape
    pad = (kernel_size - 1) // 2
    g = gaussian_taps(kernel_size, sigma, pred.dtype)
    kernel = (g[:, None] * g[None, :]).expand(C, 1, kernel_size, kernel_size).contiguous()
    p = F.pad(pred, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    out = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=C).split(N)
    mu_p2, mu_t2, mu_pt = out[0] * out[0], out[1] * out[1], out[0] * out[1]
    s_p = torch.clamp(out[2] - mu_p2, min=0.0)
    s_t = torch.clamp(out[3] - mu_t2, min=0.0)
    s_pt = out[4] - mu_pt
    m = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_p2 + mu_t2 + c1) * (s_p + s_t + c2))
    m = m[..., pad:-pad, pad:-pad]
    return m.reshape(N, -1).mean(-1)


def psnr_from_sse(sse: float, numel: int, data_range: float = 1.0) -> float:
    """torchmetrics `_psnr_compute` (base 10, dim=None)."""
    return (2 * math.log(data_range) - math.log(sse / numel)) * (10 / math.log(10))


def validation_metrics(sr_batches, hr_batches):
    """reference trainer.py:60-68: sr = (1 + G(lr))/2 and hr = (1 + hr)/2 per batch -> (mean SSIM over images, PSNR over
    the whole set).  Inputs are the [-1,1] generator outputs / targets, fp32 NCHW."""
    sse, numel, ssims = 0.0, 0, []
    for sr, hr in zip(sr_batches, hr_batches):
        a, b = (1.0 + sr) / 2.0, (1.0 + hr) / 2.0
        sse += float(((a - b).double() ** 2).sum())
        numel += a.numel()
        ssims.append(ssim_per_image(a, b))
    return float(torch.cat(ssims).mean()), psnr_from_sse(sse, numel)


# --------------------------------------------------------------------------- f4: crop + antialiased bicubic downscale
def _cubic_aa(x: float, a: float = -0.5) -> float:
    """ATen UpSample.h `bicubic_aa_filter`: Keys cubic with a = -0.5 (the PIL convention; NOT the -0.75 of non-AA bicubic)."""
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def aa_bicubic_taps(in_size: int, out_size: int):
    """ATen UpSampleKernel.cpp `_compute_indices_min_size_weights_aa` (align_corners=False): per output index the first
    input index, the tap count and the normalised float32 taps.  Returns (xmin[int32 out], xsize[int32 out], w[float32 out x K])."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    K = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xsize = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, K), np.float32)
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(0, int(center - support + 0.5))
        size = min(in_size, int(center + support + 0.5)) - lo
        taps = np.array([_cubic_aa((j + lo - center + 0.5) * invscale) for j in range(size)], np.float32)
        total = np.float32(taps.sum(dtype=np.float32))
        xmin[i], xsize[i] = lo, size
        w[i, :size] = taps / total
    return xmin, xsize, w


def resize_aa_bicubic(img: Tensor, out_h: int, out_w: int) -> Tensor:
    """Separable restatement (horizontal pass first, then vertical, as ATen's separable kernel does for the last two
    dims): [.., H, W] float32 -> [.., out_h, out_w]."""
    H, W = img.shape[-2:]
    xm, xs, xw = aa_bicubic_taps(W, out_w)
    ym, ys, yw = aa_bicubic_taps(H, out_h)
    tmp = torch.zeros(img.shape[:-1] + (out_w,), dtype=torch.float32)
    for j in range(out_w):
        taps = torch.from_numpy(xw[j, :xs[j]])
        tmp[..., j] = (img[..., xm[j]:xm[j] + xs[j]].float() * taps).sum(-1)
    out = torch.zeros(img.shape[:-2] + (out_h, out_w), dtype=torch.float32)
    for i in range(out_h):
        taps = torch.from_numpy(yw[i, :ys[i]])
        out[..., i, :] = (tmp[..., ym[i]:ym[i] + ys[i], :] * taps[:, None]).sum(-2)
    return out


def crop_and_downscale(image_u8: np.ndarray, crop_y: int, crop_x: int, lr_size: int, scale: int, use_torch: bool = True):
    """reference dataloader.py:24-38 with the two random.randint draws passed in: uint8 CHW image ->
    (lr [3,lr,lr], hr [3,lr*scale,lr*scale]) float32 in [-1,1] (LR may overshoot: the resize works on floats, no clamp)."""
    hr_size = lr_size * scale
    hr = torch.tensor(np.ascontiguousarray(image_u8[:, crop_y:crop_y + hr_size, crop_x:crop_x + hr_size]), dtype=torch.float32)
    if use_torch:   # exactly what torchvision's v2.Resize dispatches to for a float tensor
        lr = F.interpolate(hr[None], size=(lr_size, lr_size), mode="bicubic", antialias=True, align_corners=False)[0]
    else:
        lr = resize_aa_bicubic(hr, lr_size, lr_size)
    return lr / 127.5 - 1.0, hr / 127.5 - 1.0


def replacement_sample_indices(num_images: int, num_samples: int, seed: int) -> Tensor:
    """torch.utils.data.RandomSampler(replacement=True, num_samples, generator=g) (reference train.py:69-80): the index
    stream it yields, 32 at a time plus the remainder (torch/utils/data/sampler.py `RandomSampler.__iter__`)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_samples // 32):
        out.append(torch.randint(high=num_images, size=(32,), dtype=torch.int64, generator=g))
    out.append(torch.randint(high=num_images, size=(num_samples % 32,), dtype=torch.int64, generator=g))
    return torch.cat(out)
