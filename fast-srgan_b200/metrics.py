"""Validation metrics of the reference Trainer (trainer.py:46-51, 53-69) on the B200 library: PSNR and SSIM
(torchmetrics 1.4.0 semantics, data_range = 1, 11x11 gaussian sigma 1.5, reduction "none") in ONE fused CUDA pass per
batch (`fsr_psnr_ssim`), state kept on the device like a torchmetrics Metric (reset / update / compute).
No CPU fallback: CPU tensors raise."""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List

import torch
import torch.distributed as dist

from . import _lib as L

KERNEL_SIZE, SIGMA = 11, 1.5


def gaussian_taps() -> List[float]:
    """The normalised 1-D window torchmetrics builds (`_gaussian`, float32 arithmetic)."""
    d = torch.arange((1 - KERNEL_SIZE) / 2, (1 + KERNEL_SIZE) / 2, 1, dtype=torch.float32)
    g = torch.exp(-torch.pow(d / SIGMA, 2) / 2)
    return (g / g.sum()).tolist()


class ValidationMetrics:
    """ssim / psnr objects of trainer.py:46-51 folded into one accumulator.

    update(sr, hr): fp32 NCHW CUDA tensors; with rescale=True (default) both are mapped (1 + v)/2 first, exactly
    trainer.py:64-66.  compute() -> {"ssim": mean over images of the per-image SSIM, "psnr": PSNR of the pooled MSE}
    (= `self.ssim.compute().mean()`, `self.psnr.compute().mean()` of trainer.py:67-68)."""

    def __init__(self, device, data_range: float = 1.0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ValidationMetrics runs on the CUDA library only - there is no CPU fallback")
        self.data_range = float(data_range)
        self._taps = (ctypes.c_float * KERNEL_SIZE)(*gaussian_taps())
        self.reset()

    def reset(self):
        self._sse = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._numel = 0
        self._ssim: List[torch.Tensor] = []

    @torch.no_grad()
    def update(self, sr: torch.Tensor, hr: torch.Tensor, rescale: bool = True):
        if sr.device.type != "cuda" or hr.device.type != "cuda":
            raise RuntimeError("ValidationMetrics.update needs CUDA tensors (no CPU fallback)")
        if sr.shape != hr.shape or sr.dim() != 4:
            raise RuntimeError(f"prediction {tuple(sr.shape)} and target {tuple(hr.shape)} must be equal NCHW shapes")
        sr, hr = sr.float().contiguous(), hr.float().contiguous()
        N, C, H, W = sr.shape
        sums = torch.zeros(N, dtype=torch.float64, device=sr.device)
        s = 0.5 if rescale else 1.0
        b = 0.5 if rescale else 0.0
        step = max(1, 65535 // C)                      # one launch covers at most 65535 (image, channel) planes (grid.z)
        for i in range(0, N, step):
            n = min(step, N - i)
            L.check(L.load().fsr_psnr_ssim(sr[i:i + n].data_ptr(), hr[i:i + n].data_ptr(), n, C, H, W, s, b, self.data_range,
                                           ctypes.cast(self._taps, ctypes.c_void_p), self._sse.data_ptr(), sums[i:i + n].data_ptr(),
                                           L.stream_ptr(sr.device)), "psnr/ssim")
        self._ssim.append(sums / float(C * (H - KERNEL_SIZE + 1) * (W - KERNEL_SIZE + 1)))
        self._numel += sr.numel()

    @staticmethod
    def pool(stat: torch.Tensor, data_range: float, sync: bool = True) -> Dict[str, float]:
        """stat = [sum of squared errors, element count, sum of per-image SSIM, image count] (float64, any device) ->
        pooled metrics; summed over ranks first when torch.distributed is initialised (what torchmetrics'
        dist_reduce_fx = "sum" / "cat" followed by .mean() amounts to)."""
        if sync and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stat)
        sse, numel, ssum, n_img = (float(v) for v in stat.tolist())
        mse = sse / numel
        psnr = float("inf") if mse == 0.0 else 10.0 * math.log10(data_range ** 2 / mse)
        return {"ssim": ssum / n_img, "psnr": psnr}

    def compute(self, sync: bool = True) -> Dict[str, float]:
        """{"ssim": mean per-image SSIM, "psnr": PSNR of the pooled MSE, "ssim_per_image": this rank's values}."""
        if not self._ssim:
            raise RuntimeError("compute() before any update()")
        ssim = torch.cat(self._ssim)
        stat = torch.stack([self._sse[0], torch.tensor(float(self._numel), dtype=torch.float64, device=self.device),
                            ssim.sum(), torch.tensor(float(ssim.numel()), dtype=torch.float64, device=self.device)])
        out = self.pool(stat, self.data_range, sync)
        out["ssim_per_image"] = ssim
        return out
