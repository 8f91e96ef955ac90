"""GPU parity of the section-8 "next" rows through the C ABI: fused PSNR/SSIM (f3) against the oracle restatement of
torchmetrics 1.4.0, crop + antialiased bicubic (f4) against the committed reference-dataloader fixture and the oracle."""
import math
import os

import numpy as np
import pytest
import torch

import aux_oracle as A

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_golden.npz")


def seeded(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)


# ragged tile edges (W-10, H-10 not multiples of 16), the minimum 11x11 image, the training geometry, a 720p-like plane
@pytest.mark.parametrize("shape", [(2, 3, 96, 96), (3, 3, 11, 11), (1, 3, 27, 43), (2, 1, 40, 26), (1, 3, 360, 640)])
def test_psnr_ssim_matches_oracle(shape):
    from fast_srgan_b200.metrics import ValidationMetrics
    hr = seeded(shape, 1) * 2 - 1
    sr = (hr + 0.2 * (seeded(shape, 2) - 0.5)).clamp(-1, 1)
    m = ValidationMetrics("cuda")
    m.update(sr.cuda(), hr.cuda())
    got = m.compute()
    ssim_ref = A.ssim_per_image((1 + sr) / 2, (1 + hr) / 2)
    _, psnr_ref = A.validation_metrics([sr], [hr])
    assert (got["ssim_per_image"].cpu().float() - ssim_ref).abs().max().item() <= 2e-5
    assert abs(got["ssim"] - float(ssim_ref.mean())) <= 2e-5
    assert abs(got["psnr"] - psnr_ref) <= 1e-4                               # dB


def test_metrics_accumulate_over_batches_and_reset():
    """trainer.py:60-68: PSNR pools the squared error over the whole loader, SSIM averages the images."""
    from fast_srgan_b200.metrics import ValidationMetrics
    batches = [(seeded((2, 3, 32, 48), 10 + i) * 2 - 1, seeded((2, 3, 32, 48), 20 + i) * 2 - 1) for i in range(3)]
    m = ValidationMetrics("cuda")
    for sr, hr in batches:
        m.update(sr.cuda(), hr.cuda())
    got = m.compute()
    ssim_ref, psnr_ref = A.validation_metrics([b[0] for b in batches], [b[1] for b in batches])
    assert abs(got["ssim"] - ssim_ref) <= 2e-5 and abs(got["psnr"] - psnr_ref) <= 1e-4
    m.reset()
    x = seeded((1, 3, 20, 20), 5).cuda()
    m.update(x, x, rescale=False)
    out = m.compute()
    assert abs(out["ssim"] - 1.0) <= 1e-6 and out["psnr"] == float("inf")
    # constant offset d on [0,1] images: PSNR = -20 log10 d exactly, whatever the tiling
    m.reset()
    y = seeded((2, 3, 50, 37), 6).cuda() * 0.5
    m.update(y + 0.125, y, rescale=False)
    assert abs(m.compute()["psnr"] - (-20 * math.log10(0.125))) <= 1e-5
    with pytest.raises(RuntimeError):
        m.update(torch.zeros(1, 3, 10, 40).cuda(), torch.zeros(1, 3, 10, 40).cuda())    # H < 11


def test_crop_resize_matches_reference_fixture():
    from fast_srgan_b200 import data
    aux = np.load(GOLD)
    imgs = [aux[f"img{i}"] for i in range(3)]
    cache = data.DeviceImageCache(imgs)
    lr, hr = data.crop_resize_batch(cache, torch.from_numpy(aux["samples"]), 24, 4)
    assert lr.shape == (12, 3, 24, 24) and hr.shape == (12, 3, 96, 96)
    assert np.abs(lr.cpu().numpy() - aux["lr"]).max() <= 5e-6                # [-1,1] units; fp32 summation order only
    for k, (idx, cy, cx) in enumerate(aux["samples"]):
        hr_ref = imgs[idx][:, cy:cy + 96, cx:cx + 96].astype(np.float32) / 127.5 - 1.0
        assert np.array_equal(hr[k].cpu().numpy(), hr_ref.astype(np.float32))          # bit-exact: one IEEE divide + subtract
    lr2, _ = data.crop_resize_batch(cache, torch.from_numpy(aux["samples_s2"]), 16, 2)
    assert np.abs(lr2.cpu().numpy() - aux["lr_s2"]).max() <= 5e-6


@pytest.mark.parametrize("lr_size,scale", [(24, 4), (32, 4), (25, 3), (48, 2), (8, 8)])
def test_crop_resize_geometries_match_oracle(lr_size, scale):
    from fast_srgan_b200 import data
    rs = np.random.RandomState(lr_size * 10 + scale)
    hr_size = lr_size * scale
    imgs = [rs.randint(0, 256, (3, hr_size + 17, hr_size + 5), dtype=np.uint8), rs.randint(0, 256, (3, hr_size, hr_size), dtype=np.uint8)]
    cache = data.DeviceImageCache(imgs)
    samples = [(0, 0, 0), (0, 17, 5), (0, 9, 2), (1, 0, 0), (0, 3, 4)]
    lr, hr = data.crop_resize_batch(cache, torch.tensor(samples, dtype=torch.int32), lr_size, scale)
    for k, (idx, cy, cx) in enumerate(samples):
        lr_ref, hr_ref = A.crop_and_downscale(imgs[idx], cy, cx, lr_size, scale)
        assert (lr[k].cpu() - lr_ref).abs().max().item() <= 5e-6
        assert torch.equal(hr[k].cpu(), hr_ref)


def test_gpu_crop_loader_shards_like_one_process():
    """W ranks at B/W draw the same images as one process at batch B (crop offsets are per-rank random, as the reference's
    per-worker `random` is); batches are fp32 NCHW in [-1,1] with LR = downscaled HR."""
    from fast_srgan_b200 import data
    rs = np.random.RandomState(0)
    imgs = [rs.randint(0, 256, (3, 100 + 7 * i, 120 + 3 * i), dtype=np.uint8) for i in range(5)]
    cache = data.DeviceImageCache(imgs)
    one = data.ShardedReplacementSampler(len(imgs), 64, 16, seed=3)
    loader = data.GpuCropLoader(cache, one, 24, 4, seed=3)
    assert len(loader) == 4
    n = 0
    for lr, hr in loader:
        assert lr.shape == (16, 3, 24, 24) and hr.shape == (16, 3, 96, 96) and lr.is_cuda
        assert hr.min().item() >= -1.0 and hr.max().item() <= 1.0
        # LR is the antialiased downscale of HR: the 4x4 box average of HR correlates strongly with it even on white
        # noise (0.88 measured; an unrelated crop would give ~0)
        box = torch.nn.functional.avg_pool2d(hr, 4)
        c = torch.corrcoef(torch.stack([box.flatten(), lr.flatten()]))[0, 1].item()
        assert c > 0.8
        n += 1
    assert n == 4
    idx_full = list(one)
    parts = [list(data.ShardedReplacementSampler(len(imgs), 64, 16, seed=3, rank=r, world=2)) for r in range(2)]
    for b in range(4):
        assert torch.equal(torch.cat([parts[0][b], parts[1][b]]), idx_full[b])


def test_trainer_validation_loop_matches_oracle():
    """Trainer.calculate_metrics_over_dataset (trainer.py:53-69) on a two-batch loader vs the oracle generator + metrics."""
    import types
    import srgan_oracle as O
    from fast_srgan_b200.metrics import ValidationMetrics
    from fast_srgan_b200.model import Generator
    sd = O.make_generator_state(64, 2, seed=1234)
    g = Generator(types.SimpleNamespace(n_filters=64, n_layers=2), compute_dtype=torch.float16)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    batches = [(seeded((2, 3, 12, 16), 30 + i) * 2 - 1, seeded((2, 3, 48, 64), 40 + i) * 2 - 1) for i in range(2)]
    m = ValidationMetrics("cuda")
    srs = []
    with torch.no_grad():
        for lr, hr in batches:
            m.update(g(lr.cuda()), hr.cuda())
            srs.append(O.generator_forward(sd, lr))
    got = m.compute()
    ssim_ref, psnr_ref = A.validation_metrics(srs, [b[1] for b in batches])
    # the generator itself is within 1e-3 of the oracle (fp16 operands); the metrics inherit that
    assert abs(got["ssim"] - ssim_ref) <= 2e-3 and abs(got["psnr"] - psnr_ref) <= 2e-2
