"""CPU tests of the section-8 "next" rows (f3 metrics, f4 data path): the oracle restatement against the committed
reference fixture (tests/golden/aux_golden.npz, written by oracle/make_aux_golden.py from the unmodified reference
dataloader), known answers for the metrics, and the product's host logic (tap table, sharded sampler)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import aux_oracle as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_golden.npz")


@pytest.fixture(scope="module")
def aux():
    return np.load(GOLD)


def test_crop_downscale_matches_reference_fixture(aux):
    imgs = [aux[f"img{i}"] for i in range(3)]
    for k, (idx, cy, cx) in enumerate(aux["samples"]):
        for use_torch in (True, False):
            lr, hr = A.crop_and_downscale(imgs[idx], int(cy), int(cx), 24, 4, use_torch=use_torch)
            assert np.abs(lr.numpy() - aux["lr"][k]).max() <= 2e-6           # fp32 summation order only
            assert abs(float(np.abs(hr.numpy()).sum()) - aux["hr_checksum"][k]) <= 1e-2
    idx, cy, cx = aux["samples_s2"][0]
    lr, _ = A.crop_and_downscale(imgs[idx], int(cy), int(cx), 16, 2, use_torch=False)
    assert np.abs(lr.numpy() - aux["lr_s2"][0]).max() <= 2e-6


@pytest.mark.parametrize("geom", [(96, 24), (64, 16), (100, 25), (90, 30), (97, 24), (33, 11), (24, 96)])
def test_tap_restatement_matches_torch_interpolate(geom):
    n_in, n_out = geom
    g = torch.Generator().manual_seed(n_in)
    x = torch.rand((2, 3, n_in, n_in), generator=g) * 255
    ref = F.interpolate(x, size=(n_out, n_out), mode="bicubic", antialias=True, align_corners=False)
    assert (A.resize_aa_bicubic(x, n_out, n_out) - ref).abs().max().item() <= 2e-4      # of 255


def test_product_tap_table_equals_oracle():
    from fast_srgan_b200 import data
    for n_in, n_out in [(96, 24), (64, 16), (100, 25), (96, 48), (90, 30), (97, 24), (24, 96), (128, 32)]:
        o, p = A.aa_bicubic_taps(n_in, n_out), data.aa_bicubic_tap_table(n_in, n_out)
        assert np.array_equal(o[0], p[0]) and np.array_equal(o[1], p[1])
        assert np.abs(o[2] - p[2]).max() <= 6e-8 and abs(float(p[2].sum(1).max()) - 1.0) <= 1e-6


def test_sampler_stream_matches_reference_fixture_and_shards(aux):
    from fast_srgan_b200.data import ShardedReplacementSampler
    ref = torch.from_numpy(aux["sampler_800_173_seed1234"])
    assert torch.equal(A.replacement_sample_indices(800, 173, 1234), ref)
    full = list(ShardedReplacementSampler(800, 173, 24, 1234))
    assert len(full) == 7 and torch.equal(torch.cat(full), ref[:168])                  # drop_last
    for world in (2, 4, 8):
        shards = [list(ShardedReplacementSampler(800, 173, 24, 1234, r, world)) for r in range(world)]
        for b in range(7):
            assert torch.equal(torch.cat([shards[r][b] for r in range(world)]), full[b])   # W ranks == one process
    with pytest.raises(ValueError):
        ShardedReplacementSampler(800, 173, 24, 1234, 0, 5)


def test_metric_known_answers():
    g = torch.Generator().manual_seed(3)
    a = torch.rand((2, 3, 40, 52), generator=g)
    assert torch.allclose(A.ssim_per_image(a, a), torch.ones(2), atol=1e-6)
    assert abs(A.psnr_from_sse(0.1 ** 2 * 1000, 1000) - 20.0) <= 1e-9               # constant offset 0.1 -> 20 dB
    # the gaussian window against scipy's separable filter with mirror boundaries (independent implementation)
    from scipy.ndimage import correlate1d
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    w = A.gaussian_taps().double().numpy()
    f = lambda z: correlate1d(correlate1d(z.double().numpy(), w, axis=-1, mode="mirror"), w, axis=-2, mode="mirror")
    mx, my = f(a), f(b)
    sxx, syy, sxy = f(a * a) - mx * mx, f(b * b) - my * my, f(a * b) - mx * my
    m = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sxx + syy + 9e-4))
    ref = m[..., 5:-5, 5:-5].reshape(2, -1).mean(-1)
    assert np.abs(A.ssim_per_image(a, b).numpy() - ref).max() <= 1e-5
    # validation_metrics pools the squared error over batches and averages SSIM over images (trainer.py:60-68)
    sr = [a * 2 - 1, b * 2 - 1]
    hr = [b * 2 - 1, b * 2 - 1]
    ssim, psnr = A.validation_metrics(sr, hr)
    mse = float(((a - b).double() ** 2).sum()) / (2 * a.numel())
    assert abs(psnr - 10 * math.log10(1 / mse)) <= 1e-4
    assert abs(ssim - float((A.ssim_per_image(a, b).sum() + 2.0) / 4)) <= 1e-6


def test_metrics_and_data_need_cuda():
    from fast_srgan_b200 import data, metrics
    with pytest.raises(RuntimeError, match="no CPU"):
        metrics.ValidationMetrics("cpu")
    with pytest.raises(RuntimeError, match="no CPU"):
        data.DeviceImageCache([np.zeros((3, 8, 8), np.uint8)], device="cpu")


def test_crop_offsets_are_in_bounds_and_seeded_per_rank():
    """GpuCropLoader.draw (dataloader.py:27-29 for a batch): offsets keep the HR crop inside every image, are reproducible
    for a (seed, rank) and differ between ranks.  Host logic only: a stub stands in for the device-resident cache."""
    import types
    from fast_srgan_b200.data import GpuCropLoader, ShardedReplacementSampler
    shapes = [(96, 96), (100, 131), (240, 97), (96, 300)]
    stub = types.SimpleNamespace(shapes=shapes, device=torch.device("cpu"))
    pin = torch.Tensor.pin_memory
    torch.Tensor.pin_memory = lambda self, *a, **k: self            # no CUDA in this container
    try:
        def loader(rank):
            return GpuCropLoader(stub, ShardedReplacementSampler(len(shapes), 64, 8, seed=1, rank=rank, world=2), 24, 4, seed=9)
        idx = torch.arange(32) % len(shapes)
        a, b, c = loader(0).draw(idx), loader(0).draw(idx), loader(1).draw(idx)
    finally:
        torch.Tensor.pin_memory = pin
    assert a.dtype == torch.int32 and tuple(a.shape) == (32, 3) and torch.equal(a, b) and not torch.equal(a, c)
    for i, cy, cx in a.tolist():
        h, w = shapes[i]
        assert 0 <= cy <= h - 96 and 0 <= cx <= w - 96
    assert all(int(a[k, 1]) == 0 and int(a[k, 2]) == 0 for k in range(32) if shapes[int(a[k, 0])] == (96, 96))
    with pytest.raises(ValueError):
        GpuCropLoader(types.SimpleNamespace(shapes=[(95, 200)], device=torch.device("cpu")),
                      ShardedReplacementSampler(1, 8, 8, seed=1), 24, 4)
