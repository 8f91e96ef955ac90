"""Host logic of the pixel-pair view for n_filters <= 32 (fast_srgan_b200/pairs.py): the expanded weights reproduce the
32-channel convolution / PixelShuffle exactly when the NHWC tensor is re-read as [N,H,W/2,64].  Pure torch on CPU."""
import torch
import torch.nn.functional as F

from fast_srgan_b200.pairs import _expand_up, _pad32, expand_pair


def to_pairs(x):                      # NCHW [N,C,H,W] -> pair grid NCHW [N,2C,H,W/2], slot (parity, c)
    N, C, H, W = x.shape
    return x.view(N, C, H, W // 2, 2).permute(0, 4, 1, 2, 3).reshape(N, 2 * C, H, W // 2)


def from_pairs(y, C):
    N, _, H, Wp = y.shape
    return y.view(N, 2, C, H, Wp).permute(0, 2, 3, 4, 1).reshape(N, C, H, 2 * Wp)


def test_expand_pair_is_the_same_convolution():
    g = torch.Generator().manual_seed(3)
    for co, ci, H, W in ((32, 32, 5, 8), (3, 32, 4, 2), (128, 32, 3, 6)):
        x = torch.randn((2, ci, H, W), generator=g, dtype=torch.float64)
        w = torch.randn((co, ci, 3, 3), generator=g, dtype=torch.float64)
        ref = F.conv2d(x, w, padding=1)
        got = from_pairs(F.conv2d(to_pairs(x), expand_pair(w), padding=1), co)
        assert torch.allclose(got, ref, atol=1e-12)
    assert (expand_pair(torch.ones(1, 1, 3, 3)) != 0).sum().item() == 2 * 9   # 6 of 12 (s, pi, po) blocks: half dense


def test_expand_up_matches_pixel_shuffle():
    """Kernel contract: GEMM column block q = 2i + j (64 values) goes to out[n, 2y+i, 2x'+j, :] of [N,2H,2W',64];
    read as 32-channel pixels that must be PixelShuffle(2)(conv(x) + b)  (model.py:39-40)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn((1, 32, 3, 4), generator=g, dtype=torch.float64)
    w = torch.randn((128, 32, 3, 3), generator=g, dtype=torch.float64)
    b = torch.randn(128, generator=g, dtype=torch.float64)
    ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=1), 2)                        # [1,32,6,8]
    w2, b2 = _expand_up(w, b)
    z = F.conv2d(to_pairs(x), w2, b2, padding=1)                                  # [1,256,3,2]
    N, _, H, Wp = z.shape
    out = torch.zeros((N, 2 * H, 2 * Wp, 64), dtype=torch.float64)
    for i in (0, 1):
        for j in (0, 1):
            out[:, i::2, j::2, :] = z[:, (2 * i + j) * 64:(2 * i + j + 1) * 64].permute(0, 2, 3, 1)
    got = out.reshape(N, 2 * H, 4 * Wp, 32).permute(0, 3, 1, 2)                   # pairs -> 32-channel pixels
    assert torch.allclose(got, ref, atol=1e-12)


def test_narrow_networks_pad_to_32_channels_exactly():
    """n_filters = 16: the upsampling conv [4F, F] padded to [128, 32] keeps reference channel 4c + q at its index, so the
    pixel-shuffled output is the true 16 channels followed by 16 exact zeros (pairs.PairGenerator._pack)."""
    g = torch.Generator().manual_seed(9)
    Fn = 16
    x = torch.randn((1, Fn, 4, 6), generator=g, dtype=torch.float64)
    w = torch.randn((4 * Fn, Fn, 3, 3), generator=g, dtype=torch.float64)
    b = torch.randn(4 * Fn, generator=g, dtype=torch.float64)
    ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=1), 2)                        # [1,16,8,12]
    w2, b2 = _expand_up(_pad32(w, 128, 32).double(), _pad32(b, 128).double())
    xp = torch.zeros((1, 32, 4, 6), dtype=torch.float64)
    xp[:, :Fn] = x
    z = F.conv2d(to_pairs(xp), w2, b2, padding=1)
    N, _, H, Wp = z.shape
    out = torch.zeros((N, 2 * H, 2 * Wp, 64), dtype=torch.float64)
    for i in (0, 1):
        for j in (0, 1):
            out[:, i::2, j::2, :] = z[:, (2 * i + j) * 64:(2 * i + j + 1) * 64].permute(0, 2, 3, 1)
    got = out.reshape(N, 2 * H, 4 * Wp, 32).permute(0, 3, 1, 2)
    assert torch.allclose(got[:, :Fn], ref, atol=1e-6) and got[:, Fn:].abs().max().item() == 0.0
