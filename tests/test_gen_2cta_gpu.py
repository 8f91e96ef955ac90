"""GPU parity of the CTA-pair general conv (csrc/conv3x3_gen_2cta.cuh: tcgen05 cta_group::2, M = 256, 128-wide slices)
against the single-CTA weight-stationary kernel (bit-identical: same tap / k / K-step order per tile; fixed-point
statistics are partition independent) and PyTorch fp32: forward stride 1 / 2, data gradient stride 1 / 2, both
epilogues, resident weights (Cin = 64), odd tile counts (clamped tail pair), short last groups."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
EPS = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}
DT = [torch.float16, torch.bfloat16]


@pytest.fixture(autouse=True)
def _setup():
    from fast_srgan_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    _lib.load().fsr_set_gen_2cta(-1)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def both(fn):
    from fast_srgan_b200 import _lib as L
    L.load().fsr_set_gen_2cta(1)
    a = fn()
    L.load().fsr_set_gen_2cta(0)
    b = fn()
    torch.cuda.synchronize()
    return a, b


FWD = [(64, 128, 24, 48, 48),      # resident weights (one K step), 432 tiles
       (128, 128, 64, 48, 48),     # VGG conv2_2 at the training shape
       (128, 512, 12, 40, 40),     # 4 slices, ragged tile edges
       (256, 256, 16, 24, 24),     # KC = 4: the weight double buffer wraps twice per group
       (512, 512, 128, 6, 6),      # one (mostly empty) tile per image
       (64, 128, 7, 37, 53),       # odd everything: odd tile count -> clamped tail pair
       (128, 256, 1, 16, 8)]       # a single tile: the pair's second CTA only has the clamped duplicate


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", FWD)
def test_gen_2cta_forward_stride1(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, cin, H, W), 1), dt)
    w = rnd((cout, cin, 3, 3), 2, (cin * 9) ** -0.5).to(dt).float()
    b = rnd((cout,), 3, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt)
    pair, single = both(lambda: ops.conv3x3_gen(x, wp, cout, bias=bp, act=L.ACT_RELU))
    assert torch.equal(pair, single)
    assert rel_err(nchw(pair), F.relu(F.conv2d(nchw(x), w, b, padding=1))) <= 2 * EPS[dt] + 1e-5
    (raw1, s1), (raw0, s0) = both(lambda: ops.conv3x3_gen(x, wp, cout, epilogue=L.EPI_RAW_STATS))
    assert torch.equal(raw1, raw0) and torch.equal(s1, s0)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", [(64, 128, 40, 96, 96), (128, 256, 24, 48, 48), (512, 512, 64, 12, 12), (256, 128, 3, 20, 12)])
def test_gen_2cta_stride2_forward_and_dgrad(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, cin, H, W), 4), dt)
    w = rnd((cout, cin, 3, 3), 5, (cin * 9) ** -0.5).to(dt).float()
    wp, _ = ops.pack_conv3x3(w, None, dt)
    xp = ops.parity_layout(x, True)
    (raw1, s1), (raw0, s0) = both(lambda: ops.conv3x3_gen(xp, wp, cout, stride=2, epilogue=L.EPI_RAW_STATS))
    assert torch.equal(raw1, raw0) and torch.equal(s1, s0)
    assert rel_err(nchw(raw1), F.conv2d(nchw(x), w, stride=2, padding=1)) <= 2 * EPS[dt] + 1e-5
    if cin % 128 == 0:                                   # the data gradient's GEMM columns are the forward input channels
        dy = nhwc(rnd((N, cout, H // 2, W // 2), 7), dt)
        wt = ops.pack_conv3x3_t(w, dt)
        d1, d0 = both(lambda: ops.conv3x3_gen(dy, wt, cin, stride=2, mode=1))
        assert torch.equal(d1, d0)
        xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
        F.conv2d(xr, w, stride=2, padding=1).backward(nchw(dy))
        assert rel_err(nchw(ops.parity_layout(d1, False)), xr.grad) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", [(128, 256, 16, 24, 40), (512, 512, 5, 12, 12), (256, 64, 8, 24, 24)])
def test_gen_2cta_dgrad_stride1(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops
    w = rnd((cout, cin, 3, 3), 6, (cout * 9) ** -0.5).to(dt).float()
    dy = nhwc(rnd((N, cout, H, W), 7), dt)
    wt = ops.pack_conv3x3_t(w, dt)
    d1, d0 = both(lambda: ops.conv3x3_gen(dy, wt, cin, stride=1, mode=1))
    assert torch.equal(d1, d0)
    xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
    F.conv2d(xr, w, padding=1).backward(nchw(dy))
    assert rel_err(nchw(d1), xr.grad) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("Fm,N,H,W", [(128, 8, 20, 24), (256, 2, 16, 8), (128, 3, 13, 21)])
def test_gen_2cta_pixel_shuffle_epilogue(dt, Fm, N, H, W):
    """UpSamplingBlock (model.py:39-40) for n_filters = 128 / 256: bias + PReLU + PixelShuffle through the pair kernel's
    5-D TMA store, against the single-CTA scatter epilogue (same bits) and PyTorch."""
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, Fm, H, W), 8), dt)
    wu = rnd((4 * Fm, Fm, 3, 3), 9, (Fm * 9) ** -0.5).to(dt).float()
    bu = rnd((4 * Fm,), 10, 0.1)
    alpha = torch.tensor([0.2], device="cuda")
    wp, bp = ops.pack_conv3x3(wu, bu, dt, ps_perm=True)
    u1, u0 = both(lambda: ops.conv3x3_gen(x, wp, 4 * Fm, epilogue=L.EPI_PS_PRELU, bias=bp, alpha=alpha))
    assert torch.equal(u1, u0)
    ref = F.prelu(F.pixel_shuffle(F.conv2d(nchw(x), wu, bu, padding=1), 2), alpha)
    assert rel_err(nchw(u1), ref) <= 2 * EPS[dt] + 1e-5
