"""GPU parity of the weight-stationary general conv (csrc/conv3x3_gen_ws.cuh) at shapes where every CTA owns several
4-tile groups (full groups, a short last group, both TMEM accumulator sets re-used, the weight double buffer wrapping,
resident weights when Cin = 64): bit-identical to the per-tile streaming kernel (same MMA order per tile) and within
output rounding of PyTorch fp32."""
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
    _lib.load().fsr_set_gen_2cta(0)          # this file compares the two single-CTA kernels (the CTA-pair kernel: test_gen_2cta_gpu.py)
    yield
    _lib.load().fsr_set_gen_ws(-1)
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
    L.load().fsr_set_gen_ws(1)
    a = fn()
    L.load().fsr_set_gen_ws(0)
    b = fn()
    return a, b


# (cin, cout, N, H, W): tiles per CTA = N*ceil(H/16)*ceil(W/8) / (148 // (cout/64))
FWD = [(64, 64, 24, 96, 96),      # resident weights, ~11.7 tiles / CTA
       (128, 512, 12, 40, 40),    # KC = 2, 18 CTAs per slice, 10 tiles / CTA: groups 4 + 4 + 2, ragged tile edges
       (256, 256, 16, 24, 24),    # KC = 4 (weight double buffer wraps twice per group), 2.6 tiles / CTA
       (64, 128, 7, 37, 53)]      # odd everything, some CTAs with 3 and some with 4 tiles


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", FWD)
def test_gen_ws_forward_stride1(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, cin, H, W), 1), dt)
    w = rnd((cout, cin, 3, 3), 2, (cin * 9) ** -0.5).to(dt).float()
    b = rnd((cout,), 3, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt)
    ws, st = both(lambda: ops.conv3x3_gen(x, wp, cout, bias=bp, act=L.ACT_RELU))
    assert torch.equal(ws, st)
    assert rel_err(nchw(ws), F.relu(F.conv2d(nchw(x), w, b, padding=1))) <= 2 * EPS[dt] + 1e-5
    (raw1, s1), (raw0, s0) = both(lambda: ops.conv3x3_gen(x, wp, cout, epilogue=L.EPI_RAW_STATS))
    assert torch.equal(raw1, raw0) and torch.equal(s1, s0)          # fixed-point statistics: partition independent


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", [(64, 128, 40, 96, 96), (128, 256, 24, 48, 48), (512, 512, 64, 12, 12)])
def test_gen_ws_stride2_forward_and_dgrad(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, cin, H, W), 4), dt)
    w = rnd((cout, cin, 3, 3), 5, (cin * 9) ** -0.5).to(dt).float()
    wp, _ = ops.pack_conv3x3(w, None, dt)
    xp = ops.parity_layout(x, True)
    (raw1, s1), (raw0, s0) = both(lambda: ops.conv3x3_gen(xp, wp, cout, stride=2, epilogue=L.EPI_RAW_STATS))
    assert torch.equal(raw1, raw0) and torch.equal(s1, s0)
    assert rel_err(nchw(raw1), F.conv2d(nchw(x), w, stride=2, padding=1)) <= 2 * EPS[dt] + 1e-5
    dy = nhwc(rnd((N, cout, H // 2, W // 2), 7), dt)
    wt = ops.pack_conv3x3_t(w, dt)
    d1, d0 = both(lambda: ops.conv3x3_gen(dy, wt, cin, stride=2, mode=1))
    assert torch.equal(d1, d0)
    xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
    F.conv2d(xr, w, stride=2, padding=1).backward(nchw(dy))
    assert rel_err(nchw(ops.parity_layout(d1, False)), xr.grad) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
def test_gen_ws_dgrad_stride1_and_pixel_shuffle(dt):
    from fast_srgan_b200 import ops, _lib as L
    cin, cout, N, H, W = 128, 256, 16, 24, 40
    w = rnd((cout, cin, 3, 3), 6, (cout * 9) ** -0.5).to(dt).float()
    dy = nhwc(rnd((N, cout, H, W), 7), dt)
    wt = ops.pack_conv3x3_t(w, dt)
    d1, d0 = both(lambda: ops.conv3x3_gen(dy, wt, cin, stride=1, mode=1))
    assert torch.equal(d1, d0)
    xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
    F.conv2d(xr, w, padding=1).backward(nchw(dy))
    assert rel_err(nchw(d1), xr.grad) <= 2 * EPS[dt] + 1e-5
    # F = 128 upsampling block: bias + PReLU + PixelShuffle scatter epilogue
    Fm = 128
    x = nhwc(rnd((8, Fm, 20, 24), 8), dt)
    wu = rnd((4 * Fm, Fm, 3, 3), 9, (Fm * 9) ** -0.5).to(dt).float()
    bu = rnd((4 * Fm,), 10, 0.1)
    alpha = torch.tensor([0.2], device="cuda")
    wp, bp = ops.pack_conv3x3(wu, bu, dt, ps_perm=True)
    u1, u0 = both(lambda: ops.conv3x3_gen(x, wp, 4 * Fm, epilogue=L.EPI_PS_PRELU, bias=bp, alpha=alpha))
    assert torch.equal(u1, u0)
    ref = F.prelu(F.pixel_shuffle(F.conv2d(nchw(x), wu, bu, padding=1), 2), alpha)
    assert rel_err(nchw(u1), ref) <= 2 * EPS[dt] + 1e-5
