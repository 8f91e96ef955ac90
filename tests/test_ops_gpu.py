"""GPU parity tests of each kernel, through the C ABI, against a plain PyTorch fp32 reference of the
same op evaluated on the same (fp16/bf16-rounded) inputs.  Tolerances are written per test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]
EPS = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}   # half-ulp relative rounding of the stored output


@pytest.fixture(autouse=True, params=[(1, 0), (1, 1), (0, 0)], ids=["halo1", "halo1-ws", "halo3"])
def _setup(request):
    """No TF32 in the PyTorch reference; every test runs in each A-operand staging mode of the conv and with the
    weight-stationary (tcgen05.mma.ws, collector re-use) MMA pairing on and off."""
    from fast_srgan_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _lib.load().fsr_set_halo_mode(request.param[0])
    _lib.load().fsr_set_ws_mode(request.param[1])
    yield
    _lib.load().fsr_set_halo_mode(1)
    _lib.load().fsr_set_ws_mode(-1)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def nhwc(x_nchw, dt):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


SHAPES = [(1, 8, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (2, 5, 7)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
def test_conv3x3_raw_stats(dt, shape):
    from fast_srgan_b200 import ops
    N, H, W = shape
    x = nhwc(rnd((N, 64, H, W), 1), dt)
    w = rnd((64, 64, 3, 3), 2, 0.05).to(dt).float()
    wp, _ = ops.pack_conv3x3(w, None, dt)
    raw, stats = ops.conv3x3_c64_raw_stats(x, wp)
    ref = F.conv2d(nchw_f32(x), w, padding=1)
    got = nchw_f32(raw)
    scale = ref.abs().max().item()
    # fp32 accumulation of exactly-representable products; only the output rounding differs
    assert (got - ref).abs().max().item() <= 2 * EPS[dt] * scale + 1e-5
    # InstanceNorm statistics are accumulated (fp32) over the STORED, rounded outputs: exact w.r.t. those
    # (up to fp32 summation order), and within rounding noise of the fp32 conv
    assert stats.dtype == torch.int64                      # fixed point: order-independent integer atomics
    stats_i = stats
    stats = ops.stats_to_float(stats).float()
    s_got = torch.stack([got.sum((2, 3)), (got * got).sum((2, 3))], dim=-1)
    assert torch.allclose(stats, s_got, rtol=1e-4, atol=1e-3)
    s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], dim=-1)
    assert torch.allclose(stats, s_ref, rtol=4 * EPS[dt], atol=4 * EPS[dt] * scale * (H * W) ** 0.5 + 1e-3)
    raw2, stats2 = ops.conv3x3_c64_raw_stats(x, wp)        # bitwise reproducible run to run
    assert torch.equal(raw2, raw) and torch.equal(stats2, stats_i)


@pytest.mark.parametrize("dt", DTYPES)
def test_conv3x3_tap_localisation(dt):
    """One non-zero tap at a time: catches descriptor / halo-offset mistakes per (r, s)."""
    from fast_srgan_b200 import ops
    N, H, W = 1, 16, 32
    x = nhwc(rnd((N, 64, H, W), 3), dt)
    for r in range(3):
        for s in range(3):
            w = torch.zeros(64, 64, 3, 3, device="cuda")
            w[:, :, r, s] = rnd((64, 64), 10 + r * 3 + s, 0.1).to(dt).float()
            wp, _ = ops.pack_conv3x3(w, None, dt)
            raw, _ = ops.conv3x3_c64_raw_stats(x, wp)
            ref = F.conv2d(nchw_f32(x), w, padding=1)
            err = (nchw_f32(raw) - ref).abs().max().item()
            assert err <= 2 * EPS[dt] * ref.abs().max().item() + 1e-5, (r, s, err)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 8, 16), (2, 11, 19)])
def test_conv3x3_pixelshuffle_prelu(dt, shape):
    from fast_srgan_b200 import ops
    N, H, W = shape
    x = nhwc(rnd((N, 64, H, W), 4), dt)
    w = rnd((256, 64, 3, 3), 5, 0.05).to(dt).float()
    b = rnd((256,), 6, 0.1)
    alpha = torch.tensor([0.2], device="cuda")
    wp, bp = ops.pack_conv3x3(w, b, dt, ps_perm=True)
    got = nchw_f32(ops.conv3x3_c64_ps_prelu(x, wp, bp, alpha))
    ref = F.prelu(F.pixel_shuffle(F.conv2d(nchw_f32(x), w, b, padding=1), 2), alpha)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 2 * EPS[dt] * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 8, 16), (2, 13, 27)])
def test_conv3x3_head_tanh(dt, shape):
    from fast_srgan_b200 import ops
    N, H, W = shape
    x = nhwc(rnd((N, 64, H, W), 7), dt)
    w = rnd((3, 64, 3, 3), 8, 0.05).to(dt).float()
    b = rnd((3,), 9, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt, cout_pad=16)
    ref = torch.tanh(F.conv2d(nchw_f32(x), w, b, padding=1))
    got = ops.conv3x3_c64_head(x, wp, bp)
    assert (got - ref).abs().max().item() <= 2e-5     # fp32 in, fp32 out: only accumulation order differs
    u8 = ops.conv3x3_c64_head(x, wp, bp, out_u8=True)
    ref8 = ((ref + 1) / 2 * 255).permute(0, 2, 3, 1)
    assert (u8.float() - ref8.floor()).abs().max().item() <= 1.0   # truncation; +-1 where fp32 noise crosses an integer
    assert ((u8.float() - ref8.floor()).abs() > 0).float().mean().item() < 1e-3


@pytest.mark.parametrize("dt", DTYPES)
def test_conv3x3_bias_relu(dt):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((2, 64, 9, 17), 11), dt)
    w = rnd((128, 64, 3, 3), 12, 0.05).to(dt).float()
    b = rnd((128,), 13, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt)
    got = nchw_f32(ops.conv3x3_c64_bias_act(x, wp, bp, act=L.ACT_RELU))
    ref = F.relu(F.conv2d(nchw_f32(x), w, b, padding=1))
    assert (got - ref).abs().max().item() <= 2 * EPS[dt] * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("u8", [False, True])
def test_neck_conv(dt, u8):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = 2, 11, 19
    w, b = rnd((64, 3, 3, 3), 14, 0.2), rnd((64,), 15, 0.1)
    alpha = torch.tensor([0.25], device="cuda")
    if u8:
        g = torch.Generator().manual_seed(16)
        img = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).cuda()
        xin, xref = img, (img.float() / 127.5 - 1.0).permute(0, 3, 1, 2)
    else:
        xin = xref = rnd((N, 3, H, W), 16)
    got = nchw_f32(ops.neck_conv3x3(xin, w, b, dt, act=L.ACT_PRELU, alpha=alpha))
    ref = F.prelu(F.conv2d(xref, w, b, padding=1), alpha)
    assert (got - ref).abs().max().item() <= 2 * EPS[dt] * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("mode", ["prelu", "none_res", "lrelu"])
def test_instnorm_apply(dt, mode):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W, C = 2, 13, 21, 64
    raw = nhwc(rnd((N, C, H, W), 17, 3.0) + 0.7, dt)
    rf = nchw_f32(raw)
    stats = ops.stats_from_float(rf.sum((2, 3)), (rf * rf).sum((2, 3)))
    alpha = torch.tensor([0.3], device="cuda")
    res = nhwc(rnd((N, C, H, W), 18), dt)
    ref = F.instance_norm(rf, eps=1e-5)
    if mode == "prelu":
        got = ops.instnorm_apply(raw, stats, act=L.ACT_PRELU, alpha=alpha)
        ref = F.prelu(ref, alpha)
    elif mode == "lrelu":
        got = ops.instnorm_apply(raw, stats, act=L.ACT_LRELU, slope=0.01)
        ref = F.leaky_relu(ref, 0.01)
    else:
        got = ops.instnorm_apply(raw, stats, residual=res)
        ref = ref + nchw_f32(res)
    assert (nchw_f32(got) - ref).abs().max().item() <= 2 * EPS[dt] * ref.abs().max().item() + 1e-4


@pytest.mark.parametrize("dt", DTYPES)
def test_pixel_shuffle_standalone(dt):
    from fast_srgan_b200 import ops
    x = rnd((2, 256, 7, 9), 19)
    got = ops.pixel_shuffle2(nhwc(x, dt))
    ref = F.pixel_shuffle(x.to(dt).float(), 2)
    assert torch.equal(nchw_f32(got), ref)     # pure data movement: bit exact


@pytest.mark.parametrize("dt", DTYPES)
def test_layout_roundtrip(dt):
    from fast_srgan_b200 import ops
    x = rnd((2, 67, 5, 9), 20)
    y = ops.nchw_to_nhwc(x, dt)
    assert torch.equal(y, x.permute(0, 2, 3, 1).to(dt))
    assert torch.equal(ops.nhwc_to_nchw(y), x.to(dt).float())
