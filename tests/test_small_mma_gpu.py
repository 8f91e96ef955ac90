"""GPU parity of the two 3-channel-sided convs on warp-level tensor-core MMAs (csrc/small_mma.cuh) through the C ABI:
against PyTorch fp32 (conv / autograd) and against the CUDA-core kernels they replace (fsr_set_small_mma(0))."""
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
    _lib.load().fsr_set_small_mma(-1)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


# ragged widths (strip tail masked), single pixel, several strips per row, more strips than resident warps
NECK_SHAPES = [(1, 1, 1), (2, 5, 7), (3, 24, 24), (2, 33, 50), (4, 96, 96), (1, 180, 320)]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", NECK_SHAPES)
@pytest.mark.parametrize("mode", ["prelu", "lrelu_c128", "vgg_relu", "u8", "nobias_none"])
def test_neck_mma(dt, shape, mode):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = shape
    cout = 128 if mode == "lrelu_c128" else 64
    w, b = rnd((cout, 3, 3, 3), 14, 0.2), rnd((cout,), 15, 0.1)
    alpha = torch.tensor([0.25], device="cuda")
    if mode == "u8":
        g = torch.Generator().manual_seed(16)
        img = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).cuda()
        xin, xref = img, (img.float() / 127.5 - 1.0).permute(0, 3, 1, 2)
    else:
        xin = xref = rnd((N, 3, H, W), 16).clamp(-1, 1)
    kw = dict(act=L.ACT_PRELU, alpha=alpha)
    ref_fn = lambda z: F.prelu(z, alpha)
    if mode == "lrelu_c128":
        kw, ref_fn = dict(act=L.ACT_LRELU, slope=0.2), (lambda z: F.leaky_relu(z, 0.2))
    elif mode == "vgg_relu":
        kw, ref_fn = dict(act=L.ACT_RELU, vgg_norm=True), F.relu
        mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
        xref = ((xref + 1) / 2 - mean) / std
    elif mode == "nobias_none":
        kw, ref_fn, b = dict(act=L.ACT_NONE), (lambda z: z), None
    L.load().fsr_set_small_mma(1)
    got = nchw(ops.neck_conv3x3(xin, w, b, dt, **kw))
    L.load().fsr_set_small_mma(0)
    old = nchw(ops.neck_conv3x3(xin, w, b, dt, **kw))
    ref = ref_fn(F.conv2d(xref, w, b, padding=1))
    tol = 2 * EPS[dt] * ref.abs().max().item() + 1e-5      # output rounding only: the hi/lo split keeps fp32-input accuracy
    assert (got - ref).abs().max().item() <= tol
    assert (got - old).abs().max().item() <= tol


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(2, 6, 8), (3, 24, 24), (5, 13, 9), (1, 1, 1), (8, 96, 96)])
@pytest.mark.parametrize("C", [64, 128])
def test_wgrad_c3_mma(dt, shape, C):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = shape
    x3 = rnd((N, 3, H, W), 27)
    d64 = rnd((N, C, H, W), 28).to(dt).float()
    # conv input side (neck): dW[c,c3,r,s] = sum x3[., c3, y+r-1, x+s-1] * dOut[., c, y, x]
    wr = torch.zeros((C, 3, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x3, wr, padding=1).backward(d64)
    # conv output side (head): dW[c3,c,r,s] = sum dpre[., c3, y, x] * x64[., c, y+r-1, x+s-1]
    wh = torch.zeros((3, C, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(d64, wh, padding=1).backward(x3)
    for mma in (1, 0):
        L.load().fsr_set_small_mma(mma)
        tol = 1e-4 if mma else 2e-4
        out0 = torch.zeros((27, C), device="cuda")
        ops.wgrad_c3(x3, nhwc(d64, dt), out0, flip=False, layout=0)
        assert rel_err(out0.view(3, 9, C).permute(2, 0, 1).reshape(C, 3, 3, 3), wr.grad) <= tol, mma
        out2 = torch.zeros((C, 3, 3, 3), device="cuda")
        ops.wgrad_c3(x3, nhwc(d64, dt), out2, flip=False, layout=2)
        assert rel_err(out2, wr.grad) <= tol, mma
        out1 = torch.zeros((3, C, 3, 3), device="cuda")
        ops.wgrad_c3(x3, nhwc(d64, dt), out1, flip=True, layout=1)
        assert rel_err(out1, wh.grad) <= tol, mma
        # accumulation semantics: a second call adds
        ops.wgrad_c3(x3, nhwc(d64, dt), out1, flip=True, layout=1)
        assert rel_err(out1, 2 * wh.grad) <= tol, mma
