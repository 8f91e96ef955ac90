"""The padded-flattened tile mapping of the small VGG19 layers (fsr_conv3x3_gen_flat + the padded max-pool kernels):
an M tile is 128 consecutive positions of the flattened (image, y', x') index of zero-bordered padded tensors, so tiles
straddle images.  Checked against PyTorch fp32 (forward, data gradient, max-pool forward / backward between the plain
and the padded layout) and, end to end, against the plain-layout VGG path (FSR_VGG_FLAT=0)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
EPS = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}
DT = [torch.float16, torch.bfloat16]


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def pad_nhwc(x):
    return F.pad(x, (0, 0, 1, 1, 1, 1)).contiguous()


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,N,H,W", [(256, 512, 128, 12, 12), (512, 512, 64, 6, 6), (128, 256, 3, 8, 8), (64, 128, 5, 2, 2),
                                            (512, 128, 7, 12, 25), (128, 128, 1, 1, 1)])
def test_flat_conv_forward_and_dgrad(dt, cin, cout, N, H, W):
    from fast_srgan_b200 import ops, _lib as L
    x = nhwc(rnd((N, cin, H, W), 1), dt)
    w = rnd((cout, cin, 3, 3), 2, (cin * 9) ** -0.5).to(dt).float()
    b = rnd((cout,), 3, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt)
    y = ops.conv3x3_gen_flat(pad_nhwc(x), wp, cout, mode=0, bias=bp, act=L.ACT_RELU)
    ref = F.relu(F.conv2d(nchw(x), w, b, padding=1))
    assert y.shape == (N, H + 2, W + 2, cout)
    assert rel_err(nchw(y[:, 1:-1, 1:-1, :]), ref) <= 2 * EPS[dt] + 1e-5
    border = y.clone()
    border[:, 1:-1, 1:-1, :] = 0
    assert border.abs().max().item() == 0.0                       # the zero border survives (next layer's padding)
    plain = ops.conv3x3_gen(x, wp, cout, bias=bp, act=L.ACT_RELU)  # same MMA order per output pixel: same bits
    assert torch.equal(y[:, 1:-1, 1:-1, :], plain)
    if cin % 128 == 0:
        dy = nhwc(rnd((N, cout, H, W), 7), dt)
        wt = ops.pack_conv3x3_t(w, dt)
        dx = ops.conv3x3_gen_flat(pad_nhwc(dy), wt, cin, mode=1)
        xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
        F.conv2d(xr, w, padding=1).backward(nchw(dy))
        assert rel_err(nchw(dx[:, 1:-1, 1:-1, :]), xr.grad) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("in_pad,out_pad", [(False, True), (True, True), (True, False)])
def test_padded_maxpool_forward_backward(dt, in_pad, out_pad):
    from fast_srgan_b200 import ops
    N, C, H, W = 3, 128, 12, 8
    x = F.relu(rnd((N, C, H, W), 5)).to(dt)
    xin = pad_nhwc(nhwc(x, dt)) if in_pad else nhwc(x, dt)
    y = ops.maxpool2_padded(xin, in_pad, out_pad)
    ref = F.max_pool2d(x.float(), 2)
    got = y[:, 1:-1, 1:-1, :] if out_pad else y
    assert torch.equal(nchw(got), ref)
    g = rnd((N, C, H // 2, W // 2), 6).to(dt)
    gin = pad_nhwc(nhwc(g, dt)) if out_pad else nhwc(g, dt)
    dx = ops.maxpool2_relu_bwd_padded(xin, gin, in_pad, out_pad)
    plain = ops.maxpool2_relu_bwd(nhwc(x, dt), nhwc(g, dt))
    assert torch.equal(dx[:, 1:-1, 1:-1, :] if in_pad else dx, plain)
    if in_pad:
        b = dx.clone()
        b[:, 1:-1, 1:-1, :] = 0
        assert b.abs().max().item() == 0.0


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(4, 3, 96, 96), (2, 3, 32, 32)])
def test_vgg_flat_path_equals_plain_path(dt, shape):
    """VGG19[:34] forward features and image gradient with the flat mapping on (default) and off: identical bits."""
    import types
    import srgan_oracle as O
    from fast_srgan_b200.model import VGG19
    v = VGG19(compute_dtype=dt)
    v.load_state_dict(O.make_vgg19_state(99))
    v = v.cuda().eval()
    net = v._engine()
    x = (torch.rand(shape, generator=torch.Generator().manual_seed(4)) * 2 - 1).cuda()
    res = {}
    for flat in ("1", "0"):
        os.environ["FSR_VGG_FLAT"] = flat
        try:
            feat, ctx = net.forward(x, save=True)
            pad = net.feat_pad
            nb = shape[0] // 2
            g = torch.Generator().manual_seed(9)
            d = (torch.randn((nb,) + tuple(feat.shape[1:] if not pad else (feat.shape[1] - 2, feat.shape[2] - 2, feat.shape[3])), generator=g) * 0.01).to(dt).cuda()
            dfeat = F.pad(d, (0, 0, 1, 1, 1, 1)).contiguous() if pad else d
            dimg = torch.zeros((nb,) + tuple(shape[1:]), device="cuda")
            net.backward(ctx, dfeat, dimg)
            torch.cuda.synchronize()
            res[flat] = ((feat[:, 1:-1, 1:-1, :] if pad else feat).clone(), dimg.clone(), pad)
        finally:
            os.environ.pop("FSR_VGG_FLAT", None)
    assert res["1"][2] and not res["0"][2]
    assert torch.equal(res["1"][0], res["0"][0])
    assert torch.equal(res["1"][1], res["0"][1])
