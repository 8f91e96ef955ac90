"""GPU parity of the res-block conv with the first InstanceNorm + PReLU fused into its load path (conv3x3_c64_kernel XF
variant, fsr_conv3x3_c64_in): bit-identical to normalise-then-convolve (the in-smem transform uses the same fp32
operations as instnorm_apply), and Generator.forward is bit-identical with the fusion on and off."""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]


@pytest.fixture(autouse=True)
def _setup():
    from fast_srgan_b200 import _lib
    yield
    _lib.load().fsr_set_fuse_in(-1)
    _lib.load().fsr_set_ws_mode(-1)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


# single tile, ragged edges, image smaller than a tile, several images per CTA range, full-size frames (97 tiles per CTA)
SHAPES = [(1, 8, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (2, 5, 7), (5, 1, 1), (4, 180, 320)]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ws", [1, 0])
def test_fused_input_conv_bit_identical(dt, shape, ws):
    from fast_srgan_b200 import ops, _lib as L
    L.load().fsr_set_ws_mode(ws)
    N, H, W = shape
    x = (rnd((N, 64, H, W), 1) * 1.7 + 0.3).permute(0, 2, 3, 1).contiguous().to(dt)
    w1 = rnd((64, 64, 3, 3), 2, 0.05)
    w2 = rnd((64, 64, 3, 3), 3, 0.05)
    alpha = torch.tensor([0.2], device="cuda")
    wp1, _ = ops.pack_conv3x3(w1, None, dt)
    wp2, _ = ops.pack_conv3x3(w2, None, dt)
    raw1, st1 = ops.conv3x3_c64_raw_stats(x, wp1)
    y1 = ops.instnorm_apply(raw1, st1, act=L.ACT_PRELU, alpha=alpha)
    ref, st_ref = ops.conv3x3_c64_raw_stats(y1, wp2)
    got, st_got = ops.conv3x3_c64_in(raw1, st1, alpha, wp2)
    assert torch.equal(got, ref)
    assert torch.equal(st_got, st_ref)
    # and against plain PyTorch fp32 on the same rounded operands (the reference's bn1 + relu1 + conv2, model.py:55-64)
    y1f = y1.float().permute(0, 3, 1, 2)
    refc = F.conv2d(y1f, w2.to(dt).float(), padding=1)
    eps = 2.0 ** -11 if dt == torch.float16 else 2.0 ** -8
    assert (got.float().permute(0, 3, 1, 2) - refc).abs().max().item() <= 2 * eps * refc.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 3, 90, 160), (3, 3, 17, 9)])
def test_generator_identical_with_and_without_fusion(dt, shape):
    import srgan_oracle as O
    from fast_srgan_b200 import _lib as L
    from fast_srgan_b200.model import Generator
    sd = O.make_generator_state(64, 8, seed=1234)
    g = Generator(types.SimpleNamespace(n_filters=64, n_layers=8), compute_dtype=dt)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand(shape, generator=gen) * 2 - 1).cuda()
    with torch.no_grad():
        L.load().fsr_set_fuse_in(1)
        a = g(x).clone()
        L.load().fsr_set_fuse_in(0)
        b = g(x).clone()
    assert torch.equal(a, b)
