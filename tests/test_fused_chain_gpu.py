"""GPU tests of the round-2 inference kernels (first run on hardware in round 2: all bit-identical at first try):
  conv3x3_up_2cta_kernel          the 64->256 upsampling conv as a tcgen05 CTA-pair (cta_group::2) kernel
  conv3x3_c64_kernel<.., XF = 2>  conv1 / bottleneck with the previous block's bn2 + skip fused into the load path
and of the fully fused residual chain of fsr_generator_forward built from them (model.py:67-69, :86-95, :115)."""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.fixture
def lib():
    from fast_srgan_b200 import _lib as L
    lib = L.load()
    yield lib
    lib.fsr_set_up_2cta(-1)
    lib.fsr_set_fuse_res(-1)
    lib.fsr_set_fuse_in(-1)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 16, 8), (1, 16, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (4, 90, 160), (1, 1, 1), (5, 7, 3)])
def test_up_conv_cta_pair_matches_single_cta(lib, dt, shape):
    """The pair kernel against the single-CTA kernel (bit-identical: same tap / k order per tile) and PyTorch fp32 on the
    same rounded operands (model.py:30-40).  Odd tile counts exercise the clamped tail pair."""
    from fast_srgan_b200 import ops
    N, H, W = shape
    x = rnd((N, 64, H, W), 1).permute(0, 2, 3, 1).contiguous().to(dt)
    w = rnd((256, 64, 3, 3), 2, 0.05).to(dt).float()
    b = rnd((256,), 3, 0.1)
    alpha = torch.tensor([0.2], device="cuda")
    wp, bp = ops.pack_conv3x3(w, b, dt, ps_perm=True)
    lib.fsr_set_up_2cta(0)
    base = ops.conv3x3_c64_ps_prelu(x, wp, bp, alpha)
    lib.fsr_set_up_2cta(1)
    got = ops.conv3x3_c64_ps_prelu(x, wp, bp, alpha)
    torch.cuda.synchronize()
    assert torch.equal(got, base)
    ref = F.prelu(F.pixel_shuffle(F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1), 2), alpha)
    eps = 2.0 ** -11 if dt == torch.float16 else 2.0 ** -8
    assert (got.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 2 * eps * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 8, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (2, 5, 7), (4, 180, 320)])
def test_conv_with_fused_norm_and_skip_matches_unfused(dt, shape):
    """fsr_conv3x3_c64_res_in (bn2 + skip of block l folded into conv1 of block l+1, model.py:65+69 -> :47-54): x_next, the
    conv output and its fixed-point statistics are bit-identical to instnorm_apply(+residual) followed by the plain conv."""
    from fast_srgan_b200 import ops
    N, H, W = shape
    x_prev = (rnd((N, 64, H, W), 1) * 1.3).permute(0, 2, 3, 1).contiguous().to(dt)
    w0 = rnd((64, 64, 3, 3), 2, 0.05)
    w1 = rnd((64, 64, 3, 3), 3, 0.05)
    wp0, _ = ops.pack_conv3x3(w0, None, dt)
    wp1, _ = ops.pack_conv3x3(w1, None, dt)
    raw2, st2 = ops.conv3x3_c64_raw_stats(x_prev, wp0)                     # stands in for c2 of the previous block
    x_next_ref = ops.instnorm_apply(raw2, st2, residual=x_prev)
    out_ref, st_ref = ops.conv3x3_c64_raw_stats(x_next_ref, wp1)
    x_next, out, stats = ops.conv3x3_c64_res_in(raw2, st2, x_prev, wp1)
    torch.cuda.synchronize()
    assert torch.equal(x_next, x_next_ref)
    assert torch.equal(out, out_ref)
    assert torch.equal(stats, st_ref)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L,shape", [(1, (2, 16, 24)), (2, (1, 37, 53)), (3, (3, 24, 24)), (8, (2, 45, 80)), (0, (1, 9, 9))])
def test_generator_forward_fused_chain_is_bit_identical(lib, dt, L, shape):
    """Generator.forward (model.py:112-117) with every switch of the fused path on (default) vs all off: same bits.
    Odd and even block counts exercise both ping-pong parities of the residual-chain buffers."""
    import srgan_oracle as O
    from fast_srgan_b200.model import Generator
    g = Generator(types.SimpleNamespace(n_filters=64, n_layers=L), compute_dtype=dt)
    g.load_state_dict(O.make_generator_state(64, L, seed=5))
    g = g.cuda().eval()
    gen = torch.Generator().manual_seed(11)
    x = (torch.rand((shape[0], 3, shape[1], shape[2]), generator=gen) * 2 - 1).cuda()
    outs = []
    for fuse_res, up2, fuse_in in ((1, 1, 1), (0, 0, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0)):
        lib.fsr_set_fuse_res(fuse_res)
        lib.fsr_set_up_2cta(up2)
        lib.fsr_set_fuse_in(fuse_in)
        with torch.no_grad():
            outs.append(g(x).clone())
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
