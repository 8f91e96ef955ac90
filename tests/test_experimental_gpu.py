"""Opt-in GPU tests of kernels that are compiled (libfsr_b200_experimental.so, its own library) but have NOT run on hardware
yet; skipped unless FSR_TEST_EXPERIMENTAL=1.  The product never loads that library.
  conv3x3_up_2cta_kernel: the 64->256 upsampling conv as a tcgen05 CTA-pair (cta_group::2) kernel."""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FSR_TEST_EXPERIMENTAL") != "1", reason="experimental kernels: opt in with FSR_TEST_EXPERIMENTAL=1")]


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 16, 8), (1, 16, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (4, 90, 160)])
def test_up_conv_cta_pair_matches_single_cta(dt, shape):
    """The pair kernel against the validated single-CTA kernel (expected bit-identical: same tap / k order per tile) and
    PyTorch fp32 on the same rounded operands (model.py:30-40)."""
    from fast_srgan_b200 import ops, _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xlib = ctypes.CDLL(os.path.join(root, "fast-srgan_b200", "libfsr_b200_experimental.so"))
    xlib.fsrx_conv3x3_up_2cta.restype = ctypes.c_int
    xlib.fsrx_conv3x3_up_2cta.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    N, H, W = shape
    x = rnd((N, 64, H, W), 1).permute(0, 2, 3, 1).contiguous().to(dt)
    w = rnd((256, 64, 3, 3), 2, 0.05).to(dt).float()
    b = rnd((256,), 3, 0.1)
    alpha = torch.tensor([0.2], device="cuda")
    wp, bp = ops.pack_conv3x3(w, b, dt, ps_perm=True)
    base = ops.conv3x3_c64_ps_prelu(x, wp, bp, alpha)
    got = torch.empty_like(base)
    rc = xlib.fsrx_conv3x3_up_2cta(x.data_ptr(), wp.data_ptr(), got.data_ptr(), bp.data_ptr(), alpha.data_ptr(), N, H, W,
                                   L.dtype_code(dt), L.stream_ptr(x.device))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(got, base)
    ref = F.prelu(F.pixel_shuffle(F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1), 2), alpha)
    eps = 2.0 ** -11 if dt == torch.float16 else 2.0 ** -8
    assert (got.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 2 * eps * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 8, 16), (2, 13, 21), (3, 24, 24), (1, 40, 72), (2, 5, 7), (4, 180, 320)])
def test_conv_with_fused_norm_and_skip_matches_unfused(dt, shape):
    """conv3x3_c64_xf2_kernel (bn2 + skip of block l folded into conv1 of block l+1, model.py:65+69 -> :47-54): x_next, the
    conv output and its fixed-point statistics are bit-identical to instnorm_apply(+residual) followed by the plain conv."""
    from fast_srgan_b200 import ops, _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xlib = ctypes.CDLL(os.path.join(root, "fast-srgan_b200", "libfsr_b200_experimental.so"))
    fn = xlib.fsrx_conv3x3_c64_res_in
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    N, H, W = shape
    x_prev = (rnd((N, 64, H, W), 1) * 1.3).permute(0, 2, 3, 1).contiguous().to(dt)
    w0 = rnd((64, 64, 3, 3), 2, 0.05)
    w1 = rnd((64, 64, 3, 3), 3, 0.05)
    wp0, _ = ops.pack_conv3x3(w0, None, dt)
    wp1, _ = ops.pack_conv3x3(w1, None, dt)
    raw2, st2 = ops.conv3x3_c64_raw_stats(x_prev, wp0)                     # stands in for c2 of the previous block
    x_next_ref = ops.instnorm_apply(raw2, st2, residual=x_prev)
    out_ref, st_ref = ops.conv3x3_c64_raw_stats(x_next_ref, wp1)
    x_next = torch.empty_like(x_prev)
    out = torch.empty_like(x_prev)
    stats = torch.zeros((N, 64, 2), dtype=torch.int64, device="cuda")
    rc = fn(raw2.data_ptr(), st2.data_ptr(), 1e-5, x_prev.data_ptr(), x_next.data_ptr(), wp1.data_ptr(), out.data_ptr(),
            stats.data_ptr(), N, H, W, L.dtype_code(dt), L.stream_ptr(x_prev.device))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(x_next, x_next_ref)
    assert torch.equal(out, out_ref)
    assert torch.equal(stats, st_ref)
