"""Thin tensor-level wrappers over the C ABI (include/fsr_b200.h).  Device memory comes from the
PyTorch caching allocator (plumbing); every FLOP runs in libfsr_b200.so.  CUDA tensors only."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fast_srgan_b200.ops: CUDA tensors only (no CPU fallback)")


def pack_conv3x3(weight: torch.Tensor, bias: Optional[torch.Tensor], dtype: torch.dtype, cout_pad: Optional[int] = None,
                 ps_perm: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """OIHW fp32 -> [9][cout_pad][cin] `dtype` (+ permuted/padded fp32 bias)."""
    _cuda(weight, bias)
    cout, cin = weight.shape[0], weight.shape[1]
    cout_pad = cout_pad or cout
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    wp = torch.empty((9, cout_pad, cin), dtype=dtype, device=w.device)
    bp = torch.empty(cout_pad, dtype=torch.float32, device=w.device) if b is not None else None
    L.check(L.load().fsr_pack_conv3x3_weight(L.ptr(w), L.ptr(b), L.ptr(wp), L.ptr(bp), cout, cin, cout_pad,
                                             int(ps_perm), L.dtype_code(dtype), L.stream_ptr(w.device)), "pack")
    return wp, bp


def conv3x3_c64_raw_stats(x: torch.Tensor, w_packed: torch.Tensor, stats: Optional[torch.Tensor] = None):
    """x NHWC [N,H,W,64] -> (raw NHWC [N,H,W,cout], stats [N,cout,2] fp32 (sum, sumsq))."""
    _cuda(x, w_packed)
    N, H, W, C = x.shape
    assert C == 64 and x.is_contiguous()
    cout = w_packed.shape[1]
    out = torch.empty((N, H, W, cout), dtype=x.dtype, device=x.device)
    if stats is None:
        stats = torch.zeros((N, cout, 2), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), None, stats.data_ptr(), None,
                                     N, H, W, cout, L.EPI_RAW_STATS, 0, 0.0, 0, L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 raw+stats")
    return out, stats


def conv3x3_c64_bias_act(x, w_packed, bias, act: int = L.ACT_NONE, slope: float = 0.0, alpha=None):
    _cuda(x, w_packed, bias)
    N, H, W, C = x.shape
    assert C == 64 and x.is_contiguous()
    cout = w_packed.shape[1]
    out = torch.empty((N, H, W, cout), dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias), None, L.ptr(alpha),
                                     N, H, W, cout, L.EPI_BIAS_ACT, act, slope, 0, L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 bias+act")
    return out


def conv3x3_c64_ps_prelu(x, w_packed, bias_packed, alpha):
    """UpSamplingBlock (model.py:39-40): x [N,H,W,64] -> [N,2H,2W,64]; weights packed with ps_perm."""
    _cuda(x, w_packed, bias_packed, alpha)
    N, H, W, C = x.shape
    assert C == 64 and w_packed.shape[1] == 256 and x.is_contiguous()
    out = torch.empty((N, 2 * H, 2 * W, 64), dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias_packed), None,
                                     alpha.data_ptr(), N, H, W, 256, L.EPI_PS_PRELU, 0, 0.0, 0, L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 ps+prelu")
    return out


def conv3x3_c64_head(x, w_packed, bias_packed, out_u8: bool = False):
    """Generator.head (model.py:102-110): x [N,H,W,64] -> fp32 NCHW [N,3,H,W] or uint8 NHWC."""
    _cuda(x, w_packed, bias_packed)
    N, H, W, C = x.shape
    assert C == 64 and w_packed.shape[1] == 16 and x.is_contiguous()
    if out_u8:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device)
    else:
        out = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias_packed), None, None,
                                     N, H, W, 16, L.EPI_HEAD_TANH, 0, 0.0, int(out_u8), L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 head")
    return out


def neck_conv3x3(x, weight, bias, dtype, act=L.ACT_PRELU, slope=0.0, alpha=None, vgg_norm=False):
    """x fp32 NCHW [N,3,H,W] or uint8 NHWC [N,H,W,3] -> NHWC `dtype` [N,H,W,cout]."""
    _cuda(x, weight, bias, alpha)
    in_u8 = x.dtype == torch.uint8
    x = x.contiguous()
    if in_u8:
        N, H, W, _ = x.shape
    else:
        N, _, H, W = x.shape
    cout = weight.shape[0]
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    out = torch.empty((N, H, W, cout), dtype=dtype, device=x.device)
    L.check(L.load().fsr_neck_conv3x3(x.data_ptr(), w.data_ptr(), L.ptr(b), L.ptr(alpha), out.data_ptr(), N, H, W, cout,
                                      act, slope, int(in_u8), int(vgg_norm), L.dtype_code(dtype),
                                      L.stream_ptr(x.device)), "neck conv")
    return out


def instnorm_apply(raw, stats, act=L.ACT_NONE, slope=0.0, alpha=None, residual=None, eps=1e-5, out=None):
    _cuda(raw, stats, alpha, residual)
    N, H, W, C = raw.shape
    if out is None:
        out = torch.empty_like(raw)
    L.check(L.load().fsr_instnorm_apply(raw.data_ptr(), stats.data_ptr(), L.ptr(residual), out.data_ptr(), L.ptr(alpha),
                                        N, H * W, C, act, slope, eps, L.dtype_code(raw.dtype),
                                        L.stream_ptr(raw.device)), "instnorm apply")
    return out


def pixel_shuffle2(x):
    """NHWC [N,H,W,4C] (reference channel order) -> [N,2H,2W,C]  (model.py:36)."""
    _cuda(x)
    N, H, W, C4 = x.shape
    out = torch.empty((N, 2 * H, 2 * W, C4 // 4), dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_pixel_shuffle2(x.data_ptr(), out.data_ptr(), N, H, W, C4 // 4, L.dtype_code(x.dtype),
                                        L.stream_ptr(x.device)), "pixel shuffle")
    return out


def nchw_to_nhwc(x: torch.Tensor, dtype: torch.dtype):
    _cuda(x)
    x = x.contiguous().float()
    N, C, H, W = x.shape
    out = torch.empty((N, H, W, C), dtype=dtype, device=x.device)
    L.check(L.load().fsr_nchw_f32_to_nhwc(x.data_ptr(), out.data_ptr(), N, C, H * W, L.dtype_code(dtype),
                                          L.stream_ptr(x.device)), "nchw->nhwc")
    return out


def nhwc_to_nchw(x: torch.Tensor):
    _cuda(x)
    N, H, W, C = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_nhwc_to_nchw_f32(x.data_ptr(), out.data_ptr(), N, C, H * W, L.dtype_code(x.dtype),
                                          L.stream_ptr(x.device)), "nhwc->nchw")
    return out
