"""Thin tensor-level wrappers over the C ABI (include/fsr_b200.h).  Device memory comes from the
PyTorch caching allocator (plumbing); every FLOP runs in libfsr_b200.so.  CUDA tensors only."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L


STAT_SUM_SCALE, STAT_SQ_SCALE = float(2 ** 24), float(2 ** 20)


def stats_to_float(stats: torch.Tensor) -> torch.Tensor:
    """[N,C,2] int64 fixed point (sum * 2^24, sumsq * 2^20) -> float64 (sum, sumsq)."""
    scale = torch.tensor([STAT_SUM_SCALE, STAT_SQ_SCALE], dtype=torch.float64, device=stats.device)
    return stats.double() / scale


def stats_from_float(sums: torch.Tensor, sumsq: torch.Tensor) -> torch.Tensor:
    """(sum, sumsq) [N,C] -> the kernels' [N,C,2] int64 fixed-point statistics buffer."""
    return torch.stack([(sums.double() * STAT_SUM_SCALE).round(), (sumsq.double() * STAT_SQ_SCALE).round()], dim=-1).to(torch.int64).contiguous()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fast_srgan_b200.ops: CUDA tensors only (no CPU fallback)")


def pack_conv3x3(weight: torch.Tensor, bias: Optional[torch.Tensor], dtype: torch.dtype, cout_pad: Optional[int] = None,
                 ps_perm: bool = False, out_w: Optional[torch.Tensor] = None,
                 out_b: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """OIHW fp32 -> [9][cout_pad][cin] `dtype` (+ permuted/padded fp32 bias).  out_w/out_b: repack in place."""
    _cuda(weight, bias)
    cout, cin = weight.shape[0], weight.shape[1]
    cout_pad = cout_pad or cout
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    wp = out_w if out_w is not None else torch.empty((9, cout_pad, cin), dtype=dtype, device=w.device)
    bp = (out_b if out_b is not None else torch.empty(cout_pad, dtype=torch.float32, device=w.device)) if b is not None else None
    L.check(L.load().fsr_pack_conv3x3_weight(L.ptr(w), L.ptr(b), L.ptr(wp), L.ptr(bp), cout, cin, cout_pad,
                                             int(ps_perm), L.dtype_code(dtype), L.stream_ptr(w.device)), "pack")
    return wp, bp


class PackPlan:
    """A fixed list of weight packs executed as ONE launch (fsr_pack_multi).  add() allocates the persistent destination
    buffer and returns it; run() re-packs everything from the current parameter values."""

    def __init__(self, dtype: torch.dtype):
        self.dtype = dtype
        self.tasks = []
        self._arr = None
        self._keep = []

    def add(self, weight, bias=None, transposed=False, ps_perm=False, flip=False, pad=None, row_scale=None):
        _cuda(weight, bias, row_scale)
        assert weight.dtype == torch.float32 and weight.is_contiguous()
        cout, cin = weight.shape[0], weight.shape[1]
        pad = pad or (cin if transposed else cout)
        shape = (9, pad, cout) if transposed else (9, pad, cin)
        out = torch.empty(shape, dtype=self.dtype, device=weight.device)
        bout = torch.empty(pad, dtype=torch.float32, device=weight.device) if (bias is not None and not transposed) else None
        flags = (L.PACK_T if transposed else 0) | (L.PACK_PS if ps_perm else 0) | (L.PACK_FLIP if flip else 0)
        self.tasks.append(L.FsrPackTask(weight.data_ptr(), out.data_ptr(), L.ptr(bias), L.ptr(bout), L.ptr(row_scale), cout, cin, pad, flags))
        self._keep += [weight, bias, row_scale, out, bout]
        self._arr = None
        return out, bout

    def run(self, device):
        if not self.tasks:
            return
        if self._arr is None:
            self._arr = [(L.FsrPackTask * len(chunk))(*chunk) for chunk in (self.tasks[i:i + 48] for i in range(0, len(self.tasks), 48))]
        for arr in self._arr:
            L.check(L.load().fsr_pack_multi(arr, len(arr), L.dtype_code(self.dtype), L.stream_ptr(device)), "pack_multi")


def conv3x3_c64_raw_stats(x: torch.Tensor, w_packed: torch.Tensor, stats: Optional[torch.Tensor] = None):
    """x NHWC [N,H,W,64] -> (raw NHWC [N,H,W,cout], stats [N,cout,2] fp32 (sum, sumsq))."""
    _cuda(x, w_packed)
    N, H, W, C = x.shape
    assert C == 64 and x.is_contiguous()
    cout = w_packed.shape[1]
    out = torch.empty((N, H, W, cout), dtype=x.dtype, device=x.device)
    if stats is None:
        stats = torch.zeros((N, cout, 2), dtype=torch.int64, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), None, stats.data_ptr(), None,
                                     N, H, W, cout, L.EPI_RAW_STATS, 0, 0.0, 0, L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 raw+stats")
    return out, stats


def conv3x3_c64_in(raw_in: torch.Tensor, stats_in: torch.Tensor, alpha: torch.Tensor, w_packed: torch.Tensor, eps: float = 1e-5,
                   stats: Optional[torch.Tensor] = None):
    """conv3x3(PReLU(InstanceNorm(raw_in))) with the normalisation fused into the conv's load path (model.py:55-64):
    raw_in NHWC [N,H,W,64] + its fixed-point statistics -> (raw NHWC [N,H,W,64], stats int64 [N,64,2])."""
    _cuda(raw_in, stats_in, alpha, w_packed)
    N, H, W, C = raw_in.shape
    assert C == 64 and raw_in.is_contiguous() and w_packed.shape[1] == 64
    out = torch.empty((N, H, W, 64), dtype=raw_in.dtype, device=raw_in.device)
    if stats is None:
        stats = torch.zeros((N, 64, 2), dtype=torch.int64, device=raw_in.device)
    L.check(L.load().fsr_conv3x3_c64_in(raw_in.data_ptr(), stats_in.data_ptr(), alpha.data_ptr(), eps, w_packed.data_ptr(),
                                        out.data_ptr(), stats.data_ptr(), N, H, W, L.dtype_code(raw_in.dtype),
                                        L.stream_ptr(raw_in.device)), "conv3x3 fused IN+PReLU input")
    return out, stats


def conv3x3_c64_res_in(raw_in: torch.Tensor, stats_in: torch.Tensor, res: torch.Tensor, w_packed: torch.Tensor, eps: float = 1e-5):
    """x_next = InstanceNorm(raw_in) + res formed in the conv's load path (model.py:65+69), then conv3x3(x_next):
    -> (x_next NHWC, raw NHWC [N,H,W,64], stats int64 [N,64,2]); bit-identical to instnorm_apply(+residual) + conv."""
    _cuda(raw_in, stats_in, res, w_packed)
    N, H, W, C = raw_in.shape
    assert C == 64 and raw_in.is_contiguous() and res.is_contiguous() and w_packed.shape[1] == 64
    x_next = torch.empty_like(raw_in)
    out = torch.empty_like(raw_in)
    stats = torch.zeros((N, 64, 2), dtype=torch.int64, device=raw_in.device)
    L.check(L.load().fsr_conv3x3_c64_res_in(raw_in.data_ptr(), stats_in.data_ptr(), eps, res.data_ptr(), x_next.data_ptr(),
                                            w_packed.data_ptr(), out.data_ptr(), stats.data_ptr(), N, H, W,
                                            L.dtype_code(raw_in.dtype), L.stream_ptr(raw_in.device)), "conv3x3 fused IN+skip input")
    return x_next, out, stats


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


def conv3x3_c64_head(x, w_packed, bias_packed, out_u8=False, out=None):
    """Generator.head (model.py:102-110): x [N,H,W,64] -> fp32 NCHW [N,3,H,W] or uint8 NHWC.
    out_u8: 0/False tanh->fp32, 1/True tanh->uint8, 2 linear fp32 store, 3 linear fp32 accumulate into `out`."""
    _cuda(x, w_packed, bias_packed)
    N, H, W, C = x.shape
    assert C == 64 and w_packed.shape[1] == 16 and x.is_contiguous()
    if out is not None:
        pass
    elif int(out_u8) == 1:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device)
    else:
        out = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_conv3x3_c64(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias_packed), None, None,
                                     N, H, W, 16, L.EPI_HEAD_TANH, 0, 0.0, int(out_u8), L.dtype_code(x.dtype),
                                     L.stream_ptr(x.device)), "conv3x3 head")
    return out


def neck_conv3x3(x, weight, bias, dtype, act=L.ACT_PRELU, slope=0.0, alpha=None, vgg_norm=False, out=None):
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
    if out is None:
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


def instnorm_apply_parity(raw, stats, act=L.ACT_NONE, slope=0.0, alpha=None, eps=1e-5):
    """instnorm_apply whose output is written directly in the parity-plane layout [N,4,H/2,W/2,C] (input of a stride-2 conv)."""
    _cuda(raw, stats, alpha)
    N, H, W, C = raw.shape
    out = torch.empty((N, 4, H // 2, W // 2, C), dtype=raw.dtype, device=raw.device)
    L.check(L.load().fsr_instnorm_apply_parity(raw.data_ptr(), stats.data_ptr(), out.data_ptr(), L.ptr(alpha), N, H, W, C, act, slope,
                                               eps, L.dtype_code(raw.dtype), L.stream_ptr(raw.device)), "instnorm apply (parity out)")
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


# ============================================================ training-step kernels (trainer.py:168-196)
def pack_conv3x3_t(weight: torch.Tensor, dtype: torch.dtype, ps_perm: bool = False, flip: bool = False,
                   row_pad: Optional[int] = None, row_scale: Optional[torch.Tensor] = None, out=None) -> torch.Tensor:
    """dgrad pack: OIHW fp32 -> [9][cin (pad)][cout(perm)] (rows = input channel, K = output channel)."""
    _cuda(weight, row_scale)
    cout, cin = weight.shape[0], weight.shape[1]
    row_pad = row_pad or cin
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    wp = out if out is not None else torch.empty((9, row_pad, cout), dtype=dtype, device=w.device)
    L.check(L.load().fsr_pack_conv3x3_weight_t(w.data_ptr(), wp.data_ptr(), cout, cin, int(ps_perm), int(flip), row_pad,
                                               L.ptr(row_scale), L.dtype_code(dtype), L.stream_ptr(w.device)), "pack_t")
    return wp


def conv3x3_gen(x, w_packed, cout, stride=1, mode=0, epilogue=L.EPI_BIAS_ACT, bias=None, act=L.ACT_NONE, slope=0.0,
                alpha=None, stats=None, hw=None):
    """General tensor-core conv (see include/fsr_b200.h: fsr_conv3x3_gen).
    mode 0 stride 1: x NHWC [N,H,W,cin]; mode 0 stride 2: x parity planes [N,4,H/2,W/2,cin];
    mode 1 stride 1: x = dY; mode 1 stride 2: x = dY [N,H/2,W/2,cin] -> parity-plane dX.  hw = (H, W) of the
    conv input (forward) / of dX (mode 1); inferred for stride 1."""
    _cuda(x, w_packed, bias, alpha)
    dt = L.dtype_code(x.dtype)
    cin = x.shape[-1]
    if stride == 1:
        N, H, W, _ = x.shape
        if epilogue == L.EPI_PS_PRELU:
            out = torch.empty((N, 2 * H, 2 * W, cout // 4), dtype=x.dtype, device=x.device)
        else:
            out = torch.empty((N, H, W, cout), dtype=x.dtype, device=x.device)
    elif mode == 0:
        N, _, H2, W2, _ = x.shape
        H, W = 2 * H2, 2 * W2
        out = torch.empty((N, H2, W2, cout), dtype=x.dtype, device=x.device)
    else:
        N, H2, W2, _ = x.shape
        H, W = 2 * H2, 2 * W2
        out = torch.empty((N, 4, H2, W2, cout), dtype=x.dtype, device=x.device)
    if epilogue == L.EPI_RAW_STATS and stats is None:
        stats = torch.zeros((N, cout, 2), dtype=torch.int64, device=x.device)
    L.check(L.load().fsr_conv3x3_gen(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias), L.ptr(stats), L.ptr(alpha),
                                     N, H, W, cin, cout, stride, mode, epilogue, act, slope, dt, L.stream_ptr(x.device)),
            "conv3x3_gen")
    return (out, stats) if epilogue == L.EPI_RAW_STATS else out


_WGRAD_WS = {}


def wgrad_workspace(device) -> torch.Tensor:
    """Per-device scratch for the split-K partial tiles of the weight-gradient kernels (24 MB; contents never persist
    between calls, calls on one stream are ordered)."""
    key = str(device)
    if key not in _WGRAD_WS:
        _WGRAD_WS[key] = torch.empty(L.load().fsr_wgrad_workspace_bytes(), dtype=torch.uint8, device=device)
    return _WGRAD_WS[key]


def conv3x3_gen_flat(xp, w_packed, cout, mode=0, bias=None, act=L.ACT_NONE, slope=0.0):
    """General conv on zero-bordered padded tensors: xp [N,H+2,W+2,cin] -> [N,H+2,W+2,cout] (borders written as zero).
    mode 0 forward, mode 1 data gradient.  See include/fsr_b200.h: fsr_conv3x3_gen_flat."""
    _cuda(xp, w_packed, bias)
    N, Hp, Wp, cin = xp.shape
    assert xp.is_contiguous()
    out = torch.empty((N, Hp, Wp, cout), dtype=xp.dtype, device=xp.device)
    L.check(L.load().fsr_conv3x3_gen_flat(xp.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias), N, Hp - 2, Wp - 2, cin, cout,
                                          mode, act, slope, L.dtype_code(xp.dtype), L.stream_ptr(xp.device)), "conv3x3_gen_flat")
    return out


def maxpool2_padded(x, in_pad: bool, out_pad: bool):
    """2x2 max-pool between plain [N,H,W,C] and padded [N,H+2,W+2,C] layouts (either side)."""
    _cuda(x)
    N, Hx, Wx, C = x.shape
    H, W = (Hx - 2, Wx - 2) if in_pad else (Hx, Wx)
    shape = (N, H // 2 + 2, W // 2 + 2, C) if out_pad else (N, H // 2, W // 2, C)
    out = torch.zeros(shape, dtype=x.dtype, device=x.device) if out_pad else torch.empty(shape, dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_maxpool2_padded(x.data_ptr(), out.data_ptr(), N, H, W, C, int(in_pad), int(out_pad), L.dtype_code(x.dtype),
                                         L.stream_ptr(x.device)), "maxpool (padded)")
    return out


def maxpool2_relu_bwd_padded(x, dout, in_pad: bool, out_pad: bool):
    """backward of [ReLU -> maxpool]: x = pre-pool activation (plain or padded), dout = pooled gradient (plain or padded)."""
    _cuda(x, dout)
    N, Hx, Wx, C = x.shape
    H, W = (Hx - 2, Wx - 2) if in_pad else (Hx, Wx)
    din = torch.zeros_like(x) if in_pad else torch.empty_like(x)
    L.check(L.load().fsr_maxpool2_relu_bwd_padded(x.data_ptr(), dout.data_ptr(), din.data_ptr(), N, H, W, C, int(in_pad), int(out_pad),
                                                  L.dtype_code(x.dtype), L.stream_ptr(x.device)), "maxpool bwd (padded)")
    return din


def conv3x3_wgrad(x, dy, dw, stride=1, ps_perm=False):
    """dw (fp32 OIHW, accumulated) += wgrad(x, dy).  stride 2: x in parity planes [N,4,H/2,W/2,cin]."""
    _cuda(x, dy, dw)
    if stride == 1:
        N, H, W, cin = x.shape
    else:
        N, _, H2, W2, cin = x.shape
        H, W = 2 * H2, 2 * W2
    cout = dy.shape[-1]
    ws = wgrad_workspace(x.device)
    L.check(L.load().fsr_conv3x3_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, cin, cout, stride, int(ps_perm),
                                       ws.data_ptr(), ws.numel(), L.dtype_code(x.dtype), L.stream_ptr(x.device)), "wgrad")
    return dw


def conv3x3_wgrad_grouped(x_arena, dy_arena, dws):
    """ONE launch for len(dws) weight gradients of identical shape: x_arena / dy_arena [G,N,H,W,C] NHWC, dws[g] fp32 OIHW (+=)."""
    import ctypes
    _cuda(x_arena, dy_arena, *dws)
    G, N, H, W, cin = x_arena.shape
    cout = dy_arena.shape[-1]
    assert len(dws) == G and x_arena.is_contiguous() and dy_arena.is_contiguous()
    ptrs = (ctypes.c_void_p * G)(*[d.data_ptr() for d in dws])
    ws = wgrad_workspace(x_arena.device)
    L.check(L.load().fsr_conv3x3_wgrad_grouped(x_arena.data_ptr(), dy_arena.data_ptr(), ptrs, G, x_arena.stride(0), dy_arena.stride(0),
                                               N, H, W, cin, cout, ws.data_ptr(), ws.numel(), L.dtype_code(x_arena.dtype),
                                               L.stream_ptr(x_arena.device)), "wgrad grouped")


def parity_layout(x, to_parity=True):
    _cuda(x)
    if to_parity:
        N, H, W, C = x.shape
        out = torch.empty((N, 4, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    else:
        N, _, H2, W2, C = x.shape
        H, W = 2 * H2, 2 * W2
        out = torch.empty((N, H, W, C), dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_parity_layout(x.data_ptr(), out.data_ptr(), N, H, W, C, int(to_parity), L.dtype_code(x.dtype),
                                       L.stream_ptr(x.device)), "parity")
    return out


def maxpool2(x):
    _cuda(x)
    N, H, W, C = x.shape
    out = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    L.check(L.load().fsr_maxpool2(x.data_ptr(), out.data_ptr(), N, H, W, C, L.dtype_code(x.dtype), L.stream_ptr(x.device)), "maxpool")
    return out


def maxpool2_relu_bwd(x, dout):
    _cuda(x, dout)
    N, H, W, C = x.shape
    din = torch.empty_like(x)
    L.check(L.load().fsr_maxpool2_relu_bwd(x.data_ptr(), dout.data_ptr(), din.data_ptr(), N, H, W, C, L.dtype_code(x.dtype),
                                           L.stream_ptr(x.device)), "maxpool bwd")
    return din


def relu_bwd(y, dy):
    _cuda(y, dy)
    dx = torch.empty_like(y)
    L.check(L.load().fsr_relu_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel(), L.dtype_code(y.dtype), L.stream_ptr(y.device)), "relu bwd")
    return dx


def add(a, b, out=None):
    _cuda(a, b)
    if out is None:
        out = torch.empty_like(a)
    L.check(L.load().fsr_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), L.dtype_code(a.dtype), L.stream_ptr(a.device)), "add")
    return out


def conv1x1_to1_fwd(x, w, b):
    """x NHWC [N,H,W,C], w fp32 [C], b fp32 [1] -> fp32 logits [N,H,W]."""
    _cuda(x, w, b)
    N, H, W, C = x.shape
    z = torch.empty((N, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_conv1x1_to1_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), z.data_ptr(), N * H * W, C,
                                         L.dtype_code(x.dtype), L.stream_ptr(x.device)), "conv1x1 fwd")
    return z


def conv1x1_to1_bwd(x, w, dz, dw=None, db=None, need_dx=True):
    _cuda(x, w, dz, dw, db)
    N, H, W, C = x.shape
    dx = torch.empty_like(x) if need_dx else None
    L.check(L.load().fsr_conv1x1_to1_bwd(x.data_ptr(), w.data_ptr(), dz.data_ptr(), L.ptr(dx), L.ptr(dw), L.ptr(db), N * H * W, C,
                                         L.dtype_code(x.dtype), L.stream_ptr(x.device)), "conv1x1 bwd")
    return dx


def bce_logits(z, noise, lab_scale, lab_shift, loss_out, dz=None, grad_scale=1.0):
    _cuda(z, noise, loss_out, dz)
    L.check(L.load().fsr_bce_logits(z.data_ptr(), noise.data_ptr(), lab_scale, lab_shift, z.numel(), loss_out.data_ptr(), L.ptr(dz),
                                    grad_scale, L.stream_ptr(z.device)), "bce")
    return loss_out


def smooth_l1(a, b, loss_acc, da=None, grad_scale=1.0):
    _cuda(a, b, loss_acc, da)
    dt = 2 if a.dtype == torch.float32 else L.dtype_code(a.dtype)
    L.check(L.load().fsr_smooth_l1(a.data_ptr(), b.data_ptr(), a.numel(), loss_acc.data_ptr(), L.ptr(da), grad_scale, dt,
                                   L.stream_ptr(a.device)), "smooth l1")
    return loss_acc


def instnorm_bwd(raw, stats, dy, act=L.ACT_NONE, slope=0.0, alpha=None, dalpha=None, eps=1e-5, out=None):
    _cuda(raw, stats, dy, alpha, dalpha)
    N, H, W, C = raw.shape
    red = torch.empty((N, C, 2), dtype=torch.float32, device=raw.device)
    draw = out if out is not None else torch.empty_like(raw)
    L.check(L.load().fsr_instnorm_bwd(raw.data_ptr(), stats.data_ptr(), dy.data_ptr(), red.data_ptr(), draw.data_ptr(), L.ptr(alpha),
                                      L.ptr(dalpha), N, H * W, C, act, slope, eps, L.dtype_code(raw.dtype), L.stream_ptr(raw.device)),
            "instnorm bwd")
    return draw


def instnorm_bwd_parity(raw, stats, dy_parity, act=L.ACT_NONE, slope=0.0, alpha=None, dalpha=None, eps=1e-5):
    """instnorm_bwd with dy in the parity-plane layout [N,4,H/2,W/2,C] (output of a stride-2 data gradient)."""
    _cuda(raw, stats, dy_parity, alpha, dalpha)
    N, H, W, C = raw.shape
    draw = torch.empty_like(raw)
    L.check(L.load().fsr_instnorm_bwd_parity(raw.data_ptr(), stats.data_ptr(), dy_parity.data_ptr(), draw.data_ptr(), L.ptr(alpha),
                                             L.ptr(dalpha), N, H, W, C, act, slope, eps, L.dtype_code(raw.dtype), L.stream_ptr(raw.device)),
            "instnorm bwd (parity dy)")
    return draw


def act_bwd(y, dy, act, slope=0.0, alpha=None, dalpha=None):
    _cuda(y, dy, alpha, dalpha)
    dv = torch.empty_like(y)
    L.check(L.load().fsr_act_bwd(y.data_ptr(), dy.data_ptr(), dv.data_ptr(), y.numel(), L.ptr(alpha), slope, act, L.ptr(dalpha),
                                 L.dtype_code(y.dtype), L.stream_ptr(y.device)), "act bwd")
    return dv


def ps_prelu_bwd(U, dU, alpha, dalpha=None):
    _cuda(U, dU, alpha, dalpha)
    N, H2, W2, C = U.shape
    H, W = H2 // 2, W2 // 2
    dconv = torch.empty((N, H, W, 4 * C), dtype=U.dtype, device=U.device)
    L.check(L.load().fsr_ps_prelu_bwd(U.data_ptr(), dU.data_ptr(), dconv.data_ptr(), N, H, W, C, alpha.data_ptr(), L.ptr(dalpha),
                                      L.dtype_code(U.dtype), L.stream_ptr(U.device)), "ps prelu bwd")
    return dconv


def tanh_bwd(y, dy):
    _cuda(y, dy)
    dpre = torch.empty_like(y)
    L.check(L.load().fsr_tanh_bwd(y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), y.numel(), L.stream_ptr(y.device)), "tanh bwd")
    return dpre


def wgrad_c3(img, act, out, flip=False, layout=0):
    """out fp32 += sum img[n,c3,y+dy,x+dx] * act[n,y,x,c]; layout 0 [27][C], 1 OIHW [3][C][3][3], 2 OIHW [C][3][3][3]."""
    _cuda(img, act, out)
    N, H, W, C = act.shape
    L.check(L.load().fsr_wgrad_c3(img.data_ptr(), act.data_ptr(), out.data_ptr(), N, H, W, C, int(flip), layout, L.dtype_code(act.dtype),
                                  L.stream_ptr(act.device)), "wgrad c3")
    return out


def bias_grad(g, db, ps_perm=False):
    _cuda(g, db)
    C = g.shape[-1]
    L.check(L.load().fsr_bias_grad(g.data_ptr(), db.data_ptr(), g.numel() // C, C, int(ps_perm), L.dtype_code(g.dtype), L.stream_ptr(g.device)), "bias grad")
    return db


def bias_grad_nchw(g, db):
    _cuda(g, db)
    N, C, H, W = g.shape
    L.check(L.load().fsr_bias_grad_nchw(g.data_ptr(), db.data_ptr(), N, C, H * W, L.stream_ptr(g.device)), "bias grad nchw")
    return db


def adamw(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8, wd=1e-2, grad_scale=1.0):
    _cuda(p, g, m, v)
    L.check(L.load().fsr_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, b1, b2, eps, wd, step, grad_scale,
                               L.stream_ptr(p.device)), "adamw")


def adamw_dev(p, g, m, v, lr, step_dev, b1=0.9, b2=0.999, eps=1e-8, wd=1e-2, grad_scale=1.0):
    """AdamW with the step counter in device memory (int32 tensor, incremented by the call) - CUDA-graph safe."""
    _cuda(p, g, m, v, step_dev)
    L.check(L.load().fsr_adamw_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, b1, b2, eps, wd,
                                   step_dev.data_ptr(), grad_scale, L.stream_ptr(p.device)), "adamw_dev")


def conv3x3_head(x, w_packed, bias_packed, out_mode: int = 0, out=None):
    """Generator.head for any channel count (cin multiple of 64): x NHWC [N,H,W,cin], w_packed [9][16][cin]."""
    _cuda(x, w_packed, bias_packed)
    N, H, W, C = x.shape
    if out is None:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device) if out_mode == 1 else \
            torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().fsr_conv3x3_head(x.data_ptr(), w_packed.data_ptr(), out.data_ptr(), L.ptr(bias_packed), N, H, W, C,
                                      out_mode, L.dtype_code(x.dtype), L.stream_ptr(x.device)), "conv3x3 head")
    return out
