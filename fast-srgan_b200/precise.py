"""Precise generator forward: `Generator(config, compute_dtype=torch.float32)`  (reference model.py:112-117).

Measured motivation (tests/test_baseline_configs_gpu.py::test_shipped_checkpoint_parity): on the
reference's shipped checkpoint fp16 operands give 3.4e-3 max-abs against the reference (bf16 2.9e-2) - the trained
residual stream reaches |x| ~ 17 and every 16-bit rounding (stored skip stream, raw conv outputs, conv operands,
weights) costs 1.0e-3 ... 1.9e-3 on its own - so north_star's 1e-3 on that fixture needs ~fp32 arithmetic.

The tensor cores still do the work (csrc/precise.cuh): activations live in fp32 NHWC and are split into two fp16
planes a = a_hi + a_lo, weights likewise, and every conv is three tcgen05 launches accumulating in fp32
(a_hi*w_hi + a_lo*w_hi + a_hi*w_lo; epilogue FSR_EPI_F32).  Everything runs in libfsr_b200.so; torch only owns memory.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib as L


def _split_weight(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = w.half().float()
    return hi, w - hi


class PreciseGenerator:
    """Functional forward over a fast_srgan_b200.model.Generator's parameters (n_filters == 64 after padding)."""

    def __init__(self, module):
        self.m = module
        self._pk: Dict[str, torch.Tensor] = {}
        self._key = None

    # ------------------------------------------------------------------ weights
    def _pack(self):
        from . import ops
        m = self.m
        key = tuple((c.weight._version, c.weight.data_ptr()) for _, c in m._conv_list())
        if key == self._key:
            return
        pk: Dict[str, torch.Tensor] = {}
        h = torch.float16
        for name, (w, b) in m._effective_weights().items():
            if name == "neck":
                pk["neck.w"], pk["neck.b"] = w.contiguous(), b.contiguous()
                continue
            hi, lo = _split_weight(w)
            kw = dict(ps_perm=True) if name.startswith("up") else (dict(cout_pad=16) if name == "head" else {})
            pk[name + ".hi"], bp = ops.pack_conv3x3(hi, b, h, **kw)
            pk[name + ".lo"], _ = ops.pack_conv3x3(lo, None, h, **kw)
            if bp is not None:
                pk[name + ".b"] = bp
        self._pk, self._key = pk, key

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _conv(hi, lo, w_hi, w_lo, cout):
        """three-product split conv: fp32 NHWC [N,H,W,cout]"""
        lib = L.load()
        N, H, W, _ = hi.shape
        out = torch.empty((N, H, W, cout), dtype=torch.float32, device=hi.device)
        st = L.stream_ptr(hi.device)
        for a, w, acc in ((hi, w_hi, 0), (lo, w_hi, 1), (hi, w_lo, 1)):
            L.check(lib.fsr_conv3x3_c64(a.data_ptr(), w.data_ptr(), out.data_ptr(), None, None, None, N, H, W, cout,
                                        L.EPI_F32, acc, 0.0, 0, L.FSR_F16, st), "conv3x3 fp32-accumulate")
        return out

    @staticmethod
    def _norm(raw, act, alpha, residual):
        """InstanceNorm (+PReLU) (+residual) on fp32 NHWC -> (fp32, hi, lo)"""
        lib = L.load()
        N, H, W, C = raw.shape
        st = L.stream_ptr(raw.device)
        stats = torch.zeros((N, C, 2), dtype=torch.int64, device=raw.device)
        L.check(lib.fsr_in_stats_f32(raw.data_ptr(), stats.data_ptr(), N, H * W, st), "in_stats_f32")
        out = torch.empty_like(raw)
        hi = torch.empty(raw.shape, dtype=torch.float16, device=raw.device)
        lo = torch.empty_like(hi)
        L.check(lib.fsr_in_apply_f32(raw.data_ptr(), stats.data_ptr(), L.ptr(residual), out.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                     L.ptr(alpha), act, N, H * W, 1e-5, st), "in_apply_f32")
        return out, hi, lo

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, out: torch.Tensor, in_u8: int, out_u8: int) -> torch.Tensor:
        m = self.m
        if m.padded_filters != 64:
            raise RuntimeError("compute_dtype=float32 (precise mode) is built for generator.n_filters <= 64")
        self._pack()
        pk, lib = self._pk, L.load()
        dev = x.device
        st = L.stream_ptr(dev)
        if in_u8:                                           # inference.py:48-51
            x = (x.to(torch.float32) / 127.5 - 1.0).permute(0, 3, 1, 2).contiguous()
        N, _, H, W = x.shape
        a0 = torch.empty((N, H, W, 64), dtype=torch.float32, device=dev)
        L.check(lib.fsr_neck_conv3x3_f32(x.data_ptr(), pk["neck.w"].data_ptr(), pk["neck.b"].data_ptr(), m.neck[1].weight.data_ptr(),
                                         a0.data_ptr(), N, H, W, st), "neck f32")                       # model.py:75-78
        a0_hi = torch.empty(a0.shape, dtype=torch.float16, device=dev)
        a0_lo = torch.empty_like(a0_hi)
        L.check(lib.fsr_split_f32(a0.data_ptr(), a0_hi.data_ptr(), a0_lo.data_ptr(), a0.numel(), st), "split")
        cur, cur_hi, cur_lo = a0, a0_hi, a0_lo
        for i, blk in enumerate(m.stem):                                                                 # model.py:67-69
            c1 = self._conv(cur_hi, cur_lo, pk[f"s{i}a.hi"], pk[f"s{i}a.lo"], 64)
            _, y_hi, y_lo = self._norm(c1, L.ACT_PRELU, blk.relu1.weight, None)
            c2 = self._conv(y_hi, y_lo, pk[f"s{i}b.hi"], pk[f"s{i}b.lo"], 64)
            cur, cur_hi, cur_lo = self._norm(c2, L.ACT_NONE, None, cur)
        cb = self._conv(cur_hi, cur_lo, pk["bott.hi"], pk["bott.lo"], 64)                                # model.py:86-95
        _, hi, lo = self._norm(cb, L.ACT_NONE, None, a0)                                                 # + long skip :115
        h, w = H, W
        for i in range(2):                                                                               # model.py:39-40
            conv = self._conv(hi, lo, pk[f"up{i}.hi"], pk[f"up{i}.lo"], 256)
            u = torch.empty((N, 2 * h, 2 * w, 64), dtype=torch.float32, device=dev)
            hi = torch.empty(u.shape, dtype=torch.float16, device=dev)
            lo = torch.empty_like(hi)
            L.check(lib.fsr_ps_prelu_f32(conv.data_ptr(), pk[f"up{i}.b"].data_ptr(), m.upsampling[i].relu.weight.data_ptr(), u.data_ptr(),
                                         hi.data_ptr(), lo.data_ptr(), N, h, w, st), "ps_prelu_f32")
            h, w = 2 * h, 2 * w
            del conv, u
        pre = out if not out_u8 else torch.empty((N, 3, h, w), dtype=torch.float32, device=dev)         # model.py:102-110
        for a, wgt, bias, mode in ((hi, pk["head.hi"], pk["head.b"], 2), (lo, pk["head.hi"], None, 3), (hi, pk["head.lo"], None, 3)):
            L.check(lib.fsr_conv3x3_c64(a.data_ptr(), wgt.data_ptr(), pre.data_ptr(), L.ptr(bias), None, None, N, h, w, 16,
                                        L.EPI_HEAD_TANH, 0, 0.0, mode, L.FSR_F16, st), "head fp32-accumulate")
        L.check(lib.fsr_tanh_f32(pre.data_ptr(), out.data_ptr() if out_u8 else None, N, h * w, st), "tanh")
        return out
