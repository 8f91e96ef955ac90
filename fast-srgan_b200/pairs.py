"""Generator forward for n_filters <= 32 on "pixel-pair rows"  (reference model.py:72-117 with generator.n_filters = 32).

The tensor-core kernels work on 128-byte pixel rows (64 two-byte channels: TMA box rows, 128B swizzle, UMMA K chunks).
Zero-padding a 32-channel network to 64 channels costs 4x the FLOPs and 2x the bytes.  Instead a 32-channel NHWC
tensor [N,H,W,32] (W even) is read as [N,H,W/2,64]: one row = two horizontally adjacent pixels, slot (parity p, channel c)
= 32p + c.  A 3x3 conv on the pixel grid is again a 3x3 conv on the pair grid: output pair x', parity po, needs the
pixels 2x'+po+dx (dx = -1,0,1), which live in the pairs x'-1, x', x'+1:

    W_pair[(po, co), (pi, ci), ky, s] = W[co, ci, ky, dx]   with dx = 2(s-1) + pi - po,   zero unless -1 <= dx <= 1

(6 of the 12 (s, pi, po) blocks are non-zero), and the zero padding of the pair grid is the zero padding of the pixel
grid.  So every 64-channel kernel of libfsr_b200.so runs unchanged at HALF the pixel count - 2x the ideal FLOPs instead
of 4x, 1x the bytes instead of 2x - with three additions (include/fsr_b200.h): the InstanceNorm sums of slots c and
32+c are folded (fsr_in_stats_fold_pair), the neck stores 32-channel pixels (fsr_neck_conv3x3_c32) and the head writes
two rgb pixels per row (fsr_conv3x3_c64_head_pair).  PixelShuffle keeps working because the 2x2 block order of the
CTA-pair kernel's epilogue maps onto pairs again (see _expand_up).
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from . import _lib as L


def expand_pair(w: torch.Tensor) -> torch.Tensor:
    """[co, ci, 3, 3] -> [2co, 2ci, 3, 3] on the pair grid (rows (po, co), columns (pi, ci))."""
    co, ci = w.shape[:2]
    out = w.new_zeros((2 * co, 2 * ci, 3, 3))
    for po in (0, 1):
        for pi in (0, 1):
            for s in (0, 1, 2):
                dx = 2 * (s - 1) + pi - po
                if -1 <= dx <= 1:
                    out[po * co:(po + 1) * co, pi * ci:(pi + 1) * ci, :, s] = w[:, :, :, dx + 1]
    return out


def _expand_up(w: torch.Tensor, b: torch.Tensor):
    """32 -> 128 conv + PixelShuffle(2) (model.py:39-40) as a 64 -> 256 conv on pairs, rows in the kernel's GEMM order.

    The kernel writes GEMM column block q = 2i + j' (64 values) to out[n, 2y+i, 2x'+j', 0:64] of a [N,2H,2W',64] tensor
    (W' = W/2 pairs).  Read as pixels that row is the out-pixel pair (4x'+2j', 4x'+2j'+1) = (2X+0, 2X+1) for the input
    pixel X = 2x'+j' - so block (i, j') holds, for input parity po = j', the conv channels 4c+2i+j at slot 32j + c."""
    E = expand_pair(w)                                   # rows po*128 + k,  k = 4c + 2i + j
    idx = torch.empty(256, dtype=torch.long)
    bidx = torch.empty(256, dtype=torch.long)
    for i in (0, 1):
        for jp in (0, 1):
            for j in (0, 1):
                for c in range(32):
                    r = (2 * i + jp) * 64 + 32 * j + c
                    idx[r] = jp * 128 + 4 * c + 2 * i + j
                    bidx[r] = 4 * c + 2 * i + j
    return E[idx.to(w.device)].contiguous(), b[bidx.to(b.device)].contiguous()


def _pad32(t: torch.Tensor, o=None, i=None) -> torch.Tensor:
    t = t.detach().float()
    shape = list(t.shape)
    if o is not None:
        shape[0] = o
    if i is not None:
        shape[1] = i
    out = torch.zeros(shape, dtype=torch.float32, device=t.device)
    out[tuple(slice(0, d) for d in t.shape)] = t
    return out


class PairGenerator:
    """Functional forward over a fast_srgan_b200.model.Generator's parameters, n_filters <= 32, even LR width."""

    def __init__(self, module):
        self.m = module
        self._pk: Dict[str, torch.Tensor] = {}
        self._key = None

    def _pack(self):
        from . import ops
        m = self.m
        dt = m.compute_dtype
        key = (dt,) + tuple((c.weight._version, c.weight.data_ptr()) for _, c in m._conv_list())
        if key == self._key:
            return
        pk: Dict[str, torch.Tensor] = {}
        for name, conv in m._conv_list():
            w, b = conv.weight, conv.bias
            if name == "neck":                               # [F,3,3,3] -> 64 rows (32 real-or-zero + 32 never stored)
                pk["neck.w"], pk["neck.b"] = _pad32(w, 64).contiguous(), _pad32(b, 64).contiguous()
            elif name.startswith("up"):                      # reference channel 4c+q keeps its index (c < F)
                w2, b2 = _expand_up(_pad32(w, 128, 32), _pad32(b, 128))
                pk[name + ".w"], pk[name + ".b"] = ops.pack_conv3x3(w2, b2, dt)
            elif name == "head":
                w2 = expand_pair(_pad32(w, 3, 32))           # rows (po, rgb)
                b2 = torch.cat([b.detach().float(), b.detach().float()])
                pk[name + ".w"], pk[name + ".b"] = ops.pack_conv3x3(w2, b2, dt, cout_pad=16)
            else:
                pk[name + ".w"], _ = ops.pack_conv3x3(expand_pair(_pad32(w, 32, 32)), None, dt)
        self._pk, self._key = pk, key

    @staticmethod
    def _fold(stats: torch.Tensor):
        L.check(L.load().fsr_in_stats_fold_pair(stats.data_ptr(), stats.shape[0], L.stream_ptr(stats.device)), "in_stats_fold_pair")
        return stats

    def forward(self, x: torch.Tensor, out: torch.Tensor, in_u8: int, out_u8: int) -> torch.Tensor:
        from . import ops
        m = self.m
        self._pack()
        pk, lib, dt = self._pk, L.load(), m.compute_dtype
        dev = x.device
        st = L.stream_ptr(dev)
        if in_u8:
            N, H, W, _ = x.shape
        else:
            N, _, H, W = x.shape
        Wp = W // 2
        a0 = torch.empty((N, H, Wp, 64), dtype=dt, device=dev)                        # = [N,H,W,32]
        L.check(lib.fsr_neck_conv3x3_c32(x.data_ptr(), pk["neck.w"].data_ptr(), pk["neck.b"].data_ptr(), m.neck[1].weight.data_ptr(),
                                         a0.data_ptr(), N, H, W, L.ACT_PRELU, 0.0, int(in_u8), L.dtype_code(dt), st), "neck c32")   # model.py:75-78
        lib.fsr_set_pair_rows(int(os.environ.get("FSR_PAIR_SKIP", "1") != "0"))    # skip the structural-zero K-steps
        try:
            return self._chain(a0, out, N, out_u8)
        finally:
            lib.fsr_set_pair_rows(0)

    def _chain(self, a0, out, N, out_u8):
        from . import ops
        m, pk, lib, dt = self.m, self._pk, L.load(), self.m.compute_dtype
        st = L.stream_ptr(a0.device)
        cur = a0
        for i, blk in enumerate(m.stem):                                               # model.py:55-69
            raw1, st1 = ops.conv3x3_c64_raw_stats(cur, pk[f"s{i}a.w"])
            self._fold(st1)
            raw2, st2 = ops.conv3x3_c64_in(raw1, st1, blk.relu1.weight, pk[f"s{i}b.w"])
            self._fold(st2)
            cur = ops.instnorm_apply(raw2, st2, residual=cur)
        raw, stb = ops.conv3x3_c64_raw_stats(cur, pk["bott.w"])                        # model.py:86-95, long skip :115
        self._fold(stb)
        cur = ops.instnorm_apply(raw, stb, residual=a0)
        for i in range(2):                                                             # model.py:39-40
            cur = ops.conv3x3_c64_ps_prelu(cur, pk[f"up{i}.w"], pk[f"up{i}.b"], m.upsampling[i].relu.weight)
        N_, h, wp, _ = cur.shape
        L.check(lib.fsr_conv3x3_c64_head_pair(cur.data_ptr(), pk["head.w"].data_ptr(), out.data_ptr(), pk["head.b"].data_ptr(),
                                              N, h, wp, int(out_u8), L.dtype_code(dt), st), "head pair")                             # model.py:102-110
        return out
