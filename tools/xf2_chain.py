"""EXPERIMENTAL A/B (next round's first GPU slot): Generator.forward with the residual chain's  bn2 + skip  fused into the
next conv's load path (libfsr_b200_experimental.so: fsrx_conv3x3_c64_res_in) against the shipped forward.

    python tools/xf2_chain.py [--batch 32] [--h 180] [--w 320] [--iters 20]

Runs the chain op by op through the C ABI (Python launch overhead included in BOTH arms: the baseline arm is the same
op-level chain with fsr_instnorm_apply), checks the two outputs are bit-identical, and prints both step times.
Chain per block l >= 1:  c1 = conv_res_in(c2[l-1], stats2[l-1], x[l-1] -> writes x[l]);  c2 = conv_in(c1, stats1)."""
import argparse
import ctypes
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_srgan_b200 import _lib as L, ops  # noqa: E402
from fast_srgan_b200.model import Generator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--h", type=int, default=180)
ap.add_argument("--w", type=int, default=320)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()

xlib = ctypes.CDLL(os.path.join(ROOT, "fast-srgan_b200", "libfsr_b200_experimental.so"))
res_in = xlib.fsrx_conv3x3_c64_res_in
res_in.restype = ctypes.c_int
res_in.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p]

dt = torch.float16
torch.manual_seed(1234)
g = Generator(types.SimpleNamespace(n_filters=64, n_layers=8), compute_dtype=dt).cuda().eval()
g._pack()
pk = g._packed
x = (torch.rand((args.batch, 3, args.h, args.w)) * 2 - 1).cuda()
N, H, W = args.batch, args.h, args.w
Lr = g.n_layers


def tail(cur):
    for i in range(2):
        cur = ops.conv3x3_c64_ps_prelu(cur, pk[f"up{i}.w"], pk[f"up{i}.b"], g.upsampling[i].relu.weight)
    return ops.conv3x3_head(cur, pk["head.w"], pk["head.b"], out_mode=0)


def chain_baseline():
    a0 = ops.neck_conv3x3(x, pk["neck.w"], pk["neck.b"], dt, act=L.ACT_PRELU, alpha=g.neck[1].weight)
    cur = a0
    for i, blk in enumerate(g.stem):
        raw1, s1 = ops.conv3x3_c64_raw_stats(cur, pk[f"s{i}a.w"])
        raw2, s2 = ops.conv3x3_c64_in(raw1, s1, blk.relu1.weight, pk[f"s{i}b.w"])
        cur = ops.instnorm_apply(raw2, s2, residual=cur)
    rawb, sb = ops.conv3x3_c64_raw_stats(cur, pk["bott.w"])
    return tail(ops.instnorm_apply(rawb, sb, residual=a0))


xbuf = [torch.empty((N, H, W, 64), dtype=dt, device="cuda") for _ in range(2)]


def conv_res_in(raw_prev, st_prev, x_prev, x_next, wp):
    out = torch.empty_like(raw_prev)
    stats = torch.zeros((N, 64, 2), dtype=torch.int64, device="cuda")
    rc = res_in(raw_prev.data_ptr(), st_prev.data_ptr(), 1e-5, x_prev.data_ptr(), x_next.data_ptr(), wp.data_ptr(), out.data_ptr(),
                stats.data_ptr(), N, H, W, L.dtype_code(dt), L.stream_ptr(raw_prev.device))
    assert rc == 0, rc
    return out, stats


def chain_xf2():
    a0 = ops.neck_conv3x3(x, pk["neck.w"], pk["neck.b"], dt, act=L.ACT_PRELU, alpha=g.neck[1].weight)
    x_prev, raw2, s2 = a0, None, None
    for i, blk in enumerate(g.stem):
        if i == 0:
            raw1, s1 = ops.conv3x3_c64_raw_stats(x_prev, pk[f"s{i}a.w"])
        else:
            x_next = xbuf[i & 1]
            raw1, s1 = conv_res_in(raw2, s2, x_prev, x_next, pk[f"s{i}a.w"])      # x[i] = IN(c2[i-1]) + x[i-1], then conv1
            x_prev = x_next
        raw2, s2 = ops.conv3x3_c64_in(raw1, s1, blk.relu1.weight, pk[f"s{i}b.w"])
    x_next = xbuf[Lr & 1]
    rawb, sb = conv_res_in(raw2, s2, x_prev, x_next, pk["bott.w"])                 # x[L], then the bottleneck conv
    return tail(ops.instnorm_apply(rawb, sb, residual=a0))


def timed(fn):
    for _ in range(3):
        y = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        y = fn()
    e1.record()
    torch.cuda.synchronize()
    return y, e0.elapsed_time(e1) / args.iters


with torch.no_grad():
    ref = g(x)
    yb, tb = timed(chain_baseline)
    yx, tx = timed(chain_xf2)
print(json.dumps({"shipped_forward_equals_op_chain": bool(torch.equal(ref, yb)), "xf2_bit_identical": bool(torch.equal(yb, yx)),
                  "ms_op_chain_baseline": tb, "ms_op_chain_xf2": tx, "batch": N, "hw": [H, W]}))
