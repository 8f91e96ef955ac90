"""Functional training engine: forward + hand-written backward of Generator / Discriminator / VGG19
on libfsr_b200 kernels, and the GAN step of reference trainer.py:168-196.

No autograd: the networks are static, so the backward pass is an explicit reverse walk over saved
activations (every FLOP in libfsr_b200.so; torch only owns memory).  Parameter gradients are
accumulated in fp32 straight into torch-layout (OIHW) views of ONE flat buffer per network, which is
what the NCCL all-reduce and the fused AdamW consume.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

D_STRIDES = (2, 1, 2, 1, 2, 1, 2)                       # reference model.py:148-183
VGG_PLAN = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512]


def vgg_conv_indices() -> List[int]:
    """torchvision `features` indices of the 16 convs kept by features[:34] (model.py:8)."""
    idx, i = [], 0
    for v in VGG_PLAN:
        if v == "M":
            i += 1
        else:
            idx.append(i)
            i += 2
    return idx


class ZeroArena:
    """InstanceNorm statistics buffers ([N,C,2] int64, accumulated with atomics by the conv epilogues) must start at zero.
    One `torch.zeros` per conv was 38 fill launches per step; the arena hands out slices of ONE buffer that the step zeroes
    once.  Sizes are learned during the first (eager) step, which falls back to individual allocations."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.off = 0
        self.need = 0

    def reset(self, device):
        if self.buf is None or self.buf.numel() < self.need:
            self.buf = torch.zeros(max(self.need, 1), dtype=torch.int64, device=device) if self.need else None
        elif self.off:
            self.buf[:self.off].zero_()
        self.off = 0

    def take(self, N: int, C: int, device) -> torch.Tensor:
        n = N * C * 2
        if self.buf is None or self.off + n > self.buf.numel():
            self.need = max(self.need, self.off + n)
            self.off += n
            return torch.zeros((N, C, 2), dtype=torch.int64, device=device)
        out = self.buf[self.off:self.off + n].view(N, C, 2)
        self.off += n
        self.need = max(self.need, self.off)
        return out


class FlatParams:
    """All parameters of a module re-pointed into one flat fp32 buffer (+ flat grad / Adam moments)."""

    def __init__(self, module: torch.nn.Module, with_optimizer: bool = True):
        params = [(n, p) for n, p in module.named_parameters()]
        self.names = [n for n, _ in params]
        dev = params[0][1].device
        # 4-element alignment per tensor so every view is 16-byte aligned
        self.offsets, off = {}, 0
        for n, p in params:
            self.offsets[n] = off
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.p: Dict[str, torch.Tensor] = {}
        self.g: Dict[str, torch.Tensor] = {}
        for n, p in params:
            o = self.offsets[n]
            view = self.flat[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view                                  # the module now aliases the flat buffer
            self.p[n] = view
            self.g[n] = self.grad[o:o + p.numel()].view_as(p)
        if with_optimizer:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)     # device-side step counter (graph safe)
        self.step_count = 0
        self.version = 0
        self.ext_version = 0                               # bumped when the parameters are changed from OUTSIDE the engine
        module._fsr_flat = self                            # lets the module's own forward reuse this aliasing

    def aliases(self, module) -> bool:
        n, p = next(iter(module.named_parameters()))
        return p.data_ptr() == self.p[n].data_ptr()

    def zero_grad(self):
        self.grad.zero_()

    def optimizer_state(self, lr: float = 1e-4) -> Dict[str, object]:
        """AdamW state in **torch.optim.AdamW.state_dict() format** (what trainer.py:149-156 and the pretrain files of
        trainer.py:131-141 hold): per-parameter {step, exp_avg, exp_avg_sq} in `module.parameters()` order plus one
        param_group with the hyper-parameters of trainer.py:33-38 - loadable by the reference's optimizer and vice versa."""
        state = {}
        for i, n in enumerate(self.names):
            o, numel = self.offsets[n], self.p[n].numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.m[o:o + numel].view_as(self.p[n]).detach().clone(),
                        "exp_avg_sq": self.v[o:o + numel].view_as(self.p[n]).detach().clone()}
        group = {"lr": lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 1e-2, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": True,
                 "params": list(range(len(self.names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state(self, state: Dict[str, object]):
        """Inverse of optimizer_state(); accepts a genuine torch.optim.AdamW state_dict of the same architecture
        (resume, trainer.py:90-94).  An empty `state` (optimizer that never stepped) resets the moments."""
        st = state["state"]
        if len(st) not in (0, len(self.names)):
            raise RuntimeError(f"optimizer state holds {len(st)} parameters, this network has {len(self.names)}")
        self.m.zero_()
        self.v.zero_()
        step = 0
        for i, n in enumerate(self.names):
            if i not in st:
                continue
            o, numel = self.offsets[n], self.p[n].numel()
            if tuple(st[i]["exp_avg"].shape) != tuple(self.p[n].shape):
                raise RuntimeError(f"optimizer state of parameter {i} ({n}) has shape {tuple(st[i]['exp_avg'].shape)}, "
                                   f"expected {tuple(self.p[n].shape)}")
            self.m[o:o + numel].copy_(st[i]["exp_avg"].reshape(-1).to(self.m.device, torch.float32))
            self.v[o:o + numel].copy_(st[i]["exp_avg_sq"].reshape(-1).to(self.v.device, torch.float32))
            step = int(float(st[i]["step"]))
        self.step_count = step
        self.step_dev.fill_(step)                          # the device-side counter the fused AdamW kernel increments

    def adamw_step(self, lr: float, grad_scale: float = 1.0):
        """torch.optim.AdamW defaults of trainer.py:33-38 (betas .9/.999, eps 1e-8, weight_decay 1e-2)."""
        self.step_count += 1
        ops.adamw_dev(self.flat, self.grad, self.m, self.v, lr, self.step_dev, grad_scale=grad_scale)
        self.version += 1


# ====================================================================================== Generator
class GeneratorNet:
    """model.py:72-117 forward (saving activations) and backward.  n_filters = 64 (the reference default) runs on the
    resident-weight 64-channel kernels; n_filters = 128, 192, ... (multiples of 64; BASELINE config #5) on the general
    convs (CTA-pair kernel for the 128-wide slices)."""

    def __init__(self, module, fp: FlatParams, dtype: torch.dtype):
        if module.n_filters % 64:
            raise RuntimeError("the training engine (backward kernels) is built for generator.n_filters in {64, 128, 192, ...}; "
                               "other widths (e.g. 32) are supported for inference only")
        self.m, self.fp, self.dt = module, fp, dtype
        self.L = module.n_layers
        self.F = module.n_filters
        self.c64 = self.F == 64
        self._packed_version = -1
        self._bwd_version = -1
        self.P: Dict[str, torch.Tensor] = {}
        self.arena: Optional[ZeroArena] = None             # set by GANEngine: statistics buffers zeroed once per step

    def _stats(self, x, cout=None):
        return self.arena.take(x.shape[0], cout or self.F, x.device) if self.arena is not None else None

    def _convs64(self):
        names = []
        for i in range(self.L):
            names += [f"stem.{i}.conv1.weight", f"stem.{i}.conv2.weight"]
        return names + ["bottleneck.0.weight"]

    def _plans(self):
        """The fixed pack lists (forward / data-gradient layouts): destination buffers are allocated ONCE - a captured CUDA
        graph keeps reading the right memory - and each list re-packs in ONE launch (ops.PackPlan / fsr_pack_multi)."""
        if getattr(self, "_plan_fwd", None) is None:
            p, P, dt = self.fp.p, self.P, self.dt
            fwd, bwd = ops.PackPlan(dt), ops.PackPlan(dt)
            for n in self._convs64():
                P[n], _ = fwd.add(p[n])
                # data-gradient pack: the 64-channel kernel takes flipped taps (forward tap table), the general kernel its own
                # dgrad tap table (mode 1) on the unflipped transposed pack
                P[n + ".t"], _ = bwd.add(p[n], transposed=True, flip=self.c64)
            for i in range(2):
                w, b = p[f"upsampling.{i}.conv.weight"], p[f"upsampling.{i}.conv.bias"]
                P[f"up{i}.w"], P[f"up{i}.b"] = fwd.add(w, b, ps_perm=True)
                P[f"up{i}.t"], _ = bwd.add(w, transposed=True, ps_perm=True)                      # gen-kernel dgrad
            P["head.w"], P["head.b"] = fwd.add(p["head.0.weight"], p["head.0.bias"], pad=16)
            P["head.t"] = torch.empty((self.F, 3, 3, 3), dtype=torch.float32, device=p["head.0.weight"].device)
            self._plan_fwd, self._plan_bwd = fwd, bwd
        return self._plan_fwd, self._plan_bwd

    def pack(self, need_bwd: bool, force: bool = False):
        """(Re)pack into PERSISTENT buffers (same addresses for the lifetime of the net): a captured CUDA graph keeps
        reading the right memory, and no allocation happens per step.  force: re-pack whatever the version counters
        say (the step calls it right after the point where the parameters change, so that the pack kernels are part of
        every captured graph, not only of those captured while the Python-side version happened to be stale)."""
        if not force and self._packed_version == self.fp.version and (not need_bwd or self._bwd_version == self.fp.version):
            return
        fwd, bwd = self._plans()
        dev = self.fp.flat.device
        if force or self._packed_version != self.fp.version:
            fwd.run(dev)
        if need_bwd:
            bwd.run(dev)
            # head dgrad = direct 3->F conv with transposed, flipped weights (K = 27)
            self.P["head.t"].copy_(self.fp.p["head.0.weight"].permute(1, 0, 2, 3).flip(2, 3))
            self._bwd_version = self.fp.version
        self._packed_version = self.fp.version

    # ---- the width-dependent kernels
    def _conv_raw(self, x, w):
        if self.c64:
            return ops.conv3x3_c64_raw_stats(x, w, stats=self._stats(x))
        return ops.conv3x3_gen(x, w, self.F, epilogue=L.EPI_RAW_STATS, stats=self._stats(x))

    def _conv_up(self, x, i):
        p, P = self.fp.p, self.P
        if self.c64:
            return ops.conv3x3_c64_ps_prelu(x, P[f"up{i}.w"], P[f"up{i}.b"], p[f"upsampling.{i}.relu.weight"])
        return ops.conv3x3_gen(x, P[f"up{i}.w"], 4 * self.F, epilogue=L.EPI_PS_PRELU, bias=P[f"up{i}.b"], alpha=p[f"upsampling.{i}.relu.weight"])

    def _dgrad(self, dy, name):
        if self.c64:
            return ops.conv3x3_c64_bias_act(dy, self.P[name + ".t"], None)
        return ops.conv3x3_gen(dy, self.P[name + ".t"], self.F, mode=1)

    def forward(self, lr_img: torch.Tensor, save: bool, out: Optional[torch.Tensor] = None):
        """out: optional fp32 NCHW [N,3,4h,4w] destination of sr (a slice of the step's [sr; hr] image batch).
        save: the inputs of the 2L+1 F->F convs are written into ONE arena [2L+1][N,h,w,F] (slot 2i = input of block i's
        conv1, 2i+1 = input of its conv2, 2L = input of the bottleneck conv) so that their weight gradients run as a
        single grouped launch in backward()."""
        self.pack(need_bwd=save)
        p, P, dt, Fm = self.fp.p, self.P, self.dt, self.F
        N, _, h, w = lr_img.shape
        Lb = self.L
        xa = torch.empty((2 * Lb + 1, N, h, w, Fm), dtype=dt, device=lr_img.device) if save else None
        slot = (lambda k: xa[k]) if save else (lambda k: None)
        a0 = ops.neck_conv3x3(lr_img, p["neck.0.weight"], p["neck.0.bias"], dt, act=L.ACT_PRELU, alpha=p["neck.1.weight"], out=slot(0))
        cur, blocks = a0, []
        for i in range(Lb):
            raw1, st1 = self._conv_raw(cur, P[f"stem.{i}.conv1.weight"])
            y1 = ops.instnorm_apply(raw1, st1, act=L.ACT_PRELU, alpha=p[f"stem.{i}.relu1.weight"], out=slot(2 * i + 1))
            raw2, st2 = self._conv_raw(y1, P[f"stem.{i}.conv2.weight"])
            nxt = ops.instnorm_apply(raw2, st2, residual=cur, out=slot(2 * i + 2))
            if save:
                blocks.append((raw1, st1, raw2, st2))
            cur = nxt
        rawb, stb = self._conv_raw(cur, P["bottleneck.0.weight"])
        xb = ops.instnorm_apply(rawb, stb, residual=a0)
        U0 = self._conv_up(xb, 0)
        U1 = self._conv_up(U0, 1)
        if self.c64:
            sr = ops.conv3x3_c64_head(U1, P["head.w"], P["head.b"], out=out)
        else:
            sr = ops.conv3x3_head(U1, P["head.w"], P["head.b"], out_mode=0, out=out)
        ctx = dict(lr=lr_img, a0=a0, xa=xa, blocks=blocks, rawb=rawb, stb=stb, xb=xb, U0=U0, U1=U1, sr=sr) if save else None
        return sr, ctx

    def backward(self, ctx, d_sr: torch.Tensor, side: Optional[torch.cuda.Stream] = None):
        """d_sr: fp32 NCHW gradient w.r.t. the generator output; accumulates into fp.g.
        side: a stream for the weight / bias gradients.  They are off the data-gradient chain's critical path: each is issued
        on `side` behind an event recorded where its operands are complete, so it fills the SMs the chain's small launches
        leave idle; the current stream waits for `side` before returning.  (All weight-gradient launches then share `side`,
        hence the single partial-sum workspace stays race-free.)"""
        p, g, P, dt, Fm = self.fp.p, self.fp.g, self.P, self.dt, self.F
        Lb = self.L
        main = torch.cuda.current_stream() if side is not None else None
        keep = []                                             # operands of side-stream launches stay referenced until the join

        def off(fn, *operands):
            if side is None:
                fn()
                return
            keep.append(operands)
            side.wait_event(main.record_event())
            with torch.cuda.stream(side):
                fn()
        dpre = ops.tanh_bwd(ctx["sr"], d_sr)                                           # model.py:109
        off(lambda: (ops.wgrad_c3(dpre, ctx["U1"], g["head.0.weight"], flip=True, layout=1),
                     ops.bias_grad_nchw(dpre, g["head.0.bias"])), dpre)
        dU = ops.neck_conv3x3(dpre, P["head.t"], None, dt, act=L.ACT_NONE)              # dgrad of model.py:103-108
        for i, (U, xin) in ((1, (ctx["U1"], ctx["U0"])), (0, (ctx["U0"], ctx["xb"]))):   # model.py:39-40
            dconv = ops.ps_prelu_bwd(U, dU, p[f"upsampling.{i}.relu.weight"], g[f"upsampling.{i}.relu.weight"])
            off(lambda xin=xin, dconv=dconv, i=i: (ops.conv3x3_wgrad(xin, dconv, g[f"upsampling.{i}.conv.weight"], ps_perm=True),
                                                   ops.bias_grad(dconv, g[f"upsampling.{i}.conv.bias"], ps_perm=True)), xin, dconv)
            dU = ops.conv3x3_gen(dconv, P[f"up{i}.t"], Fm, mode=1)
        dxb = dU
        # output gradients of the 2L+1 F->F convs, same slot order as the input arena of forward()
        xa = ctx["xa"]
        da = torch.empty_like(xa)
        ops.instnorm_bwd(ctx["rawb"], ctx["stb"], dxb, out=da[2 * Lb])                  # model.py:94
        dcur = self._dgrad(da[2 * Lb], "bottleneck.0.weight")
        # all 2L+1 weight gradients of the residual chain: grouped tcgen05 launches (<= 148 (group, cin, cout) pairs each);
        # with a side stream the upper half of the chain is launched as soon as its output gradients exist
        names = self._convs64()
        per = max(1, 148 // ((Fm // 64) ** 2))

        def wgrad_slots(k0, k1):
            for a in range(k0, k1, per):
                b = min(a + per, k1)
                ops.conv3x3_wgrad_grouped(xa[a:b], da[a:b], [g[n] for n in names[a:b]])
        half = Lb // 2 if side is not None else 0                                       # blocks >= half: slots 2*half .. 2L
        for i in reversed(range(Lb)):                                                   # model.py:67-69
            raw1, st1, raw2, st2 = ctx["blocks"][i]
            ops.instnorm_bwd(raw2, st2, dcur, out=da[2 * i + 1])
            dy1 = self._dgrad(da[2 * i + 1], f"stem.{i}.conv2.weight")
            ops.instnorm_bwd(raw1, st1, dy1, act=L.ACT_PRELU, alpha=p[f"stem.{i}.relu1.weight"],
                             dalpha=g[f"stem.{i}.relu1.weight"], out=da[2 * i])
            din = self._dgrad(da[2 * i], f"stem.{i}.conv1.weight")
            dcur = ops.add(din, dcur)                                                   # + skip (model.py:69)
            if i == half and half > 0:
                off(lambda: wgrad_slots(2 * half, 2 * Lb + 1))
        off(lambda: wgrad_slots(0, 2 * half if half > 0 else 2 * Lb + 1))
        da0 = ops.add(dcur, dxb)                                                        # + long skip (model.py:115)
        dv = ops.act_bwd(ctx["a0"], da0, L.ACT_PRELU, alpha=p["neck.1.weight"], dalpha=g["neck.1.weight"])
        off(lambda: (ops.wgrad_c3(ctx["lr"], dv, g["neck.0.weight"], flip=False, layout=2),           # model.py:76
                     ops.bias_grad(dv, g["neck.0.bias"])), dv)
        if side is not None:
            main.wait_stream(side)
        keep.clear()


# ====================================================================================== Discriminator
class DiscriminatorNet:
    """model.py:139-193 forward (saving activations) and backward."""

    def __init__(self, module, fp: FlatParams, dtype: torch.dtype):
        self.m, self.fp, self.dt = module, fp, dtype
        F_ = module.n_filters
        self.widths = [(F_, F_), (F_, 2 * F_), (2 * F_, 2 * F_), (2 * F_, 4 * F_), (4 * F_, 4 * F_), (4 * F_, 8 * F_), (8 * F_, 8 * F_)]
        self._packed_version = -1
        self._bwd_version = -1
        self.P: Dict[str, torch.Tensor] = {}
        self.arena: Optional[ZeroArena] = None

    def pack(self, need_bwd: bool, force: bool = False):
        """(Re)pack into persistent buffers, one launch per layout (CUDA-graph safe, see GeneratorNet.pack)."""
        if not force and self._packed_version == self.fp.version and (not need_bwd or self._bwd_version == self.fp.version):
            return
        p, P, dt = self.fp.p, self.P, self.dt
        if getattr(self, "_plan_fwd", None) is None:
            fwd, bwd = ops.PackPlan(dt), ops.PackPlan(dt)
            for i in range(7):
                w = p[f"stem.{i}.conv.weight"]
                P[f"w{i}"], _ = fwd.add(w)
                P[f"t{i}"], _ = bwd.add(w, transposed=True)
            P["neck.t"], _ = bwd.add(p["neck.0.weight"], transposed=True, flip=True, pad=16)        # 64 -> 3 image gradient
            self._plan_fwd, self._plan_bwd = fwd, bwd
        dev = self.fp.flat.device
        if force or self._packed_version != self.fp.version:
            self._plan_fwd.run(dev)
        if need_bwd:
            self._plan_bwd.run(dev)
            self._bwd_version = self.fp.version
        self._packed_version = self.fp.version

    def forward(self, img: torch.Tensor, save: bool):
        self.pack(need_bwd=save)
        p, P, dt = self.fp.p, self.P, self.dt
        d0 = ops.neck_conv3x3(img, p["neck.0.weight"], p["neck.0.bias"], dt, act=L.ACT_LRELU, slope=0.2)
        cur, layers = d0, []
        cur_is_parity = False
        for i, s in enumerate(D_STRIDES):
            cout = self.widths[i][1]
            xin = cur if (s == 1 or cur_is_parity) else ops.parity_layout(cur, True)
            raw, st = ops.conv3x3_gen(xin, P[f"w{i}"], cout, stride=s, epilogue=L.EPI_RAW_STATS,
                                      stats=self.arena.take(img.shape[0], cout, img.device) if self.arena is not None else None)
            # the next stride-2 block reads its input in parity-plane layout: written directly by the normalise pass
            cur_is_parity = i + 1 < len(D_STRIDES) and D_STRIDES[i + 1] == 2
            if cur_is_parity:
                act = ops.instnorm_apply_parity(raw, st, act=L.ACT_LRELU, slope=0.01)
            else:
                act = ops.instnorm_apply(raw, st, act=L.ACT_LRELU, slope=0.01)
            if save:
                layers.append((xin, raw, st))
            cur = act
        z = ops.conv1x1_to1_fwd(cur, p["stem.7.weight"].view(-1), p["stem.7.bias"])
        ctx = dict(img=img, d0=d0, layers=layers, act6=cur) if save else None
        return z, ctx

    def backward(self, ctx, dz: torch.Tensor, wgrad: bool, d_img: Optional[torch.Tensor], defer: Optional[list] = None):
        """dz fp32 [N,6,6].  wgrad: accumulate parameter gradients (D step).  d_img: fp32 NCHW image-gradient
        accumulator (G step; the wasted D weight gradients of trainer.py:195 are skipped - never consumed).
        defer: a list - the weight / bias gradient launches (off the data-gradient chain's critical path) are not issued
        but appended as closures; the caller runs them later, e.g. on a side stream (GANEngine._seg_d_update)."""
        p, g, P = self.fp.p, self.fp.g, self.P

        def later(fn):
            if defer is None:
                fn()
            else:
                defer.append(fn)
        dcur = ops.conv1x1_to1_bwd(ctx["act6"], p["stem.7.weight"].view(-1), dz,
                                   g["stem.7.weight"].view(-1) if wgrad else None, g["stem.7.bias"] if wgrad else None)
        dcur_parity = False
        for i in reversed(range(7)):
            xin, raw, st = ctx["layers"][i]
            s = D_STRIDES[i]
            if dcur_parity:      # gradient from a stride-2 data gradient: read in its parity-plane layout (no re-layout pass)
                draw = ops.instnorm_bwd_parity(raw, st, dcur, act=L.ACT_LRELU, slope=0.01)
            else:
                draw = ops.instnorm_bwd(raw, st, dcur, act=L.ACT_LRELU, slope=0.01)
            if wgrad:
                later(lambda xin=xin, draw=draw, i=i, s=s: ops.conv3x3_wgrad(xin, draw, g[f"stem.{i}.conv.weight"], stride=s))
            if i == 0 and not (wgrad or d_img is not None):
                break
            dcur = ops.conv3x3_gen(draw, P[f"t{i}"], self.widths[i][0], stride=s, mode=1)
            dcur_parity = s == 2 and i > 0 and raw.shape[1] * raw.shape[2] * 4 <= 4096    # consumer plane = 4x this layer's output
            if s == 2 and not dcur_parity:
                dcur = ops.parity_layout(dcur, False)
        dv = ops.act_bwd(ctx["d0"], dcur, L.ACT_LRELU, slope=0.2)
        if wgrad:
            img = ctx["img"]
            later(lambda: ops.wgrad_c3(img, dv, g["neck.0.weight"], flip=False, layout=2))
            later(lambda: ops.bias_grad(dv, g["neck.0.bias"]))
        if d_img is not None:
            ops.conv3x3_c64_head(dv, P["neck.t"], None, out_u8=3, out=d_img)


# ====================================================================================== VGG19[:34]
class VGGNet:
    """model.py:5-23: renorm + vgg19.features[:34] (frozen).  forward + data gradient only."""

    def __init__(self, module, dtype: torch.dtype):
        self.m, self.dt = module, dtype
        self.idx = vgg_conv_indices()
        self.P: Dict[str, torch.Tensor] = {}
        self._packed = False
        self._version = -1

    def pack(self, need_bwd: bool):
        ver = getattr(self.m, "_weights_version", 0)
        if ver != self._version:                     # perceptual_network.load_state_dict() after the first step: repack
            self.P.clear()
            self._packed, self._version = False, ver
        if self._packed and (not need_bwd or "bwd" in self.P):
            return
        sd = {k: v for k, v in self.m.state_dict().items()}
        P, dt = self.P, self.dt
        for j, i in enumerate(self.idx):
            w, b = sd[f"vgg.{i}.weight"], sd[f"vgg.{i}.bias"]
            if j == 0:
                P["w0"], P["b0"] = w.float().contiguous(), b.float().contiguous()
                if need_bwd:
                    scale = (0.5 / sd["std"].view(-1)).float().contiguous()          # d/dx of ((x+1)/2 - mean)/std
                    P["t0"] = ops.pack_conv3x3_t(w, dt, flip=True, row_pad=16, row_scale=scale)
            else:
                if f"w{j}" not in P:
                    P[f"w{j}"], P[f"b{j}"] = ops.pack_conv3x3(w, b, dt)
                if need_bwd:
                    P[f"t{j}"] = ops.pack_conv3x3_t(w, dt)
        if need_bwd:
            P["bwd"] = torch.empty(0)
        self._packed = True

    FLAT_MAX_W = 25          # conv3x3_gen_flat: the padded-flattened tile mapping needs W + 2 <= 27 rows of lead / tail per box

    def forward(self, img: torch.Tensor, save: bool):
        """-> (features, ctx).  From the first pooled resolution with W <= 25 on (12x12 and 6x6 at the training shape:
        conv4_x, conv5_x) activations are kept in the zero-bordered PADDED layout [N,H+2,W+2,C] and the convs run in the
        flat tile mapping (fsr_conv3x3_gen_flat); `self.feat_pad` tells whether the returned features are padded."""
        self.pack(need_bwd=save)
        P, dt = self.P, self.dt
        acts, j = [], 0
        cur = None
        pad = False
        use_flat = os.environ.get("FSR_VGG_FLAT", "1") != "0"
        for v in VGG_PLAN:
            if v == "M":
                wo = (cur.shape[2] - (2 if pad else 0)) // 2
                out_pad = use_flat and wo <= self.FLAT_MAX_W and cur.shape[3] % 128 == 0
                pooled = ops.maxpool2_padded(cur, pad, out_pad) if (pad or out_pad) else ops.maxpool2(cur)
                if save:
                    acts[-1] = (acts[-1][0], True, acts[-1][2])
                cur, pad = pooled, out_pad
            else:
                if j == 0:
                    cur = ops.neck_conv3x3(img, P["w0"], P["b0"], dt, act=L.ACT_RELU, vgg_norm=True)
                elif pad:
                    cur = ops.conv3x3_gen_flat(cur, P[f"w{j}"], v, mode=0, bias=P[f"b{j}"], act=L.ACT_RELU)
                else:
                    cur = ops.conv3x3_gen(cur, P[f"w{j}"], v, bias=P[f"b{j}"], act=L.ACT_RELU)
                if save:
                    acts.append((cur, False, pad))
                j += 1
        self.feat_pad = pad
        return cur, (dict(acts=acts) if save else None)

    def backward(self, ctx, dfeat: torch.Tensor, d_img: torch.Tensor):
        """dfeat: gradient w.r.t. relu5_3 features (NHWC dtype; padded iff the forward returned padded features) of the
        FIRST dfeat.shape[0] images of the forward batch (the step runs VGG once on [sr; hr] and differentiates the sr half
        only); accumulates the image gradient into d_img (fp32 NCHW)."""
        P = self.P
        acts = ctx["acts"]
        widths = [v for v in VGG_PLAN if v != "M"]
        dcur = dfeat
        nb = dfeat.shape[0]
        dcur_pad = acts[-1][2]
        for j in reversed(range(len(acts))):
            a, pooled, a_pad = acts[j]
            a = a[:nb]
            if pooled:
                da = ops.maxpool2_relu_bwd_padded(a, dcur, a_pad, dcur_pad) if (a_pad or dcur_pad) else ops.maxpool2_relu_bwd(a, dcur)
            else:
                da = ops.relu_bwd(a, dcur)
            if j == 0:
                ops.conv3x3_c64_head(da, P["t0"], None, out_u8=3, out=d_img)
            elif a_pad:
                dcur = ops.conv3x3_gen_flat(da, P[f"t{j}"], widths[j - 1], mode=1)
            else:
                dcur = ops.conv3x3_gen(da, P[f"t{j}"], widths[j - 1], mode=1)
            dcur_pad = a_pad


# ====================================================================================== GAN step
class GANEngine:
    """One iteration of trainer.py:168-196 (`train_step`) and of the pre-training loop (:104-111).

    Schedule of a step (same results as the reference's order; what differs is only WHEN things run):
      * `G(lr)` is evaluated ONCE.  The reference evaluates it at :173 (detached, for the discriminator step) and again
        at :185; between the two only the discriminator is updated, the forward is deterministic and InstanceNorm keeps
        no running statistics, so both evaluations are the same tensor bit for bit (asserted in
        tests/test_train_step_gpu.py::test_generator_forward_is_identical_before_and_after_the_d_step).
      * two streams.  After G(lr) the SIDE stream takes everything that touches the discriminator - D([sr; hr]) forward
        and backward with its weight gradients (:172-180), the gradient exchange (all-reduce over ranks), AdamW (:181),
        the re-pack of its weights, then D(sr) through the UPDATED discriminator and its data gradient (:186-188) -
        while the MAIN stream runs the two VGG19 passes, the content loss and the VGG data gradient (:190-192), none of
        which touch the discriminator (SURVEY.md 8e "legal overlap windows").  They join before the generator's
        backward (:195), whose weight / bias gradients are again issued on the side stream behind per-operand events
        while the main stream walks the data-gradient chain.  Same arithmetic as one stream: results are bit-identical
        to the single-stream order except that the residual chain's grouped weight gradient is launched in two halves
        (a different, still fixed, split of one sum).  A/B switches: FSR_TRAIN_OVERLAP, FSR_D_SIDE, FSR_ADV_SIDE,
        FSR_G_WGRAD_SIDE.
      * the whole step - collectives included, issued through libfsr_b200's fsr_nccl_* - is captured into ONE CUDA graph
        per input shape after two eager warm-up steps."""

    def __init__(self, generator, discriminator, vgg, lr_g: float, lr_d: float, dtype: torch.dtype = torch.bfloat16,
                 loss_scale: Optional[float] = None, process_group=None):
        self.dt = dtype
        self.gp, self.dp = FlatParams(generator), FlatParams(discriminator)
        self.G = GeneratorNet(generator, self.gp, dtype)
        self.D = DiscriminatorNet(discriminator, self.dp, dtype)
        self.V = VGGNet(vgg, dtype)
        self.lr_g, self.lr_d = lr_g, lr_d
        # fp16 gradients (1/numel-scaled losses) would underflow: static loss scale; bf16 needs none
        self.S = float(loss_scale) if loss_scale is not None else (4096.0 if dtype == torch.float16 else 1.0)
        self.pg = process_group
        import os
        self.use_graph = os.environ.get("FSR_GRAPH", "1") != "0"
        self.defer_d_wgrad = os.environ.get("FSR_DEFER_D_WGRAD", "1") != "0"
        self.adv_on_side = os.environ.get("FSR_ADV_SIDE", "1") != "0"
        self.d_on_side = os.environ.get("FSR_D_SIDE", "1") != "0"
        self.g_wgrad_side = os.environ.get("FSR_G_WGRAD_SIDE", "1") != "0"
        self._d_sr_adv = None
        self._d_deferred = None
        self.overlap = os.environ.get("FSR_TRAIN_OVERLAP", "1") != "0"
        self._graphs: Dict = {}
        self._side: Optional[torch.cuda.Stream] = None
        self.arena = ZeroArena()
        self.G.arena = self.D.arena = self.arena
        self.comm = None                                    # distributed.FlatComm: NCCL through the C ABI
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        if self.world > 1:
            from .distributed import FlatComm
            self.comm = FlatComm(self.gp.flat.device, process_group)
            # replicas must START identical (they stay identical because every rank applies the same reduced gradient):
            # rank 0's parameters and Adam moments win, whatever each rank's RNG produced at construction
            for fp in (self.gp, self.dp):
                for buf in (fp.flat, fp.m, fp.v):
                    self.comm.broadcast(buf)
                fp.version += 1
            if not self.comm.native:
                # torch.distributed fallback (FSR_NCCL_CAPI=0 / no libnccl): ProcessGroupNCCL collectives are not captured
                # into the step graph (measured: capture on a side stream hangs) - the step runs eagerly instead
                self.use_graph = False

    def _allreduce(self, flat_grad: torch.Tensor):
        if self.world > 1:
            self.comm.allreduce(flat_grad)                  # NCCL sum over NVLink on the current stream; 1/world is folded into AdamW

    # ------------------------------------------------------------------ public steps
    def train_step(self, lr_img: torch.Tensor, hr_img: torch.Tensor, noise: Dict[str, torch.Tensor]):
        """lr_img [B,3,h,w], hr_img [B,3,4h,4w] fp32 NCHW in [-1,1] (this rank's shard);
        noise = {"d_real","d_fake","g_real"}: uniform [0,1) tensors shaped like D's output (trainer.py:175,176,187).

        With use_graph (default) the ~450 launches of a step, both gradient all-reduces included, are captured ONCE per
        input shape into one CUDA graph (after two eager warm-up steps) and replayed; inputs are copied into static
        buffers; the AdamW step counters live in device memory."""
        lr_img, hr_img = lr_img.contiguous().float(), hr_img.contiguous().float()
        B = lr_img.shape[0]
        ins = (lr_img, hr_img, noise["d_real"].reshape(B, -1).contiguous().float(),
               noise["d_fake"].reshape(B, -1).contiguous().float(), noise["g_real"].reshape(B, -1).contiguous().float())
        if not self.use_graph:
            return self._step(ins)
        key = (tuple(lr_img.shape), tuple(hr_img.shape))
        ext = (self.gp.ext_version, self.dp.ext_version, getattr(self.V.m, "_weights_version", 0))
        if ext != getattr(self, "_ext_seen", ext):
            # load_state_dict / load_checkpoints since the capture: captured graphs only re-pack what the step itself
            # changes (and VGG packs are re-allocated) -> drop them; the next two steps run eagerly, then re-capture
            self._graphs.clear()
        self._ext_seen = ext
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = dict(calls=0, graph=None)
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= 2:                                        # eager warm-up: allocator, func attributes, packs
                return self._step(ins)
            st["inputs"] = [t.clone() for t in ins]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                                   # capture only: nothing executes here
                self._step(st["inputs"])
            st["graph"], st["out"] = g, self._out
            for fp in (self.gp, self.dp):
                fp.step_count -= 1                                      # undo the bookkeeping of the (non-executing) capture
        for dst, src in zip(st["inputs"], ins):
            dst.copy_(src, non_blocking=True)
        st["graph"].replay()
        for fp in (self.gp, self.dp):
            fp.step_count += 1
            fp.version += 1
        self.G.m._packed_key = None
        return st["out"]

    def _step(self, ins):
        """The whole iteration on the current stream (+ the side stream of the overlap window)."""
        d_side = self.overlap and self.d_on_side
        self._seg_gfwd(ins)
        if not d_side:
            self._seg_dstep(ins, defer=self.overlap and self.defer_d_wgrad)
        main = torch.cuda.current_stream()
        if self.overlap:
            if self._side is None:
                self._side = torch.cuda.Stream()
            side = self._side
            side.wait_stream(main)                                      # sr = G(lr) (d_on_side) / the D gradients are complete
            with torch.cuda.stream(side):
                if d_side:
                    self._seg_dstep(ins, defer=False)                   # the whole discriminator step beside the VGG passes
                self._seg_d_update()
                if self.adv_on_side:
                    self._seg_adv(ins)                                  # D(sr) through the updated D: independent of VGG too
            self._seg_content(ins)                                      # VGG passes: independent of the discriminator
            main.wait_stream(side)                                      # updated + re-packed D before D(sr) at :186
            self._d_deferred = None                                     # the side stream's operands may be recycled now
            if not self.adv_on_side:
                self._seg_adv(ins)
        else:
            self._seg_d_update()
            self._seg_content(ins)
            self._seg_adv(ins)
        self._seg_g(ins)
        self._allreduce(self.gp.grad)
        self._seg_opt(ins)
        return self._out

    def _run_segments(self, ins, graphs=None):                          # kept for diagnostics / older callers
        return self._step(ins)

    # ------------------------------------------------------------------ segments
    def _seg_d(self, ins, defer: bool = False):                         # diagnostics / tests: both halves on this stream
        self._seg_gfwd(ins)
        self._seg_dstep(ins, defer)

    def _seg_gfwd(self, ins):
        """G(lr) (saved for the generator step) into the first half of the image batch X = [sr; hr] (trainer.py:173 == :185)."""
        lr_img, hr_img, n_real, n_fake, _ = ins
        S = self.S
        self._losses = losses = torch.zeros(4, dtype=torch.float32, device=lr_img.device)   # real, fake, adv bce, content sum
        self.dp.zero_grad()
        self.gp.zero_grad()
        self.arena.reset(lr_img.device)                                 # every InstanceNorm statistics buffer of the step
        self.G.pack(need_bwd=True, force=True)                          # G changed at the end of the previous step
        # ONE image batch X = [sr; hr] feeds the discriminator step and the VGG pass: D(hr) and D(sr) share their weights
        # (:172, :174) and so do VGG(sr) and VGG(hr) (:190-191), every op is per-sample (InstanceNorm included), so one
        # launch over 2B images is the same arithmetic as two launches over B - with half the launches
        B = lr_img.shape[0]
        X = torch.empty((2 * B,) + tuple(hr_img.shape[1:]), dtype=torch.float32, device=hr_img.device)
        self._X = X
        self._sr, self._ctx_g = self.G.forward(lr_img, save=True, out=X[:B])      # :173 == :185 (see class docstring)
        X[B:].copy_(hr_img)

    def _seg_dstep(self, ins, defer: bool = False):
        """the discriminator step up to its gradient (trainer.py:171-180) on X = [sr; hr]."""
        _, _, n_real, n_fake, _ = ins
        S, losses, X = self.S, self._losses, self._X
        B = X.shape[0] // 2
        z, ctx = self.D.forward(X, save=True)                           # :172 (second half) and :174 (first half)
        dz = torch.empty_like(z)
        ops.bce_logits(z[B:], n_real, 0.3, 0.8, losses[0:1], dz[B:], grad_scale=0.5 * S)     # :175,177,179
        ops.bce_logits(z[:B], n_fake, 0.3, 0.0, losses[1:2], dz[:B], grad_scale=0.5 * S)     # :176,178,179
        # :180 - the weight gradients are off the data-gradient chain: they are issued in _seg_d_update (side stream of the
        # overlap window), where they fill the SMs the small VGG layers leave idle
        self._d_deferred = [] if defer else None
        self.D.backward(ctx, dz, wgrad=True, d_img=None, defer=self._d_deferred)

    def _seg_d_update(self):
        """discriminator weight gradients (deferred from _seg_d) + gradient exchange + discriminator AdamW (trainer.py:181)
        + re-pack of its weights (side stream)."""
        for fn in (self._d_deferred or ()):
            fn()
        self._allreduce(self.dp.grad)
        self.dp.adamw_step(self.lr_d, grad_scale=1.0 / (self.S * self.world))
        self.D.pack(need_bwd=True, force=True)

    def _seg_content(self, ins):
        """content loss and its gradient w.r.t. sr (trainer.py:190-192 and their part of :195)."""
        S, losses, sr = self.S, self._losses, self._sr
        B = sr.shape[0]
        feats, ctx_v = self.V.forward(self._X, save=True)               # :190 (first half) and :191 (second half)
        fake_f, real_f = feats[:B], feats[B:]
        # padded features (flat tile mapping of the <= 12x12 layers): both halves carry the same zero border, which adds
        # nothing to the SmoothL1 sum or its gradient; the mean is taken over the REAL element count
        pd = 2 if self.V.feat_pad else 0
        nfeat = fake_f.shape[0] * (fake_f.shape[1] - pd) * (fake_f.shape[2] - pd) * fake_f.shape[3]
        dfeat = torch.empty_like(fake_f)
        ops.smooth_l1(fake_f, real_f, losses[3:4], dfeat, grad_scale=0.5 * S / nfeat)   # :192,194
        self._d_sr = torch.zeros_like(sr)
        self.V.backward(ctx_v, dfeat, self._d_sr)                       # :195 (VGG branch)
        self._nfeat = float(nfeat)

    def _seg_adv(self, ins):
        """adversarial loss through the UPDATED discriminator and its gradient w.r.t. sr (trainer.py:186-188 and the D branch
        of :195) into its own buffer: it may run on the side stream while the VGG branch fills self._d_sr."""
        n_g = ins[4]
        S, losses, sr = self.S, self._losses, self._sr
        z, ctx_d = self.D.forward(sr, save=True)                        # :186 (updated D)
        dz = torch.empty_like(z)
        ops.bce_logits(z, n_g, 0.3, 0.7, losses[2:3], dz, grad_scale=0.5 * 0.1 * S)          # :187-188,194
        self._d_sr_adv = torch.zeros_like(sr)
        self.D.backward(ctx_d, dz, wgrad=False, d_img=self._d_sr_adv)   # :195 (D branch; its wasted wgrad is skipped)

    def _seg_adv_and_g(self, ins):                                       # diagnostics / tests: both halves on this stream
        self._seg_adv(ins)
        self._seg_g(ins)

    def _seg_g(self, ins):
        """the generator backward from the summed image gradient (trainer.py:195, :194's sum of the two losses)."""
        losses, sr = self._losses, self._sr
        self._d_sr.add_(self._d_sr_adv)                                 # VGG branch + D branch (fp32: a + b as before)
        self.G.backward(self._ctx_g, self._d_sr, side=self._side if (self.overlap and self.g_wgrad_side) else None)
        self._out = dict(loss_real=losses[0], loss_fake=losses[1], adv_loss=0.1 * losses[2], content_loss=losses[3] / self._nfeat, sr=sr)

    def _seg_opt(self, ins):
        self.gp.adamw_step(self.lr_g, grad_scale=1.0 / (self.S * self.world))   # :196
        self.G.m._packed_key = None                                        # the module's inference cache is stale now

    def pretrain_step(self, lr_img: torch.Tensor, hr_img: torch.Tensor):
        """trainer.py:104-111: generator-only SmoothL1 warm-up."""
        S = self.S
        lr_img, hr_img = lr_img.contiguous().float(), hr_img.contiguous().float()
        self.gp.zero_grad()
        self.arena.reset(lr_img.device)
        sr, ctx = self.G.forward(lr_img, save=True)
        loss = torch.zeros(1, dtype=torch.float32, device=lr_img.device)
        d_sr = torch.empty_like(sr)
        ops.smooth_l1(sr, hr_img, loss, d_sr, grad_scale=S / sr.numel())
        self.G.backward(ctx, d_sr)
        self._allreduce(self.gp.grad)
        self.gp.adamw_step(self.lr_g, grad_scale=1.0 / (S * self.world))
        self.G.m._packed_key = None
        return dict(loss=loss[0] / float(sr.numel()), sr=sr)
