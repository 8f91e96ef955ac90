"""Drop-in mirror of the reference's model.py surface on the B200 kernels.

`Generator(config)`, `Discriminator(config)`, `VGG19()` keep the reference constructor
signatures (model.py:73, :140, :6), forward(x: fp32 NCHW) -> fp32 NCHW, and IDENTICAL
state_dict keys/shapes, so reference checkpoints load unchanged
(including the `_orig_mod.` prefix handled like inference.py:31-32).

The arithmetic never touches torch.nn.functional: parameters live in small holder modules and
forward() drives libfsr_b200.so (hand-written sm_100a kernels) through ctypes.  There is no
CPU fallback - a non-CUDA input or a missing library raises.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import _lib as L


def _default_dtype() -> torch.dtype:
    return torch.bfloat16 if os.environ.get("FSR_DTYPE", "fp16").lower() in ("bf16", "bfloat16") else torch.float16


# ---------------------------------------------------------------------------------------------------------------------
# Autograd bridge (SURVEY.md 8b "Callers": trainer.py:110,180,195 call loss.backward() THROUGH the modules).
# forward() saves the engine's activation context, backward() runs the engine's hand-written reverse walk (every FLOP in
# libfsr_b200) and hands the parameter gradients to autograd, so a reference-style loop
#     loss = criterion(discriminator(generator(lr)), labels); loss.backward(); optimizer.step()
# with stock torch.optim.AdamW over module.parameters() runs unmodified.  Trainer.train_step (engine.GANEngine) remains
# the fast path: one CUDA graph, flat AdamW, no per-parameter gradient copies.
class _NetFn(torch.autograd.Function):
    """Generator / Discriminator: inputs (x, *parameters) so that autograd routes gradients to the parameters."""

    @staticmethod
    def forward(ctx, module, x, *params):
        net = module._net_for_autograd()
        net.fp.version += 1                  # an external optimizer may have stepped the (aliased) parameters in place
        y, c = net.forward(x.contiguous().float(), save=True)
        ctx.net, ctx.c, ctx.kind = net, c, module._fsr_kind
        ctx.x_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        net.fp.zero_grad()
        dy = dy.contiguous().float()
        dx = None
        if ctx.kind == "G":
            net.backward(ctx.c, dy)          # image gradient of the generator input is never needed (model.py:75: leaf)
        else:
            if ctx.needs_input_grad[1]:
                dx = torch.zeros(ctx.x_shape, dtype=torch.float32, device=dy.device)
            net.backward(ctx.c, dy, wgrad=True, d_img=dx)
        grads = tuple(net.fp.g[n].clone() for n in net.fp.names)
        return (None, dx) + grads


class _VggFn(torch.autograd.Function):
    """VGG19[:34] is frozen (model.py:9-10): only the image gradient flows back (trainer.py:195 -> sr)."""

    @staticmethod
    def forward(ctx, module, x):
        net = module._engine()
        feat, c = net.forward(x.contiguous().float(), save=x.requires_grad)
        ctx.net, ctx.c, ctx.x_shape, ctx.pad = net, c, tuple(x.shape), net.feat_pad
        from . import ops
        if net.feat_pad:                                   # flat tile mapping: drop the zero border of the padded layout
            feat = feat[:, 1:-1, 1:-1, :].contiguous()
        return ops.nhwc_to_nchw(feat)

    @staticmethod
    def backward(ctx, dfeat):
        from . import ops
        if ctx.c is None:
            return None, None
        dx = torch.zeros(ctx.x_shape, dtype=torch.float32, device=dfeat.device)
        d = ops.nchw_to_nhwc(dfeat.contiguous().float(), ctx.net.dt)
        if ctx.pad:
            dp = torch.zeros((d.shape[0], d.shape[1] + 2, d.shape[2] + 2, d.shape[3]), dtype=d.dtype, device=d.device)
            dp[:, 1:-1, 1:-1, :] = d
            d = dp
        ctx.net.backward(ctx.c, d, dx)
        return None, dx


class _Conv(torch.nn.Module):
    """Parameter holder with torch.nn.Conv2d's names, shapes and default init (kaiming_uniform a=sqrt(5))."""

    def __init__(self, cin: int, cout: int, k: int = 3, bias: bool = True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(cout, cin, k, k))
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(cin * k * k)
            self.bias = torch.nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class _Slope(torch.nn.Module):
    """torch.nn.PReLU() holder: one shared slope, init 0.25 (model.py:37,56,77)."""

    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.full((1,), 0.25))


class _Empty(torch.nn.Module):
    """Parameter-less placeholder keeping Sequential indices aligned with the reference."""


class ResidualBlock(torch.nn.Module):
    """model.py:43-69 - conv1, bn1(IN), relu1(PReLU), conv2, bn2(IN), + x."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = _Conv(in_channels, out_channels, bias=False)
        self.bn1 = _Empty()
        self.relu1 = _Slope()
        self.conv2 = _Conv(in_channels, out_channels, bias=False)
        self.bn2 = _Empty()


class UpSamplingBlock(torch.nn.Module):
    """model.py:26-40 - conv(F->4F), PixelShuffle(2), PReLU."""

    def __init__(self, config):
        super().__init__()
        self.conv = _Conv(config.n_filters, config.n_filters * 4, bias=True)
        self.phase_shift = _Empty()
        self.relu = _Slope()


class SimpleBlock(torch.nn.Module):
    """model.py:120-136 - conv(stride s, no bias), IN, LeakyReLU(0.01)."""

    def __init__(self, in_channels: int, out_channels: int, stride: int):
        super().__init__()
        self.stride = stride
        self.conv = _Conv(in_channels, out_channels, bias=False)
        self.bn = _Empty()
        self.act = _Empty()


def _strip_prefix(state_dict):
    return {k.replace("_orig_mod.", ""): v for k, v in state_dict.items()}


class Generator(torch.nn.Module):
    """Reference model.py:72-117 on B200 kernels.

    forward(x) : fp32 NCHW [N,3,H,W] in [-1,1] -> fp32 NCHW [N,3,4H,4W]   (model.py:112-117)
    compute_dtype: torch.float16 (default: 1e-3 of the reference on random-init weights), torch.bfloat16 (1.5e-2), or
    torch.float32 = precise mode (fp32 storage, split fp16 operands: 1e-3 on the reference's shipped checkpoint too).
    super_resolve_u8(img) : uint8 NHWC -> uint8 NHWC, the inference.py:48-56 pipeline fused into
    the neck load and the head store.
    """

    def __init__(self, config, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.n_filters = int(config.n_filters)
        self.n_layers = int(config.n_layers)
        F_ = self.n_filters
        self.neck = torch.nn.Sequential(_Conv(3, F_, bias=True), _Slope())
        self.stem = torch.nn.Sequential(*[ResidualBlock(F_, F_) for _ in range(self.n_layers)])
        self.bottleneck = torch.nn.Sequential(_Conv(F_, F_, bias=False), _Empty())
        self.upsampling = torch.nn.Sequential(UpSamplingBlock(config), UpSamplingBlock(config))
        self.head = torch.nn.Sequential(_Conv(F_, 3, bias=True), _Empty())
        self.compute_dtype = compute_dtype or _default_dtype()
        self.l2_group = int(os.environ.get("FSR_L2_GROUP", "0"))   # images per L2-resident group (0 = all)
        self._packed: Dict[str, torch.Tensor] = {}
        self._packed_key = None
        self._ws = None
        self._precise = None
        self._pair = None

    # -- checkpoints saved from torch.compile'd modules carry `_orig_mod.` (inference.py:30-33)
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        res = super().load_state_dict(_strip_prefix(state_dict), strict=strict, assign=assign)
        fp = getattr(self, "_fsr_flat", None)
        if fp is not None:                 # a training engine aliases these parameters: its weight packs are stale now
            fp.version += 1
            fp.ext_version += 1
        return res

    # ------------------------------------------------------------------ weight packing
    def _conv_list(self):
        convs = [("neck", self.neck[0])]
        for i, blk in enumerate(self.stem):
            convs += [(f"s{i}a", blk.conv1), (f"s{i}b", blk.conv2)]
        convs += [("bott", self.bottleneck[0]), ("up0", self.upsampling[0].conv), ("up1", self.upsampling[1].conv),
                  ("head", self.head[0])]
        return convs

    def _effective_weights(self):
        """fp32 OIHW tensors padded with zero channels up to a multiple of 64 (the kernels work on 64-channel = 128-byte
        pixel rows).  Zero channels stay exactly zero through conv / InstanceNorm ((0-0)*rsqrt(0+eps)) / PReLU / pixel
        shuffle, so the result equals the unpadded network; for n_filters = 32 this costs 4x the ideal FLOPs (reported
        as such in the config sweep)."""
        F_, Fp = self.n_filters, self.padded_filters

        def pad(t, o, i=None):
            t = t.detach().float()
            shape = list(t.shape)
            shape[0] = o
            if i is not None:
                shape[1] = i
            out = torch.zeros(shape, dtype=torch.float32, device=t.device)
            out[tuple(slice(0, d) for d in t.shape)] = t
            return out

        eff = {}
        for name, conv in self._conv_list():
            w, b = conv.weight, conv.bias
            if name == "neck":
                eff[name] = (pad(w, Fp), pad(b, Fp))
            elif name.startswith("up"):
                eff[name] = (pad(w, 4 * Fp, Fp), pad(b, 4 * Fp))        # reference channel 4c+q keeps its index (c < F)
            elif name == "head":
                eff[name] = (pad(w, 3, Fp), b.detach().float())
            else:
                eff[name] = (pad(w, Fp, Fp), None)
        return eff

    def _pack(self):
        """(Re)pack OIHW fp32 parameters into the kernel layout when they changed."""
        dev = self.neck[0].weight.device
        key = (str(dev), self.compute_dtype) + tuple((c.weight._version, c.weight.data_ptr()) for _, c in self._conv_list())
        if key == self._packed_key:
            return
        from . import ops
        Fp, dt = self.padded_filters, self.compute_dtype
        pk: Dict[str, torch.Tensor] = {}
        for name, (w, b) in self._effective_weights().items():
            if name == "neck":
                pk["neck.w"], pk["neck.b"] = w.contiguous(), b.contiguous()
            elif name.startswith("up"):
                pk[name + ".w"], pk[name + ".b"] = ops.pack_conv3x3(w, b, dt, ps_perm=True)
            elif name == "head":
                pk[name + ".w"], pk[name + ".b"] = ops.pack_conv3x3(w, b, dt, cout_pad=16)
            else:
                pk[name + ".w"], _ = ops.pack_conv3x3(w, None, dt)
        self._packed, self._packed_key = pk, key

    @property
    def padded_filters(self) -> int:
        return (self.n_filters + 63) // 64 * 64

    def _params_struct(self) -> "L.FsrGeneratorParams":
        self._pack()
        pk = self._packed
        P = L.FsrGeneratorParams()
        P.n_filters, P.n_layers, P.dtype = self.padded_filters, self.n_layers, L.dtype_code(self.compute_dtype)
        P.neck_w, P.neck_b = pk["neck.w"].data_ptr(), pk["neck.b"].data_ptr()
        P.neck_alpha = self.neck[1].weight.data_ptr()
        for i, blk in enumerate(self.stem):
            P.stem_w1[i] = pk[f"s{i}a.w"].data_ptr()
            P.stem_alpha[i] = blk.relu1.weight.data_ptr()
            P.stem_w2[i] = pk[f"s{i}b.w"].data_ptr()
        P.bott_w = pk["bott.w"].data_ptr()
        for i in range(2):
            P.up_w[i] = pk[f"up{i}.w"].data_ptr()
            P.up_b[i] = pk[f"up{i}.b"].data_ptr()
            P.up_alpha[i] = self.upsampling[i].relu.weight.data_ptr()
        P.head_w, P.head_b = pk["head.w"].data_ptr(), pk["head.b"].data_ptr()
        return P

    def _forward_wide(self, x: torch.Tensor, out: torch.Tensor, in_u8: int, out_u8: int):
        """n_filters > 64 (multiples of 64 after padding): the same graph on the general-channel kernels."""
        from . import ops
        self._pack()
        pk, dt, Fp = self._packed, self.compute_dtype, self.padded_filters
        a0 = ops.neck_conv3x3(x, pk["neck.w"], pk["neck.b"], dt, act=L.ACT_PRELU, alpha=self.neck[1].weight)
        cur = a0
        for i, blk in enumerate(self.stem):
            raw, st = ops.conv3x3_gen(cur, pk[f"s{i}a.w"], Fp, epilogue=L.EPI_RAW_STATS)
            y = ops.instnorm_apply(raw, st, act=L.ACT_PRELU, alpha=blk.relu1.weight)
            raw, st = ops.conv3x3_gen(y, pk[f"s{i}b.w"], Fp, epilogue=L.EPI_RAW_STATS)
            cur = ops.instnorm_apply(raw, st, residual=cur)
        raw, st = ops.conv3x3_gen(cur, pk["bott.w"], Fp, epilogue=L.EPI_RAW_STATS)
        cur = ops.instnorm_apply(raw, st, residual=a0)
        for i in range(2):
            cur = ops.conv3x3_gen(cur, pk[f"up{i}.w"], 4 * Fp, epilogue=L.EPI_PS_PRELU, bias=pk[f"up{i}.b"],
                                  alpha=self.upsampling[i].relu.weight)
        return ops.conv3x3_head(cur, pk["head.w"], pk["head.b"], out_mode=out_u8, out=out)

    def _workspace(self, N, H, W, dev):
        need = L.load().fsr_generator_workspace_bytes(N, H, W, self.padded_filters, self.n_layers)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws, need

    def _run(self, x: torch.Tensor, out: torch.Tensor, N, H, W, in_u8, out_u8):
        if self.n_layers > L.FSR_MAX_LAYERS:
            raise RuntimeError(f"n_layers > {L.FSR_MAX_LAYERS} not supported")
        if self.padded_filters > 512:
            raise RuntimeError("generator.n_filters > 512 is not supported by this build")
        if self.compute_dtype == torch.float32:
            # precise mode: fp32 storage, fp16 hi+lo split operands on the tensor cores (precise.py)
            if self._precise is None:
                from .precise import PreciseGenerator
                self._precise = PreciseGenerator(self)
            return self._precise.forward(x, out, in_u8, out_u8)
        if self.padded_filters != 64:
            return self._forward_wide(x, out, in_u8, out_u8)
        if self.n_filters <= 32 and W % 2 == 0 and os.environ.get("FSR_PAIR32", "1") != "0":
            # two 32-channel pixels per 128-byte row (pairs.py): half the zero-padding cost of the 64-channel chain
            if self._pair is None:
                from .pairs import PairGenerator
                self._pair = PairGenerator(self)
            return self._pair.forward(x, out, in_u8, out_u8)
        P = self._params_struct()
        ws, need = self._workspace(N, H, W, x.device)
        rc = L.load().fsr_generator_forward(P, x.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), N, H, W,
                                            in_u8, out_u8, self.l2_group, L.stream_ptr(x.device))
        L.check(rc, "fsr_generator_forward")
        return out

    @staticmethod
    def _require_cuda(x):
        if not x.is_cuda:
            raise RuntimeError("fast_srgan_b200.Generator runs on CUDA (sm_100a) tensors only - no CPU fallback")

    _fsr_kind = "G"

    def _net_for_autograd(self):
        from .engine import FlatParams, GeneratorNet
        fp = getattr(self, "_fsr_flat", None)
        if fp is None or not fp.aliases(self):               # first use, or .to(device) re-allocated the parameters
            fp = FlatParams(self, with_optimizer=False)
            self._train_net = None
        if getattr(self, "_train_net", None) is None or self._train_net.fp is not fp:
            dt = self.compute_dtype if self.compute_dtype != torch.float32 else torch.bfloat16
            self._train_net = GeneratorNet(self, fp, dt)
        return self._train_net

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._require_cuda(x)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training through autograd (trainer.py:108,185): hand-written forward + backward kernels behind a Function
            return _NetFn.apply(self, x, *self.parameters())
        x = x.contiguous().float()
        N, C, H, W = x.shape
        if C != 3:
            raise RuntimeError(f"expected 3 input channels, got {C}")
        out = torch.empty((N, 3, 4 * H, 4 * W), dtype=torch.float32, device=x.device)
        return self._run(x, out, N, H, W, 0, 0)

    @torch.no_grad()
    def super_resolve_u8(self, img: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 NHWC [N,H,W,3] -> uint8 NHWC [N,4H,4W,3]; fuses inference.py:48-51 and :54-56."""
        self._require_cuda(img)
        if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
            raise RuntimeError("expected uint8 NHWC [N,H,W,3]")
        img = img.contiguous()
        N, H, W, _ = img.shape
        if out is None:
            out = torch.empty((N, 4 * H, 4 * W, 3), dtype=torch.uint8, device=img.device)
        return self._run(img, out, N, H, W, 1, 1)


class Discriminator(torch.nn.Module):
    """Reference model.py:139-193 on B200 kernels: neck conv + LeakyReLU(0.2), seven SimpleBlocks
    (strides 2,1,2,1,2,1,2; widths F,2F,2F,4F,4F,8F,8F), 1x1 conv -> patch-logit map [N,1,H/16,W/16]."""

    def __init__(self, config, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.config = config
        self.n_filters = F_ = int(config.n_filters)
        self.neck = torch.nn.Sequential(_Conv(3, F_, bias=True), _Empty())
        widths = [(F_, F_, 2), (F_, 2 * F_, 1), (2 * F_, 2 * F_, 2), (2 * F_, 4 * F_, 1), (4 * F_, 4 * F_, 2),
                  (4 * F_, 8 * F_, 1), (8 * F_, 8 * F_, 2)]
        self.stem = torch.nn.Sequential(*[SimpleBlock(ci, co, s) for ci, co, s in widths], _Conv(8 * F_, 1, k=1, bias=True))
        self.compute_dtype = compute_dtype or _default_dtype()
        self._net = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        res = super().load_state_dict(_strip_prefix(state_dict), strict=strict, assign=assign)
        fp = getattr(self, "_fsr_flat", None)
        if fp is not None:                 # a training engine aliases these parameters: its weight packs are stale now
            fp.version += 1
            fp.ext_version += 1
        return res

    def _engine(self):
        from .engine import DiscriminatorNet, FlatParams
        fp = getattr(self, "_fsr_flat", None)
        if fp is None or not fp.aliases(self):               # first use, or .to(device) re-allocated the parameters
            if self.n_filters != 64:
                raise RuntimeError("this build of libfsr_b200 supports discriminator.n_filters == 64 only")
            fp = FlatParams(self, with_optimizer=False)
            self._net = None
        if self._net is None or self._net.fp is not fp:
            self._net = DiscriminatorNet(self, fp, self.compute_dtype)
        return self._net

    _fsr_kind = "D"

    def _net_for_autograd(self):
        return self._engine()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        Generator._require_cuda(x)
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            return _NetFn.apply(self, x, *self.parameters()).unsqueeze(1)      # trainer.py:172,174,186
        net = self._engine()
        net.fp.version += 1            # parameters may have been changed by the caller (load_state_dict / optimizer)
        z, _ = net.forward(x.contiguous().float(), save=False)
        return z.unsqueeze(1)


class VGG19(torch.nn.Module):
    """Reference model.py:5-23: frozen vgg19.features[:34] (-> relu5_3) behind the [-1,1] -> ImageNet renorm.
    ImageNet weights are loaded with load_state_dict (keys vgg.{idx}.weight/bias); there is no download here."""

    def __init__(self, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        from .engine import VGG_PLAN, vgg_conv_indices
        mods: Dict[str, torch.nn.Module] = {}
        cin = 3
        for idx, cout in zip(vgg_conv_indices(), [v for v in VGG_PLAN if v != "M"]):
            conv = _Conv(cin, cout, bias=True)
            torch.nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")   # torchvision's VGG init
            torch.nn.init.zeros_(conv.bias)
            mods[str(idx)] = conv
            cin = cout
        self.vgg = torch.nn.ModuleDict(mods)
        for p in self.parameters():
            p.requires_grad = False
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self.compute_dtype = compute_dtype or _default_dtype()
        self._net = None
        self.weights_loaded = False       # True once load_state_dict() ran (the reference loads IMAGENET1K_V1, model.py:8)
        self._weights_version = 0         # bumped by every load: the engine's packed copies are re-made (engine.VGGNet.pack)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts the reference's keys (`vgg.{idx}.weight/bias`, `mean`, `std`; model.py:8-18) and torchvision's own
        vgg19 checkpoint keys (`features.{idx}.*`, e.g. a locally cached vgg19-dcbb9e9d.pth)."""
        sd = _strip_prefix(state_dict)
        if any(k.startswith("features.") for k in sd):
            keep = {f"features.{i}." for i in self.vgg.keys()}
            sd = {k.replace("features.", "vgg."): v for k, v in sd.items() if k[:k.rfind(".") + 1] in keep}
            sd.setdefault("mean", self.mean)
            sd.setdefault("std", self.std)
        res = super().load_state_dict(sd, strict=strict, assign=assign)
        self.weights_loaded = True
        self._weights_version += 1
        return res

    def _engine(self):
        from .engine import VGGNet
        if self._net is None:
            self._net = VGGNet(self, self.compute_dtype)
        return self._net

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        Generator._require_cuda(x)
        if torch.is_grad_enabled() and x.requires_grad:
            return _VggFn.apply(self, x)                                         # trainer.py:190
        net = self._engine()
        feat, _ = net.forward(x.contiguous().float(), save=False)
        from . import ops
        if net.feat_pad:
            feat = feat[:, 1:-1, 1:-1, :].contiguous()
        return ops.nhwc_to_nchw(feat)
