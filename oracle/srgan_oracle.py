"""CPU oracle for the Fast-SRGAN hot path (TEST INFRASTRUCTURE - never shipped, never timed as product).

This is a functional fp32 restatement, in plain PyTorch CPU ops, of the reference's
hot path: Generator / Discriminator / VGG19 forward (reference model.py) and the GAN
step body (reference trainer.py:168-196).  Every function cites the reference
file:line it follows.  It takes *state dicts* (reference key names, OIHW fp32) so no
module-construction RNG order is involved.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product package (fast-srgan_b200/) must not.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The oracle is pinned instead against the reference itself:
oracle/make_golden.py imports /root/reference/model.py unmodified in the build
container, runs it on seeded inputs/weights and (a) asserts this restatement matches
it to <=1e-5, (b) writes the reference's outputs to tests/golden/*.npz.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# VGG19 `features[:34]` layer plan (reference model.py:8): conv indices and pools.
# numbers = conv out-channels, "M" = 2x2 max-pool.  Ends at relu5_3 (index 33).
VGG19_PLAN = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M",
              512, 512, 512, 512, "M", 512, 512, 512]
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # reference model.py:13
IMAGENET_STD = (0.229, 0.224, 0.225)    # reference model.py:17


def vgg19_conv_indices():
    """torchvision `features` indices of the conv layers kept by features[:34]."""
    idx, i = [], 0
    for v in VGG19_PLAN:
        if v == "M":
            i += 1
        else:
            idx.append(i)
            i += 2  # conv + relu
    return idx


# --------------------------------------------------------------------------- weights
def _fill(gen: torch.Generator, shape, std: float) -> Tensor:
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def make_generator_state(n_filters: int = 64, n_layers: int = 8, seed: int = 1234) -> StateDict:
    """Deterministic random Generator weights with the reference key names (model.py:72-110).

    Scale ~ PyTorch's default conv init (std = 1/sqrt(3*fan_in)); PReLU alpha = 0.25
    perturbed so that alpha-gradient bugs are visible.  NOT the module RNG order.
    """
    g = torch.Generator().manual_seed(seed)
    Fm = n_filters
    sd: StateDict = {}

    def conv(name, co, ci, bias):
        std = 1.0 / math.sqrt(3.0 * ci * 9)
        sd[name + ".weight"] = _fill(g, (co, ci, 3, 3), std)
        if bias:
            sd[name + ".bias"] = _fill(g, (co,), std)

    conv("neck.0", Fm, 3, True)
    sd["neck.1.weight"] = torch.tensor([0.25]) + 0.05 * torch.randn(1, generator=g)
    for i in range(n_layers):
        conv(f"stem.{i}.conv1", Fm, Fm, False)
        sd[f"stem.{i}.relu1.weight"] = torch.tensor([0.25]) + 0.05 * torch.randn(1, generator=g)
        conv(f"stem.{i}.conv2", Fm, Fm, False)
    conv("bottleneck.0", Fm, Fm, False)
    for i in range(2):
        conv(f"upsampling.{i}.conv", 4 * Fm, Fm, True)
        sd[f"upsampling.{i}.relu.weight"] = torch.tensor([0.25]) + 0.05 * torch.randn(1, generator=g)
    conv("head.0", 3, Fm, True)
    return sd


def make_discriminator_state(n_filters: int = 64, seed: int = 4321) -> StateDict:
    """Deterministic random Discriminator weights, reference key names (model.py:139-189)."""
    g = torch.Generator().manual_seed(seed)
    Fm = n_filters
    sd: StateDict = {}
    std = 1.0 / math.sqrt(3.0 * 3 * 9)
    sd["neck.0.weight"] = _fill(g, (Fm, 3, 3, 3), std)
    sd["neck.0.bias"] = _fill(g, (Fm,), std)
    widths = [(Fm, Fm), (Fm, 2 * Fm), (2 * Fm, 2 * Fm), (2 * Fm, 4 * Fm),
              (4 * Fm, 4 * Fm), (4 * Fm, 8 * Fm), (8 * Fm, 8 * Fm)]
    for i, (ci, co) in enumerate(widths):
        sd[f"stem.{i}.conv.weight"] = _fill(g, (co, ci, 3, 3), 1.0 / math.sqrt(3.0 * ci * 9))
    sd["stem.7.weight"] = _fill(g, (1, 8 * Fm, 1, 1), 1.0 / math.sqrt(3.0 * 8 * Fm))
    sd["stem.7.bias"] = _fill(g, (1,), 0.05)
    return sd


def make_vgg19_state(seed: int = 99, width_div: int = 1) -> StateDict:
    """Deterministic random VGG19[:34] weights (keys `vgg.{idx}.weight/bias`, model.py:8).

    ImageNet weights cannot be downloaded (no network) so parity uses random init,
    He-normal so activations keep O(1) scale through 15 ReLU convs.  `width_div`
    shrinks channel counts for cheap CPU tests (1 = the real 64..512 widths).
    """
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    ci = 3
    for idx, co in zip(vgg19_conv_indices(), [v for v in VGG19_PLAN if v != "M"]):
        co = co // width_div
        sd[f"vgg.{idx}.weight"] = _fill(g, (co, ci, 3, 3), math.sqrt(2.0 / (ci * 9)))
        sd[f"vgg.{idx}.bias"] = _fill(g, (co,), 0.05)
        ci = co
    sd["mean"] = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    sd["std"] = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return sd


# --------------------------------------------------------------------------- storage-rounding emulation
# The B200 engine stores activations / packed weights in fp16|bf16 (fp32 accumulation).  To separate "what any
# 16-bit-storage implementation must lose" from implementation error, the oracle can emulate exactly that: with
# `storage_rounding(dtype)` active every conv weight, conv output and activation tensor is rounded to `dtype`
# (straight-through in backward) while all arithmetic stays in the caller's precision (fp64 in the tests).
_STORAGE_DTYPE = None


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt):
        return x.to(dt).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


class storage_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _STORAGE_DTYPE
        self.prev, _STORAGE_DTYPE = _STORAGE_DTYPE, self.dtype

    def __exit__(self, *a):
        global _STORAGE_DTYPE
        _STORAGE_DTYPE = self.prev


def _q(t: Tensor) -> Tensor:
    return t if _STORAGE_DTYPE is None else _RoundSTE.apply(t, _STORAGE_DTYPE)


def _conv(x, w, b=None, stride=1, padding=1, quant_w=True):
    return _q(F.conv2d(x, _q(w) if quant_w else w, b, stride=stride, padding=padding))


# --------------------------------------------------------------------------- building blocks
def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """torch.nn.InstanceNorm2d defaults: affine=False, biased variance, eps=1e-5
    (reference model.py:55,65,94,132)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def prelu(x: Tensor, alpha: Tensor) -> Tensor:
    """Single shared-slope PReLU (reference model.py:37,56,77)."""
    return torch.where(x >= 0, x, alpha.view(1, 1, 1, 1) * x)


def pixel_shuffle2(x: Tensor) -> Tensor:
    """out[n,c,2h+i,2w+j] = in[n,4c+2i+j,h,w]  (torch.nn.PixelShuffle(2), model.py:36)."""
    n, c4, h, w = x.shape
    c = c4 // 4
    return x.view(n, c, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, c, 2 * h, 2 * w)


# --------------------------------------------------------------------------- networks
def residual_block(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """reference model.py:67-69:  bn2(conv2(relu1(bn1(conv1(x))))) + x."""
    y = _q(prelu(instance_norm(_conv(x, sd[p + "conv1.weight"])), sd[p + "relu1.weight"]))
    return _q(instance_norm(_conv(y, sd[p + "conv2.weight"])) + x)


def generator_forward(sd: StateDict, x: Tensor, n_layers: Optional[int] = None) -> Tensor:
    """reference model.py:112-117."""
    if n_layers is None:
        n_layers = sum(1 for k in sd if k.startswith("stem.") and k.endswith("conv1.weight"))
    # (the 3-channel neck runs on CUDA cores with fp32 weights; only its output is stored rounded)
    residual = _q(prelu(F.conv2d(x, sd["neck.0.weight"], sd["neck.0.bias"], padding=1), sd["neck.1.weight"]))
    y = residual
    for i in range(n_layers):
        y = residual_block(sd, f"stem.{i}.", y)
    y = _q(instance_norm(_conv(y, sd["bottleneck.0.weight"])) + residual)             # model.py:115
    for i in range(2):                                                                  # model.py:39-40
        y = F.conv2d(y, _q(sd[f"upsampling.{i}.conv.weight"]), sd[f"upsampling.{i}.conv.bias"], padding=1)
        y = _q(prelu(pixel_shuffle2(y), sd[f"upsampling.{i}.relu.weight"]))             # fused epilogue: rounded once
    y = F.conv2d(y, _q(sd["head.0.weight"]), sd["head.0.bias"], padding=1)              # model.py:102-110 (fp32 out)
    return torch.tanh(y)


D_STRIDES = (2, 1, 2, 1, 2, 1, 2)  # reference model.py:148-183


def discriminator_forward(sd: StateDict, x: Tensor) -> Tensor:
    """reference model.py:191-193 (neck :143-146, SimpleBlock :135-136, 1x1 conv :184-186)."""
    y = _q(F.leaky_relu(F.conv2d(x, sd["neck.0.weight"], sd["neck.0.bias"], padding=1), 0.2))
    for i, s in enumerate(D_STRIDES):
        y = _conv(y, sd[f"stem.{i}.conv.weight"], stride=s)
        y = _q(F.leaky_relu(instance_norm(y), 0.01))    # torch.nn.LeakyReLU() default slope
    return F.conv2d(y, sd["stem.7.weight"], sd["stem.7.bias"])


def vgg19_forward(sd: StateDict, x: Tensor) -> Tensor:
    """reference model.py:20-23 followed by torchvision vgg19.features[:34]."""
    y = (x + 1.0) / 2.0
    y = (y - sd["mean"]) / sd["std"]
    convs = iter(vgg19_conv_indices())
    for v in VGG19_PLAN:
        if v == "M":
            y = F.max_pool2d(y, kernel_size=2, stride=2)
        else:
            i = next(convs)
            y = _q(F.relu(F.conv2d(y, _q(sd[f"vgg.{i}.weight"]) if i > 0 else sd[f"vgg.{i}.weight"], sd[f"vgg.{i}.bias"], padding=1)))
    return y


# --------------------------------------------------------------------------- losses / optimiser
def bce_with_logits_mean(z: Tensor, t: Tensor) -> Tensor:
    """torch.nn.BCEWithLogitsLoss() (trainer.py:41): mean(max(z,0) - z*t + log1p(exp(-|z|)))."""
    return (z.clamp_min(0) - z * t + torch.log1p(torch.exp(-z.abs()))).mean()


def smooth_l1_mean(a: Tensor, b: Tensor) -> Tensor:
    """torch.nn.SmoothL1Loss() beta=1 (trainer.py:43)."""
    d = (a - b).abs()
    return torch.where(d < 1.0, 0.5 * d * d, d - 0.5).mean()


class AdamWState:
    """torch.optim.AdamW defaults used by trainer.py:33-38 (betas .9/.999, eps 1e-8, wd 1e-2)."""

    def __init__(self, params: StateDict, lr: float):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, 0.9, 0.999, 1e-8, 1e-2
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    @torch.no_grad()
    def step(self, params: StateDict, grads: StateDict):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            p.mul_(1.0 - self.lr * self.wd)
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def _leaf(sd: StateDict, skip=()) -> StateDict:
    return {k: (v.detach().clone().requires_grad_(k not in skip)) for k, v in sd.items()}


def gan_step(g_sd: StateDict, d_sd: StateDict, vgg_sd: StateDict, lr_img: Tensor, hr_img: Tensor,
             noise: Dict[str, Tensor], opt_g: AdamWState, opt_d: AdamWState):
    """One iteration of the reference GAN loop body, trainer.py:168-196, with the three
    `rand_like` label-noise draws (trainer.py:175,176,187) supplied by the caller as
    noise["d_real"], noise["d_fake"], noise["g_real"] (uniform [0,1) tensors shaped like
    the discriminator output) so CPU and CUDA runs consume identical labels.

    Updates g_sd / d_sd in place; returns dict(losses..., g_grads, d_grads).
    """
    # ---- discriminator step (trainer.py:171-181)
    d = _leaf(d_sd)
    y_real = discriminator_forward(d, hr_img)
    with torch.no_grad():
        sr = generator_forward(g_sd, lr_img)                       # :173 (.detach())
    y_fake = discriminator_forward(d, sr)
    real_labels = 0.3 * noise["d_real"] + 0.8                       # :175
    fake_labels = 0.3 * noise["d_fake"]                             # :176
    loss_real = bce_with_logits_mean(y_real, real_labels)           # :177
    loss_fake = bce_with_logits_mean(y_fake, fake_labels)           # :178
    d_loss = 0.5 * loss_real + 0.5 * loss_fake                      # :179
    d_loss.backward()                                               # :180
    d_grads = {k: v.grad for k, v in d.items() if v.grad is not None}
    opt_d.step(d_sd, d_grads)                                       # :181

    # ---- generator step (trainer.py:184-196)
    gl = _leaf(g_sd)
    sr = generator_forward(gl, lr_img)                              # :185
    y_fake = discriminator_forward(d_sd, sr)                        # :186 (updated D)
    adv = 1e-1 * bce_with_logits_mean(y_fake, 0.3 * noise["g_real"] + 0.7)   # :187-188
    fake_f = vgg19_forward(vgg_sd, sr)                              # :190
    with torch.no_grad():
        real_f = vgg19_forward(vgg_sd, hr_img)                      # :191
    content = smooth_l1_mean(fake_f, real_f)                        # :192
    g_loss = 0.5 * adv + 0.5 * content                              # :194
    g_loss.backward()                                               # :195
    g_grads = {k: v.grad for k, v in gl.items() if v.grad is not None}
    opt_g.step(g_sd, g_grads)                                       # :196
    return dict(loss_real=loss_real.detach(), loss_fake=loss_fake.detach(), adv_loss=adv.detach(),
                content_loss=content.detach(), g_grads=g_grads, d_grads=d_grads, sr=sr.detach())


def pretrain_step(g_sd: StateDict, lr_img: Tensor, hr_img: Tensor, opt_g: AdamWState):
    """reference trainer.py:104-111 (generator-only SmoothL1 warm-up)."""
    gl = _leaf(g_sd)
    loss = smooth_l1_mean(generator_forward(gl, lr_img), hr_img)
    loss.backward()
    g_grads = {k: v.grad for k, v in gl.items() if v.grad is not None}
    opt_g.step(g_sd, g_grads)
    return dict(loss=loss.detach(), g_grads=g_grads)


def to_uint8_image(sr: Tensor) -> Tensor:
    """reference inference.py:54-56: ((y+1)/2*255) -> numpy astype(uint8) (truncation), NHWC."""
    y = ((sr + 1.0) / 2.0).permute(0, 2, 3, 1) * 255
    return y.to(torch.uint8)  # float->uint8 truncates toward zero like numpy astype for in-range values


def from_uint8_image(img_u8_nhwc: Tensor) -> Tensor:
    """reference inference.py:48-51: uint8 HWC / 127.5 - 1 -> NCHW fp32."""
    return (img_u8_nhwc.to(torch.float32) / 127.5 - 1.0).permute(0, 3, 1, 2).contiguous()
