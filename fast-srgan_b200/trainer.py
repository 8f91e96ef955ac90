"""Mirror of the reference's trainer.py `Trainer` on the B200 engine.

`Trainer(config)` builds Generator / Discriminator / VGG19 like trainer.py:15-51 and exposes
  train_step(lr, hr, noise=None)  - one iteration of the GAN loop body trainer.py:168-196
  pretrain_step(lr, hr)           - one iteration of trainer.py:104-111
  pretrain(...) / train(...)      - the loops of trainer.py:89-141 / :158-233 over a dataloader
  save_checkpoints(step)          - the four files of trainer.py:143-156
  calculate_metrics_over_dataset(dl) - SSIM / PSNR over a validation loader, trainer.py:53-69, on the fused CUDA
                                    metric kernel (metrics.py) instead of torchmetrics
TensorBoard logging (trainer.py:17,70-78,198-233) is observability, not compute: scalars are returned to the
caller / collected in `self.history` instead (out of scope).
Multi-GPU: construct under torchrun after torch.distributed.init_process_group("nccl"); each rank feeds its
own shard of the mini-batch, gradients are summed with one NCCL all-reduce per network per step.
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, Optional

import torch

from .engine import GANEngine
from .metrics import ValidationMetrics
from .model import VGG19, Discriminator, Generator


class Trainer:
    def __init__(self, config, compute_dtype: Optional[torch.dtype] = None, vgg_state_dict=None):
        self.config = config
        dev = torch.device(config.training.device if str(config.training.device).startswith("cuda") else "cuda")
        if not torch.cuda.is_available():
            raise RuntimeError("fast_srgan_b200.Trainer needs a CUDA (sm_100a) device - there is no CPU fallback")
        dt = compute_dtype or (torch.float16 if os.environ.get("FSR_TRAIN_DTYPE", "bf16") == "fp16" else torch.bfloat16)
        self.generator = Generator(config=config.generator, compute_dtype=dt).to(dev)
        self.discriminator = Discriminator(config=config.discriminator, compute_dtype=dt).to(dev)
        self.perceptual_network = VGG19(compute_dtype=dt).to(dev)      # trainer.py:22
        # The reference's VGG19() downloads IMAGENET1K_V1 (model.py:8).  Here the weights come, in this order, from
        # `vgg_state_dict`, `config.training.vgg_weights` (a file), torchvision's local hub cache - never a silent
        # random init: training the content loss against random features is not the reference's objective.
        vgg_state_dict = vgg_state_dict if vgg_state_dict is not None else self._find_vgg_weights(config)
        if vgg_state_dict is not None:
            self.perceptual_network.load_state_dict(vgg_state_dict)
        else:
            warnings.warn("fast_srgan_b200.Trainer: no VGG19 weights given (vgg_state_dict=, config.training.vgg_weights, or "
                          "~/.cache/torch/hub/checkpoints/vgg19-dcbb9e9d.pth) - the perceptual network is RANDOMLY "
                          "initialised, unlike the reference (model.py:8 loads IMAGENET1K_V1).  Call "
                          "trainer.perceptual_network.load_state_dict(...) before training for real.", RuntimeWarning, stacklevel=2)
        self.perceptual_network.eval()
        self.device = dev
        self._dtype = dt
        self._engine: Optional[GANEngine] = None
        from .distributed import rank_and_world
        self.rank, self.world = rank_and_world()
        self._noise_gen = torch.Generator(device=dev)                  # per-rank label noise: every shard draws its own
        self._noise_gen.manual_seed(int(getattr(getattr(config, "experiment", None), "seed", 0) or 0) + self.rank)
        self.metrics = ValidationMetrics(dev, data_range=1.0)          # trainer.py:46-51
        self.history: list = []                                        # (phase, step, dict) instead of the SummaryWriter

    @staticmethod
    def _find_vgg_weights(config):
        cand = [getattr(getattr(config, "training", None), "vgg_weights", None),
                os.path.join(os.environ.get("TORCH_HOME", os.path.expanduser("~/.cache/torch")), "hub", "checkpoints", "vgg19-dcbb9e9d.pth")]
        for path in cand:
            if path and os.path.exists(str(path)):
                return torch.load(str(path), map_location="cpu")
        return None

    @property
    def engine(self) -> GANEngine:
        """Built lazily so that load_state_dict() on the modules before the first step is honoured."""
        if self._engine is None:
            t = self.config.training
            self._engine = GANEngine(self.generator, self.discriminator, self.perceptual_network,
                                     lr_g=float(t.generator_lr), lr_d=float(t.discriminator_lr), dtype=self._dtype)
        return self._engine

    def _label_noise(self, B: int, hw) -> Dict[str, torch.Tensor]:
        # trainer.py:175,176,187 draw torch.rand_like(y); here a per-rank CUDA generator (pass `noise` for parity runs)
        return {k: torch.rand((B, 1) + tuple(hw), generator=self._noise_gen, device=self.device) for k in ("d_real", "d_fake", "g_real")}

    def train_step(self, lr_images: torch.Tensor, hr_images: torch.Tensor, *, noise: Optional[Dict[str, torch.Tensor]] = None):
        lr_images = lr_images.to(self.device, non_blocking=True)       # trainer.py:168-170
        hr_images = hr_images.to(self.device, non_blocking=True)
        if noise is None:
            noise = self._label_noise(lr_images.shape[0], (hr_images.shape[2] // 16, hr_images.shape[3] // 16))
        else:
            noise = {k: v.to(self.device) for k, v in noise.items()}
        return self.engine.train_step(lr_images, hr_images, noise)

    def pretrain_step(self, lr_images: torch.Tensor, hr_images: torch.Tensor):
        return self.engine.pretrain_step(lr_images.to(self.device, non_blocking=True), hr_images.to(self.device, non_blocking=True))

    @torch.no_grad()
    def calculate_metrics_over_dataset(self, dataloader, phase: str = "GAN", step: int = 0) -> Dict[str, float]:
        """trainer.py:53-69 `_calculate_metrics_over_dataset`: generator in eval mode over the whole loader,
        sr = (1 + G(lr))/2 against (1 + hr)/2; returns {"ssim": mean per-image SSIM, "psnr": PSNR of the pooled MSE}."""
        was_training = self.generator.training
        self.generator.eval()
        self.metrics.reset()
        for lr_images, hr_images in dataloader:
            lr_images = lr_images.to(self.device, non_blocking=True)
            hr_images = hr_images.to(self.device, non_blocking=True)
            self.metrics.update(self.generator(lr_images), hr_images, rescale=True)   # trainer.py:64-66
        out = self.metrics.compute()
        res = {"ssim": out["ssim"], "psnr": out["psnr"]}
        self.history.append((phase, step, res))
        self.generator.train(was_training)
        return res

    def _every(self, name: str, step: int) -> bool:
        n = int(getattr(self.config.training, name, 0) or 0)
        return n > 0 and step % n == 0

    def pretrain(self, train_dataloader, val_dataloader=None):
        if os.path.exists("runs/pretrain.pt"):                                             # trainer.py:90-94
            ck = torch.load("runs/pretrain.pt", map_location="cpu")
            self.generator.load_state_dict(ck["model"])
            self.engine.gp.load_optimizer_state(ck["optimizer"])
            return None
        if val_dataloader is not None:
            self.calculate_metrics_over_dataset(val_dataloader, "Pretrain", 0)            # trainer.py:95
        last = None
        for step, (lr_images, hr_images) in enumerate(train_dataloader, start=1):          # trainer.py:99-111
            last = self.pretrain_step(lr_images, hr_images)
            if val_dataloader is not None and self._every("checkpoint_iter", step):
                self.calculate_metrics_over_dataset(val_dataloader, "Pretrain", step)     # trainer.py:126
        if self.rank == 0:                                                                 # trainer.py:131-141
            os.makedirs("runs", exist_ok=True)
            e, t = self.engine, self.config.training
            torch.save({"model": self.generator.state_dict(), "optimizer": e.gp.optimizer_state(float(t.generator_lr))},
                       "runs/pretrain_generator.pt")
            torch.save({"model": self.discriminator.state_dict(), "optimizer": e.dp.optimizer_state(float(t.discriminator_lr))},
                       "runs/pretrain_discriminator.pt")
        self._barrier()
        return last

    def _barrier(self):
        if self.world > 1:
            torch.distributed.barrier()

    def train(self, train_dataloader, val_dataloader=None):
        if val_dataloader is not None:
            self.calculate_metrics_over_dataset(val_dataloader, "GAN", 0)                 # trainer.py:159
        last = None
        for step, (lr_images, hr_images) in enumerate(train_dataloader, start=1):          # trainer.py:165-196
            last = self.train_step(lr_images, hr_images)
            if self._every("checkpoint_iter", step):
                if val_dataloader is not None:
                    self.calculate_metrics_over_dataset(val_dataloader, "GAN", step)      # trainer.py:231
                self.save_checkpoints(step)                                               # trainer.py:232
        return last

    def save_checkpoints(self, step: int):
        """trainer.py:143-156: generator / discriminator state dicts + optimizer states (torch.optim.AdamW state_dict
        format) under runs/<name>/.  Replicas are identical, so rank 0 alone writes; the others wait at a barrier."""
        save_dir = os.path.join("runs", self.config.experiment.name)
        if self.rank == 0:
            os.makedirs(save_dir, exist_ok=True)
            torch.save(self.generator.state_dict(), os.path.join(save_dir, f"generator_epoch_{step}.pt"))
            torch.save(self.discriminator.state_dict(), os.path.join(save_dir, f"discriminator_epoch_{step}.pt"))
            e, t = self.engine, self.config.training
            for name, fp, lr in (("generator", e.gp, t.generator_lr), ("discriminator", e.dp, t.discriminator_lr)):
                torch.save(fp.optimizer_state(float(lr)), os.path.join(save_dir, f"{name}_optim_epoch_{step}.pt"))
        self._barrier()

    def load_checkpoints(self, step: int, save_dir: Optional[str] = None):
        """Resume from the four files save_checkpoints(step) wrote (the reference only reloads its pretrain checkpoint,
        trainer.py:90-94; here both networks and both AdamW states)."""
        save_dir = save_dir or os.path.join("runs", self.config.experiment.name)
        e = self.engine                                      # build first: the modules' parameters alias its flat buffers
        self.generator.load_state_dict(torch.load(os.path.join(save_dir, f"generator_epoch_{step}.pt"), map_location="cpu"))
        self.discriminator.load_state_dict(torch.load(os.path.join(save_dir, f"discriminator_epoch_{step}.pt"), map_location="cpu"))
        for name, fp in (("generator", e.gp), ("discriminator", e.dp)):
            fp.load_optimizer_state(torch.load(os.path.join(save_dir, f"{name}_optim_epoch_{step}.pt"), map_location="cpu"))
            fp.version += 1                                  # parameters changed: weight packs are stale
            fp.ext_version += 1                              # ... and captured CUDA graphs with them
