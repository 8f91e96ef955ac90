"""The reference's own training loop shape (trainer.py:104-111 and :171-196): `loss.backward()` THROUGH the modules and
stock `torch.optim.AdamW(module.parameters())`.  The modules' autograd bridge (model.py `_NetFn` / `_VggFn`) must make
that loop produce what the fast path (Trainer.train_step: flat buffers, one CUDA graph) produces."""
import types

import pytest
import torch

import srgan_oracle as O

pytestmark = pytest.mark.gpu


def ns(**k):
    return types.SimpleNamespace(**k)


def _modules(dt):
    from fast_srgan_b200.model import VGG19, Discriminator, Generator
    g = Generator(ns(n_filters=64, n_layers=8), compute_dtype=dt)
    d = Discriminator(ns(n_filters=64), compute_dtype=dt)
    v = VGG19(compute_dtype=dt)
    g.load_state_dict(O.make_generator_state(64, 8, 1234))
    d.load_state_dict(O.make_discriminator_state(64, 4321))
    v.load_state_dict(O.make_vgg19_state(99))
    return g.cuda(), d.cuda(), v.cuda().eval()


def _inputs(B=4):
    gen = torch.Generator().manual_seed(31)
    lr = (torch.rand((B, 3, 24, 24), generator=gen) * 2 - 1).cuda()
    hr = (torch.rand((B, 3, 96, 96), generator=gen) * 2 - 1).cuda()
    noise = {k: torch.rand((B, 1, 6, 6), generator=gen).cuda() for k in ("d_real", "d_fake", "g_real")}
    return lr, hr, noise


def test_reference_style_gan_step_matches_train_step():
    dt = torch.bfloat16
    lr_img, hr_img, noise = _inputs()
    # ---- the reference loop body, verbatim apart from the injected label noise (trainer.py:171-196)
    generator, discriminator, perceptual_network = _modules(dt)
    optim_g = torch.optim.AdamW(generator.parameters(), lr=1e-4)
    optim_d = torch.optim.AdamW(discriminator.parameters(), lr=1e-4)
    gan_loss, l1_loss = torch.nn.BCEWithLogitsLoss(), torch.nn.SmoothL1Loss()
    optim_d.zero_grad(set_to_none=True)
    y_real = discriminator(hr_img)
    fake_hr_images = generator(lr_img).detach()
    y_fake = discriminator(fake_hr_images)
    loss_real = gan_loss(y_real, 0.3 * noise["d_real"] + 0.8)
    loss_fake = gan_loss(y_fake, 0.3 * noise["d_fake"])
    (0.5 * loss_real + 0.5 * loss_fake).backward()
    optim_d.step()
    optim_g.zero_grad(set_to_none=True)
    fake_hr_images = generator(lr_img)
    y_fake = discriminator(fake_hr_images)
    adv = 1e-1 * gan_loss(y_fake, 0.3 * noise["g_real"] + 0.7)
    content = l1_loss(perceptual_network(fake_hr_images), perceptual_network(hr_img))
    (0.5 * adv + 0.5 * content).backward()
    optim_g.step()
    torch.cuda.synchronize()
    assert y_fake.shape == (4, 1, 6, 6) and all(p.grad is not None for p in generator.parameters())

    # ---- the fast path on the same weights / inputs
    from fast_srgan_b200.trainer import Trainer
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=8), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=dt, vgg_state_dict=O.make_vgg19_state(99))
    tr.generator.load_state_dict(O.make_generator_state(64, 8, 1234))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    out = tr.train_step(lr_img, hr_img, noise=noise)
    torch.cuda.synchronize()
    for name, a, b in (("loss_real", loss_real, out["loss_real"]), ("loss_fake", loss_fake, out["loss_fake"]),
                       ("adv", adv, out["adv_loss"]), ("content", content, out["content_loss"])):
        print(f"{name}: autograd loop {a.item():.6f}  train_step {b.item():.6f}")
        assert abs(a.item() - b.item()) <= 2e-3 * max(1.0, abs(b.item()))
    sd0 = {"g": O.make_generator_state(64, 8, 1234), "d": O.make_discriminator_state(64, 4321)}
    for tag, mod, fast in (("g", generator, tr.generator), ("d", discriminator, tr.discriminator)):
        worst, agrees = 0.0, []
        for (k, p), (_, q) in zip(mod.named_parameters(), fast.named_parameters()):
            worst = max(worst, (p - q).abs().max().item())
            if p.numel() > 1:
                u0, u1 = p.detach().cpu() - sd0[tag][k], q.detach().cpu() - sd0[tag][k]
                agrees.append((torch.sign(u0) == torch.sign(u1)).float().mean().item())
        print(f"{tag}: max |param_autograd - param_train_step| {worst:.3e}, update-sign agreement mean {sum(agrees)/len(agrees):.4f} min {min(agrees):.4f}")
        assert worst <= 2.05e-4 and min(agrees) >= 0.9


def test_reference_style_pretrain_step():
    """trainer.py:104-111 through autograd: SmoothL1(G(lr), hr).backward(); AdamW.step() - loss falls, parameters move."""
    generator, _, _ = _modules(torch.bfloat16)
    lr_img, hr_img, _ = _inputs()
    hr_img = torch.nn.functional.interpolate(lr_img, scale_factor=4, mode="bilinear")
    optim = torch.optim.AdamW(generator.parameters(), lr=1e-3)
    losses = []
    for _ in range(20):
        optim.zero_grad(set_to_none=True)
        loss = torch.nn.functional.smooth_l1_loss(generator(lr_img), hr_img)
        loss.backward()
        optim.step()
        losses.append(loss.item())
    print("pretrain losses", losses[0], "->", losses[-1])
    assert losses[-1] < losses[0]
    generator.eval()
    with torch.no_grad():                      # the inference path sees the parameters the optimizer stepped
        y = generator(lr_img)
    assert torch.isfinite(y).all()
