"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only mount)

What it does
  1. imports /root/reference/model.py and trainer.py as-is (torchmetrics / tensorboard are
     stubbed in sys.modules because they are not installed; model.vgg19 is wrapped so the
     ImageNet download - impossible without network - becomes weights=None),
  2. loads the oracle's deterministic state dicts into the reference modules,
  3. runs the reference forward passes and ONE genuine iteration of Trainer.train()
     (trainer.py:158-196) with torch.rand_like patched to return the committed label noise,
  4. asserts oracle/srgan_oracle.py reproduces every result (<=2e-5 abs), and
  5. writes the reference's outputs as golden fixtures.

The GPU box has no /root/reference: tests there use the fixtures + the oracle.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import srgan_oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    # --- stubs for absent third-party modules (not part of the hot path)
    tm = types.ModuleType("torchmetrics")
    tmi = types.ModuleType("torchmetrics.image")

    class _Metric:
        def __init__(self, *a, **k):
            pass

        def to(self, *_):
            return self

        def reset(self):
            pass

        def update(self, *a):
            pass

        def compute(self):
            return torch.zeros(1)

    tmi.PeakSignalNoiseRatio = _Metric
    tmi.StructuralSimilarityIndexMeasure = _Metric
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.image"] = tmi
    tb = types.ModuleType("torch.utils.tensorboard.writer")

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return lambda *a, **k: None

    tb.SummaryWriter = _Writer
    sys.modules["torch.utils.tensorboard.writer"] = tb
    import model  # the reference's model.py

    _orig = model.vgg19
    model.vgg19 = lambda weights=None: _orig(weights=None)
    import trainer  # the reference's trainer.py
    return model, trainer


def ns(**k):
    return types.SimpleNamespace(**k)


def seeded(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * 2 - 1


def check(name, a, b, tol=2e-5, rel=None):
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    print(f"  oracle vs reference  {name:32s} max-abs {err:.3e}  (ref abs-max {scale:.3e})")
    if rel is not None:
        assert err <= rel * scale + 1e-9, (name, err, scale)
    else:
        assert err <= tol, (name, err)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model, trainer = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    out = {}

    # ---------------- generator, full width (F=64, L=8) small frame
    for tag, Fm, L, shape in (("g64x8", 64, 8, (1, 3, 20, 24)), ("g32x2", 32, 2, (2, 3, 9, 13))):
        sd = O.make_generator_state(Fm, L, seed=1234)
        ref = model.Generator(ns(n_filters=Fm, n_layers=L))
        ref.load_state_dict(sd)
        ref.eval()
        x = seeded(shape, 7)
        with torch.no_grad():
            y_ref = ref(x)
            y_or = O.generator_forward(sd, x)
        check(f"generator {tag}", y_or, y_ref)
        out[f"{tag}_y"] = y_ref.numpy()

    # ---------------- discriminator
    dsd = O.make_discriminator_state(64, seed=4321)
    dref = model.Discriminator(ns(n_filters=64))
    dref.load_state_dict(dsd)
    xd = seeded((2, 3, 96, 96), 11)
    with torch.no_grad():
        yd_ref = dref(xd)
        yd_or = O.discriminator_forward(dsd, xd)
    check("discriminator", yd_or, yd_ref)
    out["d64_y"] = yd_ref.numpy()

    # ---------------- VGG19 (random init - ImageNet weights need network)
    vsd = O.make_vgg19_state(seed=99)
    vref = model.VGG19()
    vref.load_state_dict(vsd)
    vref.eval()
    xv = seeded((1, 3, 32, 32), 13)
    with torch.no_grad():
        yv_ref = vref(xv)
        yv_or = O.vgg19_forward(vsd, xv)
    check("vgg19", yv_or, yv_ref, tol=1e-4)
    out["vgg_y"] = yv_ref.numpy()

    # ---------------- one genuine Trainer.train() iteration (trainer.py:158-196), fp32 and fp64
    # NOTE (measured here): these gradients are ill-conditioned - InstanceNorm over 6x6 / 12x12
    # planes divides by tiny per-plane sigmas - so the reference's own fp32 run differs from its
    # fp64 run by up to ~1.5e-2 (max-abs / abs-max) on some tensors.  Both are recorded: the
    # fp64 run of the reference is the ground truth, the fp32 run shows the reference's noise.
    B = 2
    gsd = O.make_generator_state(64, 8, seed=1234)
    lr_img, hr_img = seeded((B, 3, 24, 24), 21), seeded((B, 3, 96, 96), 22)
    gn = torch.Generator().manual_seed(23)
    noise = {k: torch.rand((B, 1, 6, 6), generator=gn) for k in ("d_real", "d_fake", "g_real")}

    def digest(prefix, tensors):
        for k, v in tensors.items():
            flat = v.reshape(-1).double()
            out[f"{prefix}/{k}/norm"] = np.float64(flat.norm().item())
            step = max(1, flat.numel() // 64)
            out[f"{prefix}/{k}/sample"] = flat[::step][:64].numpy().copy()

    for tag, dt in (("step32", torch.float32), ("step64", torch.float64)):
        torch.set_default_dtype(dt)
        cfg = ns(experiment=ns(name="golden"),
                 generator=ns(n_filters=64, n_layers=8), discriminator=ns(n_filters=64, n_layers=7),
                 training=ns(compiled=False, device="cpu", generator_lr=1e-4, discriminator_lr=1e-4,
                             log_iter=10**9, checkpoint_iter=10**9))
        tr = trainer.Trainer(cfg)
        cast = lambda sd: {k: v.to(dt) for k, v in sd.items()}
        tr.generator.load_state_dict(cast(gsd))
        tr.discriminator.load_state_dict(cast(dsd))
        tr.perceptual_network.load_state_dict(cast(vsd))
        queue = [noise["d_real"].to(dt), noise["d_fake"].to(dt), noise["g_real"].to(dt)]
        real_rand_like = torch.rand_like
        torch.rand_like = lambda t, *a, **k: queue.pop(0).clone()
        captured = {}
        real_bwd = torch.Tensor.backward

        def spy_backward(self, *a, **k):
            real_bwd(self, *a, **k)
            which = "d" if "d" not in captured else "g"
            net = tr.discriminator if which == "d" else tr.generator
            captured[which] = {k_: p.grad.detach().clone() for k_, p in net.named_parameters() if p.grad is not None}

        torch.Tensor.backward = spy_backward
        try:
            tr.train([(lr_img.to(dt), hr_img.to(dt))], [])
        finally:
            torch.rand_like = real_rand_like
            torch.Tensor.backward = real_bwd
            torch.set_default_dtype(torch.float32)
        g_after = {k: v.detach().clone() for k, v in tr.generator.state_dict().items()}
        d_after = {k: v.detach().clone() for k, v in tr.discriminator.state_dict().items()}

        # oracle step on the same inputs, same dtype
        og, od, ov = cast(gsd), cast(dsd), cast(vsd)
        og = {k: v.clone() for k, v in og.items()}
        od = {k: v.clone() for k, v in od.items()}
        res = O.gan_step(og, od, ov, lr_img.to(dt), hr_img.to(dt), {k: v.to(dt) for k, v in noise.items()},
                         O.AdamWState(og, 1e-4), O.AdamWState(od, 1e-4))
        grad_rel = 1e-9 if dt == torch.float64 else 0.25  # fp32: reference noise, see NOTE
        for k in captured["d"]:
            check(f"{tag} d_grad {k}", res["d_grads"][k], captured["d"][k], rel=grad_rel)
        for k in captured["g"]:
            check(f"{tag} g_grad {k}", res["g_grads"][k], captured["g"][k], rel=grad_rel)
        # after one AdamW step from zero state the update is lr*sign(g) (m/sqrt(v)=+-1): tiny grads can flip
        after_tol = 1e-12 if dt == torch.float64 else 2.1e-4
        for k in g_after:
            check(f"{tag} g_after {k}", og[k], g_after[k], tol=after_tol)
        for k in d_after:
            check(f"{tag} d_after {k}", od[k], d_after[k], tol=after_tol)
        digest(f"{tag}_d_grad", captured["d"])
        digest(f"{tag}_g_grad", captured["g"])
        digest(f"{tag}_g_after", g_after)
        digest(f"{tag}_d_after", d_after)
        for k in ("loss_real", "loss_fake", "adv_loss", "content_loss"):
            out[f"{tag}_{k}"] = np.float64(res[k].item())
            print(f"  {tag} {k} = {res[k].item():.9f}")

    # ---------------- shipped checkpoint sanity anchor - reference only, weights do not travel
    ck = {k.replace("_orig_mod.", ""): v for k, v in torch.load(f"{REF}/models/model.pt", map_location="cpu").items()}
    gck = model.Generator(ns(n_filters=64, n_layers=8))
    gck.load_state_dict(ck)
    gck.eval()
    torch.manual_seed(0)
    xa = torch.rand(2, 3, 90, 160) * 2 - 1
    with torch.no_grad():
        ya, yo = gck(xa), O.generator_forward(ck, xa)
    check("generator(model.pt) 90x160", yo, ya, tol=5e-5)
    out["ckpt_anchor"] = np.array([ya.mean().item(), ya.std().item(), ya.abs().max().item()])
    print("  checkpoint anchor mean/std/absmax", out["ckpt_anchor"])

    np.savez_compressed(os.path.join(GOLD, "reference_golden.npz"), **out)
    print("wrote", os.path.join(GOLD, "reference_golden.npz"),
          os.path.getsize(os.path.join(GOLD, "reference_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
