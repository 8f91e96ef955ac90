"""GPU parity of Discriminator / VGG19 forward and of one GAN train step (trainer.py:168-196) against the
CPU oracle (and, through it, the committed fp64 run of the unmodified reference).

Gradient tolerances: SURVEY/oracle measurements show the reference's OWN fp32 gradients differ from its fp64
run by up to ~1.5e-2 (max-abs / abs-max) per tensor, because InstanceNorm over 6x6..12x12 planes is
ill-conditioned.  The engine computes with bf16|fp16 operands and fp32 accumulation, so parity is asserted in
relative L2 norm per tensor against the fp64 oracle with the bounds written below."""
import types

import pytest
import torch

import srgan_oracle as O

pytestmark = pytest.mark.gpu


def seeded(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * 2 - 1


def ns(**k):
    return types.SimpleNamespace(**k)


@pytest.mark.parametrize("dt,tol", [(torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_discriminator_forward(golden, dt, tol):
    from fast_srgan_b200.model import Discriminator
    sd = O.make_discriminator_state(64, seed=4321)
    d = Discriminator(ns(n_filters=64), compute_dtype=dt)
    d.load_state_dict(sd)
    d = d.cuda()
    x = seeded((2, 3, 96, 96), 11)
    with torch.no_grad():
        y = d(x.cuda()).cpu()
    err = (y.numpy() - golden["d64_y"]).__abs__().max()
    print(f"discriminator {dt}: max-abs vs reference golden {err:.3e} (|y| max {abs(golden['d64_y']).max():.3f})")
    assert y.shape == (2, 1, 6, 6) and err <= tol


@pytest.mark.parametrize("dt,tol", [(torch.float16, 3e-3), (torch.bfloat16, 4e-2)])
def test_vgg_forward(golden, dt, tol):
    from fast_srgan_b200.model import VGG19
    sd = O.make_vgg19_state(seed=99)
    v = VGG19(compute_dtype=dt)
    v.load_state_dict(sd)
    v = v.cuda()
    x = seeded((1, 3, 32, 32), 13)
    with torch.no_grad():
        y = v(x.cuda()).cpu()
    ref = golden["vgg_y"]
    err = abs(y.numpy() - ref).max() / abs(ref).max()
    print(f"vgg19 {dt}: rel max err vs reference golden {err:.3e}")
    assert y.shape == ref.shape and err <= tol


_ORACLE_CACHE = {}


def _oracle_step(B, storage=None):
    """fp64 oracle of one step; storage=dtype additionally rounds every stored tensor like a 16-bit engine would.
    (cached: the CPU fp64 step is the slowest part of the GPU test-suite)"""
    if (B, storage) not in _ORACLE_CACHE:
        _ORACLE_CACHE[(B, storage)] = _oracle_step_uncached(B, storage)
    res, og, od, lr_img, hr_img, noise = _ORACLE_CACHE[(B, storage)]
    return res, {k: v.clone() for k, v in og.items()}, {k: v.clone() for k, v in od.items()}, lr_img, hr_img, noise


def _oracle_step_uncached(B, storage=None):
    import contextlib
    c = lambda sd: {k: v.double().clone() for k, v in sd.items()}
    og, od, ov = c(O.make_generator_state(64, 8, 1234)), c(O.make_discriminator_state(64, 4321)), c(O.make_vgg19_state(99))
    lr_img, hr_img = seeded((B, 3, 24, 24), 21), seeded((B, 3, 96, 96), 22)
    gn = torch.Generator().manual_seed(23)
    noise = {k: torch.rand((B, 1, 6, 6), generator=gn) for k in ("d_real", "d_fake", "g_real")}
    with (O.storage_rounding(storage) if storage is not None else contextlib.nullcontext()):
        res = O.gan_step(og, od, ov, lr_img.double(), hr_img.double(), {k: v.double() for k, v in noise.items()},
                         O.AdamWState(og, 1e-4), O.AdamWState(od, 1e-4))
    return res, og, od, lr_img, hr_img, noise


def _run_step(dt, B=2):
    from fast_srgan_b200.trainer import Trainer
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=8), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=dt)
    tr.generator.load_state_dict(O.make_generator_state(64, 8, 1234))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    tr.perceptual_network.load_state_dict(O.make_vgg19_state(99))
    res, og, od, lr_img, hr_img, noise = _oracle_step(B)
    out = tr.train_step(lr_img, hr_img, noise=noise)
    torch.cuda.synchronize()
    return tr, out, res, og, od


@pytest.mark.parametrize("dt,ltol", [(torch.bfloat16, 5e-3), (torch.float16, 1e-3)])
def test_gan_train_step_vs_fp64_oracle(golden, dt, ltol):
    """Losses: tight.  Gradients: these are ill-conditioned (InstanceNorm over 6x6..24x24 planes followed by
    kinked activations): ROUNDING THE STORED ACTIVATIONS ALONE - emulated in the fp64 oracle with
    storage_rounding(dtype), no other change - already moves them by ~10 % (fp16) / ~30 % (bf16) in relative L2
    (the reference's own fp32 run is ~1.5e-2 from its fp64 run).  The engine must be as good as that emulation:
        err_engine(tensor) <= 1.6 * err_emulation(tensor) + 0.03     (both measured against the fp64 oracle)."""
    tr, out, res, og, od = _run_step(dt)
    for k in ("loss_real", "loss_fake", "adv_loss", "content_loss"):
        ref = float(golden[f"step64_{k}"])         # the reference's own fp64 run (oracle == reference to 1e-9)
        got = out[k].item()
        print(f"{k}: engine {got:.6f} reference-fp64 {ref:.6f}")
        assert abs(got - ref) <= ltol * max(1.0, abs(ref))
    emu, _, _, _, _, _ = _oracle_step(2, storage=dt)
    e = tr.engine
    rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    for name, fp, key in (("D", e.dp, "d_grads"), ("G", e.gp, "g_grads")):
        scal = max([v.abs().item() for v in res[key].values() if v.numel() == 1] + [1e-30])
        for k, gref in res[key].items():
            got = fp.g[k].double().cpu() / e.S
            err, err_emu = rel(got, gref), rel(emu[key][k], gref)
            cos = (got.flatten() @ gref.flatten() / (got.norm() * gref.norm()).clamp_min(1e-30)).item()
            print(f"{name} grad {k:28s} engine {err:.3e}  16-bit-storage emulation {err_emu:.3e}  cos {cos:.4f}")
            if gref.numel() > 1:
                assert err <= 1.6 * err_emu + 0.03, (name, k, err, err_emu)
                assert cos >= (0.85 if dt == torch.bfloat16 else 0.95), (name, k, cos)
            else:
                # single PReLU slopes: a sum with heavy cancellation - one number has no norm to average over, so
                # bound its absolute error by the scale of this network's slope gradients
                aerr, aemu = (got - gref).abs().item(), (emu[key][k] - gref).abs().item()
                assert aerr <= max(3.0 * aemu, 0.1 * scal), (name, k, aerr, aemu, scal)
    # parameters after AdamW: the first step moves every weight by ~lr*sign(g) (m/sqrt(v) = +-1)
    for name, fp, ref_after, sd0 in (("D", e.dp, od, O.make_discriminator_state(64, 4321)),
                                     ("G", e.gp, og, O.make_generator_state(64, 8, 1234))):
        for k, pref in ref_after.items():
            got = fp.p[k].double().cpu()
            assert (got - pref).abs().max().item() <= 2.05e-4           # |delta| <= 2*lr (+wd) even if a sign flips
            upd_ref, upd_got = pref - sd0[k].double(), got - sd0[k].double()
            agree = (torch.sign(upd_ref) == torch.sign(upd_got)).double().mean().item()
            if pref.numel() > 1:
                assert agree >= 0.8, (name, k, agree)


def test_generator_forward_is_identical_before_and_after_the_d_step():
    """The engine evaluates G(lr) once per step where the reference evaluates it twice (trainer.py:173 and :185).  That is
    result-identical because only the discriminator changes in between and the forward is bitwise deterministic:
    checked here by running the discriminator half of a step between two evaluations."""
    tr, out, res, og, od = _run_step(torch.bfloat16)
    e = tr.engine
    g = torch.Generator().manual_seed(3)
    lr_img = (torch.rand((4, 3, 24, 24), generator=g) * 2 - 1).cuda()
    hr_img = (torch.rand((4, 3, 96, 96), generator=g) * 2 - 1).cuda()
    n = [torch.rand((4, 36), generator=g).cuda() for _ in range(3)]
    a, _ = e.G.forward(lr_img, save=False)
    a = a.clone()
    e._seg_d((lr_img, hr_img, n[0], n[1], n[2]))
    e._seg_d_update()                                   # discriminator AdamW: the only parameter change between :173 and :185
    b, _ = e.G.forward(lr_img, save=True)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(e._sr, b)


def test_gan_train_step_generator_width_128_vs_fp64_oracle():
    """BASELINE config #5 widths in TRAINING: generator.n_filters = 128 runs the same step on the general-channel kernels
    (CTA-pair conv for the 128-wide slices, grouped weight gradients with 4 (cin, cout) pairs per layer, F-channel
    pixel-shuffle backward).  Losses and generator gradients against the fp64 oracle."""
    import contextlib
    from fast_srgan_b200.trainer import Trainer
    dt, B, Fm, Lb = torch.float16, 2, 128, 2
    c = lambda sd: {k: v.double().clone() for k, v in sd.items()}
    og, od, ov = c(O.make_generator_state(Fm, Lb, 77)), c(O.make_discriminator_state(64, 4321)), c(O.make_vgg19_state(99))
    lr_img, hr_img = seeded((B, 3, 24, 24), 21), seeded((B, 3, 96, 96), 22)
    gn = torch.Generator().manual_seed(23)
    noise = {k: torch.rand((B, 1, 6, 6), generator=gn) for k in ("d_real", "d_fake", "g_real")}
    res = O.gan_step(og, od, ov, lr_img.double(), hr_img.double(), {k: v.double() for k, v in noise.items()},
                     O.AdamWState(og, 1e-4), O.AdamWState(od, 1e-4))
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=Fm, n_layers=Lb), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=dt, vgg_state_dict=O.make_vgg19_state(99))
    tr.generator.load_state_dict(O.make_generator_state(Fm, Lb, 77))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    out = tr.train_step(lr_img, hr_img, noise=noise)
    torch.cuda.synchronize()
    for k in ("loss_real", "loss_fake", "adv_loss", "content_loss"):
        print(f"{k}: engine {out[k].item():.6f} oracle {res[k].item():.6f}")
        assert abs(out[k].item() - res[k].item()) <= 1.5e-3 * max(1.0, abs(res[k].item()))
    e = tr.engine
    for k, gref in res["g_grads"].items():
        got = e.gp.g[k].double().cpu() / e.S
        if gref.numel() > 1:
            cos = (got.flatten() @ gref.flatten() / (got.norm() * gref.norm()).clamp_min(1e-30)).item()
            rel = ((got - gref).norm() / gref.norm().clamp_min(1e-30)).item()
            print(f"G(F=128) grad {k:28s} rel-L2 {rel:.3e} cos {cos:.4f}")
            assert cos >= 0.95, (k, cos)


def test_train_step_updates_inference_weights():
    """After a step the Generator module (inference path) must see the updated parameters."""
    tr, out, res, og, od = _run_step(torch.bfloat16)
    x = seeded((1, 3, 24, 24), 5)
    with torch.no_grad():
        y = tr.generator(x.cuda()).cpu()
        ref = O.generator_forward({k: v.float() for k, v in og.items()}, x)
    assert (y - ref).abs().max().item() <= 3e-2


def test_pretrain_step_runs_and_reduces_loss():
    from fast_srgan_b200.trainer import Trainer
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=2), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-3, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=torch.bfloat16)
    tr.generator.load_state_dict(O.make_generator_state(64, 2, 7))
    lr_img = seeded((4, 3, 16, 16), 1)
    hr_img = torch.nn.functional.interpolate(lr_img, scale_factor=4, mode="bilinear")
    l0 = tr.pretrain_step(lr_img, hr_img)["loss"].item()
    for _ in range(30):
        l1 = tr.pretrain_step(lr_img, hr_img)["loss"].item()
    print("pretrain loss", l0, "->", l1)
    assert l1 < l0


def test_cuda_graph_train_step_matches_eager():
    """Steps 3+ replay a captured CUDA graph (device-side AdamW step counter): same trajectory as eager execution up to
    the usual Adam sign flips of noise-level gradients (<= 2*lr per step per weight)."""
    from fast_srgan_b200.trainer import Trainer

    def run(use_graph):
        cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=2), discriminator=ns(n_filters=64, n_layers=7),
                 training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
        tr = Trainer(cfg, compute_dtype=torch.bfloat16)
        tr.generator.load_state_dict(O.make_generator_state(64, 2, 1234))
        tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
        tr.perceptual_network.load_state_dict(O.make_vgg19_state(99))
        tr.engine.use_graph = use_graph
        losses = []
        for step in range(5):
            g = torch.Generator().manual_seed(100 + step)
            lr_img, hr_img = torch.rand((2, 3, 24, 24), generator=g) * 2 - 1, torch.rand((2, 3, 96, 96), generator=g) * 2 - 1
            noise = {k: torch.rand((2, 1, 6, 6), generator=g) for k in ("d_real", "d_fake", "g_real")}
            out = tr.train_step(lr_img, hr_img, noise=noise)
            losses.append([out[k].item() for k in ("loss_real", "loss_fake", "adv_loss", "content_loss")])
        torch.cuda.synchronize()
        e = tr.engine
        assert e.gp.step_count == 5 and int(e.gp.step_dev.item()) == 5 and int(e.dp.step_dev.item()) == 5
        # the packed weight buffers the (captured) kernels read must hold the CURRENT parameters: packs are in-place into
        # persistent buffers, never re-allocated (a graph would otherwise keep reading stale/freed memory)
        from fast_srgan_b200 import ops
        wd, _ = ops.pack_conv3x3(e.dp.p["stem.3.conv.weight"], None, torch.bfloat16)
        assert torch.equal(wd, e.D.P["w3"])
        wg, _ = ops.pack_conv3x3(e.gp.p["stem.1.conv2.weight"], None, torch.bfloat16)
        # G was updated by the last AdamW after its last pack: the buffer holds the weights used by the last forward
        assert wg.shape == e.G.P["stem.1.conv2.weight"].shape
        return torch.tensor(losses), e.gp.flat.clone(), e.dp.flat.clone()

    le, ge, de = run(False)
    lg, gg, dg = run(True)
    print("eager losses", le[-1].tolist(), "graph losses", lg[-1].tolist())
    assert (le - lg).abs().max().item() <= 2e-2
    assert (ge - gg).abs().max().item() <= 5 * 2.05e-4 and (de - dg).abs().max().item() <= 5 * 2.05e-4


def test_backward_is_bitwise_reproducible_in_its_large_reductions():
    """Round 2: the weight gradients (two-stage split-K reduction in fixed order) and the InstanceNorm backward (in-block
    sums) use no floating-point atomics, the small cross-block sums (bias, PReLU-slope and 3-channel first-layer weight
    gradients, loss sums) go through fixed-point integer atomics (fsr_common.cuh DetRed), and the forward was already
    bitwise deterministic - so EVERY gradient of both networks is IDENTICAL run to run."""
    from fast_srgan_b200.trainer import Trainer
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=3), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=torch.bfloat16, vgg_state_dict=O.make_vgg19_state(99))
    tr.generator.load_state_dict(O.make_generator_state(64, 3, 1234))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    e = tr.engine
    e.use_graph = False
    B = 8
    g = torch.Generator().manual_seed(5)
    lr = (torch.rand((B, 3, 24, 24), generator=g) * 2 - 1).cuda()
    hr = (torch.rand((B, 3, 96, 96), generator=g) * 2 - 1).cuda()
    noise = [torch.rand((B, 36), generator=g).cuda() for _ in range(3)]
    ins = (lr, hr, noise[0], noise[1], noise[2])

    def grads():
        e._seg_d(ins)
        e._seg_content(ins)
        e._seg_adv_and_g(ins)                       # no optimizer step: parameters stay put, gradients are recomputed
        torch.cuda.synchronize()
        return e.dp.grad.clone(), e.gp.grad.clone()

    d1, g1 = grads()
    d2, g2 = grads()
    assert torch.equal(d1, d2) and torch.equal(g1, g2), "a gradient differs run to run"
    for i in range(7):
        k = f"stem.{i}.conv.weight"
        o, n = e.dp.offsets[k], e.dp.p[k].numel()
        assert torch.equal(d1[o:o + n], d2[o:o + n]), k
    for k in e.G._convs64() + ["upsampling.0.conv.weight", "upsampling.1.conv.weight"]:
        o, n = e.gp.offsets[k], e.gp.p[k].numel()
        assert torch.equal(g1[o:o + n], g2[o:o + n]), k


def test_gradients_shard_exactly_over_batch():
    """Size-independent property behind the multi-GPU path: every op is per-sample, losses are batch means,
    so grad(full batch) == mean over shards of grad(shard).  Checked on the discriminator step at BASELINE-like size."""
    from fast_srgan_b200.trainer import Trainer
    cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=2), discriminator=ns(n_filters=64, n_layers=7),
             training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
    tr = Trainer(cfg, compute_dtype=torch.bfloat16)
    tr.generator.load_state_dict(O.make_generator_state(64, 2, 1234))
    tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
    tr.perceptual_network.load_state_dict(O.make_vgg19_state(99))
    e = tr.engine
    B = 16
    g = torch.Generator().manual_seed(77)
    lr = (torch.rand((B, 3, 24, 24), generator=g) * 2 - 1).cuda()
    hr = (torch.rand((B, 3, 96, 96), generator=g) * 2 - 1).cuda()
    n1, n2, n3 = (torch.rand((B, 36), generator=g).cuda() for _ in range(3))

    def d_grad(sl):
        e._seg_d((lr[sl].contiguous(), hr[sl].contiguous(), n1[sl].contiguous(), n2[sl].contiguous(), n3[sl].contiguous()))
        torch.cuda.synchronize()
        return e.dp.grad.clone()

    full = d_grad(slice(0, B))
    halves = 0.5 * (d_grad(slice(0, B // 2)) + d_grad(slice(B // 2, B)))
    rel = ((full - halves).norm() / full.norm()).item()
    again = d_grad(slice(0, B))
    noise = ((full - again).norm() / full.norm()).item()
    print("full-batch vs mean-of-shards D gradient: rel-L2", rel, " run-to-run", noise)
    # forward and backward are bitwise reproducible (fixed-point statistics, fixed-order / integer reductions): run-to-run
    # noise is exactly 0; sharding changes the order of the batch sum inside the weight gradients only (measured 2e-3)
    assert noise == 0.0 and rel <= 6e-3


def test_training_is_bitwise_reproducible_across_trainers():
    """A restart reproduces a run: two independently constructed Trainers (same seeds, same inputs) take four GAN steps
    and end with bit-identical generator / discriminator parameters, optimizer moments and reported losses
    (trainer.py:168-196).  Both the captured-graph path and the eager path are exercised."""
    from fast_srgan_b200.trainer import Trainer
    B = 8
    g = torch.Generator().manual_seed(11)
    lr = (torch.rand((B, 3, 24, 24), generator=g) * 2 - 1).cuda()
    hr = (torch.rand((B, 3, 96, 96), generator=g) * 2 - 1).cuda()
    noise = {k: torch.rand((B, 1, 6, 6), generator=g).cuda() for k in ("d_real", "d_fake", "g_real")}

    def run(use_graph):
        cfg = ns(experiment=ns(name="t", seed=0), generator=ns(n_filters=64, n_layers=3), discriminator=ns(n_filters=64, n_layers=7),
                 training=ns(device="cuda", generator_lr=1e-4, discriminator_lr=1e-4))
        tr = Trainer(cfg, compute_dtype=torch.bfloat16, vgg_state_dict=O.make_vgg19_state(99))
        tr.generator.load_state_dict(O.make_generator_state(64, 3, 1234))
        tr.discriminator.load_state_dict(O.make_discriminator_state(64, 4321))
        tr.engine.use_graph = use_graph
        losses = []
        for _ in range(4):
            out = tr.train_step(lr, hr, noise=noise)
            losses.append([float(out[k]) for k in ("loss_real", "loss_fake", "adv_loss", "content_loss")])
        torch.cuda.synchronize()
        e = tr.engine
        return e.gp.flat.clone(), e.dp.flat.clone(), e.gp.m.clone(), e.dp.v.clone(), losses

    a = run(True)
    b = run(True)
    c = run(False)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    for x, y in zip(a[:4], c[:4]):
        assert torch.equal(x, y)
    assert a[4] == b[4] == c[4]
