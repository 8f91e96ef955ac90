def test_config_loader_schema_and_overrides():
    from fast_srgan_b200 import config
    c = config.load(overrides=["training.batch_size=64", "generator.n_layers=4", "experiment.name=x"])
    assert c.generator.n_filters == 64 and c.generator.n_layers == 4 and c.training.batch_size == 64
    assert isinstance(c.training.generator_lr, float) and abs(c.training.generator_lr - 1e-4) < 1e-12   # PyYAML gives '1e-4'
    assert c.experiment.name == "x" and c.training.compiled is False


def test_flat_params_optimizer_state_round_trip():
    """Checkpoint / resume host logic (trainer.py:143-156, 90-94) on CPU: the AdamW state is written in
    torch.optim.AdamW.state_dict() format, survives save -> load into a freshly built network, loads into a GENUINE
    torch.optim.AdamW over the same parameters (and back), and a layout mismatch is refused."""
    import io
    import types

    import pytest
    import torch
    from fast_srgan_b200.engine import FlatParams
    from fast_srgan_b200.model import Discriminator, Generator
    d1 = Discriminator(types.SimpleNamespace(n_filters=64))
    f1 = FlatParams(d1)
    f1.m.normal_(); f1.v.uniform_(); f1.step_count = 7
    buf = io.BytesIO()
    torch.save(f1.optimizer_state(lr=1e-4), buf)
    buf.seek(0)
    d2 = Discriminator(types.SimpleNamespace(n_filters=64))
    f2 = FlatParams(d2)
    f2.load_optimizer_state(torch.load(buf))
    assert f2.step_count == 7 and int(f2.step_dev.item()) == 7
    for n in f1.names:
        o, k = f1.offsets[n], f1.p[n].numel()
        assert torch.equal(f2.m[o:o + k], f1.m[o:o + k]) and torch.equal(f2.v[o:o + k], f1.v[o:o + k])
    # the reference's optimizer (trainer.py:33-38) accepts the file ...
    opt = torch.optim.AdamW(d2.parameters(), lr=1e-4)
    opt.load_state_dict(f1.optimizer_state(lr=1e-4))
    p0 = next(iter(d2.parameters()))
    assert torch.equal(opt.state[p0]["exp_avg"], f1.m[:p0.numel()].view_as(p0)) and float(opt.state[p0]["step"]) == 7.0
    # ... and its own state_dict loads here
    f3 = FlatParams(Discriminator(types.SimpleNamespace(n_filters=64)))
    f3.load_optimizer_state(opt.state_dict())
    assert f3.step_count == 7 and torch.equal(f3.m[:p0.numel()], f1.m[:p0.numel()])
    g = FlatParams(Generator(types.SimpleNamespace(n_filters=64, n_layers=2)))
    with pytest.raises(RuntimeError):
        g.load_optimizer_state(f1.optimizer_state())


def test_vgg19_accepts_torchvision_keys_and_versions_its_weights():
    """Trainer must not train against a silently random VGG, and a load after the first use must be seen."""
    import torch
    from fast_srgan_b200.model import VGG19
    v = VGG19()
    assert v.weights_loaded is False and v._weights_version == 0
    tv = {}
    for idx, conv in v.vgg.items():
        tv[f"features.{idx}.weight"] = torch.full_like(conv.weight, 0.5)
        tv[f"features.{idx}.bias"] = torch.full_like(conv.bias, -1.0)
    tv["features.34.weight"] = torch.zeros(512, 512, 3, 3)          # layers past features[:34] are ignored
    tv["classifier.0.weight"] = torch.zeros(8, 8)
    v.load_state_dict(tv)
    assert v.weights_loaded and v._weights_version == 1
    assert float(v.vgg["0"].weight[0, 0, 0, 0]) == 0.5 and float(v.vgg["32"].bias[0]) == -1.0
    v.load_state_dict(v.state_dict())                              # the reference's own key names (vgg.{idx}.*, mean, std)
    assert v._weights_version == 2
