def test_config_loader_schema_and_overrides():
    from fast_srgan_b200 import config
    c = config.load(overrides=["training.batch_size=64", "generator.n_layers=4", "experiment.name=x"])
    assert c.generator.n_filters == 64 and c.generator.n_layers == 4 and c.training.batch_size == 64
    assert isinstance(c.training.generator_lr, float) and abs(c.training.generator_lr - 1e-4) < 1e-12   # PyYAML gives '1e-4'
    assert c.experiment.name == "x" and c.training.compiled is False


def test_flat_params_optimizer_state_round_trip():
    """Checkpoint / resume host logic (trainer.py:143-156, 90-94) on CPU: the flat AdamW state survives save -> load into a
    freshly built network and a layout mismatch is refused."""
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
    torch.save(f1.optimizer_state(), buf)
    buf.seek(0)
    f2 = FlatParams(Discriminator(types.SimpleNamespace(n_filters=64)))
    f2.load_optimizer_state(torch.load(buf))
    assert f2.step_count == 7 and int(f2.step_dev.item()) == 7
    assert torch.equal(f2.m, f1.m) and torch.equal(f2.v, f1.v)
    g = FlatParams(Generator(types.SimpleNamespace(n_filters=64, n_layers=2)))
    with pytest.raises(RuntimeError):
        g.load_optimizer_state(f1.optimizer_state())
