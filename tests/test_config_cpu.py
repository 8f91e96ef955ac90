def test_config_loader_schema_and_overrides():
    from fast_srgan_b200 import config
    c = config.load(overrides=["training.batch_size=64", "generator.n_layers=4", "experiment.name=x"])
    assert c.generator.n_filters == 64 and c.generator.n_layers == 4 and c.training.batch_size == 64
    assert isinstance(c.training.generator_lr, float) and abs(c.training.generator_lr - 1e-4) < 1e-12   # PyYAML gives '1e-4'
    assert c.experiment.name == "x" and c.training.compiled is False
