"""n_filters <= 32 generators on pixel-pair rows (fast_srgan_b200/pairs.py; reference model.py:72-117 with
generator.n_filters = 32, BASELINE configs[4]) against the oracle, and against the zero-padded 64-channel path."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import srgan_oracle as O  # noqa: E402


def seeded(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def make(Fm, L, dt, seed=77):
    from fast_srgan_b200.model import Generator
    sd = O.make_generator_state(Fm, L, seed=seed)
    g = Generator(types.SimpleNamespace(n_filters=Fm, n_layers=L), compute_dtype=dt)
    g.load_state_dict(sd)
    return g.cuda().eval(), sd


@pytest.mark.parametrize("F,L,shape", [(32, 2, (2, 3, 24, 40)), (32, 8, (1, 3, 45, 80)), (32, 1, (3, 3, 7, 2)),
                                       (32, 3, (1, 3, 1, 6)), (16, 2, (1, 3, 20, 18)), (32, 4, (1, 3, 37, 254))])
def test_pair_generator_vs_oracle(F, L, shape):
    g, sd = make(F, L, torch.float16)
    x = seeded(shape, 41)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
        ref = O.generator_forward(sd, x)
    assert g._pair is not None, "pair path not taken"
    err = (y - ref).abs().max().item()
    print(f"pair generator F={F} L={L} {shape}: max-abs {err:.3e}")
    assert y.shape == ref.shape and err <= 1e-3            # north_star tolerance


def test_pair_generator_bf16_and_padded_path_agree(monkeypatch):
    g, sd = make(32, 4, torch.bfloat16)
    x = seeded((2, 3, 30, 44), 43)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
        ref = O.generator_forward(sd, x)
    assert (y - ref).abs().max().item() <= 1.5e-2         # bf16 bound of tests/test_generator_gpu.py
    g16, _ = make(32, 4, torch.float16)
    with torch.no_grad():
        a = g16(x.cuda()).cpu()
        monkeypatch.setenv("FSR_PAIR32", "0")
        b = g16(x.cuda()).cpu()
    d = (a - b).abs().max().item()
    print(f"pair vs zero-padded 64-channel path: {d:.3e}")
    assert d <= 1e-3


def test_pair_generator_uint8_pipeline():
    """inference.py:48-56 through the pair path: uint8 in -> uint8 out within +-1 LSB of the oracle's."""
    g, sd = make(32, 2, torch.float16)
    gen = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 24, 32, 3), generator=gen, dtype=torch.uint8)
    with torch.no_grad():
        got = g.super_resolve_u8(img.cuda()).cpu()
        x = (img.float() / 127.5 - 1.0).permute(0, 3, 1, 2).contiguous()
        ref = ((O.generator_forward(sd, x).permute(0, 2, 3, 1) + 1.0) / 2.0 * 255.0).clamp(0, 255).to(torch.uint8)
    diff = (got.int() - ref.int()).abs()
    assert got.shape == ref.shape and diff.max().item() <= 1 and (diff > 0).float().mean().item() < 0.02


def test_pair_generator_is_deterministic_and_batch_invariant():
    g, _ = make(32, 3, torch.float16)
    x = seeded((3, 3, 24, 40), 5).cuda()
    with torch.no_grad():
        a = g(x).clone()
        b = g(x).clone()
        c = g(x[1:2]).clone()
    assert torch.equal(a, b) and torch.equal(a[1:2], c)
