"""GPU parity of the whole Generator.forward (model.py:112-117) against the CPU oracle and the
committed reference goldens.  north_star tolerance: 4x SR output within 1e-3 max-abs of the
reference on random-init weights (fp16 operands + fp32 accumulation; bf16 has its own bound)."""
import types

import numpy as np
import pytest
import torch

import srgan_oracle as O

pytestmark = pytest.mark.gpu
TOL = {torch.float16: 1e-3, torch.bfloat16: 1.5e-2}


def seeded(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * 2 - 1


def make(Fm, L, dt, seed=1234):
    from fast_srgan_b200.model import Generator
    sd = O.make_generator_state(Fm, L, seed=seed)
    g = Generator(types.SimpleNamespace(n_filters=Fm, n_layers=L), compute_dtype=dt)
    g.load_state_dict(sd)
    return g.cuda().eval(), sd


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_generator_vs_reference_golden(golden, dt):
    g, sd = make(64, 8, dt)
    x = seeded((1, 3, 20, 24), 7)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
    err = np.abs(y.numpy() - golden["g64x8_y"]).max()
    print(f"generator {dt} vs reference golden: max-abs {err:.3e}")
    assert err <= TOL[dt]


@pytest.mark.parametrize("shape,L", [((2, 3, 24, 24), 8), ((1, 3, 37, 53), 4), ((3, 3, 16, 16), 0), ((1, 3, 90, 160), 8)])
def test_generator_vs_oracle_shapes(shape, L):
    g, sd = make(64, L, torch.float16)
    x = seeded(shape, 31)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
        ref = O.generator_forward(sd, x)
    err = (y - ref).abs().max().item()
    print(f"generator {shape} L={L}: max-abs {err:.3e}")
    assert y.shape == ref.shape and err <= 1e-3


def test_generator_l2_groups_identical():
    """Processing the residual chain in L2-sized image groups must not change a single bit (InstanceNorm statistics
    are accumulated with order-independent fixed-point integer atomics)."""
    g, _ = make(64, 3, torch.float16)
    x = seeded((5, 3, 24, 40), 5).cuda()
    with torch.no_grad():
        g.l2_group = 0
        a = g(x).clone()
        g.l2_group = 2
        b = g(x).clone()
    assert torch.equal(a, b)


def test_generator_uint8_pipeline():
    """inference.py:48-56 fused path: uint8 in -> uint8 out equals the oracle's uint8 up to +-1 LSB."""
    g, sd = make(64, 8, torch.float16)
    gen = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 24, 32, 3), generator=gen, dtype=torch.uint8)
    out = g.super_resolve_u8(img.cuda()).cpu()
    ref = O.to_uint8_image(O.generator_forward(sd, O.from_uint8_image(img)))
    diff = (out.int() - ref.int()).abs()
    assert out.shape == ref.shape and diff.max().item() <= 1
    assert (diff > 0).float().mean().item() < 0.02


def test_generator_batch_independence_fullsize():
    """Size-independent property at a BASELINE-sized frame: InstanceNorm is per-sample, so a
    frame's output does not depend on its batch neighbours."""
    g, _ = make(64, 8, torch.float16)
    x = seeded((3, 3, 180, 320), 9).cuda()
    with torch.no_grad():
        full = g(x)
        single = g(x[1:2])
    assert torch.equal(full[1:2], single)                   # bitwise: integer (fixed-point) statistics atomics
    assert torch.isfinite(full).all() and full.abs().max().item() <= 1.0


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_generator_f32_vs_reference_golden(golden, dt):
    """n_filters = 32 (BASELINE config #5): zero-padded to 64-channel rows; must equal the unpadded reference network."""
    g, sd = make(32, 2, dt)
    x = seeded((2, 3, 9, 13), 7)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
    err = np.abs(y.numpy() - golden["g32x2_y"]).max()
    print(f"generator F=32 {dt} vs reference golden: max-abs {err:.3e}")
    assert err <= TOL[dt]


@pytest.mark.parametrize("F,L,shape", [(128, 2, (2, 3, 24, 40)), (128, 4, (1, 3, 37, 21)), (32, 8, (1, 3, 45, 80))])
def test_generator_other_widths_vs_oracle(F, L, shape):
    """n_filters = 128 runs on the general-channel kernels (res convs, pixel-shuffle epilogue, K-looped head)."""
    g, sd = make(F, L, torch.float16)
    x = seeded(shape, 17)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
        ref = O.generator_forward(sd, x)
        u8 = g.super_resolve_u8(((x.permute(0, 2, 3, 1) + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).cuda()).cpu()
    err = (y - ref).abs().max().item()
    print(f"generator F={F} L={L} {shape}: max-abs {err:.3e}")
    assert y.shape == ref.shape and err <= 1e-3 and u8.shape == (shape[0], 4 * shape[2], 4 * shape[3], 3)
