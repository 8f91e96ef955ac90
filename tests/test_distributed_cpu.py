"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch / label-noise sharding, the flat gradient
buffer layout and the all-reduce, checked with the oracle: mean over ranks of shard gradients == full-batch gradient
(InstanceNorm is per-sample, so batch sharding is exact - SURVEY.md 8e)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import types
    import srgan_oracle as O
    from fast_srgan_b200 import distributed as D
    from fast_srgan_b200.engine import FlatParams
    from fast_srgan_b200.model import Discriminator
    torch.set_num_threads(2)
    r, _, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    B = 4
    g = torch.Generator().manual_seed(3)
    hr = torch.rand((B, 3, 32, 32), generator=g) * 2 - 1
    noise = {"d_real": torch.rand((B, 1, 2, 2), generator=g)}
    dsd = O.make_discriminator_state(64, seed=4321)

    def d_grads(x, n):
        d = {k: v.double().clone().requires_grad_(True) for k, v in dsd.items()}
        O.bce_with_logits_mean(O.discriminator_forward(d, x.double()), (0.3 * n + 0.8).double()).backward()
        return {k: v.grad for k, v in d.items()}

    # this rank's shard -> gradients into the flat buffer the engine all-reduces
    mod = Discriminator(types.SimpleNamespace(n_filters=64))
    mod.load_state_dict(dsd)
    fp = FlatParams(mod, with_optimizer=False)
    assert fp.aliases(mod) and fp.numel >= sum(p.numel() for p in mod.parameters())
    gl = d_grads(D.shard_batch(hr, rank, world), D.shard_noise(noise, rank, world)["d_real"])
    for k, v in gl.items():
        fp.g[k].copy_(v.float())
    comm = D.FlatComm(torch.device("cpu"))         # CPU / gloo: the torch.distributed fallback of the engine's exchange
    assert not comm.native and (comm.rank, comm.world) == (rank, world)
    comm.allreduce(fp.grad)
    fp.grad.mul_(1.0 / world)                      # the factor the engine folds into AdamW
    probe = torch.full((5,), float(rank + 1))
    comm.broadcast(probe)                          # replica initialisation: rank 0's buffer wins
    assert torch.equal(probe, torch.ones(5))
    full = d_grads(hr, noise["d_real"])
    worst = max(((fp.g[k].double() - full[k]).norm() / full[k].norm()).item() for k in full)
    ret[rank] = worst
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert len(ret) == world and max(ret.values()) <= 1e-5, dict(ret)


def test_shard_ranges():
    sys.path.insert(0, ROOT)
    from fast_srgan_b200 import distributed as D
    assert [D.shard_range(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]
    try:
        D.shard_range(10, 0, 4)
        assert False
    except ValueError:
        pass


def _aux_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import aux_oracle as A
    from fast_srgan_b200 import distributed as D
    from fast_srgan_b200.data import ShardedReplacementSampler
    from fast_srgan_b200.metrics import ValidationMetrics
    torch.set_num_threads(2)
    D.init_from_env("gloo")
    # (1) data path: every rank draws the same global index stream and keeps its slice of each global batch
    mine = list(ShardedReplacementSampler(37, 8 * 5 + 3, 8, seed=11, rank=rank, world=world))
    gathered = [None] * world
    dist.all_gather_object(gathered, [b.tolist() for b in mine])
    ref = A.replacement_sample_indices(37, 43, 11)
    ok_sampler = all(sum((gathered[r][b] for r in range(world)), []) == ref[8 * b:8 * b + 8].tolist() for b in range(5))
    # (2) validation metrics: each rank's partial sums (its shard of the validation set) pooled by one all-reduce
    g = torch.Generator().manual_seed(5)
    sr = torch.rand((4, 3, 24, 28), generator=g) * 2 - 1
    hr = torch.rand((4, 3, 24, 28), generator=g) * 2 - 1
    a, b = (1 + D.shard_batch(sr, rank, world)) / 2, (1 + D.shard_batch(hr, rank, world)) / 2
    ss = A.ssim_per_image(a, b)
    stat = torch.tensor([float(((a - b).double() ** 2).sum()), float(a.numel()), float(ss.double().sum()), float(ss.numel())],
                        dtype=torch.float64)
    out = ValidationMetrics.pool(stat, 1.0, sync=True)
    ssim_ref, psnr_ref = A.validation_metrics([sr], [hr])
    ret[rank] = (ok_sampler, abs(out["ssim"] - ssim_ref), abs(out["psnr"] - psnr_ref))
    dist.destroy_process_group()


def test_two_rank_sampler_shards_and_metric_pooling():
    """SURVEY 8e for the f3 / f4 rows: W ranks at B/W see exactly one process's batches; pooled PSNR / SSIM equal the
    single-process values (torchmetrics' sum / cat reductions)."""
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_aux_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert len(ret) == world
    for ok, d_ssim, d_psnr in ret.values():
        assert ok and d_ssim <= 1e-6 and d_psnr <= 1e-6
