"""ABI re-entrancy: two host threads, each with its OWN context (fsr_ctx_create / fsr_ctx_bind: option
overrides + internal side streams), stream, generator and buffers, run Generator.forward concurrently; every result
must equal what the same generator produces alone.  One thread uses the round-1 kernels (no fused input transform, single-
CTA upsampling conv, 3 sub-batches on the context's side streams), the other the defaults."""
import ctypes
import threading
import types

import pytest
import torch

import srgan_oracle as O

pytestmark = pytest.mark.gpu


def test_two_contexts_two_streams_concurrently():
    from fast_srgan_b200 import _lib as L
    from fast_srgan_b200.model import Generator
    lib = L.load()
    gens, xs, refs = [], [], []
    for seed in (1, 2):
        g = Generator(types.SimpleNamespace(n_filters=64, n_layers=4), compute_dtype=torch.float16)
        g.load_state_dict(O.make_generator_state(64, 4, seed=seed))
        g = g.cuda().eval()
        gen = torch.Generator().manual_seed(10 + seed)
        x = (torch.rand((6, 3, 40, 56), generator=gen) * 2 - 1).cuda()
        with torch.no_grad():
            refs.append(g(x).clone())                      # alone, process defaults (all variants are bit-identical)
        gens.append(g)
        xs.append(x)
    torch.cuda.synchronize()
    opts = [{L.OPT_FUSE_IN: 0, L.OPT_UP_2CTA: 0, L.OPT_OVERLAP_STREAMS: 3}, {}]
    outs, errs = [[], []], []

    def worker(i):
        try:
            ctx = ctypes.c_void_p()
            L.check(lib.fsr_ctx_create(ctypes.byref(ctx)))
            for k, v in opts[i].items():
                L.check(lib.fsr_ctx_set(ctx, k, v))
            L.check(lib.fsr_ctx_bind(ctx))
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream), torch.no_grad():
                for _ in range(25):
                    outs[i].append(gens[i](xs[i]).clone())
            stream.synchronize()
            lib.fsr_ctx_bind(None)
            L.check(lib.fsr_ctx_destroy(ctx))
        except Exception as exc:                           # surfaced in the main thread
            errs.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    torch.cuda.synchronize()
    for i in range(2):
        assert len(outs[i]) == 25
        for y in outs[i]:
            assert torch.equal(y, refs[i])
