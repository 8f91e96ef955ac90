"""GPU parity tests of the training-step kernels (general conv forward / dgrad / wgrad, norm and activation
backward, losses, AdamW) through the C ABI against PyTorch fp32 autograd on the same rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
EPS = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}
DT = [torch.float16, torch.bfloat16]


@pytest.fixture(autouse=True, params=[1, 0], ids=["gen-ws", "gen-stream"])
def _setup(request):
    """Every test runs with the general conv weight-stationary over 4-tile groups (conv3x3_gen_ws.cuh) and with the
    per-tile weight-streaming kernel (conv3x3_gen.cuh)."""
    from fast_srgan_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _lib.load().fsr_set_gen_ws(request.param)
    yield
    _lib.load().fsr_set_gen_ws(-1)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(dt)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (2, 13, 21)), (128, 64, (1, 24, 24)), (64, 128, (2, 16, 8)),
                                            (256, 512, (2, 12, 12)), (512, 512, (3, 6, 6))])
def test_gen_conv_fwd_stride1(dt, cin, cout, shape):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = shape
    x = nhwc(rnd((N, cin, H, W), 1), dt)
    w = rnd((cout, cin, 3, 3), 2, (cin * 9) ** -0.5).to(dt).float()
    b = rnd((cout,), 3, 0.1)
    wp, bp = ops.pack_conv3x3(w, b, dt)
    got = nchw(ops.conv3x3_gen(x, wp, cout, bias=bp, act=L.ACT_RELU))
    ref = F.relu(F.conv2d(nchw(x), w, b, padding=1))
    assert rel_err(got, ref) <= 2 * EPS[dt] + 1e-5
    raw, st = ops.conv3x3_gen(x, wp, cout, epilogue=L.EPI_RAW_STATS)
    ref2 = F.conv2d(nchw(x), w, padding=1)
    g2 = nchw(raw)
    assert rel_err(g2, ref2) <= 2 * EPS[dt] + 1e-5
    s_got = torch.stack([g2.sum((2, 3)), (g2 * g2).sum((2, 3))], dim=-1)
    assert torch.allclose(ops.stats_to_float(st).float(), s_got, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (2, 16, 24)), (128, 128, (1, 48, 48)), (512, 512, (2, 12, 12))])
def test_gen_conv_fwd_stride2(dt, cin, cout, shape):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = shape
    x = nhwc(rnd((N, cin, H, W), 4), dt)
    w = rnd((cout, cin, 3, 3), 5, (cin * 9) ** -0.5).to(dt).float()
    wp, _ = ops.pack_conv3x3(w, None, dt)
    xp = ops.parity_layout(x, True)
    assert torch.equal(ops.parity_layout(xp, False), x)
    raw, st = ops.conv3x3_gen(xp, wp, cout, stride=2, epilogue=L.EPI_RAW_STATS)
    ref = F.conv2d(nchw(x), w, stride=2, padding=1)
    assert rel_err(nchw(raw), ref) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,shape,stride", [(64, 64, (2, 13, 21), 1), (64, 128, (1, 24, 24), 1), (256, 128, (2, 12, 12), 1),
                                                   (64, 64, (2, 16, 24), 2), (128, 256, (1, 24, 24), 2)])
def test_gen_conv_dgrad(dt, cin, cout, shape, stride):
    from fast_srgan_b200 import ops
    N, H, W = shape
    w = rnd((cout, cin, 3, 3), 6, (cout * 9) ** -0.5).to(dt).float()
    dy = nhwc(rnd((N, cout, H // stride, W // stride), 7), dt)
    xr = torch.zeros((N, cin, H, W), device="cuda", requires_grad=True)
    F.conv2d(xr, w, stride=stride, padding=1).backward(nchw(dy))
    wt = ops.pack_conv3x3_t(w, dt)
    dx = ops.conv3x3_gen(dy, wt, cin, stride=stride, mode=1)
    if stride == 2:
        dx = ops.parity_layout(dx, False)
    assert rel_err(nchw(dx), xr.grad) <= 2 * EPS[dt] + 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cin,cout,shape,stride", [(64, 64, (2, 13, 21), 1), (64, 64, (4, 24, 24), 1), (128, 64, (1, 16, 8), 1),
                                                   (64, 256, (2, 24, 24), 1), (64, 64, (2, 16, 24), 2), (128, 256, (2, 24, 24), 2)])
def test_conv_wgrad(dt, cin, cout, shape, stride):
    from fast_srgan_b200 import ops
    N, H, W = shape
    x = nhwc(rnd((N, cin, H, W), 8), dt)
    dy = nhwc(rnd((N, cout, H // stride, W // stride), 9), dt)
    wr = torch.zeros((cout, cin, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(nchw(x), wr, stride=stride, padding=1).backward(nchw(dy))
    dw = torch.zeros((cout, cin, 3, 3), device="cuda")
    ops.conv3x3_wgrad(ops.parity_layout(x, True) if stride == 2 else x, dy, dw, stride=stride)
    assert rel_err(dw, wr.grad) <= 1e-4          # exact products, fp32 accumulation: only summation order differs
    ops.conv3x3_wgrad(ops.parity_layout(x, True) if stride == 2 else x, dy, dw, stride=stride)   # accumulates
    assert rel_err(dw, 2 * wr.grad) <= 1e-4


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("G,shape", [(1, (2, 24, 24)), (5, (3, 24, 24)), (17, (4, 24, 24)), (3, (2, 13, 21))])
def test_conv_wgrad_grouped_and_deterministic(dt, G, shape):
    """ONE grouped launch for G weight gradients of identical shape (the generator's 2L+1 residual-chain convs) ==
    G separate launches == PyTorch autograd; and the two-stage (no atomics) reduction is bitwise reproducible."""
    from fast_srgan_b200 import ops
    N, H, W = shape
    xa = torch.stack([nhwc(rnd((N, 64, H, W), 20 + g), dt) for g in range(G)])
    da = torch.stack([nhwc(rnd((N, 64, H, W), 60 + g), dt) for g in range(G)])
    dws = [torch.zeros((64, 64, 3, 3), device="cuda") for _ in range(G)]
    ops.conv3x3_wgrad_grouped(xa, da, dws)
    again = [torch.zeros((64, 64, 3, 3), device="cuda") for _ in range(G)]
    ops.conv3x3_wgrad_grouped(xa, da, again)
    for g in range(G):
        wr = torch.zeros((64, 64, 3, 3), device="cuda", requires_grad=True)
        F.conv2d(nchw(xa[g]), wr, padding=1).backward(nchw(da[g]))
        single = torch.zeros((64, 64, 3, 3), device="cuda")
        ops.conv3x3_wgrad(xa[g], da[g], single)
        assert rel_err(dws[g], wr.grad) <= 1e-4 and rel_err(single, wr.grad) <= 1e-4
        assert torch.equal(dws[g], again[g])            # fixed-order reduction: run-to-run identical


@pytest.mark.parametrize("dt", DT)
def test_conv_wgrad_ps_perm(dt):
    from fast_srgan_b200 import ops
    N, H, W = 2, 12, 16
    x = nhwc(rnd((N, 64, H, W), 10), dt)
    dy_ref = rnd((N, 256, H, W), 11).to(dt).float()              # reference channel order
    perm = torch.tensor([4 * (col % 64) + col // 64 for col in range(256)], device="cuda")
    dy_perm = nhwc(dy_ref[:, perm], dt)                           # GEMM column order
    wr = torch.zeros((256, 64, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(nchw(x), wr, padding=1).backward(dy_ref)
    dw = torch.zeros((256, 64, 3, 3), device="cuda")
    ops.conv3x3_wgrad(x, dy_perm, dw, ps_perm=True)
    assert rel_err(dw, wr.grad) <= 1e-4


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mode", ["none", "prelu", "lrelu"])
def test_instnorm_bwd(dt, mode):
    from fast_srgan_b200 import ops, _lib as L
    N, C, H, W = 2, 64, 12, 20
    raw = nhwc(rnd((N, C, H, W), 12, 2.0) + 0.5, dt)
    dy = nhwc(rnd((N, C, H, W), 13), dt)
    rr = nchw(raw).requires_grad_(True)
    alpha = torch.tensor([0.3], device="cuda", requires_grad=True)
    y = F.instance_norm(rr, eps=1e-5)
    if mode == "prelu":
        y = F.prelu(y, alpha)
    elif mode == "lrelu":
        y = F.leaky_relu(y, 0.01)
    y.backward(nchw(dy))
    rf = nchw(raw)
    stats = ops.stats_from_float(rf.sum((2, 3)), (rf * rf).sum((2, 3)))
    dalpha = torch.zeros(1, device="cuda")
    act = {"none": L.ACT_NONE, "prelu": L.ACT_PRELU, "lrelu": L.ACT_LRELU}[mode]
    draw = ops.instnorm_bwd(raw, stats, dy, act=act, slope=0.01, alpha=alpha.detach(), dalpha=dalpha)
    assert rel_err(nchw(draw), rr.grad) <= 4 * EPS[dt] + 1e-4
    if mode == "prelu":
        assert abs(dalpha.item() - alpha.grad.item()) <= 1e-3 * max(1.0, abs(alpha.grad.item()))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(3, 128, 24, 24), (2, 64, 48, 48), (5, 512, 12, 12)])
def test_instnorm_bwd_with_parity_layout_gradient(dt, shape):
    """fsr_instnorm_bwd_parity (dy read straight from a stride-2 data gradient's parity planes) == re-layout + fsr_instnorm_bwd,
    and fsr_instnorm_apply_parity == fsr_instnorm_apply + re-layout: same bits."""
    from fast_srgan_b200 import ops, _lib as L
    N, C, H, W = shape
    raw = nhwc(rnd(shape, 31), dt)
    dy = nhwc(rnd(shape, 32, 0.1), dt)
    stats = ops.stats_from_float(raw.float().sum(dim=(1, 2)), (raw.float() ** 2).sum(dim=(1, 2)))
    a = ops.instnorm_bwd(raw, stats, dy, act=L.ACT_LRELU, slope=0.01)
    b = ops.instnorm_bwd_parity(raw, stats, ops.parity_layout(dy, True), act=L.ACT_LRELU, slope=0.01)
    assert torch.equal(a, b)
    y = ops.instnorm_apply(raw, stats, act=L.ACT_LRELU, slope=0.01)
    yp = ops.instnorm_apply_parity(raw, stats, act=L.ACT_LRELU, slope=0.01)
    assert torch.equal(ops.parity_layout(y, True), yp)


@pytest.mark.parametrize("dt", DT)
def test_vgg_pool_and_relu_bwd(dt):
    from fast_srgan_b200 import ops
    x = F.relu(rnd((2, 64, 8, 12), 14)).to(dt).float()
    xn = nhwc(x, dt)
    got = ops.maxpool2(xn)
    assert torch.equal(nchw(got), F.max_pool2d(x, 2))
    dout = nhwc(rnd((2, 64, 4, 6), 15), dt)
    pre = rnd((2, 64, 8, 12), 14).to(dt).float().requires_grad_(True)      # same values before ReLU
    F.max_pool2d(F.relu(pre), 2).backward(nchw(dout))
    din = ops.maxpool2_relu_bwd(xn, dout)
    assert torch.equal(nchw(din), pre.grad.to(dt).float())
    y = F.relu(rnd((2, 64, 5, 8), 16)).to(dt)
    dy = rnd((2, 64, 5, 8), 17).to(dt)
    assert torch.equal(ops.relu_bwd(nhwc(y.float(), dt), nhwc(dy.float(), dt)), nhwc((dy.float() * (y.float() > 0)), dt))


@pytest.mark.parametrize("dt", DT)
def test_conv1x1_and_losses(dt):
    from fast_srgan_b200 import ops
    N, H, W, C = 3, 6, 6, 512
    x = nhwc(rnd((N, C, H, W), 18), dt)
    w = rnd((C,), 19, 0.05)
    b = rnd((1,), 20)
    z = ops.conv1x1_to1_fwd(x, w, b)
    xr = nchw(x).requires_grad_(True)
    wr = w.clone().view(1, C, 1, 1).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    zr = F.conv2d(xr, wr, br)
    assert rel_err(z.view(N, 1, H, W), zr.detach()) <= 1e-5
    noise = torch.rand((N, H, W), device="cuda")
    t = 0.3 * noise + 0.8
    loss_ref = F.binary_cross_entropy_with_logits(zr.view(N, H, W), t)
    (0.5 * loss_ref).backward()
    loss = torch.zeros(1, device="cuda")
    dz = torch.empty_like(z)
    ops.bce_logits(z, noise, 0.3, 0.8, loss, dz, grad_scale=0.5)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5
    dw, db = torch.zeros(C, device="cuda"), torch.zeros(1, device="cuda")
    dx = ops.conv1x1_to1_bwd(x, w, dz, dw, db)
    assert rel_err(dw, wr.grad.view(-1)) <= 1e-4 and rel_err(db, br.grad) <= 1e-4
    assert rel_err(nchw(dx), xr.grad) <= 2 * EPS[dt] + 1e-6
    a, bb = nhwc(rnd((N, C, H, W), 21, 1.5), dt), nhwc(rnd((N, C, H, W), 22), dt)
    ar = a.float().requires_grad_(True)
    lr = F.smooth_l1_loss(ar, bb.float())
    lr.backward()
    acc = torch.zeros(1, device="cuda")
    da = torch.empty_like(a)
    # a loss scale keeps the 1/numel gradient out of the fp16 subnormal range (the trainer does the same)
    S = 4096.0
    ops.smooth_l1(a, bb, acc, da, grad_scale=S / a.numel())
    assert abs(acc.item() / a.numel() - lr.item()) <= 1e-5
    assert rel_err(da.float() / S, ar.grad) <= 2 * EPS[dt] + 1e-7


@pytest.mark.parametrize("dt", DT)
def test_upsample_head_neck_backward_glue(dt):
    from fast_srgan_b200 import ops, _lib as L
    N, H, W = 2, 6, 8
    alpha = torch.tensor([0.2], device="cuda")
    conv = rnd((N, 256, H, W), 23).to(dt).float().requires_grad_(True)       # reference channel order
    ar = alpha.clone().requires_grad_(True)
    U = F.prelu(F.pixel_shuffle(conv, 2), ar)
    dU = rnd((N, 64, 2 * H, 2 * W), 24).to(dt).float()
    U.backward(dU)
    dalpha = torch.zeros(1, device="cuda")
    dconv = ops.ps_prelu_bwd(nhwc(U.detach(), dt), nhwc(dU, dt), alpha, dalpha)
    perm = torch.tensor([4 * (col % 64) + col // 64 for col in range(256)], device="cuda")
    assert rel_err(nchw(dconv), conv.grad[:, perm]) <= 2 * EPS[dt] + 1e-6
    assert abs(dalpha.item() - ar.grad.item()) <= 2e-2 * max(1.0, abs(ar.grad.item()))    # u/alpha re-derivation from rounded u
    # tanh backward + 3-channel weight gradients
    y = torch.tanh(rnd((N, 3, H, W), 25))
    dy = rnd((N, 3, H, W), 26)
    assert rel_err(ops.tanh_bwd(y, dy), dy * (1 - y * y)) <= 1e-6
    # neck wgrad: img = conv input (3ch), act = dOut (64ch)
    x3 = rnd((N, 3, H, W), 27)
    dout = rnd((N, 64, H, W), 28).to(dt).float()
    wr = torch.zeros((64, 3, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x3, wr, padding=1).backward(dout)
    out = torch.zeros((27, 64), device="cuda")
    ops.wgrad_c3(x3, nhwc(dout, dt), out, flip=False)
    assert rel_err(out.view(3, 9, 64).permute(2, 0, 1).reshape(64, 3, 3, 3), wr.grad) <= 1e-4
    # head wgrad: img = dpre (3ch, conv OUTPUT side), act = conv input x (64ch)
    x64 = rnd((N, 64, H, W), 29).to(dt).float()
    wh = torch.zeros((3, 64, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x64, wh, padding=1).backward(dy)
    out2 = torch.zeros((27, 64), device="cuda")
    ops.wgrad_c3(dy, nhwc(x64, dt), out2, flip=True)
    assert rel_err(out2.view(3, 9, 64).permute(0, 2, 1).reshape(3, 64, 3, 3), wh.grad) <= 1e-4
    # bias grads
    db = torch.zeros(64, device="cuda")
    ops.bias_grad(nhwc(dout, dt), db)
    assert rel_err(db, dout.sum((0, 2, 3))) <= 1e-4
    db3 = torch.zeros(3, device="cuda")
    ops.bias_grad_nchw(dy, db3)
    assert rel_err(db3, dy.sum((0, 2, 3))) <= 1e-4


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mode", ["prelu", "lrelu"])
def test_act_bwd(dt, mode):
    from fast_srgan_b200 import ops, _lib as L
    v = rnd((2, 64, 9, 11), 40).requires_grad_(True)
    alpha = torch.tensor([0.3], device="cuda", requires_grad=True)
    y = F.prelu(v, alpha) if mode == "prelu" else F.leaky_relu(v, 0.2)
    dy = rnd((2, 64, 9, 11), 41).to(dt).float()
    yq = y.detach().to(dt).float()
    y.backward(dy)
    dalpha = torch.zeros(1, device="cuda")
    act = L.ACT_PRELU if mode == "prelu" else L.ACT_LRELU
    dv = ops.act_bwd(nhwc(yq, dt), nhwc(dy, dt), act, slope=0.2, alpha=alpha.detach(), dalpha=dalpha)
    assert rel_err(nchw(dv), v.grad) <= 2 * EPS[dt] + 1e-6
    if mode == "prelu":
        assert abs(dalpha.item() - alpha.grad.item()) <= 4 * EPS[dt] * dy.abs().mul(yq.abs()).sum().item() / 0.3 + 1e-4


def test_adamw_matches_torch():
    from fast_srgan_b200 import ops
    p = rnd((1000,), 30)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-4)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = rnd((1000,), 30 + step)
        ref.grad = g.clone()
        opt.step()
        ops.adamw(p, g, m, v, 1e-4, step)
    assert (p - ref.detach()).abs().max().item() <= 1e-6
