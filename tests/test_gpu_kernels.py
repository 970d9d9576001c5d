"""GPU (-m gpu): every libclk kernel, called through the C ABI, against the stock PyTorch CPU fp32 op it
replaces, on the same bf16-rounded operands.

Tolerances (per kernel): outputs stored as bf16 carry one bf16 rounding (rel-L2 <= 3e-3);
fp32 outputs (weight gradients, logits, statistics) <= 1e-5 .. 1e-4; integer outputs bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 3e-3


@pytest.fixture(scope="module")
def ops(lib_built):
    from continual_learning_b200 import _lib, ops as _ops
    _lib.ensure_device(0)
    return _ops


def rel(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rnd(g, *shape, scale=1.0):
    return torch.randn(*shape, generator=g) * scale


def bfr(x):  # bf16-rounded fp32 (CPU)
    return x.to(torch.bfloat16).float()


def to_nhwc_dev(x):  # CPU NCHW fp32 -> CUDA NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def from_nhwc(x):  # CUDA NHWC -> CPU NCHW fp32
    return x.float().cpu().permute(0, 3, 1, 2)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,K,N", [(128, 64, 64), (1000, 128, 128), (4096, 256, 256), (300, 64, 512), (1, 64, 64)])
def test_gemm_fprop_bias_relu_stats(ops, P, K, N):
    g = gen(P + K + N)
    a, w, b = bfr(rnd(g, P, K)), bfr(rnd(g, N, K, scale=0.1)), rnd(g, N)
    s_sum = torch.zeros(N, device="cuda", dtype=torch.float64)
    s_sq = torch.zeros(N, device="cuda", dtype=torch.float64)
    out = ops.gemm_fprop(a.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda(), b.cuda(), N, relu=True,
                         stats=(s_sum, s_sq))
    ref = torch.relu(a @ w.t() + b)
    assert rel(out, ref) <= BF16_TOL
    q = out.float().cpu().double()
    assert rel(s_sum, q.sum(0)) <= 1e-6 and rel(s_sq, (q * q).sum(0)) <= 1e-6


@pytest.fixture(params=["generic", "halo", "pair"])
def conv_kernel(request):
    """run the conv3x3 forward/dgrad cases through all three tcgen05 kernels: the generic per-tap-TMA kernel, the
    persistent halo kernel and the CTA-pair (cta_group::2, M = 256) halo kernel.  The library picks between them by
    image size and channel count; the knobs force one."""
    from continual_learning_b200 import _lib
    _lib.set_tuning("conv3_v2", {"generic": 0, "halo": 2, "pair": 4}[request.param])
    _lib.set_tuning("conv3_pair", 0 if request.param == "halo" else 1)
    yield request.param
    _lib.set_tuning("conv3_v2", 4)
    _lib.set_tuning("conv3_pair", 1)


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(2, 16, 16, 64, 0, 64), (3, 8, 24, 128, 0, 256), (2, 16, 16, 64, 64, 128),
                                              (5, 4, 4, 128, 128, 64), (1, 32, 32, 64, 0, 64), (33, 2, 2, 64, 0, 128),
                                              (1, 1, 1, 64, 0, 64), (2, 3, 5, 64, 64, 64), (1, 64, 64, 64, 0, 64),
                                              (2, 40, 72, 64, 64, 128)])
def test_conv3x3_fprop_with_folded_concat(ops, conv_kernel, n, h, w, c0, c1, co):
    g = gen(n * 100 + h + co)
    x, wt, b = bfr(rnd(g, n, c0 + c1, h, w)), rnd(g, co, c0 + c1, 3, 3, scale=0.05), rnd(g, co)
    xh = to_nhwc_dev(x)
    x0 = xh[..., :c0].contiguous()
    x1 = xh[..., c0:].contiguous() if c1 else None
    wf, _ = ops.pack_conv3x3(wt.cuda())
    s_sum = torch.zeros(co, device="cuda", dtype=torch.float64)
    s_sq = torch.zeros(co, device="cuda", dtype=torch.float64)
    y = ops.conv3x3_fprop(x0, x1, wf, b.cuda(), relu=True, stats=(s_sum, s_sq))
    ref = torch.relu(F.conv2d(x, bfr(wt), b, padding=1))  # torch.cat folded: channels of x are [x0 | x1]
    assert rel(from_nhwc(y), ref) <= BF16_TOL
    q = y.float().cpu().double().reshape(-1, co)
    assert rel(s_sum, q.sum(0)) <= 1e-6 and rel(s_sq, (q * q).sum(0)) <= 1e-6


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(2, 16, 16, 64, 0, 64), (3, 8, 24, 128, 0, 256), (2, 16, 16, 64, 64, 128),
                                              (2, 3, 5, 64, 64, 64), (1, 34, 18, 128, 0, 512)])
def test_conv3x3_fprop_eval_applies_running_stat_batchnorm_in_the_epilogue(ops, conv_kernel, n, h, w, c0, c1, co):
    """module.eval() path (trainer.py:271): conv -> ReLU -> BatchNorm(running stats) in ONE launch against the three
    stock ops in fp32."""
    g = gen(n * 10 + h + co + 5)
    x, wt, b = bfr(rnd(g, n, c0 + c1, h, w)), rnd(g, co, c0 + c1, 3, 3, scale=0.05), rnd(g, co)
    gamma, beta = rnd(g, co), rnd(g, co)
    rmean, rvar = rnd(g, co, scale=0.3), torch.rand(co, generator=g) + 0.5
    scale = gamma / torch.sqrt(rvar + 1e-5)
    shift = beta - rmean * scale
    xh = to_nhwc_dev(x)
    x0 = xh[..., :c0].contiguous()
    x1 = xh[..., c0:].contiguous() if c1 else None
    wf, _ = ops.pack_conv3x3(wt.cuda())
    z = ops.conv3x3_fprop_eval(x0, x1, wf, b.cuda(), scale.cuda(), shift.cuda(), relu=True)
    ref = F.batch_norm(torch.relu(F.conv2d(x, bfr(wt), b, padding=1)), rmean, rvar, gamma, beta, training=False)
    assert rel(from_nhwc(z), ref) <= BF16_TOL


def test_gemm_fprop_eval_stem(ops):
    g = gen(77)
    P, K, N = 1000, 64, 64
    a, w, b = bfr(rnd(g, P, K)), bfr(rnd(g, N, K, scale=0.1)), rnd(g, N)
    scale, shift = rnd(g, N), rnd(g, N)
    out = ops.gemm_fprop_eval(a.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda(), b.cuda(), N, scale.cuda(),
                              shift.cuda(), relu=True)
    assert rel(out, torch.relu(a @ w.t() + b) * scale + shift) <= BF16_TOL


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(2, 16, 16, 64, 0, 64), (2, 8, 8, 128, 128, 128), (1, 16, 16, 256, 0, 64),
                                              (2, 6, 10, 64, 64, 64), (1, 64, 48, 128, 0, 64), (2, 33, 17, 64, 64, 128),
                                              # 64 gradient channels out of 64 / 128 in: the paired-tap kernel (one and
                                              # two K chunks), widths that are not multiples of its 30-pixel segments
                                              (2, 6, 10, 64, 0, 128), (1, 33, 61, 64, 0, 128), (3, 9, 31, 64, 0, 64),
                                              (1, 1, 1, 64, 0, 64)])
def test_conv3x3_dgrad_split_destinations(ops, conv_kernel, n, h, w, c0, c1, co):
    g = gen(7 + n + h + co)
    dy, wt = bfr(rnd(g, n, co, h, w)), rnd(g, co, c0 + c1, 3, 3, scale=0.05)
    _, wd = ops.pack_conv3x3(wt.cuda())
    dx0, dx1 = ops.conv3x3_dgrad(to_nhwc_dev(dy), wd, c0, c1)
    ref = F.conv_transpose2d(dy, bfr(wt), padding=1)  # dgrad of a stride-1 pad-1 conv
    got = from_nhwc(dx0) if dx1 is None else torch.cat([from_nhwc(dx0), from_nhwc(dx1)], 1)
    assert rel(got, ref) <= BF16_TOL


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(2, 16, 16, 64, 0, 64), (2, 8, 8, 128, 0, 256), (4, 16, 16, 64, 64, 128),
                                              (8, 32, 32, 64, 0, 64), (3, 4, 4, 128, 0, 128), (2, 5, 3, 64, 0, 64),
                                              (2, 40, 24, 128, 128, 256)])
@pytest.mark.parametrize("variant", ["pair", "halo", "generic"])
def test_conv3x3_wgrad_and_unpack(ops, variant, n, h, w, c0, c1, co):
    from continual_learning_b200 import _lib
    _lib.set_tuning("wgrad_v2", {"generic": 0, "halo": 1, "pair": 2}[variant])
    try:
        _check_conv3x3_wgrad(ops, n, h, w, c0, c1, co)
        if variant != "generic":  # deterministic split-K form: partial buffers summed in a fixed order
            _check_conv3x3_wgrad_split(ops, n, h, w, c0, c1, co)
    finally:
        _lib.set_tuning("wgrad_v2", 2)


def _check_conv3x3_wgrad_split(ops, n, h, w, c0, c1, co):
    g = gen(17 + n + h + co)
    x, dy = bfr(rnd(g, n, c0 + c1, h, w)), bfr(rnd(g, n, co, h, w, scale=0.1))
    xh = to_nhwc_dev(x)
    x0 = xh[..., :c0].contiguous()
    x1 = xh[..., c0:].contiguous() if c1 else None
    dyh = to_nhwc_dev(dy)
    parts = ops.conv3x3_wgrad_split(dyh, x0, x1)
    assert parts.shape[0] == ops.conv3x3_wgrad_splits(co, c0 + c1, n, h, w)
    wref = torch.zeros(co, c0 + c1, 3, 3, requires_grad=True)
    F.conv2d(x, wref, padding=1).backward(dy)
    got = parts.sum(0).reshape(3, 3, c0 + c1, co).permute(3, 2, 0, 1)
    assert rel(got, wref.grad) <= 2e-5
    assert torch.equal(parts, ops.conv3x3_wgrad_split(dyh, x0, x1))  # bit-reproducible


def _check_conv3x3_wgrad(ops, n, h, w, c0, c1, co):
    g = gen(11 + n + h + co)
    x, dy = bfr(rnd(g, n, c0 + c1, h, w)), bfr(rnd(g, n, co, h, w, scale=0.1))
    xh = to_nhwc_dev(x)
    x0 = xh[..., :c0].contiguous()
    x1 = xh[..., c0:].contiguous() if c1 else None
    dw = ops.conv3x3_wgrad(to_nhwc_dev(dy), x0, x1)
    wref = torch.zeros(co, c0 + c1, 3, 3, requires_grad=True)
    F.conv2d(x, wref, padding=1).backward(dy)
    grad = torch.empty(co, c0 + c1, 3, 3, device="cuda")
    ops.unpack_wgrad(dw, grad, co, c0 + c1, 9, co, c0 + c1, transposed=True)
    assert rel(grad, wref.grad) <= 2e-5  # fp32 accumulate, fp32 output, split-K order only


@pytest.mark.parametrize("P,cu,ct", [(1000, 64, 64), (5000, 128, 64), (777, 64, 128), (63, 64, 64)])
def test_gemm_wgrad(ops, P, cu, ct):
    g = gen(P)
    u, t = bfr(rnd(g, P, cu)), bfr(rnd(g, P, ct))
    out = ops.gemm_wgrad(u.to(torch.bfloat16).cuda(), t.to(torch.bfloat16).cuda())
    assert rel(out, u.t() @ t) <= 2e-5


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 8, 8, 128, 64), (2, 4, 4, 256, 128), (3, 2, 6, 64, 64), (1, 16, 16, 1024, 512)])
def test_conv_transpose_2x2_fprop_dgrad_wgrad(ops, n, h, w, ci, co):
    g = gen(n + h + ci)
    x, wt, b = bfr(rnd(g, n, ci, h, w)), rnd(g, ci, co, 2, 2, scale=0.05), rnd(g, co)
    wf, wd = ops.pack_convT(wt.cuda())
    xh = to_nhwc_dev(x)
    y = ops.convT_fprop(xh, wf, b.cuda())
    assert rel(from_nhwc(y), F.conv_transpose2d(x, bfr(wt), b, stride=2)) <= BF16_TOL
    dy = bfr(rnd(g, n, co, 2 * h, 2 * w))
    dyh = to_nhwc_dev(dy)
    dx = ops.convT_dgrad(dyh, wd)
    assert rel(from_nhwc(dx), F.conv2d(dy, bfr(wt), stride=2)) <= BF16_TOL
    dw = ops.convT_wgrad(xh, dyh)
    wref = torch.zeros(ci, co, 2, 2, requires_grad=True)
    F.conv_transpose2d(x, wref, stride=2).backward(dy)
    grad = torch.empty(ci, co, 2, 2, device="cuda")
    ops.unpack_wgrad(dw, grad, ci, co, 4, ci, co)
    assert rel(grad, wref.grad) <= 2e-5


@pytest.mark.parametrize("n,h,w,ci,nc", [(2, 16, 16, 64, 21), (1, 8, 8, 64, 2), (3, 4, 4, 128, 16)])
def test_head_1x1_fprop_dgrad_wgrad(ops, n, h, w, ci, nc):
    g = gen(n + nc)
    x, wt, b = bfr(rnd(g, n, h, w, ci)), rnd(g, nc, ci, 1, 1, scale=0.1), rnd(g, nc)
    wf, wd = ops.pack_head(wt.cuda())
    xd = x.to(torch.bfloat16).cuda()
    lg = ops.gemm_fprop(xd, wf, b.cuda(), nc, out_f32=True)
    w2 = bfr(wt).reshape(nc, ci)
    assert lg.shape[-1] == nc and rel(lg, x @ w2.t() + b) <= 1e-5  # fp32 logits
    dl = torch.zeros(n, h, w, 64)
    dl[..., :nc] = bfr(rnd(g, n, h, w, nc))
    dld = dl.to(torch.bfloat16).cuda()
    assert rel(ops.gemm_fprop(dld, wd, None, ci), dl[..., :nc] @ w2) <= BF16_TOL
    dw = ops.gemm_wgrad(dld, xd)
    assert rel(dw[:nc], dl[..., :nc].reshape(-1, nc).t() @ x.reshape(-1, ci)) <= 2e-5
    assert float(dw[nc:].abs().max()) == 0.0  # padded logit channels carry no gradient


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (3, 8, 8, 256), (2, 4, 4, 1024), (1, 2, 2, 64)])
def test_batchnorm_train_forward_backward_pool(ops, n, h, w, c):
    g = gen(n + c)
    y = bfr(torch.relu(rnd(g, n, c, h, w)))
    P = n * h * w
    yd = to_nhwc_dev(y)
    dev = "cuda"
    s_sum, s_sq = (torch.zeros(c, device=dev, dtype=torch.float64) for _ in range(2))
    ops.bn_stats(yd, s_sum, s_sq)
    gamma, beta = rnd(g, c) * 0.2 + 1.0, rnd(g, c) * 0.2
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    mean, invstd, scale, shift = (torch.empty(c, device=dev) for _ in range(4))
    ops.bn_finalize(s_sum, s_sq, gamma.cuda(), beta.cuda(), rm, rv, mean, invstd, scale, shift, P)
    z = ops.bn_apply(yd, scale, shift)
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    yin = y.clone().requires_grad_(True)
    zr = bn(yin)
    assert rel(from_nhwc(z), zr) <= BF16_TOL
    assert rel(rm, bn.running_mean) <= 1e-5 and rel(rv, bn.running_var) <= 1e-5
    # fused apply + 2x2 max-pool with window index (first max wins, models/unet.py:12)
    z2, pooled, idx = ops.bn_apply_pool(yd, scale, shift)
    assert torch.equal(z2, z)
    pr, ir = F.max_pool2d(from_nhwc(z2), 2, 2, return_indices=True)
    assert torch.equal(from_nhwc(pooled), pr)
    hh = ir // w - 2 * torch.arange(h // 2).view(1, 1, -1, 1)
    ww = ir % w - 2 * torch.arange(w // 2).view(1, 1, 1, -1)
    assert torch.equal(idx.cpu().permute(0, 3, 1, 2).long(), hh * 2 + ww)
    # backward: BN (batch statistics) + ReLU mask + bias gradient
    dz = bfr(rnd(g, n, c, h, w))
    dzd = to_nhwc_dev(dz)
    s1, s2 = (torch.zeros(c, device=dev, dtype=torch.float64) for _ in range(2))
    ops.bn_bwd_reduce(dzd, yd, s1, s2)
    dgamma, dbeta, ka, kb, kc = (torch.empty(c, device=dev) for _ in range(5))
    ops.bn_bwd_finalize(s1, s2, gamma.cuda(), mean, invstd, dgamma, dbeta, ka, kb, kc, P)
    dbias = torch.zeros(c, device=dev, dtype=torch.float64)
    dpre = ops.bn_relu_bwd_apply(dzd, yd, ka, kb, kc, dbias)
    zr.backward(dz)
    refd = yin.grad * (y > 0)
    assert rel(from_nhwc(dpre), refd) <= BF16_TOL
    assert rel(dgamma, bn.weight.grad) <= 1e-4 and rel(dbeta, bn.bias.grad) <= 1e-4
    assert rel(dbias, refd.sum((0, 2, 3))) <= 1e-4
    # max-pool backward fused with the skip-gradient add
    dp, skip = bfr(rnd(g, n, c, h // 2, w // 2)), bfr(rnd(g, n, c, h, w))
    din = ops.maxpool_bwd_add(to_nhwc_dev(dp), idx, to_nhwc_dev(skip))
    assert rel(from_nhwc(din), F.max_unpool2d(dp, ir, 2, 2) + skip) <= BF16_TOL
    # ... and with the BatchNorm-backward reductions of the pooled layer (din is its dz): same tensor, and the sums of
    # a separate clk_bn_bwd_reduce pass over (din, y)
    r1, r2 = (torch.zeros(c, device=dev, dtype=torch.float64) for _ in range(2))
    din2 = ops.maxpool_bwd_add_reduce(to_nhwc_dev(dp), idx, to_nhwc_dev(skip), yd, r1, r2)
    assert torch.equal(din2, din)
    w1, w2 = (torch.zeros(c, device=dev, dtype=torch.float64) for _ in range(2))
    ops.bn_bwd_reduce(din, yd, w1, w2)
    assert rel(r1, w1) <= 1e-5 and rel(r2, w2) <= 1e-5   # fp32 partial sums in a different order, fp64 totals
    dd, yy = din.double().reshape(-1, c), yd.double().reshape(-1, c)
    assert rel(r1, dd.sum(0)) <= 1e-5 and rel(r2, (dd * yy).sum(0)) <= 1e-5


def test_batchnorm_eval_mode_uses_running_stats(ops):
    g = gen(3)
    c, n, h, w = 128, 2, 8, 8
    y = bfr(torch.relu(rnd(g, n, c, h, w)))
    gamma, beta = rnd(g, c) * 0.2 + 1.0, rnd(g, c) * 0.2
    rm, rv = rnd(g, c) * 0.1, torch.rand(c, generator=g) + 0.5
    dev = "cuda"
    mean, invstd, scale, shift = (torch.empty(c, device=dev) for _ in range(4))
    rmd, rvd = rm.cuda(), rv.cuda()
    ops.bn_finalize(None, None, gamma.cuda(), beta.cuda(), rmd, rvd, mean, invstd, scale, shift, n * h * w, training=False)
    z = ops.bn_apply(to_nhwc_dev(y), scale, shift)
    ref = F.batch_norm(y, rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert rel(from_nhwc(z), ref) <= BF16_TOL
    assert torch.equal(rmd.cpu(), rm) and torch.equal(rvd.cpu(), rv)  # eval never touches the buffers


def test_maxpool_tie_breaks_to_first_and_propagates_nan(ops):
    y = torch.zeros(1, 2, 2, 64)
    y[0, :, :, 0] = 1.0                       # four-way tie -> index 0
    y[0, 0, 1, 1] = 5.0; y[0, 1, 0, 1] = 5.0  # tie between window positions 1 and 2 -> 1
    y[0, 1, 1, 2] = float("nan")              # NaN wins
    _, pooled, idx = ops.bn_apply_pool(y.to(torch.bfloat16).cuda(), None, None)
    assert idx[0, 0, 0, 0].item() == 0 and idx[0, 0, 0, 1].item() == 1 and idx[0, 0, 0, 2].item() == 3
    assert torch.isnan(pooled[0, 0, 0, 2].float()).item()


@pytest.mark.parametrize("P,c,cold", [(1000, 21, 0), (4096, 21, 16), (513, 2, 0), (255, 7, 7)])
def test_fused_ce_kd_loss_forward_backward(ops, P, c, cold):
    g = gen(P + c)
    z = rnd(g, P, c, scale=2.0).requires_grad_(True)
    lab = torch.randint(0, c, (P,), generator=g)
    zo = rnd(g, P, cold, scale=2.0) if cold else None
    T, lam = 2.0, 1.0
    acc, dl = ops.ce_kd_loss(z.detach().cuda(), lab.cuda(), None if zo is None else zo.cuda(), T=T, lam=lam)
    loss = F.cross_entropy(z, lab)
    if cold:
        loss = loss + lam * T * T * F.kl_div(F.log_softmax(z[:, :cold] / T, 1), F.softmax(zo / T, 1), reduction="sum") / P
    loss.backward()
    got = float((acc[0] + (lam * T * T * acc[1] if cold else 0)) / P)
    assert abs(got - float(loss.detach())) <= 1e-5 * abs(float(loss.detach()))
    assert rel(dl[:, :c], z.grad) <= BF16_TOL
    assert float(dl[:, c:].float().abs().max()) == 0.0


@pytest.mark.parametrize("P,c,cold", [(1000, 21, 0), (4096, 21, 16), (513, 2, 0), (255, 7, 7), (128 * 500 + 3, 21, 16)])
def test_fused_head_loss_backward_matches_the_separate_ops(ops, P, c, cold):
    """clk_head_loss_bwd == head GEMM -> nn.CrossEntropyLoss (+ KL) -> autograd of both, in fp32 on the same
    bf16-rounded operands; the gradient w.r.t. the logits is rounded to bf16 before the two backward GEMMs (as in the
    unfused path), hence the bf16 tolerance on dz / dW / db."""
    g = gen(P + c + cold)
    z = bfr(rnd(g, P, 64))
    w = bfr(rnd(g, c, 64, scale=0.2))
    b = rnd(g, c)
    y = torch.randint(0, c, (P,), generator=g)
    zo = rnd(g, P, cold, scale=2.0) if cold else None
    T, lam = 2.0, 0.7
    wf, wd = ops.pack_head(w.view(c, 64, 1, 1).cuda())
    loss_acc, dz, dw, db = ops.head_loss_bwd(z.to(torch.bfloat16).cuda(), wf, wd, b.cuda(), y.cuda(), c,
                                              old_logits=None if zo is None else zo.cuda(), T=T, lam=lam)
    zr = z.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    logits = zr @ wr.t() + br
    ce = F.cross_entropy(logits, y, reduction="sum")
    kd = torch.zeros(())
    if cold:
        kd = F.kl_div(F.log_softmax(logits[:, :cold] / T, 1), F.softmax(zo / T, 1), reduction="sum")
    ((ce + lam * T * T * kd) / P).backward()
    assert abs(float(loss_acc[0]) - float(ce)) <= 1e-4 * abs(float(ce))
    if cold:
        assert abs(float(loss_acc[1]) - float(kd)) <= 1e-3 * abs(float(kd)) + 1e-6
    assert rel(dz, zr.grad) <= 6e-3
    assert rel(dw[:c], wr.grad) <= 6e-3
    assert rel(db[:c], br.grad) <= 6e-3
    assert float(dw[c:].abs().max()) == 0.0 if c < 64 else True


@pytest.mark.parametrize("n,nc", [(100001, 22), (65536, 21), (7, 3), (1, 2), (0, 5)])
def test_confusion_matrix_bit_exact(ops, n, nc):
    g = gen(n + nc)
    t = torch.randint(-1, nc + 1, (n,), generator=g)   # includes out-of-range targets (masked)
    p = torch.randint(0, nc, (n,), generator=g)
    conf = ops.confusion_matrix(t.cuda(), p.cuda(), nc)
    m = (t >= 0) & (t < nc)
    ref = torch.bincount(nc * t[m] + p[m], minlength=nc * nc)  # metrics.py:34-37 without .float()
    assert torch.equal(conf.cpu(), ref)


def test_confusion_matrix_flags_out_of_range_prediction(ops):
    t = torch.tensor([0, 1, 2, 2], device="cuda")
    p = torch.tensor([0, 1, 5, 1], device="cuda")
    err = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.confusion_matrix(t, p, 3, err_flag=err)
    assert int(err) == 1
    import continual_learning_b200 as clk
    with pytest.raises(RuntimeError):
        clk.metrics.eval_metrics(t.view(1, 2, 2), p.view(1, 2, 2), 3)


def test_argmax_confusion_fused(ops):
    g = gen(5)
    lg = rnd(g, 5000, 21)
    lg[10] = lg[10, 3]  # all-equal row: first index wins
    lab = torch.randint(0, 21, (5000,), generator=g)
    pred, conf, correct = ops.argmax_confusion(lg.cuda(), lab.cuda(), nc=22, want_pred=True)
    rp = lg.argmax(1)
    assert torch.equal(pred.cpu(), rp) and int(pred[10]) == 0
    assert torch.equal(conf.cpu(), torch.bincount(22 * lab + rp, minlength=484))
    assert int(correct) == int((rp == lab).sum())


def test_fused_adam_matches_torch_adam(ops):
    from continual_learning_b200.optim import FusedAdam
    g = gen(9)
    shapes = [(64, 3, 3, 3), (64,), (256, 256, 3, 3), (21, 64, 1, 1), (7,)]
    ps = [torch.nn.Parameter(rnd(g, *s).cuda()) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    o1 = FusedAdam(ps, lr=1e-3, betas=(0.5, 0.99))
    o2 = torch.optim.Adam(qs, lr=1e-3, betas=(0.5, 0.99))  # trainer.py:108-110
    for _ in range(3):
        for a, b in zip(ps, qs):
            gr = rnd(g, *a.shape)
            a.grad, b.grad = gr.cuda(), gr.clone()
        o1.step()
        o2.step()
    for a, b in zip(ps, qs):
        assert rel(a, b) <= 1e-6
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert set(sd1["state"][0]) == set(sd2["state"][0])  # step / exp_avg / exp_avg_sq: checkpoints interchange


def test_layout_and_packing_kernels(ops):
    g = gen(13)
    x = rnd(g, 2, 3, 16, 16)
    a = ops.im2col_stem(x.cuda())
    ref = F.unfold(x, 3, padding=1).permute(0, 2, 1).reshape(2, 16, 16, 27)
    assert torch.equal(a[..., :27].float().cpu(), bfr(ref)) and float(a[..., 27:].float().abs().max()) == 0.0
    wt = rnd(g, 128, 64, 3, 3)
    wf, wd = ops.pack_conv3x3(wt.cuda())
    assert torch.equal(wf.float().cpu(), bfr(wt.permute(2, 3, 0, 1).reshape(9, 128, 64)))
    assert torch.equal(wd.float().cpu(), bfr(wt.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 64, 128)))
    xx = bfr(rnd(g, 2, 8, 8, 64))
    assert torch.equal(ops.nhwc_to_nchw_f32(xx.to(torch.bfloat16).cuda()).cpu(), xx.permute(0, 3, 1, 2))
    assert torch.equal(ops.nchw_to_nhwc_bf16(x.cuda(), 8)[..., :3].float().cpu(), bfr(x.permute(0, 2, 3, 1)))


def test_unsupported_shapes_are_rejected_not_miscomputed(ops):
    from continual_learning_b200._lib import ClkError
    x = torch.zeros(1, 4, 4, 48, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(9, 64, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ClkError, match="UNSUPPORTED_SHAPE"):
        ops.conv3x3_fprop(x, None, w, None)
    with pytest.raises(ClkError, match="UNSUPPORTED_SHAPE"):
        ops.im2col_stem(torch.zeros(1, 8, 4, 4, device="cuda"))


@pytest.mark.parametrize("P,c,nc", [(1000, 21, 22), (128 * 300 + 5, 21, 21), (513, 2, 2), (64, 7, 9)])
def test_fused_head_argmax_confusion_equals_the_separate_ops(ops, P, c, nc):
    """clk_head_argmax_confusion == 1x1 head GEMM (fp32 logits) -> clk_argmax_confusion, bit for bit (same MMA, same
    tie rule), and the counts equal numpy's on those predictions."""
    g = gen(P + c)
    z = bfr(rnd(g, P, 64)).to(torch.bfloat16).cuda()
    w = bfr(rnd(g, c, 64, scale=0.2))
    b = rnd(g, c).cuda()
    y = torch.randint(0, nc, (P,), generator=g).cuda()
    wf, _ = ops.pack_head(w.view(c, 64, 1, 1).cuda())
    logits = ops.gemm_fprop(z, wf, b, c, out_f32=True)
    pred_ref, conf_ref, ok_ref = ops.argmax_confusion(logits, y, nc, want_pred=True)
    pred, conf, ok = ops.head_argmax_confusion(z, wf, b, y, c, nc=nc, want_pred=True)
    assert torch.equal(pred, pred_ref) and torch.equal(conf, conf_ref) and torch.equal(ok, ok_ref)
    want = np.bincount((nc * y.cpu().numpy() + pred.cpu().numpy()), minlength=nc * nc)
    assert np.array_equal(conf.cpu().numpy(), want)
    # accumulating form, no prediction map
    _, conf2, ok2 = ops.head_argmax_confusion(z, wf, b, y, c, nc=nc, conf=conf.clone(), correct=ok.clone())
    assert torch.equal(conf2, 2 * conf_ref) and torch.equal(ok2, 2 * ok_ref)


@pytest.mark.parametrize("n,cin,h,w", [(2, 3, 16, 16), (1, 3, 8, 48), (3, 1, 24, 16), (2, 2, 40, 32), (16, 3, 256, 256)])
def test_stem_conv_direct_kernel(ops, n, cin, h, w):
    """enc1.0 (models/unet.py:50) as one launch — the im2col tile built in shared memory from the fp32 NCHW input —
    against F.conv2d on the bf16-rounded operands, the BatchNorm statistics of its epilogue against a pass over the
    stored tensor, the inference (affine) epilogue, and against the two-launch im2col + GEMM path it replaces."""
    g = gen(n + h + w + cin)
    x = rnd(g, n, cin, h, w)
    wt, b = rnd(g, 64, cin, 3, 3, scale=0.2), rnd(g, 64)
    wf = ops.pack_stem(wt.cuda())
    s1, s2 = (torch.zeros(64, device="cuda", dtype=torch.float64) for _ in range(2))
    y = ops.stem_conv(x.cuda(), wf, b.cuda(), relu=True, stats=(s1, s2))
    ref = torch.relu(F.conv2d(bfr(x), bfr(wt), b, padding=1))
    assert rel(from_nhwc(y), ref) <= BF16_TOL
    yd = y.double().reshape(-1, 64)
    assert rel(s1, yd.sum(0)) <= 1e-6 and rel(s2, (yd * yd).sum(0)) <= 1e-6
    # same numbers as im2col + GEMM (identical bf16 operands, fp32 accumulation order aside)
    y2 = ops.gemm_fprop(ops.im2col_stem(x.cuda()), wf, b.cuda(), 64, relu=True)
    assert rel(y, y2) <= 1e-3
    scale, shift = rnd(g, 64), rnd(g, 64)
    z = ops.stem_conv(x.cuda(), wf, b.cuda(), relu=True, scale=scale.cuda(), shift=shift.cuda())
    assert rel(from_nhwc(z), ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)) <= BF16_TOL
