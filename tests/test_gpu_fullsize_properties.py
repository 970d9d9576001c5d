"""GPU (-m gpu): the CUDA path at BASELINE.json's FULL sizes (batch 16, 256x256 and the per-layer shapes of that
step), where the CPU oracle would take minutes: size-independent properties instead of element-wise comparison.

  * adjoint identities of the three conv3x3 kernels:  <conv(x), dy> = <x, dgrad(dy)> = <W, wgrad(x, dy)>
    (forward, data gradient and weight gradient are three independent kernels; the identity holds for the exact
    operator, so it bounds every one of them; tolerance = bf16 output rounding averaged over >= 10^6 terms);
  * checksums: BatchNorm statistics from the conv epilogue == a separate pass over the stored tensor; the fused
    head kernel's logit gradients sum to zero over the classes (softmax - onehot), so does the bias gradient;
    confusion matrix: total = pixels, row sums = label histogram;
  * the whole step: CUDA-graph replay == eager launches (same kernels, atomics reorder fp32 sums), loss finite and
    decreasing on a fixed batch, and the inference kernels agree with the training kernels in eval mode.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops(lib_built):
    from continual_learning_b200 import _lib, ops as _ops
    _lib.ensure_device(0)
    return _ops


def dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(16, 256, 256, 64, 0, 64), (16, 256, 256, 64, 64, 64), (16, 128, 128, 128, 0, 128),
                                              (16, 64, 64, 256, 256, 256), (16, 32, 32, 512, 0, 512),
                                              (16, 16, 16, 1024, 0, 1024)])
def test_conv3x3_adjoint_identities_at_benchmark_layer_shapes(ops, n, h, w, c0, c1, co):
    g = torch.Generator(device="cuda").manual_seed(n + h + co)
    x = torch.randn(n, h, w, c0 + c1, device="cuda", generator=g).to(bf16)
    dy = torch.randn(n, h, w, co, device="cuda", generator=g).to(bf16)
    wt = (torch.randn(co, c0 + c1, 3, 3, device="cuda", generator=g) * 0.05).to(bf16).float()
    wf, wd = ops.pack_conv3x3(wt)
    x0 = x[..., :c0].contiguous()
    x1 = x[..., c0:].contiguous() if c1 else None
    y = ops.conv3x3_fprop(x0, x1, wf, None, relu=False)
    dx0, dx1 = ops.conv3x3_dgrad(dy, wd, c0, c1)
    dw = ops.conv3x3_wgrad(dy, x0, x1)                       # [9][Cin][Cout]
    a = dot(y, dy)
    b = dot(x0, dx0) + (dot(x1, dx1) if c1 else 0.0)
    c = dot(dw, wt.permute(2, 3, 1, 0).reshape(9, c0 + c1, co))
    scale = float(y.double().norm() * dy.double().norm())    # Cauchy-Schwarz scale of the inner product
    assert abs(a - b) <= 2e-4 * scale and abs(a - c) <= 2e-4 * scale, (a, b, c, scale)


def test_epilogue_statistics_equal_a_separate_pass_at_full_size(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(16, 256, 256, 64, device="cuda", generator=g).to(bf16)
    wt = torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05
    wf, _ = ops.pack_conv3x3(wt)
    bias = torch.randn(64, device="cuda", generator=g)
    s1 = torch.zeros(64, device="cuda", dtype=torch.float64)
    q1 = torch.zeros(64, device="cuda", dtype=torch.float64)
    y = ops.conv3x3_fprop(x, None, wf, bias, relu=True, stats=(s1, q1))
    s2, q2 = torch.zeros_like(s1), torch.zeros_like(q1)
    ops.bn_stats(y, s2, q2)
    yd = y.double().reshape(-1, 64)
    assert torch.allclose(s1, s2, rtol=1e-6) and torch.allclose(q1, q2, rtol=1e-6)
    assert torch.allclose(s1, yd.sum(0), rtol=1e-6) and torch.allclose(q1, (yd * yd).sum(0), rtol=1e-6)


def test_fused_head_checksums_at_full_size(ops):
    P, c = 16 * 256 * 256, 21
    g = torch.Generator(device="cuda").manual_seed(4)
    z = torch.randn(P, 64, device="cuda", generator=g).to(bf16)
    w = (torch.randn(c, 64, 1, 1, device="cuda", generator=g) * 0.2)
    b = torch.randn(c, device="cuda", generator=g)
    y = torch.randint(0, c, (P,), device="cuda", generator=g)
    wf, wd = ops.pack_head(w)
    loss_acc, dz, dw, db = ops.head_loss_bwd(z, wf, wd, b, y, c)
    # sum over classes of (softmax - onehot) is 0 for every pixel -> the bias gradient sums to ~0 (bf16 rounding of
    # 21 terms of magnitude <= 1/P each)
    assert abs(float(db[:c].sum())) <= 1e-3 * float(db[:c].abs().sum())
    assert float(dw[c:].abs().max()) == 0.0 and float(db[c:].abs().max()) == 0.0
    # loss against torch on the same bf16 operands (fp32 GEMM on the device)
    logits = z.float() @ w.view(c, 64).to(bf16).float().t() + b
    ce = torch.nn.functional.cross_entropy(logits, y, reduction="sum")
    assert abs(float(loss_acc[0]) - float(ce)) <= 1e-4 * float(ce)
    # statistics kernel on the same head: counts add up
    pred, conf, ok = ops.head_argmax_confusion(z, wf, b, y, c, nc=c, want_pred=True)
    assert int(conf.sum()) == P and int(ok) == int((pred == y).sum())
    assert torch.equal(conf.view(c, c).sum(1), torch.bincount(y, minlength=c))
    assert torch.equal(pred, logits.argmax(1)) or float((pred != logits.argmax(1)).float().mean()) < 1e-4


def test_full_size_step_graph_equals_eager_and_learns(lib_built):
    import continual_learning_b200 as clk
    from oracle.data import structured_batch
    from oracle.unet_ref import make_state_dict
    sd = make_state_dict(0)
    x, y = structured_batch(1, 16, 256, 256)
    x, y = x.cuda(), y.cuda()
    runs = []
    for use_graph in (False, True):
        m = clk.UNet(21).cuda()
        m.load_state_dict(sd)
        m.train()
        ts = clk.TrainStep(m, clk.FusedAdam(m.parameters(), lr=1e-4, betas=(0.5, 0.99)), use_graph=use_graph)
        runs.append([float(ts.step(x, y)) for _ in range(6)])
    np.testing.assert_allclose(runs[0], runs[1], rtol=2e-3)
    assert all(np.isfinite(runs[0])) and runs[0][-1] < runs[0][0]  # the fixed batch is being fitted
    # eval mode: inference kernels (BatchNorm folded into the conv epilogue, one head+argmax kernel) against the
    # training kernels in eval mode (separate BatchNorm passes, materialised logits)
    m.eval()
    with torch.no_grad():
        pred_a = m(x).argmax(1)
    pred_b, _, _ = m.evaluate_batch(x, y, want_pred=True)
    logits_c = m.engine.forward(x, training=False, save_for_backward=True)
    pred_c = logits_c.argmax(-1)
    m.engine.release()
    assert torch.equal(pred_a, pred_b)
    assert float((pred_a != pred_c).float().mean()) < 5e-3  # one bf16 rounding per unit differs between the two paths
