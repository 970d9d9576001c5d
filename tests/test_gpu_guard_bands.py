"""GPU (-m gpu): no kernel writes outside its output tensor.

compute-sanitizer is not available on the GPU pool, so this is the bounds check: every kernel that writes an
activation tensor (TMA stores from swizzled staging tiles, pixel-shuffle stores of the ConvTranspose2d, plain 16-byte
vector stores) is given an output carved out of the middle of a larger buffer pre-filled with a sentinel; the bands in
front of and behind the output must come back untouched, and the result must be bit-identical to the same call into a
fresh tensor.  Shapes include ragged tiles (widths that are not multiples of the 30 / 16 / 128-pixel tile widths, pixel
counts that are not multiples of 128) and the shapes that select each kernel (64 output channels -> the row-tap
kernel, Cout % 128 == 0 -> the CTA-pair halo kernels, the direct stem kernel, linear and pixel-shuffle TMA stores).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16
GUARD = 8192  # elements on either side (multiple of 128 bytes for every dtype used, keeps the TMA base alignment)


@pytest.fixture(scope="module")
def ops(lib_built):
    from continual_learning_b200 import _lib, ops as o
    _lib.ensure_device(0)
    return o


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rnd(g, *shape, scale=1.0, dtype=torch.float32):
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


class Banded:
    """an output tensor of `shape` inside a sentinel-filled buffer."""

    def __init__(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        self.n = n
        if dtype.is_floating_point:
            self.buf = torch.full((n + 2 * GUARD,), 12345.0, dtype=dtype, device="cuda")
        else:
            self.buf = torch.full((n + 2 * GUARD,), 113, dtype=dtype, device="cuda")
        self.sentinel = self.buf[0].clone()
        self.out = self.buf[GUARD:GUARD + n].view(shape)

    def check(self, fresh, exact=True):
        torch.cuda.synchronize()
        assert bool((self.buf[:GUARD] == self.sentinel).all()), "wrote in front of the output tensor"
        assert bool((self.buf[GUARD + self.n:] == self.sentinel).all()), "wrote behind the output tensor"
        if exact:
            assert torch.equal(self.out, fresh), "result differs from the same call into a fresh tensor"
        else:  # fp32 reductions in arrival order
            d = (self.out.double() - fresh.double()).norm() / (fresh.double().norm() + 1e-30)
            assert float(d) <= 1e-5, "result differs from the same call into a fresh tensor"


@pytest.mark.parametrize("n,cin,h,w", [(2, 3, 16, 16), (1, 3, 8, 48), (3, 1, 24, 16), (1, 3, 40, 80)])
def test_stem_conv_stays_inside_its_output(ops, n, cin, h, w):
    g = gen(1)
    x, wt, b = rnd(g, n, cin, h, w), rnd(g, 64, cin, 3, 3, scale=0.2), rnd(g, 64)
    wf = ops.pack_stem(wt)
    band = Banded((n, h, w, 64), bf16)
    ops.stem_conv(x, wf, b, relu=True, out=band.out)
    band.check(ops.stem_conv(x, wf, b, relu=True))


@pytest.mark.parametrize("n,h,w,c0,c1,co", [
    (2, 16, 16, 64, 0, 64),     # row-tap kernel, one ragged 30-pixel tile
    (1, 24, 40, 64, 64, 64),    # row-tap kernel, folded concat, 40 = 30 + 10
    (1, 6, 70, 128, 0, 64),     # row-tap kernel, rows not a multiple of the 4-row tile
    (2, 16, 16, 64, 0, 128),    # pair kernel
    (3, 8, 24, 128, 0, 256),    # pair kernel, BN = 256
    (1, 4, 4, 256, 256, 512),   # deep layer, 16 pixels per image
    (2, 3, 5, 64, 64, 64)])     # odd sizes, 15 pixels per image
def test_conv3x3_fprop_and_eval_stay_inside_their_outputs(ops, n, h, w, c0, c1, co):
    g = gen(2)
    x0 = rnd(g, n, h, w, c0, dtype=bf16)
    x1 = rnd(g, n, h, w, c1, dtype=bf16) if c1 else None
    wt, b = rnd(g, co, c0 + c1, 3, 3, scale=0.05), rnd(g, co)
    wf, _ = ops.pack_conv3x3(wt)
    s1, s2 = (torch.zeros(co, device="cuda", dtype=torch.float64) for _ in range(2))
    band = Banded((n, h, w, co), bf16)
    ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(s1, s2), out=band.out)
    band.check(ops.conv3x3_fprop(x0, x1, wf, b, relu=True))
    scale, shift = rnd(g, co), rnd(g, co)
    band = Banded((n, h, w, co), bf16)
    ops.conv3x3_fprop_eval(x0, x1, wf, b, scale, shift, relu=True, out=band.out)
    band.check(ops.conv3x3_fprop_eval(x0, x1, wf, b, scale, shift, relu=True))


@pytest.mark.parametrize("n,h,w,c0,c1,co", [
    (2, 16, 16, 64, 0, 64),     # row-tap kernel
    (1, 24, 40, 64, 0, 128),    # row-tap kernel, two chunks of output channels on the K side
    (2, 8, 8, 128, 128, 128),   # pair kernel, two destinations
    (1, 6, 70, 64, 64, 64),     # two 64-channel destinations
    (1, 4, 4, 512, 0, 1024)])
def test_conv3x3_dgrad_stays_inside_both_destinations(ops, n, h, w, c0, c1, co):
    g = gen(3)
    dy = rnd(g, n, h, w, co, dtype=bf16)
    wt = rnd(g, co, c0 + c1, 3, 3, scale=0.05)
    _, wd = ops.pack_conv3x3(wt)
    b0 = Banded((n, h, w, c0), bf16)
    b1 = Banded((n, h, w, c1), bf16) if c1 else None
    ops.conv3x3_dgrad(dy, wd, c0, c1, out0=b0.out, out1=b1.out if c1 else None)
    f0, f1 = ops.conv3x3_dgrad(dy, wd, c0, c1)
    b0.check(f0)
    if c1:
        b1.check(f1)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 8, 8, 128, 64), (2, 4, 4, 256, 128), (3, 2, 6, 64, 64), (1, 16, 16, 1024, 512),
                                         (1, 8, 128, 128, 64), (1, 3, 256, 128, 64), (1, 5, 10, 128, 64)])
def test_conv_transpose_fprop_and_dgrad_stay_inside_their_outputs(ops, n, h, w, ci, co):
    g = gen(4)
    x = rnd(g, n, h, w, ci, dtype=bf16)
    wt, b = rnd(g, ci, co, 2, 2, scale=0.05), rnd(g, co)
    wf, wd = ops.pack_convT(wt)
    band = Banded((n, 2 * h, 2 * w, co), bf16)
    ops.convT_fprop(x, wf, b, out=band.out)
    band.check(ops.convT_fprop(x, wf, b))
    dy = rnd(g, n, 2 * h, 2 * w, co, dtype=bf16)
    band = Banded((n, h, w, ci), bf16)
    ops.convT_dgrad(dy, wd, out=band.out)
    band.check(ops.convT_dgrad(dy, wd))


@pytest.mark.parametrize("P,K,N", [(128, 64, 64), (1000, 128, 128), (300, 64, 512), (1, 64, 64), (4097, 64, 64)])
def test_gemm_fprop_stays_inside_its_output(ops, P, K, N):
    g = gen(5)
    a, w, b = rnd(g, P, K, dtype=bf16), rnd(g, N, K, scale=0.1, dtype=bf16), rnd(g, N)
    band = Banded((P, N), bf16)
    ops.gemm_fprop(a, w, b, N, relu=True, out=band.out)
    band.check(ops.gemm_fprop(a, w, b, N, relu=True))


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (3, 8, 8, 256), (1, 2, 2, 64), (1, 6, 10, 128)])
def test_batchnorm_and_pool_kernels_stay_inside_their_outputs(ops, n, h, w, c):
    g = gen(6)
    y = rnd(g, n, h, w, c, dtype=bf16)
    scale, shift = rnd(g, c), rnd(g, c)
    band = Banded((n, h, w, c), bf16)
    ops.bn_apply(y, scale, shift, out=band.out)
    band.check(ops.bn_apply(y, scale, shift))
    bz, bp = Banded((n, h, w, c), bf16), Banded((n, h // 2, w // 2, c), bf16)
    fz, fp, fidx = ops.bn_apply_pool(y, scale, shift)
    bidx = Banded(tuple(fidx.shape), fidx.dtype)
    ops.bn_apply_pool(y, scale, shift, z=bz.out, pooled=bp.out, idx=bidx.out)
    bz.check(fz)
    bp.check(fp)
    bidx.check(fidx)
    # pool backward + skip add (+ the BatchNorm-backward sums), and the BatchNorm + ReLU backward apply
    dpooled, skip = rnd(g, n, h // 2, w // 2, c, dtype=bf16), rnd(g, n, h, w, c, dtype=bf16)
    band = Banded((n, h, w, c), bf16)
    ops.maxpool_bwd_add(dpooled, fidx, skip, out=band.out)
    band.check(ops.maxpool_bwd_add(dpooled, fidx, skip))
    if (c // 8) in (8, 16, 32, 64, 128, 256):
        s1, s2 = (torch.zeros(c, device="cuda", dtype=torch.float64) for _ in range(2))
        t1, t2 = (torch.zeros(c, device="cuda", dtype=torch.float64) for _ in range(2))
        band = Banded((n, h, w, c), bf16)
        ops.maxpool_bwd_add_reduce(dpooled, fidx, skip, y, s1, s2, out=band.out)
        band.check(ops.maxpool_bwd_add_reduce(dpooled, fidx, skip, y, t1, t2))
    ka, kb, kc = rnd(g, c), rnd(g, c), rnd(g, c)
    d1, d2 = (torch.zeros(c, device="cuda", dtype=torch.float64) for _ in range(2))
    band = Banded((n, h, w, c), bf16)
    ops.bn_relu_bwd_apply(skip, y, ka, kb, kc, d1, out=band.out)
    band.check(ops.bn_relu_bwd_apply(skip, y, ka, kb, kc, d2))


@pytest.mark.parametrize("n,h,w,c0,c1,co", [(2, 16, 16, 64, 0, 64), (2, 8, 8, 128, 0, 256), (4, 16, 16, 64, 64, 128),
                                              (1, 6, 70, 64, 64, 64), (1, 4, 4, 512, 0, 1024), (2, 3, 5, 64, 0, 64)])
def test_weight_gradient_kernels_stay_inside_their_accumulators(ops, n, h, w, c0, c1, co):
    """the split-K vector REDs of the conv3x3 weight gradient, and its deterministic per-split partial buffers."""
    g = gen(7)
    dy = rnd(g, n, h, w, co, dtype=bf16)
    x0 = rnd(g, n, h, w, c0, dtype=bf16)
    x1 = rnd(g, n, h, w, c1, dtype=bf16) if c1 else None
    band = Banded((9, c0 + c1, co), torch.float32)
    band.out.zero_()
    ops.conv3x3_wgrad(dy, x0, x1, out=band.out)
    band.check(ops.conv3x3_wgrad(dy, x0, x1), exact=False)
    fresh = ops.conv3x3_wgrad_split(dy, x0, x1)
    band = Banded(tuple(fresh.shape), torch.float32)
    ops.conv3x3_wgrad_split(dy, x0, x1, out=band.out)
    band.check(fresh)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 8, 8, 128, 64), (3, 2, 6, 64, 64), (1, 16, 16, 1024, 512), (1, 5, 10, 128, 64)])
def test_conv_transpose_and_gemm_weight_gradients_stay_inside_their_accumulators(ops, n, h, w, ci, co):
    g = gen(8)
    x, dy = rnd(g, n, h, w, ci, dtype=bf16), rnd(g, n, 2 * h, 2 * w, co, dtype=bf16)
    band = Banded((4, ci, co), torch.float32)
    band.out.zero_()
    ops.convT_wgrad(x, dy, out=band.out)
    band.check(ops.convT_wgrad(x, dy), exact=False)
    u, t = rnd(g, n * h * w, 64, dtype=bf16), rnd(g, n * h * w, 64, dtype=bf16)
    band = Banded((64, 64), torch.float32)
    band.out.zero_()
    ops.gemm_wgrad(u, t, out=band.out)
    band.check(ops.gemm_wgrad(u, t), exact=False)


@pytest.mark.parametrize("P,nc,cold", [(1000, 21, 0), (4096, 21, 16), (513, 2, 0), (128 * 37 + 3, 21, 16)])
def test_fused_head_loss_backward_stays_inside_its_outputs(ops, P, nc, cold):
    g = gen(9)
    z = rnd(g, P, 64, dtype=bf16)
    wt, b = rnd(g, nc, 64, 1, 1, scale=0.2), rnd(g, nc)
    wf, wd = ops.pack_head(wt)
    labels = torch.randint(0, nc, (P,), generator=g).cuda()
    old = rnd(g, P, cold) if cold else None
    _, fdz, fdw, _ = ops.head_loss_bwd(z, wf, wd, b, labels, nc, old_logits=old)
    bdz, bdw = Banded((P, 64), bf16), Banded((64, 64), torch.float32)
    bdw.out.zero_()
    ops.head_loss_bwd(z, wf, wd, b, labels, nc, old_logits=old, dz=bdz.out, dw=bdw.out)
    bdz.check(fdz)
    bdw.check(fdw, exact=False)
