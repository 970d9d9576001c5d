"""Tensor-level wrappers over the C ABI (one per clk_* entry point).

Activations are NHWC bf16 torch tensors `[N, H, W, C]`; weights come packed (see `pack_conv3x3`,
`pack_convT`, `pack_head`).  Everything here only allocates outputs and forwards pointers — the
arithmetic is all in libclk.so.
"""
import torch

from . import _lib

bf16 = torch.bfloat16


def _dev(t):
    _lib.ensure_device(t.device.index)


# ------------------------------------------------------------------ layout
def nchw_to_nhwc_bf16(x, cpad=None):
    _dev(x)
    n, c, h, w = x.shape
    cpad = c if cpad is None else cpad
    y = torch.empty((n, h, w, cpad), device=x.device, dtype=bf16)
    _lib.call("clk_nchw_f32_to_nhwc_bf16", x.contiguous(), y, n, c, h, w, cpad)
    return y


def nhwc_to_nchw_f32(x, c=None):
    _dev(x)
    n, h, w, ldc = x.shape
    c = ldc if c is None else c
    y = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    _lib.call("clk_nhwc_to_nchw_f32", x, 1 if x.dtype == torch.float32 else 0, y, n, c, h, w, ldc)
    return y


def im2col_stem(x):
    """x fp32 NCHW [N,Cin,H,W] -> bf16 [N,H,W,64] with k = c*9 + r*3 + s."""
    _dev(x)
    n, c, h, w = x.shape
    a = torch.empty((n, h, w, 64), device=x.device, dtype=bf16)
    _lib.call("clk_im2col3x3_stem", x.contiguous(), a, n, c, h, w)
    return a


def stem_conv_supported(cin, h, w, cout):
    return cin * 9 <= 32 and h % 8 == 0 and w % 16 == 0 and cout == 64


def stem_conv(x, wf, bias, relu=True, stats=None, scale=None, shift=None, out=None):
    """enc1.0 in one launch: x fp32 NCHW [N,Cin,H,W] -> bf16 NHWC [N,H,W,64] = relu(conv3x3(x) + bias) (+ BatchNorm
    statistics, or * scale + shift in inference); the im2col tile only ever exists in shared memory."""
    _dev(x)
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, 64), device=x.device, dtype=bf16) if out is None else out
    s_sum, s_sq = (None, None) if stats is None else stats
    _lib.call("clk_stem_conv3x3_fprop", x.contiguous(), c, wf, bias, y, s_sum, s_sq, scale, shift, n, h, w,
              1 if relu else 0)
    return y


# ------------------------------------------------------------------ weights
def pack_conv3x3(w, out_f=None, out_d=None):
    """w fp32 [Cout,Cin,3,3] -> (fprop pack bf16 [9,Cout,Cin], dgrad pack bf16 [9,Cin,Cout] taps reversed)."""
    _dev(w)
    co, ci = w.shape[0], w.shape[1]
    wf = torch.empty((9, co, ci), device=w.device, dtype=bf16) if out_f is None else out_f
    wd = torch.empty((9, ci, co), device=w.device, dtype=bf16) if out_d is None else out_d
    _lib.call("clk_pack_w", w, wf, wd, co, ci, 9, co, ci, ci, co, 1)
    return wf, wd


def pack_convT(w, out_f=None, out_d=None):
    """w fp32 [Cin,Cout,2,2] -> (fprop pack bf16 [4*Cout,Cin], dgrad pack bf16 [4,Cin,Cout])."""
    _dev(w)
    ci, co = w.shape[0], w.shape[1]
    wf = torch.empty((4 * co, ci), device=w.device, dtype=bf16) if out_f is None else out_f
    wd = torch.empty((4, ci, co), device=w.device, dtype=bf16) if out_d is None else out_d
    _lib.call("clk_pack_w", w, wd, wf, ci, co, 4, ci, co, co, ci, 0)
    return wf, wd


def pack_stem(w, out_f=None):
    """w fp32 [Cout,Cin,3,3] (Cin*9 <= 64) -> bf16 [Cout,64] (k = c*9+r*3+s, zero padded)."""
    _dev(w)
    co = w.shape[0]
    k = w.shape[1] * 9
    wf = torch.zeros((co, 64), device=w.device, dtype=bf16) if out_f is None else out_f
    _lib.call("clk_pack_w", w, wf, None, co, k, 1, co, 64, 0, 0, 0)
    return wf


def pack_head(w, out_f=None, out_d=None):
    """w fp32 [ncls,Cin,1,1] -> (fprop pack bf16 [32,Cin] zero padded rows, dgrad pack bf16 [Cin,64] zero padded cols)."""
    _dev(w)
    nc, ci = w.shape[0], w.shape[1]
    assert nc <= 32
    wf = torch.zeros((32, ci), device=w.device, dtype=bf16) if out_f is None else out_f
    wd = torch.zeros((ci, 64), device=w.device, dtype=bf16) if out_d is None else out_d
    _lib.call("clk_pack_w", w, wf, wd, nc, ci, 1, 32, ci, ci, 64, 0)
    return wf, wd


def unpack_wgrad(dpacked, grad, a, b, t, lda, ldb, alpha=1.0, accumulate=False, transposed=False):
    """grad[A][B][T] = alpha * D[T][lda][ldb]  (transposed: D is [T][ldb][lda], the conv3x3 wgrad layout)."""
    _lib.call("clk_unpack_wgrad", dpacked, grad, a, b, t, lda, ldb, float(alpha), 1 if accumulate else 0,
              1 if transposed else 0)
    return grad


# ------------------------------------------------------------------ implicit GEMMs
def conv3x3_fprop(x0, x1, wf, bias, relu=True, stats=None, out=None):
    _dev(x0)
    n, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[3]
    cout = wf.shape[1]
    y = torch.empty((n, h, w, cout), device=x0.device, dtype=bf16) if out is None else out
    s_sum, s_sq = (None, None) if stats is None else stats
    _lib.call("clk_conv3x3_fprop", x0, c0, x1, c1, wf, bias, y, s_sum, s_sq, n, h, w, cout, 1 if relu else 0)
    return y


def conv3x3_fprop_eval(x0, x1, wf, bias, scale, shift, relu=True, out=None):
    """inference: z = relu(conv(x) + bias) * scale + shift in one launch (BatchNorm with running statistics)."""
    _dev(x0)
    n, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[3]
    cout = wf.shape[1]
    z = torch.empty((n, h, w, cout), device=x0.device, dtype=bf16) if out is None else out
    _lib.call("clk_conv3x3_fprop_eval", x0, c0, x1, c1, wf, bias, z, scale, shift, n, h, w, cout, 1 if relu else 0)
    return z


def conv3x3_dgrad(dy, wd, c0, c1=0, out0=None, out1=None):
    _dev(dy)
    n, h, w, cout = dy.shape
    dx0 = torch.empty((n, h, w, c0), device=dy.device, dtype=bf16) if out0 is None else out0
    dx1 = None
    if c1:
        dx1 = torch.empty((n, h, w, c1), device=dy.device, dtype=bf16) if out1 is None else out1
    _lib.call("clk_conv3x3_dgrad", dy, cout, wd, dx0, c0, dx1, c1, n, h, w)
    return dx0, dx1


def conv3x3_wgrad(dy, x0, x1=None, out=None):
    """returns fp32 [9, Cin, Cout] — tap, input channel, output channel (accumulates into `out` when given)."""
    _dev(dy)
    n, h, w, cout = dy.shape
    c0 = x0.shape[3]
    c1 = 0 if x1 is None else x1.shape[3]
    dw = torch.zeros((9, c0 + c1, cout), device=dy.device, dtype=torch.float32) if out is None else out
    _lib.call("clk_conv3x3_wgrad", dy, cout, x0, c0, x1, c1, dw, n, h, w)
    return dw


def conv3x3_wgrad_splits(cout, cin, n, h, w):
    """number of split-K partial buffers clk_conv3x3_wgrad_split writes for this shape (host-side query)."""
    k = _lib.load().clk_conv3x3_wgrad_splits(cout, cin, n, h, w)
    if k < 0:
        _lib.check(k)
    return k


def conv3x3_wgrad_split(dy, x0, x1=None, out=None):
    """deterministic split-K weight gradient: returns fp32 [nsplit, 9, Cin, Cout] partial sums (plain stores)."""
    _dev(dy)
    n, h, w, cout = dy.shape
    c0 = x0.shape[3]
    c1 = 0 if x1 is None else x1.shape[3]
    if out is None:
        out = torch.empty((conv3x3_wgrad_splits(cout, c0 + c1, n, h, w), 9, c0 + c1, cout), device=dy.device,
                          dtype=torch.float32)
    _lib.call("clk_conv3x3_wgrad_split", dy, cout, x0, c0, x1, c1, out, n, h, w)
    return out


def gemm_fprop(a, w, bias, n_store, out_f32=False, relu=False, stats=None, out=None):
    """a bf16 [..., K] (rows = pixels), w bf16 [Npad, K]; returns [..., n_store or Npad]."""
    _dev(a)
    k = a.shape[-1]
    p = a.numel() // k
    npad = w.shape[0]
    if out is None:
        ldo = n_store
        out = torch.empty((*a.shape[:-1], ldo), device=a.device, dtype=torch.float32 if out_f32 else bf16)
    ldo = out.shape[-1]
    s_sum, s_sq = (None, None) if stats is None else stats
    _lib.call("clk_gemm_fprop", a, k, w, bias, out, ldo, n_store, 1 if out_f32 else 0, 1 if relu else 0, s_sum,
              s_sq, p, npad)
    return out


def gemm_fprop_eval(a, w, bias, n_store, scale, shift, relu=True, out=None):
    """inference form of gemm_fprop (bf16 output): relu(a w^T + bias) * scale + shift."""
    _dev(a)
    k = a.shape[-1]
    p = a.numel() // k
    if out is None:
        out = torch.empty((*a.shape[:-1], n_store), device=a.device, dtype=bf16)
    _lib.call("clk_gemm_fprop_eval", a, k, w, bias, out, out.shape[-1], n_store, 1 if relu else 0, scale, shift, p,
              w.shape[0])
    return out


def gemm_wgrad(u, t, out=None):
    """out fp32 [CU, CT] += u^T t over pixels."""
    _dev(u)
    cu, ct = u.shape[-1], t.shape[-1]
    p = u.numel() // cu
    if out is None:
        out = torch.zeros((cu, ct), device=u.device, dtype=torch.float32)
    _lib.call("clk_gemm_wgrad", u, cu, t, ct, out, out.shape[0], out.shape[1], p)
    return out


def convT_fprop(x, wf, bias, out=None):
    _dev(x)
    n, h, w, cin = x.shape
    cout = wf.shape[0] // 4
    y = torch.empty((n, 2 * h, 2 * w, cout), device=x.device, dtype=bf16) if out is None else out
    _lib.call("clk_convT2x2_fprop", x, wf, bias, y, n, h, w, cin, cout)
    return y


def convT_dgrad(dy, wd, out=None):
    _dev(dy)
    n, h2, w2, cout = dy.shape
    cin = wd.shape[1]
    dx = torch.empty((n, h2 // 2, w2 // 2, cin), device=dy.device, dtype=bf16) if out is None else out
    _lib.call("clk_convT2x2_dgrad", dy, wd, dx, n, h2 // 2, w2 // 2, cin, cout)
    return dx


def convT_wgrad(x, dy, out=None):
    """returns fp32 [4, Cin, Cout]."""
    _dev(x)
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    dw = torch.zeros((4, cin, cout), device=x.device, dtype=torch.float32) if out is None else out
    _lib.call("clk_convT2x2_wgrad", x, dy, dw, n, h, w, cin, cout)
    return dw


# ------------------------------------------------------------------ BatchNorm / pool
def bn_stats(y, s_sum, s_sq):
    c = y.shape[-1]
    _lib.call("clk_bn_stats", y, s_sum, s_sq, y.numel() // c, c)


def bn_finalize(s_sum, s_sq, gamma, beta, rmean, rvar, mean, invstd, scale, shift, count, eps=1e-5, momentum=0.1,
                training=True):
    _lib.call("clk_bn_finalize", s_sum, s_sq, gamma, beta, rmean, rvar, mean, invstd, scale, shift, gamma.numel(),
              float(count), float(eps), float(momentum), 1 if training else 0)


def bn_apply(y, scale, shift, out=None):
    c = y.shape[-1]
    z = torch.empty_like(y) if out is None else out
    _lib.call("clk_bn_apply", y, z, scale, shift, y.numel() // c, c)
    return z


def bn_apply_pool(y, scale, shift, z=None, pooled=None, idx=None):
    n, h, w, c = y.shape
    if scale is not None and z is None:
        z = torch.empty_like(y)
    if pooled is None:
        pooled = torch.empty((n, h // 2, w // 2, c), device=y.device, dtype=bf16)
    if idx is None:
        idx = torch.empty((n, h // 2, w // 2, c), device=y.device, dtype=torch.uint8)
    _lib.call("clk_bn_apply_pool", y, z, pooled, idx, scale, shift, n, h, w, c)
    return z, pooled, idx


def maxpool_bwd_add(dpooled, idx, skip, out=None):
    n, ho, wo, c = dpooled.shape
    din = torch.empty((n, 2 * ho, 2 * wo, c), device=dpooled.device, dtype=bf16) if out is None else out
    _lib.call("clk_maxpool_bwd_add", dpooled, idx, skip, din, n, 2 * ho, 2 * wo, c)
    return din


def maxpool_bwd_add_reduce(dpooled, idx, skip, y, s1, s2, out=None):
    """maxpool_bwd_add + the BatchNorm-backward reductions of the pooled layer in one pass (s1 += sum din,
    s2 += sum din * y)."""
    n, ho, wo, c = dpooled.shape
    din = torch.empty((n, 2 * ho, 2 * wo, c), device=dpooled.device, dtype=bf16) if out is None else out
    _lib.call("clk_maxpool_bwd_add_reduce", dpooled, idx, skip, y, din, s1, s2, n, 2 * ho, 2 * wo, c)
    return din


def bn_bwd_reduce(dz, y, s1, s2):
    c = y.shape[-1]
    _lib.call("clk_bn_bwd_reduce", dz, y, s1, s2, y.numel() // c, c)


def bn_bwd_finalize(s1, s2, gamma, mean, invstd, dgamma, dbeta, ka, kb, kc, count, training=True, accumulate=False):
    _lib.call("clk_bn_bwd_finalize", s1, s2, gamma, mean, invstd, dgamma, dbeta, ka, kb, kc, gamma.numel(),
              float(count), 1 if training else 0, 1 if accumulate else 0)


def bn_relu_bwd_apply(dz, y, ka, kb, kc, dbias, out=None):
    c = y.shape[-1]
    dpre = torch.empty_like(y) if out is None else out
    _lib.call("clk_bn_relu_bwd_apply", dz, y, dpre, ka, kb, kc, dbias, y.numel() // c, c)
    return dpre


def channel_sum(g, out):
    c = g.shape[-1]
    _lib.call("clk_channel_sum", g, out, g.numel() // c, c)


def f64_to_f32(src, dst, n=None, ld_group=0, groups=1, alpha=1.0, accumulate=False):
    n = dst.numel() if n is None else n
    _lib.call("clk_f64_to_f32", src, dst, n, ld_group, groups, float(alpha), 1 if accumulate else 0)


# ------------------------------------------------------------------ loss / metrics / optimiser
def ce_kd_loss(logits, labels, old_logits=None, T=2.0, lam=1.0, gscale=None, dlogits=None, loss_acc=None,
               err_flag=None, ldd=64):
    """logits fp32 [..., C] (rows = pixels), labels int64; returns (loss_acc f64[2], dlogits bf16 [..., ldd])."""
    c = logits.shape[-1]
    p = logits.numel() // c
    cold = 0 if old_logits is None else old_logits.shape[-1]
    if dlogits is None:
        dlogits = torch.empty((*logits.shape[:-1], ldd), device=logits.device, dtype=bf16)
    if loss_acc is None:
        loss_acc = torch.zeros(2, device=logits.device, dtype=torch.float64)
    gscale = 1.0 / p if gscale is None else gscale
    _lib.call("clk_ce_kd_loss", logits, old_logits, labels, p, c, cold, float(T), float(lam), float(gscale),
              dlogits, dlogits.shape[-1], loss_acc, err_flag)
    return loss_acc, dlogits


def head_loss_bwd(z, wf, wd, bias, labels, num_classes, old_logits=None, T=2.0, lam=1.0, gscale=None, dz=None, dw=None,
                  dbias=None, loss_acc=None, err_flag=None):
    """1x1 head + CE (+ distillation) + head backward in one launch (the logits never leave tensor memory).
    z bf16 [..., 64]; returns (loss_acc f64[2], dz bf16 like z, dw f32 [64, 64] (rows = class), dbias f64[64])."""
    _dev(z)
    cin = z.shape[-1]
    p = z.numel() // cin
    cold = 0 if old_logits is None else old_logits.shape[-1]
    dev = z.device
    dz = torch.empty_like(z) if dz is None else dz
    dw = torch.zeros((64, cin), device=dev, dtype=torch.float32) if dw is None else dw
    dbias = torch.zeros(64, device=dev, dtype=torch.float64) if dbias is None else dbias
    loss_acc = torch.zeros(2, device=dev, dtype=torch.float64) if loss_acc is None else loss_acc
    gscale = 1.0 / p if gscale is None else gscale
    _lib.call("clk_head_loss_bwd", z, wf, wd, bias, labels, old_logits, p, cin, num_classes, cold, float(T), float(lam),
              float(gscale), dz, dw, dbias, loss_acc, err_flag)
    return loss_acc, dz, dw, dbias


def head_argmax_confusion(z, wf, bias, labels, num_classes, nc=None, want_pred=False, conf=None, correct=None):
    """1x1 head + argmax + correct count + confusion matrix in one launch (the logits never leave tensor memory).
    z bf16 [..., 64]; returns (pred int64 or None, conf int64 [nc*nc] or None, correct int64 [1])."""
    _dev(z)
    cin = z.shape[-1]
    p = z.numel() // cin
    pred = torch.empty(z.shape[:-1], device=z.device, dtype=torch.int64) if want_pred else None
    if nc is not None and conf is None:
        conf = torch.zeros(nc * nc, device=z.device, dtype=torch.int64)
    if correct is None:
        correct = torch.zeros(1, device=z.device, dtype=torch.int64)
    _lib.call("clk_head_argmax_confusion", z, wf, bias, labels, p, cin, num_classes,
              nc if nc is not None else num_classes, pred, conf if nc is not None else None, correct)
    return pred, conf, correct


def confusion_matrix(target, pred, nc, conf=None, err_flag=None):
    if conf is None:
        conf = torch.zeros(nc * nc, device=target.device, dtype=torch.int64)
    _lib.call("clk_confusion_matrix", target, pred, target.numel(), nc, conf, err_flag)
    return conf


def argmax_confusion(logits, labels, nc=None, want_pred=False, conf=None, correct=None):
    c = logits.shape[-1]
    p = logits.numel() // c
    pred = torch.empty(logits.shape[:-1], device=logits.device, dtype=torch.int64) if want_pred else None
    if nc is not None and conf is None:
        conf = torch.zeros(nc * nc, device=logits.device, dtype=torch.int64)
    if correct is None:
        correct = torch.zeros(1, device=logits.device, dtype=torch.int64)
    _lib.call("clk_argmax_confusion", logits, labels, p, c, nc if nc is not None else 0, pred, conf, correct)
    return pred, conf, correct
