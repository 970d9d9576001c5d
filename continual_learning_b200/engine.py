"""UNetEngine — launch schedule of the U-Net forward / backward over libclk kernels.

Reference path: models/unet.py:74-92 (forward) and the autograd graph behind trainer.py:175.
Data layout in HBM: every activation is NHWC bf16; per conv->ReLU->BN unit the engine keeps
`y = relu(conv(x)+b)` and `z = BN(y)` (plus the 2x2-pooled `z` and its window index for the encoder
outputs).  Skip-connection concats (models/unet.py:83-87) are never materialised: the conv kernels
walk the channels of the two source tensors back to back.  fp32 master parameters stay in the
nn.Module; bf16 packed operand copies are refreshed whenever a parameter version changes.
"""
import os

import torch

from . import _lib, ops

bf16 = torch.bfloat16
f32 = torch.float32
f64 = torch.float64


class _Unit:
    """one Conv3x3(+bias) -> ReLU -> BatchNorm2d triple (models/unet.py:13-15)."""

    def __init__(self, conv, bn, c0, c1, cout, stem=False):
        self.conv, self.bn, self.c0, self.c1, self.cout, self.stem = conv, bn, c0, c1, cout, stem
        self.x0 = self.x1 = self.y = self.z = self.pooled = self.idx = None


class UNetEngine:
    def __init__(self, module):
        self.m = module
        m = module
        c = m.conv_dim
        if c % 64 != 0:
            raise ValueError("the sm_100a kernels need conv_dim to be a multiple of 64")
        if m.in_dim * 9 > 64:
            raise ValueError("the stem kernel needs in_dim*9 <= 64")
        if m.num_classes > 32:
            raise ValueError("the head kernel needs num_classes <= 32")
        e1 = m.enc1
        self.units = [_Unit(e1[0], e1[2], 64, 0, c, stem=True), _Unit(e1[3], e1[5], c, 0, c)]
        for blk, ci, co in ((m.enc2, c, 2 * c), (m.enc3, 2 * c, 4 * c), (m.enc4, 4 * c, 8 * c)):
            b = blk.block
            self.units += [_Unit(b[1], b[3], ci, 0, co), _Unit(b[4], b[6], co, 0, co)]
        self.convT = []
        first = True
        for blk, ci, cm, co in ((m.dec1, 8 * c, 16 * c, 8 * c), (m.dec2, 16 * c, 8 * c, 4 * c),
                                (m.dec3, 8 * c, 4 * c, 2 * c), (m.dec4, 4 * c, 2 * c, c)):
            b = blk.block
            c0, c1 = (ci, 0) if first else (ci // 2, ci // 2)
            first = False
            self.units += [_Unit(b[0], b[2], c0, c1, cm), _Unit(b[3], b[5], cm, 0, cm)]
            self.convT.append((b[6], cm, co))
        self.units += [_Unit(m.last[0], m.last[2], c, c, c), _Unit(m.last[3], m.last[5], c, 0, c)]
        self.head = m.last[6]
        self._dev = None
        self._wver = None
        self._job_ptrs = None
        self._side = None
        self._pending = []
        self.use_side_stream = True
        # split-K weight gradients: False = fp32 vector REDs into one L2-resident buffer per layer (fastest);
        # True = per-split partial buffers summed in a fixed order (that kernel is bit-reproducible; the STEP is not:
        # BatchNorm statistics still use atomics) , ~0.15 ms/step slower
        self.deterministic = False
        self._forked = False
        self.side_pack = True    # late-layer weight packing on the side stream
        self.training_fwd = True
        self.save_for_backward = True
        self.logits = None
        self._stem_direct = False
        self._stem_cols = None
        self.generation = 0  # bumped by every forward (stale-backward detection in unet._UNetFn)

    # ------------------------------------------------------------------ persistent buffers
    def _setup(self, dev):
        if self._dev == dev:
            return
        self._dev = dev
        self.params = list(self.m.parameters())
        c_dim = self.m.conv_dim
        n = sum(p.numel() for p in self.params)
        self.G = torch.zeros(n, device=dev, dtype=f32)  # flat .grad storage, PyTorch layouts
        self.gview = {}
        off = 0
        for p in self.params:
            self.gview[p] = self.G[off:off + p.numel()].view_as(p)
            off += p.numel()
        # packed fp32 weight-gradient accumulators + packed bf16 operand copies
        sizes = []
        for u in self.units:
            sizes.append(64 * 64 if u.stem else 9 * u.cout * (u.c0 + u.c1))
        for (_, cm, co) in self.convT:
            sizes.append(4 * cm * co)
        sizes.append(64 * c_dim)
        self.Gp = torch.zeros(sum(sizes), device=dev, dtype=f32)
        views, off = [], 0
        for s in sizes:
            views.append(self.Gp[off:off + s])
            off += s
        nu = len(self.units)
        for i, u in enumerate(self.units):
            u.gp = views[i].view(64, 64) if u.stem else views[i].view(9, u.c0 + u.c1, u.cout)
            u.part, u.nsplit = None, 0
            if u.stem:
                u.wf = torch.zeros((u.cout, 64), device=dev, dtype=bf16)
                u.wd = None
            else:
                u.wf = torch.empty((9, u.cout, u.c0 + u.c1), device=dev, dtype=bf16)
                u.wd = torch.empty((9, u.c0 + u.c1, u.cout), device=dev, dtype=bf16)
        self.tgp, self.twf, self.twd = [], [], []
        for j, (_, cm, co) in enumerate(self.convT):
            self.tgp.append(views[nu + j].view(4, cm, co))
            self.twf.append(torch.empty((4 * co, cm), device=dev, dtype=bf16))
            self.twd.append(torch.empty((4, cm, co), device=dev, dtype=bf16))
        self.hgp = views[-1].view(64, c_dim)
        self.hwf = torch.zeros((32, self.m.conv_dim), device=dev, dtype=bf16)
        self.hwd = torch.zeros((self.m.conv_dim, 64), device=dev, dtype=bf16)
        # fp64 per-channel accumulators: forward (sum, sq) and backward (s1, s2, dbias) per unit,
        # dbias per convT, dbias of the head
        fw = sum(2 * u.cout for u in self.units)
        bw = sum(3 * u.cout for u in self.units) + sum(co for (_, _, co) in self.convT) + 64
        self.acc_f = torch.zeros(fw, device=dev, dtype=f64)
        self.acc_b = torch.zeros(bw, device=dev, dtype=f64)
        of, ob = 0, 0
        for u in self.units:
            u.s_sum, u.s_sq = self.acc_f[of:of + u.cout], self.acc_f[of + u.cout:of + 2 * u.cout]
            of += 2 * u.cout
            u.s1, u.s2, u.dbias = (self.acc_b[ob + k * u.cout:ob + (k + 1) * u.cout] for k in range(3))
            ob += 3 * u.cout
            # fp32 per-channel vectors: mean, invstd, scale, shift, kA, kB, kC
            u.vec = torch.empty((7, u.cout), device=dev, dtype=f32)
        self.tdbias = []
        for (_, _, co) in self.convT:
            self.tdbias.append(self.acc_b[ob:ob + co])
            ob += co
        self.hdbias = self.acc_b[ob:ob + 64]
        self._wver = None
        self._job_ptrs = None
        self._wg_shape = None

    def _ensure_wgrad_bufs(self, n, h, w):
        """split-K partial buffers of the conv3x3 weight gradients: [nsplit][9][Cin][Cout] fp32 per layer; the number
        of splits depends on the feature-map size (clk_conv3x3_wgrad_splits)."""
        if not self.deterministic or self._wg_shape == (n, h, w):
            return
        self._wg_shape = (n, h, w)
        down = [1, 1, 2, 2, 4, 4, 8, 8, 16, 16, 8, 8, 4, 4, 2, 2, 1, 1]
        for u, d in zip(self.units, down):
            if u.stem:
                continue
            cin = u.c0 + u.c1
            u.nsplit = ops.conv3x3_wgrad_splits(u.cout, cin, n, h // d, w // d)
            u.part = torch.empty((u.nsplit, 9, cin, u.cout), device=self._dev, dtype=f32)
        self._job_ptrs = None  # the unpack table points at the partial buffers

    def _pack_weights(self):
        ptrs = tuple(p.data_ptr() for p in self.params)
        if ptrs != self._job_ptrs:  # parameters moved, or the split-K buffers were re-sized
            self._build_jobs()
            self._job_ptrs = ptrs
            self._wver = None
        ver = (_lib.param_epoch,) + tuple(p._version for p in self.params) + tuple(p.data_ptr() for p in self.params[:2])
        if ver == self._wver:
            return
        # the first two layers' weights on this stream; the other 21 layers (99.9 % of the bytes) on the side
        # stream, overlapping im2col + enc1 (forward() joins before enc2)
        tab, n, tiles = self.pack_jobs[0]
        _lib.call("clk_pack_w_multi", tab, n, tiles, 9)
        tab, n, tiles = self.pack_jobs[1]
        if self.side_pack:
            with self._fork(None):
                _lib.call("clk_pack_w_multi", tab, n, tiles, 9)
        else:
            _lib.call("clk_pack_w_multi", tab, n, tiles, 9)
        self._wver = ver

    # ------------------------------------------------------------------ batched job tables
    @staticmethod
    def _tiles(a, b):
        return ((a + 31) // 32) * ((b + 31) // 32), (b + 31) // 32

    def _build_jobs(self):
        """int64[16] rows for clk_pack_w_multi / clk_unpack_wgrad_multi / clk_f64_to_f32_multi (include/clk.h)."""
        import struct
        one = struct.unpack("q", struct.pack("d", 1.0))[0]
        dev = self._dev
        c = self.m.conv_dim
        nc = self.m.num_classes
        pack, t0 = [[], []], [0, 0]
        # gradient groups in the order the backward pass finishes them: 0 = head + decoder, 1 = enc4, 2 = enc3..enc1
        NG = 3
        unpack = [[] for _ in range(NG)]
        reduce_, rb0 = [[] for _ in range(NG)], [0] * NG
        cvt = [[] for _ in range(NG)]
        ut0 = [0] * NG

        def add_pack(src, ab, ba, A, B, T, ldA, ldB, ldB2, ldA2, rev, grp=1):
            n, tb = self._tiles(A, B)
            pack[grp].append([src.data_ptr(), ab.data_ptr() if ab is not None else 0,
                              ba.data_ptr() if ba is not None else 0, A, B, T, ldA, ldB, ldB2, ldA2, rev, t0[grp], tb,
                              0, 0, 0])
            t0[grp] += n

        def add_unpack(g, D, grad, A, B, T, ldA, ldB, transposed=0, nsplit=1, sstride=0):
            n, tb = self._tiles(A, B)
            unpack[g].append([D.data_ptr(), grad.data_ptr(), A, B, T, ldA, ldB, one, 0, ut0[g], tb, transposed, nsplit,
                              sstride, 0, 0])
            ut0[g] += n

        def add_cvt(g, src, dst, n):
            cvt[g].append([src.data_ptr(), dst.data_ptr(), n, 0, 1, one, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])

        for i, u in enumerate(self.units):
            g = 0 if i >= 8 else (1 if i >= 6 else 2)
            w = u.conv.weight
            if u.stem:
                k = self.m.in_dim * 9
                add_pack(w, u.wf, None, u.cout, k, 1, u.cout, 64, 0, 0, 0, grp=0)
                add_unpack(g, u.gp, self.gview[w], u.cout, k, 1, 64, 64)
            else:
                ci = u.c0 + u.c1
                add_pack(w, u.wf, u.wd, u.cout, ci, 9, u.cout, ci, ci, u.cout, 1, grp=0 if i < 2 else 1)
                add_unpack(g, u.part if self.deterministic else u.gp, self.gview[w], u.cout, ci, 9, u.cout, ci,
                           transposed=1)
                if self.deterministic and u.nsplit > 1:  # sum the split-K partials into split 0 first (one batched launch per group)
                    nvec = 9 * ci * u.cout // 4
                    reduce_[g].append([u.part.data_ptr(), nvec, u.nsplit, nvec, rb0[g]] + [0] * 11)
                    rb0[g] += (nvec + 255) // 256
            add_cvt(g, u.dbias, self.gview[u.conv.bias], u.cout)
        for j, (mod, cm, co) in enumerate(self.convT):
            add_pack(mod.weight, self.twd[j], self.twf[j], cm, co, 4, cm, co, co, cm, 0)
            add_unpack(0, self.tgp[j], self.gview[mod.weight], cm, co, 4, cm, co)
            add_cvt(0, self.tdbias[j], self.gview[mod.bias], co)
        add_pack(self.head.weight, self.hwf, self.hwd, nc, c, 1, 32, c, c, 64, 0)
        add_unpack(0, self.hgp, self.gview[self.head.weight], nc, c, 1, 64, c)
        add_cvt(0, self.hdbias, self.gview[self.head.bias], nc)

        def dev_table(rows):
            return torch.tensor(rows, dtype=torch.int64).to(dev)

        self.pack_jobs = [(dev_table(pack[g]), len(pack[g]), t0[g]) for g in range(2)]
        self.unpack_jobs = [(dev_table(unpack[g]), len(unpack[g]), ut0[g]) for g in range(NG)]
        self.cvt_jobs = [(dev_table(cvt[g]), len(cvt[g])) for g in range(NG)]
        self.reduce_jobs = [((dev_table(reduce_[g]) if reduce_[g] else None), len(reduce_[g]), rb0[g]) for g in range(NG)]

    def _flush_grads(self, g):
        self._join()  # weight gradients run on the side stream
        tab, n, blocks = self.reduce_jobs[g]
        if n:
            _lib.call("clk_reduce_partials_multi", tab, n, blocks)
        tab, n, tiles = self.unpack_jobs[g]
        _lib.call("clk_unpack_wgrad_multi", tab, n, tiles, 9)
        tab, n = self.cvt_jobs[g]
        _lib.call("clk_f64_to_f32_multi", tab, n)

    # ------------------------------------------------------------------ forward
    def _unit_fwd(self, u, x0, x1, training, pool=False):
        n, h, w = x0.shape[0], x0.shape[1], x0.shape[2]
        if u.stem and self._stem_direct:  # x0 is the fp32 NCHW input itself
            h, w = x0.shape[2], x0.shape[3]
        u.x0, u.x1 = x0, x1
        if not training and not self.save_for_backward:
            # inference: BatchNorm with running statistics is an affine map known before the conv runs ->
            # applied in the conv epilogue, no pre-BN tensor and no separate BN pass
            bn = u.bn
            ops.bn_finalize(u.s_sum, u.s_sq, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                            u.vec[0], u.vec[1], u.vec[2], u.vec[3], n * h * w, eps=bn.eps, momentum=0.0,
                            training=False)
            bias = u.conv.bias.detach()
            if u.stem and self._stem_direct:
                u.z = ops.stem_conv(x0, u.wf, bias, relu=True, scale=u.vec[2], shift=u.vec[3])
            elif u.stem:
                u.z = ops.gemm_fprop_eval(x0, u.wf, bias, u.cout, u.vec[2], u.vec[3], relu=True)
            else:
                u.z = ops.conv3x3_fprop_eval(x0, x1, u.wf, bias, u.vec[2], u.vec[3], relu=True)
            u.y, u.pooled, u.idx = None, None, None
            if pool:
                _, u.pooled, u.idx = ops.bn_apply_pool(u.z, None, None)
            return u.z
        stats = (u.s_sum, u.s_sq) if training else None
        bias = u.conv.bias.detach()
        if u.stem and self._stem_direct:
            u.y = ops.stem_conv(x0, u.wf, bias, relu=True, stats=stats)  # x0 = the fp32 NCHW input itself
        elif u.stem:
            u.y = ops.gemm_fprop(x0, u.wf, bias, u.cout, relu=True, stats=stats)
        else:
            u.y = ops.conv3x3_fprop(x0, x1, u.wf, bias, relu=True, stats=stats)
        bn = u.bn
        mean, invstd, scale, shift = u.vec[0], u.vec[1], u.vec[2], u.vec[3]
        ops.bn_finalize(u.s_sum, u.s_sq, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, mean,
                        invstd, scale, shift, n * h * w, eps=bn.eps,
                        momentum=0.1 if bn.momentum is None else bn.momentum, training=training)
        if pool:
            u.z, u.pooled, u.idx = ops.bn_apply_pool(u.y, scale, shift)
        else:
            u.z, u.pooled, u.idx = ops.bn_apply(u.y, scale, shift), None, None
        return u.z

    def forward(self, x, training=True, save_for_backward=True, head=True):
        """x: fp32 NCHW CUDA tensor. Returns fp32 logits [N, H, W, num_classes] (NHWC memory).
        save_for_backward=False (with training=False) selects the fused inference kernels.
        head=False stops before the 1x1 head and returns its bf16 input [N, H, W, 64] (TrainStep fuses the head with
        the loss and the head's backward, see `head_loss_backward`)."""
        self.save_for_backward = save_for_backward
        self.generation += 1
        if not x.is_cuda:
            raise RuntimeError("continual_learning_b200.UNet runs on CUDA (sm_100a) only: there is no CPU fallback")
        _lib.ensure_device(x.device.index)
        n, cin, h, w = x.shape
        if h % 16 or w % 16:
            raise ValueError("input H and W must be multiples of 16 (four 2x2 pools, models/unet.py:76-80)")
        if cin != self.m.in_dim:
            raise ValueError(f"expected {self.m.in_dim} input channels, got {cin}")
        self._setup(x.device)
        self._ensure_wgrad_bufs(n, h, w)
        self._pack_weights()
        self.training_fwd = training
        if training:
            self.acc_f.zero_()
        U = self.units
        # the stem runs as ONE kernel that builds its im2col tiles in shared memory; the K = 64 im2col matrix is only
        # materialised (on the weight-gradient side stream, in the backward pass) for the stem's wgrad
        xf = x.float().contiguous()
        self._stem_direct = ops.stem_conv_supported(cin, h, w, U[0].cout) and os.environ.get("CLK_STEM_DIRECT", "1") != "0"
        a = xf if self._stem_direct else ops.im2col_stem(xf)
        self._stem_cols = None
        if self._stem_direct and save_for_backward:
            # the stem's weight gradient still wants the K = 64 im2col matrix: written on the side stream while the
            # (tensor-bound) first layers run, instead of at the tail of the backward pass
            with self._fork(xf):
                self._stem_cols = ops.im2col_stem(xf)
        z = self._unit_fwd(U[0], a, None, training)
        self._unit_fwd(U[1], z, None, training, pool=True)
        self._join()  # packed weights of the remaining layers (side stream)
        for k in (2, 4, 6):
            z = self._unit_fwd(U[k], U[k - 1].pooled, None, training)
            self._unit_fwd(U[k + 1], z, None, training, pool=True)
        skips = [U[7], U[5], U[3], U[1]]
        up = None
        self.tin = []
        for j in range(4):
            k = 8 + 2 * j
            if j == 0:
                z = self._unit_fwd(U[k], U[7].pooled, None, training)
            else:
                z = self._unit_fwd(U[k], skips[j - 1].z, up, training)
            z = self._unit_fwd(U[k + 1], z, None, training)
            self.tin.append(z)
            up = ops.convT_fprop(z, self.twf[j], self.convT[j][0].bias.detach())
        z = self._unit_fwd(U[16], U[1].z, up, training)
        z = self._unit_fwd(U[17], z, None, training)
        if training:
            bufs = [u.bn.num_batches_tracked for u in U if u.bn.num_batches_tracked is not None]
            if bufs:
                torch._foreach_add_(bufs, 1)
        if not head:
            self.logits = None
            return z
        self.logits = ops.gemm_fprop(z, self.hwf, self.head.bias.detach(), self.m.num_classes, out_f32=True)
        return self.logits

    def head_loss_backward(self, labels, loss_acc, old_logits=None, T=2.0, lam=1.0, err_flag=None, after_decoder=None,
                           after_group=None):
        """after forward(head=False): 1x1 head + CrossEntropy (+ distillation) + the head's backward in one launch,
        then the rest of the backward pass.  Adds {sum CE, sum KL} to loss_acc (f64[2]); returns the gradient views."""
        for g in self.head_loss_backward_segments(labels, loss_acc, old_logits, T, lam, err_flag):
            self._group_done(g, after_decoder, after_group)
        return [self.gview[p] for p in self.params]

    def head_loss_backward_segments(self, labels, loss_acc, old_logits=None, T=2.0, lam=1.0, err_flag=None):
        """generator form: yields the index of each gradient group (0 = head + decoder, 1 = enc4, 2 = enc3..enc1) as
        soon as its gradients are final in the flat buffer `G` — a data-parallel caller all-reduces that group while
        the rest of the backward pass runs (and may capture each segment in its own CUDA graph)."""
        U = self.units
        z = U[17].z
        self._begin_backward()
        n, h, w = z.shape[0], z.shape[1], z.shape[2]
        _, dz, _, _ = ops.head_loss_bwd(z, self.hwf, self.hwd, self.head.bias.detach(), labels, self.m.num_classes,
                                        old_logits=old_logits, T=T, lam=lam, gscale=1.0 / (n * h * w), dw=self.hgp,
                                        dbias=self.hdbias, loss_acc=loss_acc, err_flag=err_flag)
        yield from self.backward_segments(None, dz_head=dz)

    @staticmethod
    def _group_done(g, after_decoder, after_group):
        if g == 0 and after_decoder is not None:
            after_decoder()
        if after_group is not None:
            after_group(g)

    # ------------------------------------------------------------------ backward
    def _unit_bwd(self, u, dz, need_dx=True, reduced=False):
        """reduced=True: the producer of dz already accumulated s1 = sum dz and s2 = sum dz * y (maxpool_bwd_add_reduce)."""
        n, h, w = dz.shape[0], dz.shape[1], dz.shape[2]
        bn = u.bn
        if not reduced:
            ops.bn_bwd_reduce(dz, u.y, u.s1, u.s2)
        ka, kb, kc = u.vec[4], u.vec[5], u.vec[6]
        ops.bn_bwd_finalize(u.s1, u.s2, bn.weight.detach(), u.vec[0], u.vec[1], self.gview[bn.weight],
                            self.gview[bn.bias], ka, kb, kc, n * h * w, training=self.training_fwd)
        dpre = ops.bn_relu_bwd_apply(dz, u.y, ka, kb, kc, u.dbias)
        # the weight gradient only feeds the optimiser: run it on the side stream so that it overlaps the dgrad of
        # this layer and the (HBM-bound) BatchNorm backward of the next one
        with self._fork(dpre):
            if u.stem:
                # (direct stem: u.x0 is the fp32 input; its im2col matrix was written during the forward pass)
                ops.gemm_wgrad(dpre, self._stem_cols if self._stem_direct else u.x0, out=u.gp)
            else:
                if self.deterministic:
                    ops.conv3x3_wgrad_split(dpre, u.x0, u.x1, out=u.part)
                else:
                    ops.conv3x3_wgrad(dpre, u.x0, u.x1, out=u.gp)
        if u.stem or not need_dx:
            return None, None
        return ops.conv3x3_dgrad(dpre, u.wd, u.c0, u.c1)

    def _convT_bwd(self, j, dy):
        mod, cm, co = self.convT[j]
        with self._fork(dy):
            ops.convT_wgrad(self.tin[j], dy, out=self.tgp[j])
            ops.channel_sum(dy, self.tdbias[j])
        return ops.convT_dgrad(dy, self.twd[j])

    # ------------------------------------------------------------------ side stream for weight gradients
    class _Fork:
        def __init__(self, eng, keep):
            self.eng, self.keep = eng, keep

        def __enter__(self):
            eng = self.eng
            if not eng.use_side_stream:
                return self
            if eng._side is None:
                eng._side = torch.cuda.Stream()
            if self.keep is not None:
                eng._pending.append(self.keep)  # keep the operand alive until the join (read on another stream)
            eng._side.wait_stream(torch.cuda.current_stream())
            eng._forked = True
            self.ctx = torch.cuda.stream(eng._side)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.eng.use_side_stream:
                self.ctx.__exit__(*exc)
            return False

    def _fork(self, keep):
        return UNetEngine._Fork(self, keep)

    def _join(self):
        if self._forked and self._side is not None:  # only when work was forked since the last join
            torch.cuda.current_stream().wait_stream(self._side)
        self._forked = False
        self._pending = []

    def _begin_backward(self):
        if not self.save_for_backward:
            raise RuntimeError("backward() after an inference forward (save_for_backward=False): the fused "
                               "conv+BatchNorm kernels keep no pre-BN activations")
        self.acc_b.zero_()
        self.Gp.zero_()

    def backward(self, dlogits, after_decoder=None, dz_head=None, after_group=None):
        """dlogits: bf16 [N, H, W, 64] (columns >= num_classes zero). Fills the flat gradient buffer and
        returns the per-parameter gradient views (PyTorch layouts) in `module.parameters()` order.
        `after_decoder()` is called once the head + decoder gradients are final (85 % of the bytes), `after_group(g)`
        after each of the three gradient groups, so a data-parallel caller can start reducing them while the rest
        of the backward pass still runs.
        dz_head: the head's input gradient when `head_loss_backward` already did the head (dlogits is then None)."""
        for g in self.backward_segments(dlogits, dz_head=dz_head):
            self._group_done(g, after_decoder, after_group)
        return [self.gview[p] for p in self.params]

    # parameters() order is enc1..enc4, dec1..dec4, last: the three gradient groups as [first, last) unit indices
    GROUP_NAMES = ("head+decoder", "enc4", "enc3..enc1")

    def group_param_ranges(self):
        """[(start, end)] element ranges of the three gradient groups inside the flat buffer `G`."""
        names = [k for k, _ in self.m.named_parameters()]
        offs = [0]
        for p in self.m.parameters():
            offs.append(offs[-1] + p.numel())
        first_dec = next(i for i, nm in enumerate(names) if nm.startswith("dec"))
        first_enc4 = next(i for i, nm in enumerate(names) if nm.startswith("enc4"))
        return [(offs[first_dec], offs[-1]), (offs[first_enc4], offs[first_dec]), (0, offs[first_enc4])]

    def backward_segments(self, dlogits, dz_head=None):
        U = self.units
        if dz_head is None:
            self._begin_backward()
            # 1x1 head (models/unet.py:72)
            with self._fork(dlogits):
                ops.gemm_wgrad(dlogits, U[17].z, out=self.hgp)
                ops.channel_sum(dlogits, self.hdbias)
            dz = ops.gemm_fprop(dlogits, self.hwd, None, self.m.conv_dim)
        else:
            dz = dz_head
        dz, _ = self._unit_bwd(U[17], dz)
        skip_grads = []
        dskip, dup = self._unit_bwd(U[16], dz)
        skip_grads.append(dskip)  # for enc1
        for j in (3, 2, 1, 0):
            k = 8 + 2 * j
            dz = self._convT_bwd(j, dup)
            dz, _ = self._unit_bwd(U[k + 1], dz)
            dskip, dup = self._unit_bwd(U[k], dz)
            if j > 0:
                skip_grads.append(dskip)  # enc2, enc3, enc4 in that order
            else:
                dpool = dskip  # gradient of the centre pool output
        self._flush_grads(0)  # head + decoder gradients -> PyTorch layout (one batched launch each)
        yield 0
        # encoder, deepest first: enc4 (U[7]) .. enc1 (U[1])
        for lvl, k in ((3, 7), (2, 5), (1, 3), (0, 1)):
            fuse = (U[k].cout // 8) in (8, 16, 32, 64, 128, 256)
            if fuse:  # max-pool backward + skip sum + the BatchNorm-backward reductions of U[k] in one pass
                dz = ops.maxpool_bwd_add_reduce(dpool, U[k].idx, skip_grads[lvl], U[k].y, U[k].s1, U[k].s2)
            else:
                dz = ops.maxpool_bwd_add(dpool, U[k].idx, skip_grads[lvl])
            dz, _ = self._unit_bwd(U[k], dz, reduced=fuse)
            dpool, _ = self._unit_bwd(U[k - 1], dz, need_dx=(k > 1))
            if k == 7:
                self._flush_grads(1)  # enc4: 11 % of the parameters, final long before the backward pass ends
                yield 1
        self._flush_grads(2)  # enc3..enc1
        yield 2

    def release(self):
        """drop the saved activations (after backward, or after an eval forward)."""
        for u in self.units:
            u.x0 = u.x1 = u.y = u.z = u.pooled = u.idx = None
        self._stem_cols = None
        self.tin = []
