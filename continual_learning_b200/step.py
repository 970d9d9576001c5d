"""TrainStep — the whole training step of trainer.py:168-176 as one call (optionally one CUDA graph):

    inputs/labels H2D -> [frozen old-model forward] -> U-Net forward -> fused CE(+KD) loss fwd+bwd
    -> U-Net backward -> [bucketed NCCL all-reduce of the gradients] -> Adam

This is the public fast path (`bench.py` e2e goes through `TrainStep.step_host`).  The drop-in
nn.Module path (UNet + CrossEntropyDistillLoss + FusedAdam driven by the reference Trainer) launches
exactly the same kernels through autograd.
"""
import os

import torch

from . import ops
from .optim import FusedAdam


class TrainStep:
    def __init__(self, model, optimizer, old_model=None, T=2.0, lam=1.0, use_graph=True, comm=None):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("TrainStep needs continual_learning_b200.FusedAdam")
        self.model, self.opt, self.old = model, optimizer, old_model
        self.T, self.lam = float(T), float(lam)
        # with a communicator the step is captured as FOUR graphs cut where a gradient group becomes final
        # (head + decoder | enc4 | enc3..enc1 | Adam); the NCCL all-reduces are launched eagerly between the replays on
        # NCCL's own stream and overlap the next segment (capturing the collectives themselves into one graph hung
        # on B200 x2 with torch 2.11 / NCCL 2.28.9).  CLK_DDP_GRAPH=0 falls back to eager launches.
        self.use_graph = use_graph and (comm is None or os.environ.get("CLK_DDP_GRAPH", "1") != "0")
        # head GEMM + loss + head backward as one kernel (needs the reference geometry: 64 channels, <= 32 classes)
        self.fused_head = (model.conv_dim == 64 and model.num_classes <= 32
                           and os.environ.get("CLK_FUSED_HEAD", "1") != "0")
        self.comm = comm  # parallel.GradAllReduce or None
        self.graph = None
        self.shape = None
        self.step_count = 0       # optimiser steps taken so far; re-read from the optimiser state when it changes
        self._opt_epoch = None    # FusedAdam.state_epoch seen at the last sync
        self.launches_per_step = 0
        self._graph_ptr = None
        self._opt_ready = False

    # ------------------------------------------------------------------ one eager step on device tensors
    def _segments(self, x, y):
        """the step up to the optimiser as a generator: yields the index of each gradient group (UNetEngine
        .backward_segments) when it is final in the flat buffer, so that the caller can all-reduce it (and, in graph
        mode, cut the CUDA graph there)."""
        eng = self.model.engine
        old_logits = None
        if self.old is not None:
            oeng = self.old.engine
            old_logits = oeng.forward(x, training=False, save_for_backward=False)
            oeng.release()
        if self.fused_head:
            # the logits never reach HBM: head GEMM, loss and the head's backward are one kernel
            eng.forward(x, training=self.model.training, head=False)
            self.loss_acc.zero_()
            yield from eng.head_loss_backward_segments(y, self.loss_acc, old_logits=old_logits, T=self.T, lam=self.lam,
                                                       err_flag=self.err_flag)
        else:
            logits = eng.forward(x, training=self.model.training)
            self.loss_acc.zero_()
            if self.dlogits is None:
                n, _, h, w = x.shape
                self.dlogits = torch.zeros((n, h, w, 64), device=x.device, dtype=torch.bfloat16)
            ops.ce_kd_loss(logits, y, old_logits, T=self.T, lam=self.lam, dlogits=self.dlogits, loss_acc=self.loss_acc,
                           err_flag=self.err_flag)
            yield from eng.backward_segments(self.dlogits)
        eng.release()
        for p in eng.params:
            v = eng.gview[p]
            if p.grad is not v:  # first step, or the module was moved (.cpu()/.cuda() in save_network)
                p.grad = v

    def _adam(self):
        gscale = 1.0 if self.comm is None else 1.0 / self.comm.world_size
        self.opt.step(grad_scale=gscale, hyper_dev=self.hyper)

    def _body(self, x, y):
        eng = self.model.engine
        for g in self._segments(x, y):
            if self.comm is not None:
                self.comm.launch_group(eng.G, g)   # overlaps the rest of the backward pass
        if self.comm is not None:
            self.comm.finish(eng.G)
        self._adam()

    def _alloc(self, x, y):
        dev = x.device
        n, _, h, w = x.shape
        self.shape = (tuple(x.shape), tuple(y.shape))
        self.x_static = torch.empty_like(x)
        self.y_static = torch.empty_like(y)
        self.dlogits = None  # only the unfused path materialises the logits gradient
        self.loss_acc = torch.zeros(2, device=dev, dtype=torch.float64)
        self.err_flag = torch.zeros(1, device=dev, dtype=torch.int32)
        self.hyper = torch.zeros(4, device=dev, dtype=torch.float32)
        self.npix = n * h * w
        self.graph = None
        self._opt_ready = False

    def _sync_step_count(self):
        """Adam's bias correction uses the optimiser's own step count: pick it up from the state on first use and
        after `load_state_dict` (resume, trainer.py:98), so that a warm exp_avg / exp_avg_sq is not corrected as t=1."""
        if self._opt_epoch == self.opt.state_epoch:
            return
        self._opt_epoch = self.opt.state_epoch
        plist = [p for g in self.opt.param_groups for p in g["params"]]
        self.step_count = int(self.opt.group_step(plist))

    def _set_hyper(self):
        gs = 1.0 if self.comm is None else 1.0 / self.comm.world_size
        vals = self.opt.hyper_values(self.step_count + 1, grad_scale=gs)
        # pageable source: the driver stages it before returning, so the host may run ahead safely
        self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))

    def step(self, x, y):
        """x fp32 [B,3,H,W], y int64 [B,H,W], both on the device. Returns the loss as a 0-dim fp64 device tensor."""
        if self.shape != (tuple(x.shape), tuple(y.shape)):
            self._alloc(x, y)
        self._sync_step_count()
        self._set_hyper()
        if not self.use_graph:
            self._body(x, y)
        else:
            self.x_static.copy_(x, non_blocking=True)
            self.y_static.copy_(y, non_blocking=True)
            ptr = self.model.engine.params[0].data_ptr() if self.model.engine._dev is not None else None
            if self.graph is not None and ptr != self._graph_ptr:
                self.graph = None  # parameters were re-allocated (module moved): capture again
            if self.graph is None:
                # warm-up eagerly on a side stream (allocator + lazy initialisation), then capture
                state = self._snapshot()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body(self.x_static, self.y_static)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                self._restore(state)
                self.model.engine._wver = None
                if self.old is not None:
                    self.old.engine._wver = None
                from . import _lib
                n0 = _lib.launch_count
                if self.comm is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._body(self.x_static, self.y_static)
                    self.graph = [g]
                else:
                    # one graph per gradient group + one for Adam, sharing a memory pool (replayed in capture order)
                    graphs, pool = [], None
                    gen = self._segments(self.x_static, self.y_static)
                    for _ in range(3):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, pool=pool):
                            grp = next(gen)
                        assert grp == len(graphs)
                        pool = g.pool()
                        graphs.append(g)
                    for _ in gen:          # host-only tail of the step body (no launches)
                        raise RuntimeError("unexpected extra gradient group")
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        self._adam()
                    graphs.append(g)
                    self.graph = graphs
                self.launches_per_step = _lib.launch_count - n0  # clk_* kernels replayed per step
                self._graph_ptr = self.model.engine.params[0].data_ptr()
                self._restore(state)
            if self.comm is None:
                self.graph[0].replay()
            else:
                G = self.model.engine.G
                for grp in range(3):
                    self.graph[grp].replay()
                    self.comm.launch_group(G, grp)   # NCCL stream: overlaps the next segment's replay
                self.comm.finish(G)
                self.graph[3].replay()
        self.step_count += 1
        self.opt.set_group_step(self.model.engine.params, self.step_count)
        loss = self.loss_acc[0] / self.npix
        if self.old is not None:
            loss = loss + (self.lam * self.T * self.T / self.npix) * self.loss_acc[1]
        return loss

    def step_host(self, x_pinned, y_pinned, prefetch=None, defer_loss=False):
        """end-to-end step from pinned HOST tensors: H2D of the inputs, the step, D2H of the loss (float).

        `prefetch=(x_next, y_next)` starts the H2D copy of the NEXT batch on a copy stream before this step's
        kernels are launched, so the PCIe transfer overlaps the compute (every batch is still copied exactly
        once, inside the caller's loop); the next call recognises its staged inputs and skips the copy.

        `defer_loss=True` keeps the host one step ahead of the device: the loss of this step is copied to pinned host
        memory asynchronously and RETURNED BY THE NEXT CALL (the first call returns None; `flush_loss()` returns the
        last one).  Every step's loss is still read back, but the device never idles waiting for the host to
        launch the next step."""
        dev = next(self.model.parameters()).device
        staged = getattr(self, "_staged", None)
        if staged is not None and staged[0] is x_pinned and staged[1] is y_pinned:
            torch.cuda.current_stream().wait_event(staged[4])
            x, y = staged[2], staged[3]
        else:
            x = x_pinned.to(dev, non_blocking=True)
            y = y_pinned.to(dev, non_blocking=True)
        self._staged = None
        if prefetch is not None:
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream()
            with torch.cuda.stream(self._copy_stream):
                xn = prefetch[0].to(dev, non_blocking=True)
                yn = prefetch[1].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self._staged = (prefetch[0], prefetch[1], xn, yn, ev)
        loss = self.step(x, y)
        x.record_stream(torch.cuda.current_stream())
        y.record_stream(torch.cuda.current_stream())
        # {loss, out-of-range-label flag}: ONE device->host read per step
        res = torch.stack((loss, self.err_flag[0].double()))
        if not defer_loss:
            return self._checked(res.cpu())
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = [torch.zeros(2, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._loss_ev = [None, None]
            self._loss_k = 0
        prev = self.flush_loss()
        k = self._loss_k & 1
        self._loss_host[k].copy_(res, non_blocking=True)
        self._loss_ev[k] = torch.cuda.Event()
        self._loss_ev[k].record()
        self._loss_pending = k
        self._loss_k += 1
        return prev

    def flush_loss(self):
        """the loss of the most recent `step_host(defer_loss=True)` call that has not been returned yet (or None)."""
        k = getattr(self, "_loss_pending", None)
        if k is None:
            return None
        self._loss_ev[k].synchronize()
        self._loss_pending = None
        return self._checked(self._loss_host[k])

    def _checked(self, res):
        """res = host {loss, label-error flag}: raise like nn.CrossEntropyLoss (trainer.py:174) on a label outside
        [0, num_classes) — such pixels contribute nothing to the loss or the gradient of that step."""
        if float(res[1]) != 0.0:
            self.err_flag.zero_()
            raise IndexError("Target out of bounds: a label outside [0, num_classes) reached the loss")
        return float(res[0])

    # the capture warm-up and the capture itself run the step body for real: undo their effect on the
    # parameters, optimiser state and BatchNorm buffers so that step k of a graph run equals step k eagerly
    def _snapshot(self):
        ts = [p.detach() for p in self.model.parameters()] + [b for b in self.model.buffers()]
        for p in self.model.parameters():
            st = self.opt.state.get(p, {})
            ts += [st[k] for k in ("exp_avg", "exp_avg_sq") if k in st]
        return [(t, t.clone()) for t in ts], {p: dict(self.opt.state.get(p, {})) for p in self.model.parameters()}

    def _restore(self, state):
        saved, opt_state = state
        with torch.no_grad():
            for t, c in saved:
                t.copy_(c)
            for p in self.model.parameters():
                st = self.opt.state.get(p)
                if st is None:
                    continue
                if not opt_state[p]:
                    # state was created during the warm-up: reset it to Adam's initial state
                    st["exp_avg"].zero_()
                    st["exp_avg_sq"].zero_()
        self.opt.set_group_step([p for p in self.model.parameters() if p in self.opt.state], self.step_count)
