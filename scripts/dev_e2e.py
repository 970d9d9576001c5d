"""Developer end-to-end check on the GPU box: CUDA U-Net step vs the CPU oracle (not a pytest)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import continual_learning_b200 as clk
from oracle import step_ref
from continual_learning_b200.synthetic import structured_batch
from oracle.unet_ref import UNetRef, clone_sd, make_state_dict, param_names


def rel(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    B, H, W = 2, 64, 64
    sd = make_state_dict(0)
    x, y = structured_batch(1, B, H, W)

    # ---------------- oracle (fp32 and matched rounding)
    sd_a = clone_sd(sd)
    loss_ref, logits_ref, grads_ref, cap_ref = step_ref.forward_backward(sd_a, x, y)
    sd_b = clone_sd(sd)
    loss_mr, logits_mr, grads_mr, cap_mr = step_ref.forward_backward(sd_b, x, y, matched_rounding=True)
    print(f"oracle loss fp32 {loss_ref:.6f}  matched-rounding {loss_mr:.6f}")

    # ---------------- CUDA module path
    model = clk.UNet(21).to(dev)
    model.load_state_dict(sd)
    model.train()
    c_loss = clk.CrossEntropyDistillLoss()
    xd, yd = x.to(dev), y.to(dev)
    out = model(xd)
    # per-layer localisation
    eng = model.engine
    names = [k for k in cap_mr]
    for u, nm in zip(eng.units, names):
        z = u.z.float().permute(0, 3, 1, 2)
        print(f"  layer {nm:16s} z rel-L2 vs matched {rel(z, cap_mr[nm]):.3e}  vs fp32 {rel(z, cap_ref[nm]):.3e}")
    loss = c_loss(out, yd)
    loss.backward()
    torch.cuda.synchronize()
    print(f"cuda loss {float(loss):.6f}  |d| vs fp32 {abs(float(loss) - loss_ref) / loss_ref:.2e} vs matched {abs(float(loss) - loss_mr) / loss_mr:.2e}")
    print(f"logits rel-L2 vs matched {rel(out, logits_mr):.3e}  vs fp32 {rel(out, logits_ref):.3e}")
    worst = 0.0
    gall, gall_mr, gall_ref = [], [], []
    for k, p in model.named_parameters():
        e = rel(p.grad, grads_mr[k])
        worst = max(worst, e)
        gall.append(p.grad.flatten().cpu())
        gall_mr.append(grads_mr[k].flatten())
        gall_ref.append(grads_ref[k].flatten())
        if e > 3e-2:
            print(f"   grad {k:28s} rel-L2 vs matched {e:.3e} vs fp32 {rel(p.grad, grads_ref[k]):.3e}")
    ga, gm, gr = torch.cat(gall), torch.cat(gall_mr), torch.cat(gall_ref)
    print(f"global grad rel-L2 vs matched {rel(ga, gm):.3e} (cos {cos(ga, gm):.6f})  vs fp32 {rel(ga, gr):.3e} (cos {cos(ga, gr):.6f})  worst tensor {worst:.3e}")
    print(f"running stats: enc1.2 mean {rel(model.enc1[2].running_mean, sd_a['enc1.2.running_mean']):.2e} "
          f"last.5 var {rel(model.last[5].running_var, sd_a['last.5.running_var']):.2e} "
          f"nbt {int(model.enc1[2].num_batches_tracked)}")

    # ---------------- stock nn.CrossEntropyLoss on our logits (generic backward path)
    model2 = clk.UNet(21).to(dev)
    model2.load_state_dict(sd)
    model2.train()
    out2 = model2(xd)
    l2 = torch.nn.CrossEntropyLoss()(out2, yd)
    l2.backward()
    g2 = torch.cat([p.grad.flatten().cpu() for p in model2.parameters()])
    print(f"stock CE path: loss {float(l2):.6f} grad rel-L2 vs fused-loss path {rel(g2, ga):.3e}")

    # ---------------- eval mode
    model.eval()
    with torch.no_grad():
        oe = model(xd)
    ref_eval = UNetRef(clone_sd(sd_a), training=False)(x)
    # sd_a running stats were updated by one oracle step; model's by one cuda step
    print(f"eval logits rel-L2 vs fp32 oracle(eval) {rel(oe, ref_eval):.3e}")
    model.train()

    # ---------------- TrainStep: eager vs graph vs oracle trajectory
    batches = [structured_batch(10 + i, B, H, W) for i in range(4)]
    sd_t = clone_sd(sd)
    traj_ref = step_ref.train_steps(sd_t, batches, lr=1e-4, betas=(0.5, 0.99))
    for use_graph in (False, True):
        m = clk.UNet(21).to(dev)
        m.load_state_dict(sd)
        m.train()
        opt = clk.FusedAdam(m.parameters(), lr=1e-4, betas=(0.5, 0.99))
        ts = clk.TrainStep(m, opt, use_graph=use_graph)
        traj = [float(ts.step(bx.to(dev), by.to(dev))) for bx, by in batches]
        wrel = rel(torch.cat([p.flatten() for p in m.parameters()]), torch.cat([sd_t[k].flatten() for k in param_names(sd_t)]))
        print(f"TrainStep graph={use_graph}: losses {['%.5f' % v for v in traj]} vs oracle {['%.5f' % v for v in traj_ref]}  weights rel-L2 {wrel:.3e}")

    # ---------------- continual step
    sd_old = make_state_dict(7, num_classes=16)
    old = clk.UNet(16).to(dev)
    old.load_state_dict(sd_old)
    old.eval()
    sd_c = clone_sd(sd)
    loss_c, _, grads_c, _ = step_ref.forward_backward(sd_c, x, y, old=(sd_old, 16), T=2.0, lam=1.0)
    m = clk.UNet(21).to(dev)
    m.load_state_dict(sd)
    m.train()
    cl = clk.CrossEntropyDistillLoss(old, T=2.0, lam=1.0)
    cl.observe(xd)
    lc = cl(m(xd), yd)
    lc.backward()
    gc = torch.cat([p.grad.flatten().cpu() for p in m.parameters()])
    gcr = torch.cat([grads_c[k].flatten() for k in param_names(sd_c)])
    print(f"continual: loss {float(lc):.6f} vs oracle {loss_c:.6f}; grad rel-L2 {rel(gc, gcr):.3e} cos {cos(gc, gcr):.6f}")

    # ---------------- first timing at the benchmark shape
    for (b, h, w) in [(16, 256, 256)]:
        m = clk.UNet(21).to(dev)
        m.train()
        opt = clk.FusedAdam(m.parameters(), lr=1e-4, betas=(0.5, 0.99))
        bx, by = structured_batch(3, b, h, w)
        bx, by = bx.to(dev), by.to(dev)
        for use_graph in (False, True):
            ts = clk.TrainStep(m, opt, use_graph=use_graph)
            for _ in range(3):
                ts.step(bx, by)
            torch.cuda.synchronize()
            t0 = time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ts.step(bx, by)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"timing B={b} {h}x{w} graph={use_graph}: {ms:.2f} ms/step (wall {(time.time() - t0) * 100:.2f}) -> {b / ms * 1000:.1f} img/s, "
                  f"{b * 289.28 / ms:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
