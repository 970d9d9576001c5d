"""Soak run: N optimiser steps of the CUDA-graph TrainStep on a small cycling set of structured batches; prints the
loss every 250 steps and checks that every loss is finite and that the model fits the data.  Developer tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import continual_learning_b200 as clk
from continual_learning_b200.synthetic import structured_batch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.manual_seed(0)
m = clk.UNet(21).cuda().train()
ts = clk.TrainStep(m, clk.FusedAdam(m.parameters(), lr=1e-4, betas=(0.5, 0.99)))
batches = [tuple(t.cuda() for t in structured_batch(100 + i, 16, 256, 256)) for i in range(4)]
losses = []
for i in range(steps):
    loss = ts.step(*batches[i % 4])
    if i % 250 == 0 or i == steps - 1:
        losses.append(float(loss))
        print(f"step {i:5d} loss {losses[-1]:.4f}", flush=True)
assert all(l == l and l < 1e3 for l in losses), "non-finite loss"
assert losses[-1] < 0.5 * losses[0], "the model is not fitting the data"
m.eval()
pred, conf, ok = m.evaluate_batch(*batches[0], nc=21, want_pred=True)
print(f"pixel accuracy on a training batch after {steps} steps: {100.0 * int(ok) / batches[0][1].numel():.2f} %")
