"""are two runs of the step bit-identical? (engine.deterministic on / off), per parameter group"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import continual_learning_b200 as clk
from continual_learning_b200.synthetic import structured_batch
from oracle.unet_ref import make_state_dict

sd = make_state_dict(2)
x, y = structured_batch(3, 4, 128, 128)
x, y = x.cuda(), y.cuda()
for det in (True, False):
    runs = []
    for _ in range(3):
        m = clk.UNet(21).cuda(); m.load_state_dict(sd); m.train()
        m.engine.deterministic = det
        out = m(x)
        clk.CrossEntropyDistillLoss()(out, y).backward()
        runs.append((out.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()},
                     {k: v.clone() for k, v in m.state_dict().items() if "running" in k}))
    for r in (1, 2):
        eq_logits = torch.equal(runs[0][0], runs[r][0])
        bad = [k for k in runs[0][1] if not torch.equal(runs[0][1][k], runs[r][1][k])]
        badb = [k for k in runs[0][2] if not torch.equal(runs[0][2][k], runs[r][2][k])]
        print(f"deterministic={det} run0 vs run{r}: logits equal {eq_logits}; {len(bad)} of {len(runs[0][1])} grads differ; "
              f"{len(badb)} BN buffers differ; first: {bad[:6]}")
