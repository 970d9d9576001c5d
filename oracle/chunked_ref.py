"""n-replica oracle: what nn.DataParallel (trainer.py:120-122) / one-process-per-GPU data parallelism
computes.  TEST INFRASTRUCTURE ONLY.

The batch is split into n equal chunks; every chunk runs through the same weights in train mode with
its OWN BatchNorm statistics (DataParallel replicas do not sync BN); the global-mean CE makes the
gradient the average of the per-chunk gradients; BN running stats follow chunk 0 (replica 0 shares
its buffers with the master module).
"""
from .step_ref import forward_backward
from .unet_ref import clone_sd


def chunked_forward_backward(sd, x, labels, n, **kw):
    xs, ys = x.chunk(n), labels.chunk(n)
    total, losses = None, []
    for r in range(n):
        sd_r = sd if r == 0 else clone_sd(sd)
        loss, _, grads, _ = forward_backward(sd_r, xs[r], ys[r], **kw)
        losses.append(loss)
        if total is None:
            total = {k: g.clone() / n for k, g in grads.items()}
        else:
            for k, g in grads.items():
                total[k] += g / n
    return sum(losses) / n, total
