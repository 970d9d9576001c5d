"""Developer timing of the validation sweep inner loop (config 5: eval-mode forward + fused argmax/confusion).

    python scripts/dev_eval.py [batch] [iters]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import continual_learning_b200 as clk
from continual_learning_b200 import ops


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    torch.manual_seed(0)
    m = clk.UNet(21).cuda().eval()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(b, 3, 256, 256, generator=g) * 2 - 1).cuda()
    y = torch.randint(0, 21, (b, 256, 256), generator=g).cuda()
    conf = torch.zeros(21 * 21, device="cuda", dtype=torch.int64)
    correct = torch.zeros(1, device="cuda", dtype=torch.int64)
    for mode in ("one head kernel", "BN folded", "separate passes"):
        def fwd():
            if mode == "one head kernel":
                m.evaluate_batch(x, y, nc=21, conf=conf, correct=correct)
                return
            logits = m.engine.forward(x, training=False, save_for_backward=mode != "BN folded")
            ops.argmax_confusion(logits, y, 21, conf=conf, correct=correct)
            m.engine.release()
        with torch.no_grad():
            for _ in range(3):
                fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fwd()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"eval forward + confusion, batch {b}, {mode}: {ms:.3f} ms/batch = {b / ms * 1e3:.0f} img/s")


if __name__ == "__main__":
    main()
