"""Golden values of the legacy per-image metrics (metrics.py:74-183) from the UNMODIFIED reference functions.
Run in the build container:  python tests/golden/make_golden_legacy_metrics.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np

import metrics as ref  # /root/reference/metrics.py


def main():
    rng = np.random.Generator(np.random.PCG64(21))
    cases = []
    for k, (h, w, ncl_gt, ncl_ev) in enumerate([(24, 20, 5, 5), (32, 32, 21, 22), (16, 40, 1, 3), (20, 20, 3, 1),
                                                  (33, 17, 8, 12), (12, 12, 2, 2)]):
        gt = rng.integers(0, ncl_gt, size=(h, w))
        gt = np.kron(gt[::4, ::4], np.ones((4, 4), dtype=np.int64))[:h, :w] if h >= 8 else gt
        ev = np.where(rng.random((h, w)) < 0.7, gt, rng.integers(0, ncl_ev, size=(h, w)))
        if k == 5:
            ev = gt.copy()  # perfect prediction
        cases.append((ev.astype(np.int64), gt.astype(np.int64)))
    rec = {"n": np.int64(len(cases))}
    for k, (ev, gt) in enumerate(cases):
        rec[f"eval{k}"], rec[f"gt{k}"] = ev, gt
        rec[f"out{k}"] = np.array([ref.pixel_accuracy(ev, gt), ref.mean_accuracy(ev, gt), ref.mean_IU(ev, gt),
                                   ref.frequency_weighted_IU(ev, gt)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "legacy_metrics.npz"), **rec)
    print("wrote legacy_metrics.npz")


if __name__ == "__main__":
    main()
