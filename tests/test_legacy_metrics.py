"""Legacy per-image metrics (metrics.py:74-183): pixel_accuracy, mean_accuracy, mean_IU,
frequency_weighted_IU.  Golden values come from the unmodified reference functions
(tests/golden/make_golden_legacy_metrics.py); float64, compared bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "legacy_metrics.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_equals_reference_goldens(gold):
    for k in range(int(gold["n"])):
        out = np.array(metrics_ref.legacy_metrics(gold[f"eval{k}"], gold[f"gt{k}"]), dtype=np.float64)
        assert np.array_equal(out, gold[f"out{k}"]), k


def test_matrix_formulas_equal_the_mask_formulas(gold):
    """the host arithmetic of the package (from one confusion matrix per image) against the mask-based oracle and the
    goldens, with the matrix counted by numpy (no GPU needed)"""
    from continual_learning_b200.metrics import _legacy_from_matrix
    nc = 22
    for k in range(int(gold["n"])):
        ev, gt = gold[f"eval{k}"], gold[f"gt{k}"]
        m = np.bincount(nc * gt.reshape(-1) + ev.reshape(-1), minlength=nc * nc).reshape(nc, nc).astype(np.int64)
        out = np.array(_legacy_from_matrix(m, ev.shape[0] * ev.shape[1]), dtype=np.float64)
        assert np.array_equal(out, gold[f"out{k}"]), k
    rng = np.random.Generator(np.random.PCG64(3))
    for _ in range(20):
        h, w = int(rng.integers(2, 40)), int(rng.integers(2, 40))
        gt = rng.integers(0, int(rng.integers(1, 22)), size=(h, w))
        ev = rng.integers(0, int(rng.integers(1, 22)), size=(h, w))
        m = np.bincount(nc * gt.reshape(-1) + ev.reshape(-1), minlength=nc * nc).reshape(nc, nc).astype(np.int64)
        assert np.array_equal(np.array(_legacy_from_matrix(m, h * w), dtype=np.float64),
                              np.array(metrics_ref.legacy_metrics(ev, gt), dtype=np.float64))


@pytest.mark.gpu
def test_device_counted_legacy_metrics_match_reference_goldens(gold, lib_built):
    from continual_learning_b200 import metrics as mt
    for k in range(int(gold["n"])):
        ev, gt = gold[f"eval{k}"], gold[f"gt{k}"]
        got = np.array([mt.pixel_accuracy(ev, gt), mt.mean_accuracy(ev, gt), mt.mean_IU(ev, gt),
                        mt.frequency_weighted_IU(ev, gt)], dtype=np.float64)
        assert np.array_equal(got, gold[f"out{k}"]), k


@pytest.mark.gpu
def test_batched_per_image_matrices_and_errors(lib_built):
    from continual_learning_b200 import metrics as mt
    rng = np.random.Generator(np.random.PCG64(9))
    gt = rng.integers(0, 21, size=(5, 64, 48))
    ev = rng.integers(0, 22, size=(5, 64, 48))
    conf = mt.per_image_conf_matrices(torch.from_numpy(ev).cuda(), torch.from_numpy(gt).cuda(), 22).cpu().numpy()
    for b in range(5):
        want = np.bincount(22 * gt[b].reshape(-1) + ev[b].reshape(-1), minlength=484).reshape(22, 22)
        assert np.array_equal(conf[b], want)
    out = mt.legacy_metrics_batched(ev, gt)
    for b in range(5):
        assert np.array_equal(out[b], np.array(metrics_ref.legacy_metrics(ev[b], gt[b]), dtype=np.float64))
    with pytest.raises(mt.EvalSegErr):
        mt.pixel_accuracy(np.zeros((4, 4), dtype=np.int64), np.zeros((4, 5), dtype=np.int64))
    bad = gt.copy()
    bad[0, 0, 0] = 99
    with pytest.raises(ValueError):
        mt.per_image_conf_matrices(ev, bad, 22)
