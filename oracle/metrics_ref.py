"""numpy / torch-CPU restatement of the used half of the reference metrics module (metrics.py:6-71).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch


def conf_matrix_int(target, prediction, num_classes):
    """metrics._fast_conf_matrix (metrics.py:32-38) without the final .float(): int64 [nc, nc],
    rows = target, cols = prediction; targets outside [0, nc) are dropped; a prediction outside
    [0, nc) on a kept target makes bincount longer than nc^2 and the reference's reshape raise."""
    t = np.asarray(target).reshape(-1).astype(np.int64)
    p = np.asarray(prediction).reshape(-1).astype(np.int64)
    mask = (t >= 0) & (t < num_classes)
    idx = num_classes * t[mask] + p[mask]
    if idx.size and (idx.min() < 0 or idx.max() >= num_classes ** 2 or
                     ((p[mask] < 0) | (p[mask] >= num_classes)).any()):
        # the reference raises for an index past nc^2; an in-range index produced by an out-of-range
        # prediction would silently land in a wrong cell there — both are treated as errors here
        raise RuntimeError("prediction outside [0, num_classes) for a kept target")
    return np.bincount(idx, minlength=num_classes ** 2).reshape(num_classes, num_classes)


def derived_metrics(matrix_f32):
    """overall_pixel_acc, per_class_pixel_acc, mean_IU_2, max_per_class_pixel_acc on a float32 matrix with the
    reference's float32 formulas and NaN filtering (metrics.py:11-29, 41-53)."""
    m = torch.as_tensor(matrix_f32, dtype=torch.float32)
    diag = torch.diag(m)
    overall = diag.sum() * 100 / m.sum()
    per_class = 100 * diag / m.sum(dim=1)
    avg_per_class = torch.mean(per_class[per_class == per_class])
    max_per_class = torch.max(per_class[per_class == per_class])
    jacc = diag / (m.sum(dim=1) + m.sum(dim=0) - diag)
    mean_iu = torch.mean(jacc[jacc == jacc])
    return overall, avg_per_class, mean_iu, max_per_class


def eval_metrics(target, prediction, num_classes):
    """metrics.eval_metrics (metrics.py:55-63): per-sample matrices summed in float32."""
    matrix = torch.zeros((num_classes, num_classes))
    for t, p in zip(target, prediction):
        matrix += torch.from_numpy(conf_matrix_int(t.numpy(), p.numpy(), num_classes)).float()
    return derived_metrics(matrix)


def pixel_acc(mask, predicted, total_train, correct_train):
    """metrics.pixel_acc (metrics.py:6-9)."""
    total_train += mask.nelement()
    return 100 * correct_train / total_train, total_train, correct_train


def mean_iu_binary(target, prediction):
    """metrics.mean_IU_ (metrics.py:67-71): IoU of the non-zero masks."""
    t = np.asarray(target) != 0
    p = np.asarray(prediction) != 0
    return np.sum(t & p) / np.sum(t | p)


# ---- legacy per-image metrics (metrics.py:74-183), restated with the reference's per-class boolean masks
def _masks(segm, cl):
    return [(segm == c) for c in cl]


def legacy_metrics(eval_segm, gt_segm):
    """(pixel_accuracy, mean_accuracy, mean_IU, frequency_weighted_IU) of one image pair, numpy [H, W]."""
    eval_segm, gt_segm = np.asarray(eval_segm), np.asarray(gt_segm)
    if eval_segm.shape != gt_segm.shape:
        raise ValueError("DiffDim: Different dimensions of matrices!")
    gt_cl = np.unique(gt_segm)                                   # extract_classes(gt)
    union = np.union1d(np.unique(eval_segm), gt_cl)              # union_classes
    # pixel_accuracy, metrics.py:74-98
    sum_n, sum_t = 0, 0
    for c in gt_cl:
        em, gm = (eval_segm == c), (gt_segm == c)
        sum_n += np.sum(np.logical_and(em, gm))
        sum_t += np.sum(gm.astype(np.float64))
    pa = 0 if sum_t == 0 else sum_n / sum_t
    # mean_accuracy, metrics.py:100-124
    acc = [0] * len(gt_cl)
    for i, c in enumerate(gt_cl):
        em, gm = (eval_segm == c), (gt_segm == c)
        t_i = np.sum(gm.astype(np.float64))
        if t_i != 0:
            acc[i] = np.sum(np.logical_and(em, gm)) / t_i
    ma = np.mean(acc)
    # mean_IU / frequency_weighted_IU, metrics.py:126-183
    iu, fw = [0] * len(union), [0] * len(union)
    for i, c in enumerate(union):
        em, gm = (eval_segm == c).astype(np.float64), (gt_segm == c).astype(np.float64)
        if np.sum(em) == 0 or np.sum(gm) == 0:
            continue
        n_ii = np.sum(np.logical_and(em, gm))
        t_i, n_ij = np.sum(gm), np.sum(em)
        iu[i] = n_ii / (t_i + n_ij - n_ii)
        fw[i] = (t_i * n_ii) / (t_i + n_ij - n_ii)
    miu = np.sum(iu) / len(gt_cl)
    fwiu = np.sum(fw) / (eval_segm.shape[0] * eval_segm.shape[1])
    return pa, ma, miu, fwiu
