"""Drop-in for the used half of the reference `metrics` module (metrics.py:6-71): same function names,
arguments, return types and error behaviour; the confusion matrix is counted by the shared-memory
privatised histogram kernel in int64 and the derived numbers use the reference's float32 formulas.
"""
import numpy as np
import torch

from . import _lib, ops


def _need_cuda(t):
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("continual_learning_b200.metrics counts on the GPU (sm_100a); no CPU fallback")
    return t.cuda()


def pixel_acc(mask, predicted, total_train, correct_train):
    """running pixel accuracy (metrics.py:6-9)."""
    total_train += mask.nelement()
    train_accuracy = 100 * correct_train / total_train
    return train_accuracy, total_train, correct_train


def conf_matrix_int64(target, prediction, num_classes, out=None):
    """int64 [nc, nc] on the device: bincount(nc*t + p) over 0 <= t < nc (metrics.py:32-38 without .float())."""
    t = _need_cuda(target).reshape(-1).contiguous().long()
    p = _need_cuda(prediction).reshape(-1).contiguous().long()
    _lib.ensure_device(t.device.index)
    err = torch.zeros(1, device=t.device, dtype=torch.int32)
    conf = ops.confusion_matrix(t, p, num_classes, conf=out, err_flag=err)
    if int(err.item()) != 0:
        # the reference's bincount/reshape raises RuntimeError for a prediction outside [0, nc)
        raise RuntimeError("prediction outside [0, num_classes) on a kept target: confusion matrix is undefined "
                           "(the reference's reshape fails here too)")
    return conf.view(num_classes, num_classes)


def _fast_conf_matrix(target, prediction, num_classes):
    return conf_matrix_int64(target, prediction, num_classes).float().cpu()


def nanmean(x):
    return torch.mean(x[x == x])


def nanmax(x):
    return torch.max(x[x == x])


def overall_pixel_acc(matrix):
    return torch.diag(matrix).sum() * 100 / matrix.sum()


def per_class_pixel_acc(conf_matrix):
    return nanmean(100 * torch.diag(conf_matrix) / conf_matrix.sum(dim=1))


def max_per_class_pixel_acc(conf_matrix):
    return nanmax(100 * torch.diag(conf_matrix) / conf_matrix.sum(dim=1))


def mean_IU_2(matrix):
    inter = torch.diag(matrix)
    return nanmean(inter / (matrix.sum(dim=1) + matrix.sum(dim=0) - inter))


def metrics_from_matrix(conf_int64):
    """(overall_acc, avg_per_class_acc, mean_IU, max_per_class_acc) as 0-dim float32 CPU tensors."""
    m = conf_int64.detach().cpu().float()
    return overall_pixel_acc(m), per_class_pixel_acc(m), mean_IU_2(m), max_per_class_pixel_acc(m)


def eval_metrics(target, prediction, num_classes):
    """metrics.eval_metrics (metrics.py:55-63). The reference sums per-sample float32 matrices; counting the
    whole batch in int64 gives the same matrix whenever every cell is < 2^24 and is exact beyond."""
    return metrics_from_matrix(conf_matrix_int64(target, prediction, num_classes))


def mean_IU_(target, prediction):
    """binary (non-zero) IoU (metrics.py:67-71) from the same histogram: classes {0, non-zero}."""
    t = torch.as_tensor(np.asarray(target)) if not torch.is_tensor(target) else target
    p = torch.as_tensor(np.asarray(prediction)) if not torch.is_tensor(prediction) else prediction
    m = conf_matrix_int64((t != 0).long(), (p != 0).long(), 2).cpu()
    inter, union = int(m[1, 1]), int(m[0, 1] + m[1, 0] + m[1, 1])
    return np.float64(inter) / np.float64(union) if union else np.float64("nan")


def predict_and_count(logits_nchw_view, labels, num_classes_hist=None):
    """argmax over the channel dim + correct-pixel count (+ confusion matrix) in one pass over the logits
    (trainer.py:183-184,279-280). `logits_nchw_view` is what UNet.forward returned."""
    z = logits_nchw_view.permute(0, 2, 3, 1)
    if not z.is_contiguous():
        z = z.contiguous()
    z = z.float()
    pred, conf, correct = ops.argmax_confusion(z, labels.contiguous(), nc=num_classes_hist, want_pred=True)
    return pred, correct, (None if conf is None else conf.view(num_classes_hist, num_classes_hist))


# ------------------------------------------------------------------------------------------------
# Legacy per-image metrics (metrics.py:74-183; commented out at trainer.py:190).  The reference builds one boolean
# mask per class and image in numpy; here ONE batched histogram launch counts a confusion matrix per image and the
# four numbers follow from its row sums t_i, column sums e_i and diagonal n_ii with the reference's float64 formulas.
class EvalSegErr(Exception):
    def __init__(self, value):
        self.value = value

    def __str__(self):
        return repr(self.value)


def per_image_conf_matrices(eval_segm, gt_segm, num_classes=22):
    """int64 [B, nc, nc] (rows = ground truth) for label maps [B, H, W] (or [H, W]); numpy or torch."""
    e = torch.as_tensor(eval_segm)
    g = torch.as_tensor(gt_segm)
    if e.dim() == 2:
        e, g = e[None], g[None]
    if e.dim() != 3 or tuple(e.shape) != tuple(g.shape):
        raise EvalSegErr("DiffDim: Different dimensions of matrices!")  # metrics.py:check_size
    e = _need_cuda(e).long().contiguous()
    g = _need_cuda(g).long().contiguous()
    _lib.ensure_device(e.device.index)
    b, n = e.shape[0], e.shape[1] * e.shape[2]
    conf = torch.zeros((b, num_classes, num_classes), device=e.device, dtype=torch.int64)
    err = torch.zeros(1, device=e.device, dtype=torch.int32)
    _lib.call("clk_confusion_matrix_batched", g, e, b, n, num_classes, conf, err)
    if int(err.item()):
        raise ValueError(f"label outside [0, {num_classes}) in the evaluated or the ground-truth map")
    return conf


def _legacy_from_matrix(m, area):
    """(pixel_accuracy, mean_accuracy, mean_IU, frequency_weighted_IU) of one image from its int64 matrix."""
    t = m.sum(axis=1).astype(np.float64)   # t_i  (float64 like the reference's mask sums)
    e = m.sum(axis=0).astype(np.float64)   # sum_j n_ji
    d = np.diagonal(m)                     # n_ii (int64 like np.sum(logical_and))
    gt_cl = [i for i in range(len(t)) if t[i] > 0]
    union = [i for i in range(len(t)) if t[i] > 0 or e[i] > 0]
    sum_n, sum_t = 0, 0
    for i in gt_cl:
        sum_n += d[i]
        sum_t += t[i]
    pa = 0 if sum_t == 0 else sum_n / sum_t
    acc = [d[i] / t[i] for i in gt_cl]
    ma = np.mean(acc)
    iu = [(d[i] / (t[i] + e[i] - d[i])) if (e[i] != 0 and t[i] != 0) else 0 for i in union]
    miu = np.sum(iu) / len(gt_cl)
    fw = [((t[i] * d[i]) / (t[i] + e[i] - d[i])) if (e[i] != 0 and t[i] != 0) else 0 for i in union]
    fwiu = np.sum(fw) / area
    return pa, ma, miu, fwiu


def legacy_metrics_batched(eval_segm, gt_segm, num_classes=22):
    """the four legacy metrics for every image of a batch: float64 array [B, 4]
    (pixel_accuracy, mean_accuracy, mean_IU, frequency_weighted_IU)."""
    conf = per_image_conf_matrices(eval_segm, gt_segm, num_classes).cpu().numpy()
    e = torch.as_tensor(eval_segm)
    area = e.shape[-2] * e.shape[-1]
    return np.array([_legacy_from_matrix(m, area) for m in conf], dtype=np.float64)


def pixel_accuracy(eval_segm, gt_segm):
    """sum_i(n_ii) / sum_i(t_i)   (metrics.py:74-98)"""
    return legacy_metrics_batched(eval_segm, gt_segm)[0, 0]


def mean_accuracy(eval_segm, gt_segm):
    """(1/n_cl) sum_i(n_ii/t_i)   (metrics.py:100-124)"""
    return legacy_metrics_batched(eval_segm, gt_segm)[0, 1]


def mean_IU(eval_segm, gt_segm):
    """(1/n_cl) * sum_i(n_ii / (t_i + sum_j(n_ji) - n_ii))   (metrics.py:126-153)"""
    return legacy_metrics_batched(eval_segm, gt_segm)[0, 2]


def frequency_weighted_IU(eval_segm, gt_segm):
    """sum_k(t_k)^(-1) * sum_i((t_i*n_ii)/(t_i + sum_j(n_ji) - n_ii))   (metrics.py:155-183)"""
    return legacy_metrics_batched(eval_segm, gt_segm)[0, 3]
