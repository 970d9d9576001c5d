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
