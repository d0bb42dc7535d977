"""Eval metrics on device: drop-ins for `utils/eval_utils.py:MIoU` and `ForegroundDetect`.

The reference computes, per validation image: `torch.max(logits, 1)` twice, three `torch.histc`, `.cpu().numpy()` copies
and a `numpy.bincount` (utils/eval_utils.py:63-97,107-117,151-155).  Here ONE kernel (csrc/metrics.cu) produces the
argmax and the confusion counts conf[label][pred]; every number the two classes report is derived from conf, and the
host only reads back the (nc+1) x nc int64 matrix.  Same constructor arguments, call signatures and return values.

`CAVP.forward_eval_metrics` (models/cavp_model.py) feeds the same classes through `update_from_confusion` without
materialising the full-resolution logits at all.
"""
import numpy
import torch

from . import _C


def argmax_confusion(logits, target, ignore_index, conf=None, want_pred=False):
    """logits [B, C, H, W] fp32 CUDA (any strides -> made contiguous), target [B, (1,) H, W] integer.
    -> (pred int64 [B, H, W] or None, conf int64 [C+1, C])."""
    if not logits.is_cuda:
        raise RuntimeError("cavp_b200.metrics runs on CUDA (sm_100a) only; there is no CPU fallback")
    B, C, H, W = logits.shape
    lg = logits.detach().float().contiguous()
    labels = target.detach().reshape(B, H, W).to(logits.device, torch.int64).contiguous()
    if conf is None:
        conf = torch.zeros(C + 1, C, dtype=torch.int64, device=logits.device)
    pred = torch.empty(B, H, W, dtype=torch.int64, device=logits.device) if want_pred else None
    _C.call("cavp_argmax_confusion", lg.data_ptr(), labels.data_ptr(), B, C, H * W, int(ignore_index),
            0 if pred is None else pred.data_ptr(), conf.data_ptr(), torch.cuda.current_stream(logits.device).cuda_stream)
    return pred, conf


class MIoU(object):
    """utils/eval_utils.py:32-97."""

    def __init__(self, num_classes, ignore_index, local_rank=0):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.local_rank = local_rank
        self.inter, self.union = 0, 0
        self.correct, self.label = 0, 0
        self.iou = numpy.array([0 for _ in range(num_classes)])
        self.acc = 0.0

    def get_metric_results(self, class_list=None):
        if class_list is None:
            return numpy.round(self.iou.mean().item(), 4), numpy.round(self.acc, 4)
        return numpy.round(self.iou[class_list].mean().item(), 4), numpy.round(self.acc, 4)

    @staticmethod
    def sample_from_confusion(conf, num_classes):
        """conf: [(C+1), C] counts (numpy / CPU tensor) -> (correct, labeled, inter[nc], union[nc]) exactly as
        batch_pix_accuracy / batch_intersection_union: histc(bins=nc, min=1, max=nc) keeps classes 0..nc-1."""
        c = numpy.asarray(conf, dtype=numpy.int64)
        C = c.shape[1]
        k = min(num_classes, C)
        labeled = c.sum()
        correct = numpy.trace(c[:C])
        inter = numpy.zeros(num_classes, dtype=numpy.float32)
        pred = numpy.zeros(num_classes, dtype=numpy.float32)
        lab = numpy.zeros(num_classes, dtype=numpy.float32)
        inter[:k] = numpy.diag(c[:C])[:k]
        pred[:k] = c.sum(0)[:k]
        lab[:k] = c[:C].sum(1)[:k]
        if num_classes > C:  # labels in [C, num_classes) were folded into row C by the kernel: not separable
            if c[C].sum() != 0:
                raise ValueError("labels >= number of logit channels present; metric num_classes must match")
        union = pred + lab - inter
        return [numpy.round(numpy.asarray(correct), 5), numpy.round(numpy.asarray(labeled), 5), numpy.round(inter, 5),
                numpy.round(union, 5)]

    def calculate_current_sample(self, output, target):
        # output => BxCxHxW (logits), target => Bx1xHxW.  (The reference rewrites target's ignore pixels to -1 in place;
        # the kernel skips both ignore_index and negative labels, so the caller's tensor is left untouched.)
        _, conf = argmax_confusion(output, target, self.ignore_index)
        return self.sample_from_confusion(conf.cpu().numpy(), self.num_classes)

    def _accumulate(self, curr):
        curr_correct, curr_label, curr_inter, curr_union = curr
        self.correct = self.correct + curr_correct
        self.label = self.label + curr_label
        self.inter = self.inter + curr_inter
        self.union = self.union + curr_union
        self.acc = 1.0 * self.correct / (numpy.spacing(1) + self.label)
        self.iou = 1.0 * self.inter / (numpy.spacing(1) + self.union)
        return self.get_metric_results()

    def __call__(self, x, y):
        return self._accumulate(self.calculate_current_sample(x, y))

    def update_from_confusion(self, conf):
        """conf of ONE validation call (e.g. from CAVP.forward_eval_metrics with a fresh matrix)."""
        c = conf.cpu().numpy() if torch.is_tensor(conf) else conf
        return self._accumulate(self.sample_from_confusion(c, self.num_classes))


class ForegroundDetect(object):
    """utils/eval_utils.py:100-155 (the confusion matrix is accumulated from the kernel's counts; the derived scores
    are computed with the reference's own formulas)."""

    def __init__(self, num_classes, ignore_class=255, local_rank=0):
        self.num_classes = num_classes
        self.ignore = ignore_class
        self.local_rank = local_rank
        self.confusion_matrix_ = numpy.zeros((num_classes, num_classes))

    def _hist_from_confusion(self, conf):
        c = numpy.asarray(conf, dtype=numpy.int64)
        C = c.shape[1]
        n = self.num_classes
        hist = numpy.zeros((n, n), dtype=numpy.int64)
        k = min(n, C)
        hist[:k, :k] = c[:k, :k]
        if n > C and c[C].sum() != 0:
            raise ValueError("labels >= number of logit channels present; cannot place them in the confusion matrix")
        return hist

    def f_beta_score(self, tp, fp, fn, beta2=1.0):
        score = ((1 + beta2) * tp) / ((1 + beta2) * tp + beta2 * fn + fp)
        return torch.nanmean(score)

    def get_metric_results(self, class_list=None):
        cm = torch.tensor(numpy.asarray(self.confusion_matrix_, dtype=numpy.float64))
        tp = torch.diag(cm)
        fp = cm.sum(dim=0) - tp
        fn = cm.sum(dim=1) - tp
        if class_list is not None:
            tp, fp, fn = tp[class_list], fp[class_list], fn[class_list]
        fdr = torch.nanmean(fp / (fp + tp))
        f1 = self.f_beta_score(tp, fp, fn, beta2=1.0)
        f_03 = self.f_beta_score(tp, fp, fn, beta2=0.3)
        return (torch.round(fdr, decimals=4).numpy(), torch.round(f1, decimals=4).numpy(),
                torch.round(f_03, decimals=4).numpy())

    def __call__(self, y_hat, y):
        _, conf = argmax_confusion(y_hat, y, -1 if self.ignore is None else self.ignore)
        self.confusion_matrix_ = self.confusion_matrix_ + self._hist_from_confusion(conf.cpu().numpy())

    def update_from_confusion(self, conf):
        c = conf.cpu().numpy() if torch.is_tensor(conf) else conf
        self.confusion_matrix_ = self.confusion_matrix_ + self._hist_from_confusion(c)
