"""CPU restatement of the reference's eval metrics (TEST INFRASTRUCTURE ONLY - never imported by the product).

Follows utils/eval_utils.py of cyh-0/CAVP:
  MIoU.calculate_current_sample / batch_pix_accuracy / batch_intersection_union   (:63-97)
  ForegroundDetect._fast_hist / __call__                                             (:107-117,151-155)
  ForegroundDetect.get_metric_results                                                 (:123-149, minus the .cuda() hop)
Pinned by tests/golden/metrics.pt, produced by oracle/make_golden_metrics.py from the unmodified reference classes.
"""
import numpy
import torch


def miou_sample(output, target, num_classes, ignore_index):
    """-> [correct, labeled, inter[nc], union[nc]] (eval_utils.py:63-97).  `target` is not modified."""
    target = target.clone()
    target[target == ignore_index] = -1
    _, predict = torch.max(output, 1)
    p1 = predict.int() + 1
    t1 = target.reshape(predict.shape).int() + 1
    pixel_labeled = (t1 > 0).sum()
    pixel_correct = ((p1 == t1) * (t1 > 0)).sum()
    p2 = (predict + 1) * (t1 > 0).long()
    t2 = target.reshape(predict.shape).long() + 1
    inter = p2 * (p2 == t2).long()
    area_inter = torch.histc(inter.float(), bins=num_classes, max=num_classes, min=1)
    area_pred = torch.histc(p2.float(), bins=num_classes, max=num_classes, min=1)
    area_lab = torch.histc(t2.float(), bins=num_classes, max=num_classes, min=1)
    area_union = area_pred + area_lab - area_inter
    return [numpy.round(pixel_correct.numpy(), 5), numpy.round(pixel_labeled.numpy(), 5),
            numpy.round(area_inter.numpy(), 5), numpy.round(area_union.numpy(), 5)]


def fast_hist(label_true, label_pred, n_class, ignore):
    """eval_utils.py:107-117."""
    mask = (label_true >= 0) & (label_true < n_class)
    if ignore is not None:
        mask = mask & (label_true != ignore)
    hist = numpy.bincount(n_class * label_true[mask].astype(int) + label_pred[mask], minlength=n_class ** 2)
    return hist.reshape(n_class, n_class)


def foreground_confusion(y_hat, y, n_class, ignore):
    """eval_utils.py:151-155: confusion matrix increment of one call."""
    pred = torch.argmax(y_hat, dim=1).numpy()
    lab = y.reshape(pred.shape).numpy()
    cm = numpy.zeros((n_class, n_class))
    for lt, lp in zip(lab, pred):
        cm += fast_hist(lt.flatten(), lp.flatten(), n_class, ignore)
    return cm


def foreground_scores(cm, class_list=None):
    """eval_utils.py:123-149 on CPU."""
    cm = torch.tensor(numpy.asarray(cm, dtype=numpy.float64))
    tp = torch.diag(cm)
    fp = cm.sum(dim=0) - tp
    fn = cm.sum(dim=1) - tp
    if class_list is not None:
        tp, fp, fn = tp[class_list], fp[class_list], fn[class_list]

    def f_beta(beta2):
        return torch.nanmean(((1 + beta2) * tp) / ((1 + beta2) * tp + beta2 * fn + fp))
    fdr = torch.nanmean(fp / (fp + tp))
    return (torch.round(fdr, decimals=4).numpy(), torch.round(f_beta(1.0), decimals=4).numpy(),
            torch.round(f_beta(0.3), decimals=4).numpy())


def metric_case(seed, B, C, H, W, ignore_frac=0.1, ignore_index=255):
    """Seeded logits / labels shared by the golden generator and the tests (distinct maxima: no argmax ties)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, C, H, W, generator=g)
    target = torch.randint(0, C, (B, 1, H, W), generator=g)
    # make the prediction agree with the label on ~half of the pixels so that every count is non-trivial
    agree = torch.rand(B, 1, H, W, generator=g) < 0.5
    boost = torch.zeros_like(logits).scatter_(1, target, 6.0)
    logits = torch.where(agree.expand_as(logits), logits + boost, logits)
    ign = torch.rand(B, 1, H, W, generator=g) < ignore_frac
    target = torch.where(ign, torch.full_like(target, ignore_index), target)
    # [B, H, W] labels: with a [B, 1, H, W] target the reference's `predict == target` broadcasts to [B, B, H, W] for
    # B > 1 (its validation loop runs at batch 1, trainer_cavp_vpo_mono.py:244-278); B == 1 keeps the 4-D form
    return logits, (target if B == 1 else target[:, 0])
