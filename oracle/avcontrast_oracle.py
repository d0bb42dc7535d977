"""CPU restatement of loss/av_contrast.py:AVContrast (TEST INFRASTRUCTURE ONLY - never imported by the product).

Follows AVContrast.forward (:83-112) and _contrastive (:20-81) line by line, minus the hard-coded `.cuda(local_rank)`
hops (:51,104).  Pinned by tests/golden/avcontrast.pt, which oracle/make_golden_avcontrast.py produces from the
UNMODIFIED reference class with torch.Tensor.cuda patched to the identity for the duration of the call.
"""
import torch
import torch.nn.functional as F


def contrastive(audio, visual, label, temperature, ignore_label=255, eps=1e-12):
    device = audio.device
    features = torch.cat((audio.unsqueeze(1), visual.unsqueeze(1)), dim=1)
    batch_target = [torch.unique(item) for item in label]
    batch_target = [item[item != ignore_label] for item in batch_target]
    batch_target = [item[item != 0] for item in batch_target]
    zero_idx = []
    for i in range(len(batch_target)):
        if len(batch_target[i]) == 0:
            zero_idx.append(i)
            batch_target[i] = torch.tensor([255], device=device)
    contrast_count = features.shape[1]
    contrast_feature = torch.cat(torch.unbind(features, dim=1), dim=0)
    anchor_feature = contrast_feature
    anchor_count = contrast_count
    batch_size = features.shape[0]
    batch_target = torch.stack(batch_target).view(-1, 1)
    mask = torch.eq(batch_target, batch_target.T).to(audio.dtype)
    for i in zero_idx:
        mask[i] = 0
    anchor_dot_contrast = torch.div(torch.matmul(anchor_feature, contrast_feature.T), temperature)
    logits_max, _ = torch.max(anchor_dot_contrast, dim=1, keepdim=True)
    logits = anchor_dot_contrast - logits_max.detach()
    mask = mask.repeat(anchor_count, contrast_count)
    logits_mask = torch.scatter(torch.ones_like(mask), 1, torch.arange(batch_size * anchor_count).view(-1, 1), 0)
    mask = mask * logits_mask
    exp_logits = torch.exp(logits) * logits_mask
    log_prob = logits - torch.log(exp_logits.sum(1, keepdim=True))
    mean_log_prob_pos = (mask * log_prob).sum(1) / (mask.sum(1) + eps)
    loss = -1.0 * mean_log_prob_pos
    return loss.view(anchor_count, batch_size).mean()


def avcontrast(f_v, f_a, labels, temperature, ignore_label=255, eps=1e-12):
    h, w = 128, 128
    f_v = F.normalize(f_v, p=2, dim=1)
    f_a = F.normalize(f_a, p=2, dim=1)
    labels = labels.unsqueeze(1).to(f_v.dtype).clone()
    labels = F.interpolate(labels, (h, w), mode="nearest")
    labels = labels.squeeze(1).long().reshape(labels.shape[0], h * w)
    mask = torch.ones_like(labels) - ((labels == 0).long() + (labels == ignore_label).long())
    masked_v = torch.mul(mask.unsqueeze(-1), f_v)
    masked_v = torch.div(masked_v.sum(1), (mask.sum(1).unsqueeze(-1) + eps))
    return contrastive(f_a, masked_v, labels, temperature, ignore_label, eps)


def case(seed, b, c, H, W, nc=22, empty=(), dtype=torch.float32):
    """Seeded inputs: one foreground rectangle of a random class per image (images in `empty` have none), a corner of
    ignore pixels, and two images sharing a class so that positives exist."""
    g = torch.Generator().manual_seed(seed)
    f_v = torch.randn(b, 128 * 128, c, generator=g, dtype=dtype)
    f_a = torch.randn(b, c, generator=g, dtype=dtype)
    labels = torch.zeros(b, H, W, dtype=torch.int64)
    cls = torch.randint(1, nc, (b,), generator=g)
    if b > 2:
        cls[1] = cls[0]
    for i in range(b):
        if i in empty:
            continue
        y0, x0 = 8 + 3 * i, 5 + 2 * i
        labels[i, y0:y0 + H // 3, x0:x0 + W // 2] = cls[i]
    labels[:, :6, :6] = 255
    return f_v, f_a, labels
