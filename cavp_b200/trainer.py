"""The train-step body of the reference trainers on the C-ABI kernels.

Restates trainer/trainer_cavp_vpo_mono.py:142-193 (epoch-0 branch; `train()` is the only trainer entry point the
reference has - there is no `step()` method):

    output_cat, ctr_feature_cat, pack_ = model_v_(image, audio, None, ow_flag)        (:168)
    output = output_cat[:B] + output_cat[B:] * 0.0                                     (:171)
    l_ctr_av = ContrastLoss(ctr[:B], pix_label, ctr[B:], shuffle_pix_label)            (:183)
    l_ce = CrossEntropyLoss(ignore_index=255)(output, pix_label)                       (:187)
    (l_ce + l_ctr_av).backward(); optimizer_v.step(); optimizer_a.step()               (:189-193)

Unlike the nn.Module path (one autograd node per model call) this drives one kernel graph for forward, both losses
and backward, so nothing full-resolution is produced for the zero-weighted shuffled half in the backward pass.
"""
import torch

from . import _C
from .engine import Graph
from .loss import InfoNCE, ce_forward, contrast_select


def shuffled_labels(pix_label, img_label, shuffle_idx):
    """trainer_cavp_vpo_mono.py:148-151,178-180: labels of the shuffled half (epoch-0 branch, host-side indexing)."""
    shuffle_img_label = img_label.clone()[shuffle_idx]
    shuffle_pix_label = pix_label.clone()[shuffle_idx]
    if_match = torch.all(torch.eq(img_label, shuffle_img_label), dim=1)
    shuffle_pix_label[~if_match] = 0
    shuffle_pix_label[if_match] = pix_label[if_match]
    return shuffle_pix_label


class StepResult:
    __slots__ = ("l_ce", "l_ctr", "out_pred", "out_fusion", "attn_v", "audio", "visual", "launches", "graph",
                 "param_grads")


def train_step(model, image, audio, pix_label, shuffle_pix_label, *, temperature=0.1, ignore_index=255, max_views=512,
               shuffle_idx=None, audio_func=False, assign_grads=True, keep_outputs=False, sel=None, labels_dev=None,
               profile=None, grad_sink=None):
    """forward + CE + ContrastLoss + backward.  Gradients land in `p.grad` (DDP-style callers all-reduce them
    afterwards, see cavp_b200.parallel).  Returns a StepResult; losses are device tensors (no host sync here)."""
    if not image.is_cuda:
        raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    m = model.module if hasattr(model, "module") else model
    dev = image.device
    B, _, H, W = image.shape
    nc = m.num_classes
    g = Graph(dev, prec=m.prec, train=True, sync_bn_group=m._sync_group(), grad_sink=grad_sink)
    g.profile = profile
    if grad_sink is not None:
        # data-parallel mode (cavp_b200.parallel): gradients are produced inside the flat buffer and each bucket's
        # all-reduce starts as soon as the backward tape has passed it (DDP-style overlap, main_vpo_mono.py:131-141)
        grad_sink.begin_step()
        from .parallel import BUCKET_MARKERS
        if len(grad_sink.bucket_range) == len(BUCKET_MARKERS) + 1:  # cavp_buckets(): audio | head | layer4 | rest
            for b, name in enumerate(BUCKET_MARKERS):
                g.callbacks[name] = (lambda b=b: grad_sink.flush_bucket(b, g.param_grads))
    g.use_weight_cache(m)
    if labels_dev is None:
        labels_dev = pix_label.to(dev, torch.int64, non_blocking=True)
    labels_dev = labels_dev.contiguous()
    logits, fusion, proj, fea_a, attn = m.build_graph(g, image, audio, shuffle_idx=shuffle_idx, audio_func=audio_func)
    # ContrastLoss anchor selection: host-side index logic on the labels (torch.randperm from the global CPU generator,
    # as in the reference).  It runs while the forward kernels enqueued above execute, and its index tensors go up
    # through pinned memory with non-blocking copies, so the step has no host<->device synchronisation point.
    if sel is None:
        sel = contrast_select(pix_label, shuffle_pix_label, (fusion.h, fusion.w), max_views, ignore_index)
    gpix = labels_sel = None
    if sel is not None:
        half, pix, labels = sel
        gpix = (pix + half * (B * fusion.h * fusion.w)).pin_memory().to(dev, non_blocking=True)
        labels_sel = labels.pin_memory().to(dev, non_blocking=True)
    # forward_cls + CrossEntropyLoss.  CE on output_cat[:B] + output_cat[B:]*0.0 == CE on the first B images (value
    # and gradient).  Callers that want the full-resolution prediction back (keep_outputs) get it materialised for all
    # rows, as the reference returns it (cavp_model.py:138-141); otherwise the upsample is fused into the loss kernels
    # and [2B, nc, H, W] is never written.
    pred = None
    if keep_outputs:
        pred = g.upsample_to_nchw(logits, nc, H, W)
        ce = ce_forward(g, pred.data_ptr(), labels_dev, B, nc, H * W, ignore_index)
    else:
        lse = g.empty(B * H * W)
        partials = g.empty(_C.query("cavp_ce_nblocks", B, H * W), 2)
        ce = g.empty(2)
        g.work(nbytes=4.0 * B * logits.h * logits.w * logits.ld + 12.0 * B * H * W)
        g.call("cavp_upsample_ce_fwd", logits.ptr, logits.ld, logits.h, logits.w, H, W, B, nc, labels_dev.data_ptr(),
               ignore_index, lse.data_ptr(), partials.data_ptr(), ce.data_ptr())
    nce = None
    if sel is not None:
        nce = InfoNCE(g, [(fusion.ptr, fusion.ld, gpix, fusion.c)], labels_sel, temperature)

    # ---- backward
    if keep_outputs:
        dpred = g.empty(B, nc, H, W)
        g.call("cavp_ce_bwd", pred.data_ptr(), labels_dev.data_ptr(), B, nc, H * W, ignore_index, ce.data_ptr(), 0,
               dpred.data_ptr())
        g.upsample_to_nchw_backward(logits, nc, dpred, n_valid=B)
    else:
        dlog, accumulate = g.grad_target(logits)
        assert not accumulate
        g.call("cavp_upsample_ce_bwd", logits.ptr, logits.ld, logits.h, logits.w, H, W, logits.n, B, nc, logits.c,
               labels_dev.data_ptr(), ignore_index, lse.data_ptr(), ce.data_ptr(), 0, dlog.ptr, dlog.ld)
    if nce is not None:
        dfus, accumulate = g.grad_target(fusion)
        if not accumulate:
            g.zero_act(dfus)
        nce.backward(None, [(dfus.ptr, dfus.ld)])
    g.backward()
    if grad_sink is not None:
        grad_sink.finish(g.param_grads)  # p.grad = views of the (averaged) flat buffer
        grad_sink.prezeroed = False
    elif assign_grads:
        for p in m.parameters():
            gr = g.param_grads.get(id(p))
            if gr is not None:
                p.grad = gr if p.grad is None else p.grad.add_(gr)
    res = StepResult()
    res.l_ce = ce[0]
    res.l_ctr = nce.loss[0] if nce is not None else torch.zeros((), device=dev)
    res.out_pred = pred if keep_outputs else None
    res.out_fusion = fusion.nchw() if keep_outputs else None
    res.attn_v = attn.unsqueeze(-1) if keep_outputs else None
    res.audio = fea_a.dense().reshape(fea_a.rows, fea_a.c, 1, 1) if keep_outputs else None
    res.visual = proj.nchw() if keep_outputs else None
    res.launches = g.launches
    res.graph = g if keep_outputs else None
    res.param_grads = g.param_grads
    return res
