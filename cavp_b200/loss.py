"""Losses of the CAVP train step on the C-ABI kernels.

  ContrastLoss      <- loss/contrastive_aud.py:7-142 (pixel InfoNCE with random anchor sampling)
  CrossEntropyLoss  <- loss/losser.py:53,60-62 (nn.CrossEntropyLoss(ignore_index=255), mean over valid pixels)

The random anchor *selection* is host-side index logic exactly as in the reference (it draws torch.randperm from the
global CPU generator, contrastive_aud.py:86,122-123) and depends on the labels only; everything that touches
embeddings / logits (normalise, gather, A x A similarity GEMM, InfoNCE rows, their gradients, CE and its gradient)
runs in kernels.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C
from .engine import Act, Graph, pad4


# ----------------------------------------------------------------------------------------------------------------
def contrast_select(gt_match, gt_shuffle, feat_hw, max_views=512, ignore_idx=255):
    """Index form of ContrastLoss.extraction_samples / foreground_random_selection (contrastive_aud.py:76-142).

    Returns (half[A], flat_pixel[A], label[A]) as CPU int64 tensors - half 0 = matched embeddings, 1 = shuffled
    embeddings, flat_pixel = b*h*w + y*w + x - or None when no foreground class has >= max_views pixels
    (the reference then returns tensor([0.])).  Consumes the global CPU RNG in the reference's order."""
    gm = F.interpolate(gt_match.detach().cpu().unsqueeze(1).float(), size=feat_hw, mode="nearest").squeeze(1).long()
    gs = F.interpolate(gt_shuffle.detach().cpu().unsqueeze(1).float(), size=feat_hw, mode="nearest").squeeze(1).long()
    gm, gs = gm.flatten(), gs.flatten()
    fg_pos = ((gm > 0) & (gm != ignore_idx)).nonzero().flatten()
    fg_lab = gm[fg_pos]
    pos, lab = [], []
    for item in torch.unique(fg_lab):
        cur = fg_pos[fg_lab == item]
        if cur.numel() < max_views:
            continue
        perm = torch.randperm(cur.numel())
        pos.append(cur[perm][:max_views])
        lab.append(torch.full((min(max_views, cur.numel()),), int(item), dtype=torch.int64))
    if not pos:
        return None
    bg_pos = (gm == 0).nonzero().flatten()
    n = int(min(max_views, fg_pos.numel(), bg_pos.numel()))
    p1 = torch.randperm(bg_pos.numel())
    p2 = torch.randperm(fg_pos.numel())
    sel_bg, sel_sh = bg_pos[p1][:n], fg_pos[p2][:n]
    n_match = sum(p.numel() for p in pos) + n
    half = torch.cat([torch.zeros(n_match, dtype=torch.int64), torch.ones(n, dtype=torch.int64)])
    pix = torch.cat(pos + [sel_bg, sel_sh])
    labels = torch.cat(lab + [gm[sel_bg], gs[sel_sh]])
    return half, pix, labels


class InfoNCE:
    """Device side of ContrastLoss: anchors = normalize(f)[selected pixels]; loss = info_nce(anchors, labels)."""

    def __init__(self, g, sources, labels, temperature):
        """sources: list of (base_ptr, ld, pix int64 device tensor) gathered in order into the anchor matrix."""
        self.g, self.sources, self.temperature = g, sources, temperature
        dev = g.device
        self.C = C = 304 if not sources else sources[0][3]
        self.A = A = sum(int(s[2].numel()) for s in sources)
        self.A4 = A4 = pad4(A)
        self.labels = labels.to(dev, torch.int64).contiguous()
        self.anchors = g.empty(A, C)
        self.inv_norm = g.empty(A)
        row = 0
        for base, ld, pix, _ in sources:
            n = int(pix.numel())
            if n:
                g.call("cavp_l2norm_gather", base, ld, pix.data_ptr(), n, C, self.anchors[row:].data_ptr(), C,
                       self.inv_norm[row:].data_ptr())
            row += n
        self.S = g.empty(A, A4)
        a_act = Act(self.anchors, A, 1, 1, C, needs_grad=False)
        s_act = Act(self.S, A, 1, 1, A)
        # S = A A^T: the anchor matrix is also the weight operand - pre-split once (3 MB) so that it is fetched by TMA and
        # the GEMM runs on the CTA-pair kernels (256-column tiles when A % 256 == 0) instead of the generic tile kernel
        g._igemm(a_act, self._presplit(self.anchors), A, C, s_act, geom=(1, 1, 1, 1, 1, 0, 1))
        self.rows = g.empty(3, A)  # rowmax, rowneg, rowmean
        self.loss = g.empty(1)
        g.call("cavp_infonce_fwd", self.S.data_ptr(), A4, self.labels.data_ptr(), A, float(temperature),
               self.rows[0].data_ptr(), self.rows[1].data_ptr(), self.rows[2].data_ptr(), self.loss.data_ptr())

    def _presplit(self, w):
        """[hi | lo] TF32 split of a K-major operand -> (tensor, lo offset) as Graph._igemm takes it"""
        sp = self.g.empty(2, w.shape[0], w.shape[1])
        self.g.call("cavp_split_tf32", w.data_ptr(), sp[0].data_ptr(), sp[1].data_ptr(), w.numel())
        return sp[0], w.numel()

    def backward(self, gscale, targets):
        """gscale: device tensor [1] (upstream gradient) or None.  targets: list of (grad_base_ptr, ld) aligned with
        `sources`; the anchor gradients are scatter-added there."""
        g, A, A4, C = self.g, self.A, self.A4, self.C
        G = g.empty(A, A4)
        g.call("cavp_infonce_bwd", self.S.data_ptr(), A4, self.labels.data_ptr(), A, float(self.temperature),
               self.rows[0].data_ptr(), self.rows[1].data_ptr(), _C.ptr(gscale), G.data_ptr(), A4)
        # dAnchors = G A + G^T A
        at = g.zeros(C, A4)
        g.call("cavp_transpose", self.anchors.data_ptr(), at.data_ptr(), A, C, C, A4, 1, 0, 0)
        d1 = g.empty(A, C)
        g._igemm(Act(G, A, 1, 1, A4, needs_grad=False), self._presplit(at), C, A4, Act(d1, A, 1, 1, C),
                 geom=(1, 1, 1, 1, 1, 0, 1))
        d2 = g.empty(A4, C)
        wsplits = Graph.wgrad_splits(A, A4, C)
        if wsplits > 1:
            g.call("cavp_zero", d2.data_ptr(), d2.numel() * 4)
        g.work(flops=2.0 * A * A4 * C, tag=f"wgrad P{A} Cout{A4} K{C} (InfoNCE G^T A) splits{wsplits}")
        g.call("cavp_igemm_wgrad", G.data_ptr(), self.anchors.data_ptr(), d2.data_ptr(), A, 1, 1, C, C, 1, 1, 1, 1, 1, 0,
               1, A4, A4, wsplits, g.prec_tf)
        g.call("cavp_add_inplace", d1.data_ptr(), d2.data_ptr(), A * C, 1.0)
        row = 0
        for (base, ld, pix, _), (gbase, gld) in zip(self.sources, targets):
            n = int(pix.numel())
            if n:
                g.call("cavp_l2norm_scatter_bwd", d1[row:].data_ptr(), self.anchors[row:].data_ptr(), C,
                       self.inv_norm[row:].data_ptr(), pix.data_ptr(), n, C, gbase, gld)
            row += n


def _nhwc_view(t):
    """logical NCHW tensor -> (contiguous [B*h*w, C] view or copy)"""
    v = t.permute(0, 2, 3, 1)
    if not v.is_contiguous():
        v = v.contiguous()
    return v.reshape(-1, t.shape[1])


class _ContrastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, em, es, sel, temperature, prec):
        half, pix, labels = sel
        dev = em.device
        g = Graph(dev, prec=prec, train=True)
        vm, vs = _nhwc_view(em), _nhwc_view(es)
        pm = pix[half == 0].to(dev).contiguous()
        ps = pix[half == 1].to(dev).contiguous()
        C = em.shape[1]
        nce = InfoNCE(g, [(vm.data_ptr(), vm.stride(0), pm, C), (vs.data_ptr(), vs.stride(0), ps, C)], labels,
                      temperature)
        ctx.nce, ctx.shapes, ctx.keep = nce, (em.shape, es.shape), (vm, vs)
        return nce.loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        nce, (sm, ss) = ctx.nce, ctx.shapes
        g = nce.g
        dm = g.zeros(sm[0] * sm[2] * sm[3], sm[1])
        ds = g.zeros(ss[0] * ss[2] * ss[3], ss[1])
        nce.backward(gout.reshape(1).contiguous().float(), [(dm.data_ptr(), sm[1]), (ds.data_ptr(), ss[1])])
        return (dm.view(sm[0], sm[2], sm[3], sm[1]).permute(0, 3, 1, 2),
                ds.view(ss[0], ss[2], ss[3], ss[1]).permute(0, 3, 1, 2), None, None, None)


class ContrastLoss(nn.Module):
    """loss/contrastive_aud.py:7-37 - same constructor and call signature."""

    def __init__(self, temperature, ignore_idx, max_views, prec=2):
        super().__init__()
        self.ignore_idx = ignore_idx
        self.ood_idx = 254
        self.eps = 1e-12
        self.temperature = temperature
        self.max_views = max_views
        self.prec = prec

    def forward(self, embeds_match, gt_match, embeds_shuffle, gt_shuffle):
        if not embeds_match.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        sel = contrast_select(gt_match, gt_shuffle, embeds_match.shape[2:], self.max_views, self.ignore_idx)
        if sel is None:
            return torch.tensor([.0], device=gt_match.device)  # contrastive_aud.py:34-35
        return _ContrastFn.apply(embeds_match, embeds_shuffle, sel, self.temperature, self.prec)


# ----------------------------------------------------------------------------------------------------------------
def ce_forward(g, logits_ptr, labels, B, C, HW, ignore_index=255):
    """-> device tensor [2] = (mean loss over valid pixels, #valid)."""
    nb = _C.query("cavp_ce_nblocks", B, HW)
    partials = g.empty(nb, 2)
    out = g.empty(2)
    g.call("cavp_ce_fwd", logits_ptr, labels.data_ptr(), B, C, HW, ignore_index, partials.data_ptr(), out.data_ptr())
    return out


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        logits = logits.contiguous()
        if labels.dtype.is_floating_point or labels.dtype == torch.bool:
            raise RuntimeError("CrossEntropyLoss: class-index labels must be an integer tensor")
        labels = labels.to(torch.int64).contiguous()  # the kernels read int64; other integer widths are widened here
        B, C = logits.shape[:2]
        HW = logits[0, 0].numel()
        g = Graph(logits.device, train=True)
        out = ce_forward(g, logits.data_ptr(), labels, B, C, HW, ignore_index)
        ctx.g, ctx.saved, ctx.meta = g, (logits, labels, out), (B, C, HW, ignore_index)
        return out[0].reshape(())

    @staticmethod
    def backward(ctx, gout):
        logits, labels, out = ctx.saved
        B, C, HW, ignore_index = ctx.meta
        d = torch.empty_like(logits)
        ctx.g.call("cavp_ce_bwd", logits.data_ptr(), labels.data_ptr(), B, C, HW, ignore_index, out.data_ptr(),
                   gout.reshape(1).contiguous().float().data_ptr(), d.data_ptr())
        return d, None, None


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss(ignore_index=255) as the reference's Losser uses it (loss/losser.py:53,60-62)."""

    def __init__(self, ignore_index=255):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, output, pix_label, pack_=None):
        if not output.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        return _CEFn.apply(output, pix_label, self.ignore_index)


class Losser(nn.Module):
    """loss/losser.py:49-62: forward returns only the cross-entropy term."""

    def __init__(self, num_classes, local_rank=0):
        super().__init__()
        self.num_classes = num_classes
        self.loss_ce = CrossEntropyLoss(ignore_index=255)
        self.local_rank = local_rank

    def forward(self, output, pix_label, pack_=None):
        return self.loss_ce(output, pix_label)


# ---------------------------------------------------------------------------------------------------------------------
# AVContrast (loss/av_contrast.py) - SURVEY.md 8(f) N4.  Dead code in the reference (constructed in loss/losser.py:57,
# never called by a trainer); provided because BASELINE.json's north_star names it.
def avcontrast_host_labels(labels, ignore_label=255, size=128):
    """Host-side index logic of AVContrast.forward / _contrastive on the label map (loss/av_contrast.py:27-40,97-106):
    nearest resize to 128 x 128 (F.interpolate(mode="nearest") index rule: src = floor(dst * in / out)), foreground mask
    (label != 0 and != ignore), per-image foreground class (-1 = none).  The reference's torch.stack(batch_target) only
    works when every image has at most one foreground class; more raises here as it does there."""
    lab = labels.detach().cpu()
    b, H, W = lab.shape
    iy = torch.div(torch.arange(size) * H, size, rounding_mode="floor").clamp_(max=H - 1)
    ix = torch.div(torch.arange(size) * W, size, rounding_mode="floor").clamp_(max=W - 1)
    small = lab[:, iy][:, :, ix].reshape(b, size * size)
    mask = ((small != 0) & (small != ignore_label))
    target = torch.full((b,), -1, dtype=torch.int32)
    for i in range(b):
        u = torch.unique(small[i])
        u = u[(u != ignore_label) & (u != 0)]
        if len(u) > 1:
            raise ValueError("AVContrast: more than one foreground class in an image (the reference's torch.stack of "
                             "per-image unique labels fails on this input too)")
        if len(u) == 1:
            target[i] = int(u[0])
    return mask.to(torch.uint8), mask.sum(1).float(), target


class _AVContrastFn(torch.autograd.Function):
    NCHUNK = 64

    @staticmethod
    def forward(ctx, f_v, f_a, mask, cnt, target, temperature, eps):
        b, hw, c = f_v.shape
        dev = f_v.device
        st = torch.cuda.current_stream(dev).cuda_stream
        fv = f_v.detach().float().contiguous()
        fa = f_a.detach().float().contiguous()
        nch = _AVContrastFn.NCHUNK
        partials = torch.empty(b, nch, 2, c, device=dev)
        _C.call("cavp_avc_colstats", fv.data_ptr(), mask.data_ptr(), b, hw, c, nch, partials.data_ptr(), st)
        feats = torch.empty(2 * b, c, device=dev)
        dfeat = torch.empty(2 * b, c, device=dev)
        nrm, msum = torch.empty(b, c, device=dev), torch.empty(b, c, device=dev)
        loss = torch.empty(1, device=dev)
        d_fa, dms, dnn = torch.empty(b, c, device=dev), torch.empty(b, c, device=dev), torch.empty(b, c, device=dev)
        _C.call("cavp_avc_loss", partials.data_ptr(), nch, b, c, fa.data_ptr(), cnt.data_ptr(), target.data_ptr(),
                float(temperature), float(eps), feats.data_ptr(), dfeat.data_ptr(), nrm.data_ptr(), msum.data_ptr(),
                loss.data_ptr(), d_fa.data_ptr(), dms.data_ptr(), dnn.data_ptr(), st)
        ctx.save_for_backward(fv, mask, d_fa, dms, dnn)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        fv, mask, d_fa, dms, dnn = ctx.saved_tensors
        b, hw, c = fv.shape
        st = torch.cuda.current_stream(fv.device).cuda_stream
        gs = g.detach().float().reshape(1).contiguous()
        dfv = torch.empty_like(fv)
        _C.call("cavp_avc_bwd", fv.data_ptr(), mask.data_ptr(), dms.data_ptr(), dnn.data_ptr(), gs.data_ptr(), b, hw, c,
                dfv.data_ptr(), st)
        dfa = torch.empty_like(d_fa)
        # d_fa * g: a [b, c] rescale on the same kernel family (bn_apply: out = y*scale + shift with scale = g)
        scale = gs.expand(c).contiguous()            # keep both alive until the launch is enqueued
        shift = torch.zeros(c, device=fv.device)
        _C.call("cavp_bn_apply", d_fa.data_ptr(), c, scale.data_ptr(), shift.data_ptr(), 0, 0, dfa.data_ptr(), c, b, c, 0,
                0.0, st)
        return dfv, dfa, None, None, None, None, None


class AVContrast(nn.Module):
    """loss/av_contrast.py:8-112: same constructor (temp1, local_rank) and forward(f_v [b, hw, c], f_a [b, c], labels
    [b, H, W]); hw must be 128*128 as in the reference (:89 hard-codes h, w = 128)."""

    def __init__(self, temp1, local_rank=0):
        super().__init__()
        self.temperature = temp1
        self.base_temperature = self.temperature
        self.ignore_label = 255
        self.eps = 1e-12
        self.local_rank = local_rank

    def forward(self, f_v, f_a, labels=None):
        if not f_v.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        b, hw, c = f_v.shape
        if hw != 128 * 128:
            raise ValueError("AVContrast expects a 128 x 128 token map (loss/av_contrast.py:89)")
        if c % 4:
            raise ValueError("channel count must be a multiple of 4")
        mask, cnt, target = avcontrast_host_labels(labels, self.ignore_label)
        dev = f_v.device
        return _AVContrastFn.apply(f_v, f_a, mask.to(dev), cnt.to(dev), target.to(dev), self.temperature, self.eps)
