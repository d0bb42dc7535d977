"""CPU restatement of the CAVP hot path (plain torch fp32 ops on a state_dict).

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  Nothing under cavp_b200/ imports it and the product
path fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors.  This restatement is
pinned against outputs of the UNMODIFIED reference imported in the build container
(oracle/make_golden.py -> tests/golden/*.pt; checked by tests/test_oracle_golden.py), with one caveat:
`timm.Mlp` (timm==0.4.9, absent from /root/reference) is restated from its published definition
and has no reference-side test => that one boundary is "parity unpinned".

Every function cites the reference file:line it follows (paths relative to /root/reference).
The model is a pure function of (state_dict, inputs); BN running-stat updates are returned in `new_buffers`.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # models/visual/deeplabv3/encoder_decoder.py:10
BN_MOM = 0.1  # encoder_decoder.py:11
RESNET50_LAYERS = (3, 4, 6, 3)  # models/visual/backbones/resnet.py:222


class State:
    """state_dict accessor with a key prefix; collects BN buffer updates in train mode."""

    def __init__(self, sd, train, new_buffers=None, prefix=""):
        self.sd = sd
        self.train = train
        self.new_buffers = {} if new_buffers is None else new_buffers
        self.prefix = prefix

    def sub(self, name):
        return State(self.sd, self.train, self.new_buffers, self.prefix + name + ".")

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def has(self, k):
        return (self.prefix + k) in self.sd


def batch_norm(st: State, x, momentum=BN_MOM, eps=BN_EPS):
    """nn.BatchNorm2d semantics: batch statistics (biased var) in train mode, running stats in eval;
    running_var updated with the unbiased variance (resnet.py:64, encoder_decoder.py:68)."""
    w, b = st["weight"], st["bias"]
    rm, rv = st["running_mean"], st["running_var"]
    if not st.train:
        return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)
    rm2, rv2 = rm.detach().clone(), rv.detach().clone()
    y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, eps)  # updates the clones in place
    st.new_buffers[st.prefix + "running_mean"] = rm2
    st.new_buffers[st.prefix + "running_var"] = rv2
    st.new_buffers[st.prefix + "num_batches_tracked"] = st["num_batches_tracked"] + 1
    return y


# Operand rounding for the bf16 row (BASELINE.json configs[2]; the reference itself has no reduced-precision path,
# SURVEY F7): with OPERAND_ROUND = torch.bfloat16 every conv / Linear sees its input and weight rounded to bf16 and
# multiplies them exactly (fp32 / fp64 accumulation) - the arithmetic of a tensor-core GEMM with bf16 operands and
# wide accumulation, i.e. what the product's bf16 kernels compute.  Layers whose channel counts are not multiples of 8
# keep full precision (the product runs those on the TF32 pipe).  Everything else (BN, LN, gate, losses) is unchanged.
OPERAND_ROUND = None


def _rnd(x, w):
    if OPERAND_ROUND is None or w.shape[0] % 8 or w.shape[1] % 8:
        return x, w
    return x.to(OPERAND_ROUND).to(x.dtype), w.to(OPERAND_ROUND).to(w.dtype)


def linear(x, w, b=None):
    x, w = _rnd(x, w)
    return F.linear(x, w, b)


def conv(st: State, x, stride=1, padding=0, dilation=1):
    bias = st["bias"] if st.has("bias") else None
    x, w = _rnd(x, st["weight"])
    return F.conv2d(x, w, bias, stride=stride, padding=padding, dilation=dilation)


# ---------------------------------------------------------------------------------------------
# A1  ResNet-50 deep-stem backbone  (resnet.py:53-98,101-201; encoder_decoder.py:14-59)
# ---------------------------------------------------------------------------------------------
def bottleneck(st: State, x, stride, dilation, has_down, down_stride):
    """resnet.py:75-98: 1x1 -> BN -> ReLU -> 3x3(stride, dilation) -> BN -> ReLU -> 1x1 -> BN -> (+res) -> ReLU."""
    out = F.relu(batch_norm(st.sub("bn1"), conv(st.sub("conv1"), x)))
    out = F.relu(batch_norm(st.sub("bn2"), conv(st.sub("conv2"), out, stride=stride, padding=dilation,
                                                dilation=dilation)))
    out = batch_norm(st.sub("bn3"), conv(st.sub("conv3"), out))
    res = x
    if has_down:
        res = batch_norm(st.sub("downsample.1"), conv(st.sub("downsample.0"), x, stride=down_stride))
    return F.relu(out + res)


def resnet_plan(dilation_flags):
    """Per-block (stride, dilation, has_down, down_stride) after _make_layer (resnet.py:155-184) and the
    layer4 re-dilation of Backbone._nostride_dilate (encoder_decoder.py:40-55: child i of layer4 gets
    dilate = 2, 4, 8 on every 3x3 conv and stride-2 convs become stride 1)."""
    plan = []
    cur_dil = 1
    inplanes = 128
    for li, (planes, nblocks, stride) in enumerate(zip((64, 128, 256, 512), RESNET50_LAYERS, (1, 2, 2, 2))):
        dilate = (False, *dilation_flags)[li]
        prev_dil = cur_dil
        if dilate:
            cur_dil *= stride
            stride = 1
        blocks = []
        has_down = stride != 1 or inplanes != planes * 4
        blocks.append(dict(stride=stride, dilation=prev_dil, has_down=has_down, down_stride=stride))
        inplanes = planes * 4
        for _ in range(1, nblocks):
            blocks.append(dict(stride=1, dilation=cur_dil, has_down=False, down_stride=1))
        plan.append(blocks)
    d = 2
    for blk in plan[3]:
        blk["stride"] = 1
        blk["down_stride"] = 1
        blk["dilation"] = d
        d *= 2
    return plan


def backbone(st: State, image, dilation_flags):
    """resnet.py:186-201 with the deep stem (resnet.py:107-121)."""
    s = st.sub("backbone")
    c1 = s.sub("conv1")
    x = conv(c1.sub("0"), image, stride=2, padding=1)
    x = F.relu(batch_norm(c1.sub("1"), x))
    x = conv(c1.sub("3"), x, padding=1)
    x = F.relu(batch_norm(c1.sub("4"), x))
    x = conv(c1.sub("6"), x, padding=1)
    x = F.relu(batch_norm(s.sub("bn1"), x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, blocks in enumerate(resnet_plan(dilation_flags)):
        for bi, cfg in enumerate(blocks):
            x = bottleneck(s.sub(f"layer{li + 1}.{bi}"), x, **cfg)
        feats.append(x)
    return feats


# ---------------------------------------------------------------------------------------------
# A2  DeepLabV3+ head  (encoder_decoder.py:62-164)
# ---------------------------------------------------------------------------------------------
def aspp(st: State, x):
    """encoder_decoder.py:137-164; LeakyReLU slope 0.01 (nn.LeakyReLU default, :135)."""
    outs = [conv(st.sub("map_convs.0"), x)]
    for i, r in enumerate((6, 12, 18)):
        outs.append(conv(st.sub(f"map_convs.{i + 1}"), x, padding=r, dilation=r))
    out = torch.cat(outs, dim=1)
    out = F.leaky_relu(batch_norm(st.sub("map_bn"), out), 0.01)
    out = conv(st.sub("red_conv"), out)
    pool = x.view(x.size(0), x.size(1), -1).mean(dim=-1).view(x.size(0), x.size(1), 1, 1)
    pool = conv(st.sub("global_pooling_conv"), pool)
    pool = F.leaky_relu(batch_norm(st.sub("global_pooling_bn"), pool), 0.01)
    pool = conv(st.sub("pool_red_conv"), pool)
    out = out + pool  # broadcast == .repeat(1,1,H,W) (:149-151)
    return F.leaky_relu(batch_norm(st.sub("red_bn"), out), 0.01)


def forward_feature(st: State, feats):
    """encoder_decoder.py:97-105."""
    f = aspp(st.sub("aspp"), feats[-1])
    low = feats[0]
    low = F.relu(batch_norm(st.sub("reduce.1"), conv(st.sub("reduce.0"), low)))
    f = F.interpolate(f, size=low.shape[-2:], mode="bilinear", align_corners=True)
    return torch.cat((f, low), dim=1)


def upsampling(st: State, x):
    """encoder_decoder.py:62-75."""
    lc = st.sub("last_conv")
    x = F.relu(batch_norm(lc.sub("1"), conv(lc.sub("0"), x, padding=1)))
    x = F.relu(batch_norm(lc.sub("4"), conv(lc.sub("3"), x, padding=1)))
    return conv(st.sub("classifier"), x)


# ---------------------------------------------------------------------------------------------
# A3  audio backbone  (models/audio/backbones/vgg.py:5-36, models/audio/audio_network.py:9-34)
# ---------------------------------------------------------------------------------------------
VGG_CFG = (64, "M", 128, "M", 256, 256, "M", 512, 512, "M")


def vgg_audio(st: State, x):
    s = st.sub("backbone")
    idx = 0
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.relu(conv(s.sub(f"features.{idx}"), x, padding=1))
            idx += 2
    x = x.permute(0, 2, 3, 1).contiguous().view(x.size(0), -1)  # vgg.py:19-22 (NCHW -> NHWC flatten)
    for i in (0, 2, 4):
        e = s.sub(f"embeddings.{i}")
        x = F.relu(linear(x, e["weight"], e["bias"]))
    return x


def resnet18_audio(st: State, x):
    """torchvision resnet18 with conv1 replaced, AdaptiveMaxPool2d and fc 512->out (audio_network.py:19-25)."""
    s = st.sub("backbone")
    x = F.relu(batch_norm(s.sub("bn1"), conv(s.sub("conv1"), x, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(4):
        for bi in range(2):
            b = s.sub(f"layer{li + 1}.{bi}")
            stride = 2 if (li > 0 and bi == 0) else 1
            idn = x
            out = F.relu(batch_norm(b.sub("bn1"), conv(b.sub("conv1"), x, stride=stride, padding=1)))
            out = batch_norm(b.sub("bn2"), conv(b.sub("conv2"), out, padding=1))
            if b.has("downsample.0.weight"):
                idn = batch_norm(b.sub("downsample.1"), conv(b.sub("downsample.0"), x, stride=stride))
            x = F.relu(out + idn)
    x = F.adaptive_max_pool2d(x, 1).flatten(1)
    fc = s.sub("fc")
    return linear(x, fc["weight"], fc["bias"])


# ---------------------------------------------------------------------------------------------
# A4-A7  fusion: projector + cross attention  (cavp_model.py:143-154; attn.py:30-39,64-106,146-162,232-244)
# ---------------------------------------------------------------------------------------------
def mlp(st: State, x):
    """timm 0.4.9 Mlp: fc1 -> GELU(erf) -> fc2 (dropout p = 0)."""
    x = F.gelu(linear(x, st["fc1.weight"], st["fc1.bias"]))
    return linear(x, st["fc2.weight"], st["fc2.bias"])


def layer_norm(st: State, x):
    return F.layer_norm(x, (x.shape[-1],), st["weight"], st["bias"], 1e-5)


def attention(st: State, x_q, x_k, x_v, num_heads=4):
    """attn.py:73-106: bias-free q/k/v Linear, per-head sigmoid((q k^T) * hd^-0.5), attn @ v, proj(+bias)."""
    B, N, C = x_q.shape
    hd = C // num_heads

    def split(x, w):
        return linear(x, w).reshape(x.shape[0], x.shape[1], num_heads, hd).permute(0, 2, 1, 3)

    q, k, v = split(x_q, st["q.weight"]), split(x_k, st["k.weight"]), split(x_v, st["v.weight"])
    attn = torch.sigmoid((q @ k.transpose(-2, -1)) * hd ** -0.5)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return linear(x, st["proj.weight"], st["proj.bias"]), attn


def cross_attention(st: State, fea_v, fea_a, live_only=True):
    """attn.py:232-244 with depth=1 Block.forward_ca (attn.py:146-162).

    fea_v [rows,C,h,w], fea_a [rows,C,1,1].  The audio<-visual branch (attn.py:161) does not influence
    any returned tensor used by CAVP.forward_fusion (cavp_model.py:151 discards f_a); it is computed only
    when live_only=False (used once in the golden check to confirm f_v/attn_v do not depend on it).
    """
    rows, C, h, w = fea_v.shape
    f_v = fea_v.flatten(2).transpose(1, 2)  # b c h w -> b (h w) c
    f_a = fea_a.flatten(2).transpose(1, 2)
    pv, pa = st.sub("patch_embed_v.proj"), st.sub("patch_embed_a.proj")
    f_v = linear(f_v, pv["weight"], pv["bias"])
    f_a = linear(f_a, pa["weight"], pa["bias"])
    blk = st.sub("blocks.0")
    f_v = layer_norm(blk.sub("norm1"), f_v)
    f_a = layer_norm(blk.sub("norm1"), f_a)
    o, attn_v = attention(blk.sub("attn"), f_v, f_a, f_a)
    f_v = f_v + o
    f_v = f_v + mlp(blk.sub("mlp"), layer_norm(blk.sub("norm2"), f_v))
    if not live_only:
        o2, _ = attention(blk.sub("attn"), f_a, f_v, f_v)
        f_a = f_a + o2
        f_a = f_a + mlp(blk.sub("mlp"), layer_norm(blk.sub("norm2"), f_a))
    f_v = layer_norm(st.sub("norm"), f_v)
    return f_v, f_a, attn_v


def forward_fusion(st: State, visual, fea_a):
    """cavp_model.py:143-154."""
    b, c, h, w = visual.shape
    tok = visual.flatten(2).transpose(1, 2)
    fea_v = mlp(st.sub("visual_projector"), tok).transpose(1, 2).reshape(b, c, h, w)
    fea_v_proj = fea_v.clone()
    fea_a4 = fea_a.unsqueeze(-1).unsqueeze(-1)
    f_v, _, attn_v = cross_attention(st.sub("cross_att"), fea_v, fea_a4)
    out = f_v.transpose(1, 2).reshape(b, c, h, w)
    return out, {"audio": fea_a4, "visual": fea_v_proj, "attn_v": attn_v}


# ---------------------------------------------------------------------------------------------
# A8/A9  CAVP.forward  (cavp_model.py:138-141,156-205)
# ---------------------------------------------------------------------------------------------
def audio_backbone(st: State, audio, kind):
    s = st.sub("audio_backbone")
    return vgg_audio(s, audio) if kind == "vgg" else resnet18_audio(s, audio)


def cavp_forward(sd, image, audio, *, dilation_flags, audio_kind="vgg", train=True, shuffle_idx=None,
                 audio_func=False):
    """Returns (out_pred, out_fusion, pack, new_buffers).  Train mode doubles the visual batch
    (cavp_model.py:181) and expects 2B audio rows, or B rows + shuffle_idx when audio_func (forward_audio :156-173)."""
    st = State(sd, train)
    input_shape = image.shape[-2:]
    feats = backbone(st.sub("backbone"), image, dilation_flags)
    fea_v = forward_feature(st.sub("segment"), feats)
    if train:
        fea_v = torch.cat((fea_v, fea_v.clone()), dim=0)
        fea_a = audio_backbone(st, audio, audio_kind)
        if audio_func:
            fea_a = torch.cat((fea_a, fea_a[shuffle_idx]), dim=0)
    else:
        fea_a = audio_backbone(st, audio, audio_kind)
    out_fusion, pack = forward_fusion(st, fea_v, fea_a)
    out = upsampling(st.sub("segment.upsample"), out_fusion.contiguous())
    out_pred = F.interpolate(out, size=input_shape, mode="bilinear", align_corners=False)
    return out_pred, out_fusion, pack, st.new_buffers


# ---------------------------------------------------------------------------------------------
# A10/A11  losses  (loss/losser.py:60-62; loss/contrastive_aud.py:17-142)
# ---------------------------------------------------------------------------------------------
def cross_entropy(output, pix_label, ignore_index=255):
    return F.cross_entropy(output, pix_label, ignore_index=ignore_index)


def contrast_select(gt_match, gt_shuffle, feat_hw, max_views=512, ignore_idx=255):
    """Index form of ContrastLoss.extraction_samples / foreground_random_selection
    (contrastive_aud.py:76-142).  Depends on labels only.  Consumes the global CPU torch RNG in the
    reference's order: one randperm per kept foreground class (ascending class id), then background,
    then shuffled-half.  Returns (half[A], flat_pixel[A], label[A]) with half 0 = matched embeddings,
    1 = shuffled embeddings, flat_pixel indexing (b*h*w + y*w + x); or None when no class qualifies."""
    gm = F.interpolate(gt_match.unsqueeze(1).float(), size=feat_hw, mode="nearest").squeeze(1).long().flatten()
    gs = F.interpolate(gt_shuffle.unsqueeze(1).float(), size=feat_hw, mode="nearest").squeeze(1).long().flatten()
    fg = (gm > 0) & (gm != ignore_idx)
    fg_pos = fg.nonzero().flatten()
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
    sel_bg = bg_pos[p1][:n]
    sel_sh = fg_pos[p2][:n]
    half = torch.cat([torch.zeros(sum(p.numel() for p in pos) + n, dtype=torch.int64),
                      torch.ones(n, dtype=torch.int64)])
    pix = torch.cat(pos + [sel_bg, sel_sh])
    labels = torch.cat(lab + [gm[sel_bg], gs[sel_sh]])
    return half, pix, labels


def info_nce(anchors, labels, temperature=0.1, eps=1e-12):
    """contrastive_aud.py:41-74 with contras_ == anchors_ (extraction_samples returns clones, :142)."""
    lab = labels.view(-1, 1)
    mask = torch.eq(lab, lab.t()).float()
    adc = torch.matmul(anchors, anchors.t()) / temperature
    logits = adc - adc.max(dim=1, keepdim=True)[0].detach()
    neg_mask = 1 - mask
    mask = mask.clone().fill_diagonal_(0.0)
    exp_logits = torch.exp(logits)
    neg = (exp_logits * neg_mask).sum(1, keepdim=True)
    log_prob = logits - torch.log(exp_logits + neg)
    mean_log_prob_pos = (mask * log_prob).sum(1) / (mask.sum(1) + eps)
    return -mean_log_prob_pos.mean()


def contrast_loss(embeds_match, gt_match, embeds_shuffle, gt_shuffle, max_views=512, temperature=0.1):
    """contrastive_aud.py:17-37."""
    # the selection is host-side index logic in the reference (CPU randperm); run it on CPU copies of the labels so the
    # same function also serves embeddings that live on a CUDA device / in fp64 (bench.py gpu_stock_baseline, bs32 tests)
    sel = contrast_select(gt_match.cpu(), gt_shuffle.cpu(), embeds_match.shape[2:], max_views)
    if sel is None:
        return torch.zeros(1, device=embeds_match.device, dtype=embeds_match.dtype)
    half, pix, labels = (t.to(embeds_match.device) for t in sel)
    em = F.normalize(embeds_match, p=2, dim=1).flatten(2).permute(0, 2, 1).reshape(-1, embeds_match.shape[1])
    es = F.normalize(embeds_shuffle, p=2, dim=1).flatten(2).permute(0, 2, 1).reshape(-1, embeds_shuffle.shape[1])
    anchors = torch.where(half.view(-1, 1) == 0, em[pix], es[pix])
    return info_nce(anchors, labels, temperature)


# ---------------------------------------------------------------------------------------------
# A12  the train step body  (trainer/trainer_cavp_vpo_mono.py:142-193, epoch-0 branch)
# ---------------------------------------------------------------------------------------------
def train_step_losses(sd, batch, shuffle_pix_label, *, dilation_flags, audio_kind="vgg", max_views=512):
    """Forward + both losses exactly as the trainer combines them; caller runs .backward()."""
    B = batch["image"].shape[0]
    out_cat, ctr_cat, pack, new_buffers = cavp_forward(sd, batch["image"], batch["audio"],
                                                        dilation_flags=dilation_flags, audio_kind=audio_kind,
                                                        train=True)
    output = out_cat[:B] + out_cat[B:] * 0.0  # trainer:171
    l_ctr = contrast_loss(ctr_cat[:B], batch["pix_label"], ctr_cat[B:], shuffle_pix_label, max_views)
    l_ce = cross_entropy(output, batch["pix_label"])
    return l_ce, l_ctr, out_cat, ctr_cat, pack, new_buffers


def sgd_step(p, g, buf, lr, momentum=0.9, weight_decay=5e-4):
    """torch.optim.SGD (main_vpo_mono.py:118-123): g += wd*p; buf = momentum*buf + g (buf = g first step); p -= lr*buf."""
    g = g + weight_decay * p
    buf = g.clone() if buf is None else momentum * buf + g
    return p - lr * buf, buf


def adam_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (main_vpo_mono.py:125), non-amsgrad, L2 weight decay folded into g."""
    if weight_decay:
        g = g + weight_decay * p
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v
