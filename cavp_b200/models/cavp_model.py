"""Drop-in replacement for the reference's `models/cavp_model.py` (CAVP, SoundBank) on B200.

Same constructor, attribute tree, state_dict keys and `forward()` contract as the reference
(models/cavp_model.py:66-205), so `main_vpo_*.py` / `main_avss_resize.py` can import it unchanged:

    from cavp_b200.models.cavp_model import CAVP, SoundBank

The leaf modules (nn.Conv2d / nn.BatchNorm2d / nn.Linear / nn.LayerNorm) are parameter holders with the reference's
names - `engine/utils.py:group_weight` and `SyncBatchNorm.convert_sync_batchnorm` keep working - while all arithmetic
runs in the sm_100a kernels of libcavp_b200.so through `cavp_b200.engine.Graph`.  No torch compute op is on the path.
"""
import os
import warnings
from types import SimpleNamespace

import torch
import torch.nn as nn

from ..engine import (ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, Act, Graph, new_act, pad4)

BN_EPS = 1e-5  # models/visual/deeplabv3/encoder_decoder.py:10
BN_MOMENTUM = 0.1  # encoder_decoder.py:11


# ======================================================================================================================
# parameter-holder module tree (names follow the reference files cited per class)
# ======================================================================================================================
class Bottleneck(nn.Module):
    """models/visual/backbones/resnet.py:53-98 (stride on the 3x3 conv)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=dilation, dilation=dilation,
                               bias=False)
        self.bn2 = nn.BatchNorm2d(planes, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.relu_inplace = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    """Deep-stem ResNet of models/visual/backbones/resnet.py:101-201 (Bottleneck, layers [3,4,6,3] for ResNet-50)."""

    def __init__(self, layers, stem_width=64, replace_stride_with_dilation=None):
        super().__init__()
        self.inplanes = stem_width * 2
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, stem_width, 3, stride=2, padding=1, bias=False),
            nn.BatchNorm2d(stem_width, eps=BN_EPS, momentum=BN_MOMENTUM),
            nn.ReLU(inplace=True),
            nn.Conv2d(stem_width, stem_width, 3, stride=1, padding=1, bias=False),
            nn.BatchNorm2d(stem_width, eps=BN_EPS, momentum=BN_MOMENTUM),
            nn.ReLU(inplace=True),
            nn.Conv2d(stem_width, stem_width * 2, 3, stride=1, padding=1, bias=False),
        )
        if replace_stride_with_dilation is None:
            replace_stride_with_dilation = [False, False, False]
        if len(replace_stride_with_dilation) != 3:
            raise ValueError("replace_stride_with_dilation should be None or a 3-element tuple, got {}".format(
                replace_stride_with_dilation))
        self.dilation = 1
        self.bn1 = nn.BatchNorm2d(stem_width * 2, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2, dilate=replace_stride_with_dilation[0])
        self.layer3 = self._make_layer(256, layers[2], stride=2, dilate=replace_stride_with_dilation[1])
        self.layer4 = self._make_layer(512, layers[3], stride=2, dilate=replace_stride_with_dilation[2])

    def _make_layer(self, planes, blocks, stride=1, dilate=False):
        previous_dilation = self.dilation
        downsample = None
        if dilate:
            self.dilation *= stride
            stride = 1
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * 4, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * 4, eps=BN_EPS, momentum=BN_MOMENTUM),
            )
        layers = [Bottleneck(self.inplanes, planes, stride, downsample, dilation=previous_dilation)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            layers.append(Bottleneck(self.inplanes, planes, dilation=self.dilation))
        return nn.Sequential(*layers)


class Backbone(nn.Module):
    """models/visual/deeplabv3/encoder_decoder.py:14-59 (layer4 re-dilated 2, 4, 8; its strides removed)."""

    def __init__(self, back_bone, pretrained_model=None, last_three_dilation_stride=None):
        super().__init__()
        if back_bone == 50:
            self.backbone = ResNet([3, 4, 6, 3], replace_stride_with_dilation=last_three_dilation_stride)
        elif back_bone == 101:
            self.backbone = ResNet([3, 4, 23, 3], replace_stride_with_dilation=last_three_dilation_stride)
        else:
            raise ValueError
        # resnet.py:224 hard-codes this path relative to the CWD and loads with strict=False (utils/pyt_utils.py:57)
        for path in ("ckpts/pretrained/resnet%d.pth" % back_bone, pretrained_model):
            if isinstance(path, str) and os.path.isfile(path):
                sd = torch.load(path, map_location="cpu")
                if "model" in sd:
                    sd = sd["model"]
                self.backbone.load_state_dict(sd, strict=False)
                break
        self.dilate = 2
        for m in self.backbone.layer4.children():
            for sub in m.modules():
                self._nostride_dilate(sub, self.dilate)
            self.dilate *= 2

    @staticmethod
    def _nostride_dilate(m, dilate):
        if isinstance(m, nn.Conv2d):
            if m.stride == (2, 2):
                m.stride = (1, 1)
            if m.kernel_size == (3, 3):
                m.dilation = (dilate, dilate)
                m.padding = (dilate, dilate)


class ASPP(nn.Module):
    """encoder_decoder.py:108-164"""

    def __init__(self, in_channels, out_channels, dilation_rates=(6, 12, 18), hidden_channels=256):
        super().__init__()
        self.map_convs = nn.ModuleList([
            nn.Conv2d(in_channels, hidden_channels, 1, bias=False),
            nn.Conv2d(in_channels, hidden_channels, 3, bias=False, dilation=dilation_rates[0], padding=dilation_rates[0]),
            nn.Conv2d(in_channels, hidden_channels, 3, bias=False, dilation=dilation_rates[1], padding=dilation_rates[1]),
            nn.Conv2d(in_channels, hidden_channels, 3, bias=False, dilation=dilation_rates[2], padding=dilation_rates[2]),
        ])
        self.map_bn = nn.BatchNorm2d(hidden_channels * 4)
        self.global_pooling_conv = nn.Conv2d(in_channels, hidden_channels, 1, bias=False)
        self.global_pooling_bn = nn.BatchNorm2d(hidden_channels)
        self.red_conv = nn.Conv2d(hidden_channels * 4, out_channels, 1, bias=False)
        self.pool_red_conv = nn.Conv2d(hidden_channels, out_channels, 1, bias=False)
        self.red_bn = nn.BatchNorm2d(out_channels)
        self.leak_relu = nn.LeakyReLU()


class Upsampling(nn.Module):
    """encoder_decoder.py:62-75"""

    def __init__(self, classifier_in_channels, num_classes, conv_in):
        super().__init__()
        self.classifier = nn.Conv2d(classifier_in_channels, num_classes, kernel_size=1, bias=True)
        self.last_conv = nn.Sequential(
            nn.Conv2d(conv_in, 256, kernel_size=3, stride=1, padding=1, bias=False),
            nn.BatchNorm2d(256, momentum=BN_MOMENTUM), nn.ReLU(),
            nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False),
            nn.BatchNorm2d(256, momentum=BN_MOMENTUM), nn.ReLU())


class DeepLabV3Plus(nn.Module):
    """encoder_decoder.py:78-106"""

    def __init__(self, num_classes, aspp_in_plane=2048, aspp_out_plane=256, classifier_in=256):
        super().__init__()
        conv_in = 112 if aspp_out_plane == 64 else 304
        self.aspp = ASPP(aspp_in_plane, aspp_out_plane, (6, 12, 18))
        self.reduce = nn.Sequential(nn.Conv2d(aspp_out_plane, 48, 1, bias=False),
                                    nn.BatchNorm2d(48, momentum=BN_MOMENTUM), nn.ReLU())
        self.upsample = Upsampling(classifier_in, num_classes, conv_in=conv_in)
        self.business_layer = [self.aspp, self.reduce, self.upsample.last_conv, self.upsample.classifier]


class Mlp(nn.Module):
    """timm==0.4.9 timm/models/layers/mlp.py:Mlp (the reference's dependency; attribute names fc1/act/fc2/drop)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class PatchEmbed(nn.Module):
    """models/attn.py:17-39"""

    def __init__(self, img_size, dim_in, embed_dim):
        super().__init__()
        self.img_size = img_size
        self.proj = nn.Linear(dim_in, embed_dim)
        self.num_patches = img_size[0] * img_size[1]
        self.norm = nn.Identity()


class Attention(nn.Module):
    """models/attn.py:41-63"""

    def __init__(self, dim, num_heads=8):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=False)
        self.k = nn.Linear(dim, dim, bias=False)
        self.v = nn.Linear(dim, dim, bias=False)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)


class Block(nn.Module):
    """models/attn.py:109-144"""

    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=nn.GELU, drop=0.0)


class CROSS_ATTENTION(nn.Module):
    """models/attn.py:185-244 (depth Blocks in "CA" mode; pos_embed_* exist but are never added, :235-236)."""

    def __init__(self, embed_dim=768, depth=2, num_heads=4, mlp_ratio=4.0, dim_in=1280):
        super().__init__()
        self.patch_embed_v = PatchEmbed((128, 128), dim_in, embed_dim)
        self.patch_embed_a = PatchEmbed((1, 1), dim_in, embed_dim)
        self.pos_embed_v = nn.Parameter(torch.zeros(1, self.patch_embed_v.num_patches, embed_dim))
        self.pos_embed_a = nn.Parameter(torch.zeros(1, self.patch_embed_a.num_patches, embed_dim))
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)


class VGG(nn.Module):
    """models/audio/backbones/vgg.py:5-36"""

    def __init__(self, out_plane_):
        super().__init__()
        layers, in_channels = [], 1
        for v in [64, "M", 128, "M", 256, 256, "M", 512, 512, "M"]:
            if v == "M":
                layers += [nn.MaxPool2d(kernel_size=2, stride=2)]
            else:
                layers += [nn.Conv2d(in_channels, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                in_channels = v
        self.features = nn.Sequential(*layers)
        self.embeddings = nn.Sequential(nn.Linear(512 * 4 * 6, 4096), nn.ReLU(True), nn.Linear(4096, 4096),
                                        nn.ReLU(True), nn.Linear(4096, out_plane_), nn.ReLU(True))


class AudioModel(nn.Module):
    """models/audio/audio_network.py:9-45"""

    def __init__(self, backbone, pretrain_path, out_plane, num_classes=2, in_plane=1):
        super().__init__()
        if backbone == "vgg":
            self.backbone = VGG(out_plane)
            if pretrain_path is not None:
                self.load_audio_model(pretrain_path)
        else:
            from torchvision.models import resnet18
            try:
                self.backbone = resnet18(True)  # audio_network.py:19 (ImageNet weights; needs the torch hub cache)
            except Exception as e:  # offline: random init
                warnings.warn(f"pretrained ResNet-18 unavailable ({type(e).__name__}); using random init")
                self.backbone = resnet18(weights=None)
            self.backbone.conv1 = nn.Conv2d(in_plane, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
            self.backbone.avgpool = nn.AdaptiveMaxPool2d((1, 1))
            self.backbone.fc = nn.Linear(512, out_plane)
        self.cls_head = nn.Linear(out_plane, num_classes)

    def load_audio_model(self, path_):
        param_dict = torch.load(path_, map_location="cpu")
        out_, in_ = self.backbone.state_dict()["embeddings.4.weight"].shape
        param_dict["embeddings.4.weight"] = torch.nn.init.kaiming_normal_(torch.zeros(out_, in_))
        param_dict["embeddings.4.bias"] = torch.zeros(out_)
        self.backbone.load_state_dict(param_dict, strict=True)


class SoundBank:
    """models/cavp_model.py:21-52 - per-class FIFO of audio features; host-side bookkeeping, kept in Python."""

    def __init__(self, out_dim=304, args=None, device=0):
        self.bank_vault = torch.zeros((args.num_classes, args.batch_size, out_dim), requires_grad=False, device=device)

    def update_bank(self, waveform, img_label):
        img_label[:, 0] = 0
        target = [item.nonzero().squeeze().cpu().view(-1, ).tolist() for item in img_label]
        for i in range(len(target)):
            item = target[i]
            if len(item) != 1:
                continue
            tmp_waveform = waveform[i, None] if len(waveform.shape) == 2 else waveform[i]
            self.queue(item[0], tmp_waveform)

    def queue(self, class_idx, fea_a):
        self.bank_vault[class_idx] = torch.cat((self.bank_vault[class_idx][1:], fea_a.detach()), dim=0)

    def overwrite_audio_feature(self, shuffle_fea_a, org_fea_a, mod_idx_map):
        for i, (idx, target_label) in enumerate(mod_idx_map.items()):
            fake_audio = self.bank_vault[None, target_label][:, 0]
            shuffle_fea_a[idx] = fake_audio
        return shuffle_fea_a


# ======================================================================================================================
# forward graph (all arithmetic = C-ABI kernels via Graph)
# ======================================================================================================================
def _input_nhwc(g, t, cpad):
    """NCHW torch tensor -> zero-padded NHWC Act (the model boundary)."""
    n, c, h, w = t.shape
    t = t.contiguous().float()
    a = new_act(n, h, w, cpad, g.device, needs_grad=False)
    g.call("cavp_nchw_to_nhwc", t.data_ptr(), a.ptr, n, c, h * w, cpad)
    return a


def _cv(conv):
    return dict(stride=conv.stride[0], pad=conv.padding[0], dil=conv.dilation[0])


def backbone_forward(g, net, x):
    """ResNet.forward (resnet.py:186-201): returns the 4 stage outputs."""
    c1 = net.conv1
    x = g.conv_bn(x, c1[0].weight, c1[1], **_cv(c1[0]), name="stem0")
    x = g.conv_bn(x, c1[3].weight, c1[4], **_cv(c1[3]), name="stem1")
    x = g.conv_bn(x, c1[6].weight, net.bn1, **_cv(c1[6]), name="stem2")
    x = g.maxpool(x, 3, 2, 1)
    feats = []
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        # layer4 / layer3 hold 63 % / 30 % of the ResNet's parameters and their backward runs first: each one's gradient
        # bucket is complete when the tape is back at its marker (cavp_b200/parallel.py:cavp_buckets)
        if layer is net.layer4:
            g.mark("layer4_grads_done")
        elif layer is net.layer3:
            g.mark("layer3_grads_done")
        for blk in layer:
            out = g.conv_bn(x, blk.conv1.weight, blk.bn1, **_cv(blk.conv1))
            out = g.conv_bn(out, blk.conv2.weight, blk.bn2, **_cv(blk.conv2))
            if blk.downsample is not None:
                res = g.conv_bn(x, blk.downsample[0].weight, blk.downsample[1], act=ACT_NONE, **_cv(blk.downsample[0]))
            else:
                res = x
            x = g.conv_bn(out, blk.conv3.weight, blk.bn3, res=res, **_cv(blk.conv3))  # relu(bn3 + residual)
        feats.append(x)
    return feats


def aspp_forward(g, aspp, x):
    """ASPP.forward (encoder_decoder.py:137-156)."""
    hid = aspp.map_convs[0].weight.shape[0]
    mc = new_act(x.n, x.h, x.w, 4 * hid, g.device)
    if g.train:
        nparts = ((x.rows + 127) // 128) * 4
        stats_t = g.empty(nparts, 2, 4 * hid)
        for i, conv in enumerate(aspp.map_convs):
            g.conv(x, conv.weight, want_stats=True, out=mc.slice(hid * i, hid),
                   stats_buf=(stats_t, stats_t.data_ptr() + 4 * hid * i, nparts, 4 * hid), **_cv(conv))
        out = g.bn_act(mc, (stats_t, stats_t.data_ptr(), nparts, 4 * hid), aspp.map_bn, act=ACT_LEAKY)
    else:
        for i, conv in enumerate(aspp.map_convs):
            g.conv(x, conv.weight, out=mc.slice(hid * i, hid), **_cv(conv))
        out = g.bn_eval(mc, aspp.map_bn, act=ACT_LEAKY)
    pool = g.global_avgpool(x)  # _global_pooling (:158-164)
    pool = g.conv_bn(pool, aspp.global_pooling_conv.weight, aspp.global_pooling_bn, act=ACT_LEAKY)
    pool, _ = g.conv(pool, aspp.pool_red_conv.weight)
    hw = x.h * x.w
    if g.train:
        out, stats = g.conv(out, aspp.red_conv.weight, want_stats=True, res=pool, res_div=hw)  # out += pool (:153)
        return g.bn_act(out, stats, aspp.red_bn, act=ACT_LEAKY)
    out, _ = g.conv(out, aspp.red_conv.weight, res=pool, res_div=hw)
    return g.bn_eval(out, aspp.red_bn, act=ACT_LEAKY)


def forward_feature(g, seg, feats):
    """DeepLabV3Plus.forward_feature (encoder_decoder.py:97-105): cat(upsampled ASPP, reduced low-level) -> 304 ch."""
    f = aspp_forward(g, seg.aspp, feats[-1])
    low = feats[0]
    c_aspp, c_low = f.c, seg.reduce[0].weight.shape[0]
    fea = new_act(low.n, low.h, low.w, c_aspp + c_low, g.device)
    g.bilinear(f, low.h, low.w, True, out=fea.slice(0, c_aspp))
    g.conv_bn(low, seg.reduce[0].weight, seg.reduce[1], act=ACT_RELU, out=fea.slice(c_aspp, c_low))
    return fea


def vgg_forward(g, vgg, x):
    """VGG.forward (vgg.py:17-23); the NCHW->NHWC transposes before the flatten are free in this layout."""
    for m in vgg.features:
        if isinstance(m, nn.Conv2d):
            x, _ = g.conv(x, m.weight, bias=m.bias, act=ACT_RELU, **_cv(m))
        elif isinstance(m, nn.MaxPool2d):
            x = g.maxpool(x, 2, 2, 0)
    x = g.flatten(x)
    for i in (0, 2, 4):
        lin = vgg.embeddings[i]
        x, _ = g.conv(x, lin.weight, bias=lin.bias, act=ACT_RELU)
    return x


def resnet18_audio_forward(g, net, x):
    """torchvision resnet18 with the replacements of audio_network.py:19-25."""
    x = g.conv_bn(x, net.conv1.weight, net.bn1, **_cv(net.conv1))
    x = g.maxpool(x, 3, 2, 1)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            out = g.conv_bn(x, blk.conv1.weight, blk.bn1, **_cv(blk.conv1))
            if blk.downsample is not None:
                res = g.conv_bn(x, blk.downsample[0].weight, blk.downsample[1], act=ACT_NONE, **_cv(blk.downsample[0]))
            else:
                res = x
            x = g.conv_bn(out, blk.conv2.weight, blk.bn2, res=res, **_cv(blk.conv2))
    x = g.global_maxpool(x)
    x, _ = g.conv(x, net.fc.weight, bias=net.fc.bias)
    return x


def fusion_forward(g, model, fea_v, fea_a):
    """CAVP.forward_fusion (cavp_model.py:143-154) + CROSS_ATTENTION.forward (attn.py:232-244) with depth-1
    Block.forward_ca (attn.py:146-162).  `fea_v` holds the B distinct visual rows; in train mode the reference
    duplicates them (cavp_model.py:181) - the duplicate half is bit-identical up to the gate, so it is computed once
    and shared (rep = rows / B).  The audio<-visual branch of forward_ca (attn.py:161) only feeds the discarded f_a
    (cavp_model.py:151) and is skipped."""
    vp, ca = model.visual_projector, model.cross_att
    blk = ca.blocks[0]
    Bq = fea_v.n
    h1, _ = g.conv(fea_v, vp.fc1.weight, bias=vp.fc1.bias, act=ACT_GELU, save_pre=True)
    proj, _ = g.conv(h1, vp.fc2.weight, bias=vp.fc2.bias)                      # pack["visual"]
    pe, _ = g.conv(proj, ca.patch_embed_v.proj.weight, bias=ca.patch_embed_v.proj.bias)
    fa, _ = g.conv(fea_a, ca.patch_embed_a.proj.weight, bias=ca.patch_embed_a.proj.bias)
    fvn = g.layernorm(pe, blk.norm1)
    fan = g.layernorm(fa, blk.norm1)
    q, _ = g.conv(fvn, blk.attn.q.weight)
    k, _ = g.conv(fan, blk.attn.k.weight)
    v, _ = g.conv(fan, blk.attn.v.weight)
    x, attn = g.gate(q, k, v, heads=blk.attn.num_heads)
    rows = x.n
    # f_v = norm1(f_v) + proj(attn @ v)   (attn.py:148-149: the residual is the *normed* f_v)
    f1, _ = g.conv(x, blk.attn.proj.weight, bias=blk.attn.proj.bias, res=fvn, res_mod=fvn.rows if rows != Bq else 0)
    n2 = g.layernorm(f1, blk.norm2)
    m1, _ = g.conv(n2, blk.mlp.fc1.weight, bias=blk.mlp.fc1.bias, act=ACT_GELU, save_pre=True)
    f2, _ = g.conv(m1, blk.mlp.fc2.weight, bias=blk.mlp.fc2.bias, res=f1)
    out = g.layernorm(f2, ca.norm)
    return out, proj, attn


def decoder_forward(g, up, fusion, num_classes):
    """Upsampling.forward (encoder_decoder.py:73-75); the classifier is padded to a multiple of 4 output channels."""
    lc = up.last_conv
    x = g.conv_bn(fusion, lc[0].weight, lc[1], **_cv(lc[0]))
    x = g.conv_bn(x, lc[3].weight, lc[4], **_cv(lc[3]))
    logits, _ = g.conv(x, up.classifier.weight, bias=up.classifier.bias, pad_cout=pad4(num_classes))
    return logits


class CAVP(nn.Module):
    """models/cavp_model.py:66-205.  Only `seg_model == "DeepLabV3Plus"` (the north-star path) is implemented."""

    def __init__(self, backbone, pretrain_path, num_classes=2, ignore_index=255, audio_backbone_pretrain_path=None,
                 visual_backbone=50, args=None, in_plane=1):
        super().__init__()
        seg_model = args.seg_model
        last_three_dilation_stride = args.last_three_dilation_stride
        if seg_model == "DeepLabV3Plus":
            self.latent_dim = 304
            self.backbone = Backbone(back_bone=backbone, pretrained_model=pretrain_path,
                                     last_three_dilation_stride=last_three_dilation_stride)
            self.segment = DeepLabV3Plus(num_classes=num_classes, aspp_in_plane=2048, aspp_out_plane=256)
        elif seg_model in ("HRNet", "OCR", "PVT"):
            raise NotImplementedError(f"seg_model {seg_model} is outside the B200 hot-path scope (DESIGN.md)")
        else:
            raise ValueError("UNKNOW BACKBONE")
        self.cross_att = CROSS_ATTENTION(dim_in=self.latent_dim, embed_dim=self.latent_dim, depth=1)
        self.visual_projector = Mlp(in_features=self.latent_dim, hidden_features=256, out_features=self.latent_dim,
                                    drop=0.0)
        self.audio_backbone = AudioModel(args.audio_backbone, audio_backbone_pretrain_path, self.latent_dim,
                                         in_plane=in_plane)
        self.memory = SoundBank(out_dim=self.latent_dim, args=args, device=args.local_rank)
        self.local_rank = args.local_rank
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.in_plane = in_plane
        self.audio_kind = "vgg" if args.audio_backbone == "vgg" else "resnet18"
        # 2 = fp32-parity (3xTF32 + promotion), 1 = plain TF32, 3 = bf16 operands / fp32 accumulate (BASELINE configs[2])
        self.prec = int(getattr(args, "cavp_prec", 2))
        if self.prec not in (1, 2, 3):
            raise ValueError("cavp_prec must be 1 (TF32), 2 (fp32 parity) or 3 (bf16 operands)")
        self._to_channels_last()

    # ---- parameter storage: conv weights live in OHWI (channels_last) so the kernels read them without copies
    def _to_channels_last(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and m.weight.shape[1] % 4 == 0:
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._to_channels_last()
        return out

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._to_channels_last()
        return out

    # ---- graph construction
    def build_graph(self, g, image, audio, *, shuffle_idx=None, audio_func=False):
        """Runs the forward kernels; returns (logits Act [rows,h,w,pad4(nc)], fusion Act, proj Act, fea_a Act, attn)."""
        train = g.train
        x = _input_nhwc(g, image, 4)
        feats = backbone_forward(g, self.backbone.backbone, x)
        # tape markers for the bucketed gradient all-reduce (cavp_b200/parallel.py:cavp_buckets): the backward pass runs
        # decoder -> fusion -> audio backbone -> DeepLab head -> ResNet, so the audio bucket is complete when the tape
        # is back at "audio_grads_done" and everything but the ResNet at "head_grads_done"
        g.mark("head_grads_done")
        fea_v = forward_feature(g, self.segment, feats)
        g.mark("audio_grads_done")
        a = _input_nhwc(g, audio, pad4(audio.shape[1]))
        if self.audio_kind == "vgg":
            fea_a = vgg_forward(g, self.audio_backbone.backbone, a)
        else:
            fea_a = resnet18_audio_forward(g, self.audio_backbone.backbone, a)
        if train and audio_func:
            fea_a = g.concat_rows(fea_a, g.gather_rows(fea_a, shuffle_idx))  # forward_audio (:156-173)
        fusion, proj, attn = fusion_forward(g, self, fea_v, fea_a)
        logits = decoder_forward(g, self.segment.upsample, fusion, self.num_classes)
        return logits, fusion, proj, fea_a, attn

    def _sync_group(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and any(
                isinstance(m, nn.SyncBatchNorm) for m in self.modules()):
            return dist.group.WORLD
        return None

    # ---- reference API
    def forward_audio_bank(self, fea_a_t, shuffle_info, ow_flag):
        """SoundBank side effects of forward_audio (:156-173); the returned features are built in build_graph."""
        if ow_flag:
            shuffle_fea_a = fea_a_t.clone().detach()[shuffle_info["shuffle_idx"]]
            self.memory.overwrite_audio_feature(shuffle_fea_a, fea_a_t, shuffle_info["mod_idx_map"])
            self.memory.update_bank(fea_a_t, shuffle_info["image_label"])

    def forward_train(self, image, audio=None, shuffle_info=None, ow_flag=False, audio_func=False):
        params = [p for p in self.parameters()]
        shuffle_idx = None
        if audio_func:
            shuffle_idx = shuffle_info["shuffle_idx"].to(image.device, torch.int64).contiguous()
        out = _CAVPFunction.apply(self, shuffle_idx, bool(audio_func), image, audio, *params)
        out_pred, out_fusion, visual, fea_a, attn_v = out
        if audio_func and ow_flag:
            self.forward_audio_bank(fea_a[:image.shape[0], :, 0, 0], shuffle_info, ow_flag)
        return out_pred, out_fusion, {"audio": fea_a, "visual": visual, "attn_v": attn_v}

    def forward_inference(self, image, audio=None):
        if not image.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        with torch.no_grad():
            g = Graph(image.device, prec=self.prec, train=False)
            g.use_weight_cache(self)
            logits, fusion, proj, fea_a, attn = self.build_graph(g, image, audio)
            pred = g.upsample_to_nchw(logits, self.num_classes, image.shape[-2], image.shape[-1])
            self.last_launches = g.launches
        return pred, fusion.nchw(), _pack(fusion.n, proj, fea_a, attn)

    def forward_eval_metrics(self, image, audio, target, conf=None, ignore_index=None, want_pred=True):
        """forward_inference + the metric epilogue of trainer.validation (trainer_cavp_vpo_mono.py:272-277:
        `MIoU(logits, label)`, `ForegroundDetect(logits, label)`) without ever writing the full-resolution logits: the
        bilinear upsample of forward_cls, the argmax and the (label, prediction) confusion counts run in one kernel.
        Returns (pred int64 [B,H,W] or None, conf int64 [(nc+1), nc]); pass `conf` back in to accumulate over a
        validation epoch and hand it to cavp_b200.metrics.MIoU / ForegroundDetect `.update_from_confusion`."""
        if not image.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        nc = self.num_classes
        with torch.no_grad():
            g = Graph(image.device, prec=self.prec, train=False)
            g.use_weight_cache(self)
            logits, fusion, proj, fea_a, attn = self.build_graph(g, image, audio)
            H, W = image.shape[-2:]
            if conf is None:
                conf = torch.zeros(nc + 1, nc, dtype=torch.int64, device=image.device)
            labels = target.reshape(logits.n, H, W).to(image.device, torch.int64).contiguous()
            pred = torch.empty(logits.n, H, W, dtype=torch.int64, device=image.device) if want_pred else None
            g.call("cavp_upsample_argmax_confusion", logits.ptr, logits.ld, logits.h, logits.w, H, W, logits.n, nc,
                   labels.data_ptr(), self.ignore_index if ignore_index is None else ignore_index,
                   0 if pred is None else pred.data_ptr(), conf.data_ptr())
            self.last_launches = g.launches
        return pred, conf

    def forward(self, image, audio=None, shuffle_info=None, ow_flag=False, eval_mode=False, audio_func=False):
        if eval_mode:
            return self.forward_inference(image, audio)
        return self.forward_train(image, audio, shuffle_info, ow_flag, audio_func=audio_func)


def _pack(rows, proj, fea_a, attn):
    visual = proj.nchw()
    if proj.n != rows:  # train mode: the reference returns the duplicated tensor (cavp_model.py:181,148)
        visual = visual.repeat(rows // proj.n, 1, 1, 1)  # plumbing copy of an auxiliary output
    return {"audio": fea_a.dense().reshape(fea_a.rows, fea_a.c, 1, 1), "visual": visual,
            "attn_v": attn.unsqueeze(-1)}


class _CAVPFunction(torch.autograd.Function):
    """Boundary between torch autograd and the kernel graph: one node for the whole model."""

    @staticmethod
    def forward(ctx, model, shuffle_idx, audio_func, image, audio, *params):
        if not image.is_cuda:
            raise RuntimeError("cavp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        g = Graph(image.device, prec=model.prec, train=model.training, sync_bn_group=model._sync_group())
        g.use_weight_cache(model)
        if not model.training:
            raise RuntimeError("forward_train called in eval mode; use eval_mode=True (reference: forward_inference)")
        logits, fusion, proj, fea_a, attn = model.build_graph(g, image, audio, shuffle_idx=shuffle_idx,
                                                              audio_func=audio_func)
        pred = g.upsample_to_nchw(logits, model.num_classes, image.shape[-2], image.shape[-1])
        ctx.g, ctx.logits, ctx.fusion, ctx.model = g, logits, fusion, model
        ctx.params = params
        pack = _pack(fusion.n, proj, fea_a, attn)
        outs = (pred, fusion.nchw(), pack["visual"], pack["audio"], pack["attn_v"])
        ctx.mark_non_differentiable(outs[2], outs[3], outs[4])
        model.last_launches = g.launches
        return outs

    @staticmethod
    def backward(ctx, dpred, dfusion, *unused):
        g, logits, fusion, model = ctx.g, ctx.logits, ctx.fusion, ctx.model
        if dfusion is not None:
            seed_fusion_grad(g, fusion, dfusion)
        if dpred is not None:
            g.upsample_to_nchw_backward(logits, model.num_classes, dpred.contiguous())
        g.backward()
        grads = tuple(g.param_grads.get(id(p)) for p in ctx.params)
        model.last_launches = g.launches
        ctx.g = None
        return (None, None, None, None, None) + grads


def seed_fusion_grad(g, fusion, dfusion):
    """Copy an incoming d(out_fusion) (logical NCHW, any strides) into the graph's own NHWC gradient buffer."""
    dst, accumulate = g.grad_target(fusion)
    nhwc = dfusion.permute(0, 2, 3, 1)
    if nhwc.is_contiguous():
        src = Act(nhwc.reshape(-1, fusion.c), fusion.n, fusion.h, fusion.w, fusion.c)
        (g.add_act if accumulate else g.copy_act)(dst, src)
    else:
        d = dfusion.contiguous()
        tmp = dst if not accumulate else new_act(fusion.n, fusion.h, fusion.w, fusion.c, g.device)
        g.call("cavp_nchw_to_nhwc", d.data_ptr(), tmp.ptr, fusion.n, fusion.c, fusion.h * fusion.w, fusion.c)
        if accumulate:
            g.add_act(dst, tmp)
