"""state_dict schema (key -> shape) of the reference CAVP(DeepLabV3Plus, ResNet-50) model.

TEST INFRASTRUCTURE ONLY.  Restated from the reference constructors so that the oracle can build its
state without importing either the reference or the product:
  models/visual/backbones/resnet.py:101-184,221-227   (deep-stem ResNet-50, Bottleneck)
  models/visual/deeplabv3/encoder_decoder.py:62-135   (Upsampling, DeepLabV3Plus, ASPP)
  models/attn.py:17-63,109-230                          (PatchEmbed, Attention, Block, CROSS_ATTENTION)
  models/audio/audio_network.py:9-28, backbones/vgg.py:5-36, torchvision resnet18
  models/cavp_model.py:84-136                           (module names)
"""
from collections import OrderedDict

import torch

from . import seeded


class _Schema:
    def __init__(self):
        self.shapes = OrderedDict()
        self.norm_prefixes = set()

    def conv(self, name, cout, cin, k, bias=False):
        self.shapes[name + ".weight"] = (cout, cin, k, k)
        if bias:
            self.shapes[name + ".bias"] = (cout,)

    def linear(self, name, cout, cin, bias=True):
        self.shapes[name + ".weight"] = (cout, cin)
        if bias:
            self.shapes[name + ".bias"] = (cout,)

    def bn(self, name, c):
        self.norm_prefixes.add(name)
        self.shapes[name + ".weight"] = (c,)
        self.shapes[name + ".bias"] = (c,)
        self.shapes[name + ".running_mean"] = (c,)
        self.shapes[name + ".running_var"] = (c,)
        self.shapes[name + ".num_batches_tracked"] = ()

    def ln(self, name, c):
        self.norm_prefixes.add(name)
        self.shapes[name + ".weight"] = (c,)
        self.shapes[name + ".bias"] = (c,)


def cavp_schema(num_classes, audio="vgg", in_plane=1):
    s = _Schema()
    p = "backbone.backbone."
    s.conv(p + "conv1.0", 64, 3, 3)
    s.bn(p + "conv1.1", 64)
    s.conv(p + "conv1.3", 64, 64, 3)
    s.bn(p + "conv1.4", 64)
    s.conv(p + "conv1.6", 128, 64, 3)
    s.bn(p + "bn1", 128)
    inpl = 128
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
        for bi in range(n):
            b = f"{p}layer{li + 1}.{bi}."
            s.conv(b + "conv1", planes, inpl, 1)
            s.bn(b + "bn1", planes)
            s.conv(b + "conv2", planes, planes, 3)
            s.bn(b + "bn2", planes)
            s.conv(b + "conv3", planes * 4, planes, 1)
            s.bn(b + "bn3", planes * 4)
            if bi == 0:
                s.conv(b + "downsample.0", planes * 4, inpl, 1)
                s.bn(b + "downsample.1", planes * 4)
            inpl = planes * 4
    g = "segment."
    # registration order in the reference: aspp, reduce, upsample(classifier, last_conv)
    s.conv(g + "aspp.map_convs.0", 256, 2048, 1)
    for i in (1, 2, 3):
        s.conv(g + f"aspp.map_convs.{i}", 256, 2048, 3)
    s.bn(g + "aspp.map_bn", 1024)
    s.conv(g + "aspp.global_pooling_conv", 256, 2048, 1)
    s.bn(g + "aspp.global_pooling_bn", 256)
    s.conv(g + "aspp.red_conv", 256, 1024, 1)
    s.conv(g + "aspp.pool_red_conv", 256, 256, 1)
    s.bn(g + "aspp.red_bn", 256)
    s.conv(g + "reduce.0", 48, 256, 1)
    s.bn(g + "reduce.1", 48)
    s.conv(g + "upsample.classifier", num_classes, 256, 1, bias=True)
    s.conv(g + "upsample.last_conv.0", 256, 304, 3)
    s.bn(g + "upsample.last_conv.1", 256)
    s.conv(g + "upsample.last_conv.3", 256, 256, 3)
    s.bn(g + "upsample.last_conv.4", 256)
    c = "cross_att."
    s.shapes[c + "pos_embed_v"] = (1, 128 * 128, 304)
    s.shapes[c + "pos_embed_a"] = (1, 1, 304)
    s.linear(c + "patch_embed_v.proj", 304, 304)
    s.linear(c + "patch_embed_a.proj", 304, 304)
    b = c + "blocks.0."
    s.ln(b + "norm1", 304)
    for n in ("q", "k", "v"):
        s.linear(b + "attn." + n, 304, 304, bias=False)
    s.linear(b + "attn.proj", 304, 304)
    s.ln(b + "norm2", 304)
    s.linear(b + "mlp.fc1", 1216, 304)
    s.linear(b + "mlp.fc2", 304, 1216)
    s.ln(c + "norm", 304)
    s.linear("visual_projector.fc1", 256, 304)
    s.linear("visual_projector.fc2", 304, 256)
    a = "audio_backbone.backbone."
    if audio == "vgg":
        cin, idx = 1, 0
        for v in (64, "M", 128, "M", 256, 256, "M", 512, 512, "M"):
            if v == "M":
                idx += 1
            else:
                s.conv(a + f"features.{idx}", v, cin, 3, bias=True)
                cin = v
                idx += 2
        s.linear(a + "embeddings.0", 4096, 512 * 4 * 6)
        s.linear(a + "embeddings.2", 4096, 4096)
        s.linear(a + "embeddings.4", 304, 4096)
    else:
        s.conv(a + "conv1", 64, in_plane, 7)
        s.bn(a + "bn1", 64)
        inpl = 64
        for li, planes in enumerate((64, 128, 256, 512)):
            for bi in range(2):
                b = f"{a}layer{li + 1}.{bi}."
                s.conv(b + "conv1", planes, inpl if bi == 0 else planes, 3)
                s.bn(b + "bn1", planes)
                s.conv(b + "conv2", planes, planes, 3)
                s.bn(b + "bn2", planes)
                if bi == 0 and li > 0:
                    s.conv(b + "downsample.0", planes, inpl, 1)
                    s.bn(b + "downsample.1", planes)
            inpl = planes
        s.linear(a + "fc", 304, 512)
    s.linear("audio_backbone.cls_head", 2, 304)
    return s


def seeded_state(num_classes, audio="vgg", in_plane=1, seed=0, requires_grad=False):
    """Same values oracle/seeded.fill_module_ writes into a real module with these keys."""
    sch = cavp_schema(num_classes, audio, in_plane)
    sd = OrderedDict()
    for key, shape in sch.shapes.items():
        dummy = torch.empty(shape)
        kind = seeded.classify(key, dummy, sch.norm_prefixes)
        if kind == "zero":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
        else:
            t = seeded.seeded_tensor(key, shape, seed, kind)
            if requires_grad:
                t.requires_grad_(True)
            sd[key] = t
    return sd
