"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY. Used by oracle/make_golden.py to produce tests/golden/*.pt and by
tests/test_reference_dropin.py (marked `reference`: skipped when /root/reference is absent, i.e. on the GPU box).

The reference needs three non-invasive shims:
  1. `timm` (pinned timm==0.4.9 in /root/reference/requirements.txt:21, not installed here):
     `Mlp`, `DropPath`, `to_2tuple`, `trunc_normal_`, `register_model`, `_cfg`.  Only `Mlp` is executed on
     the hot path (models/attn.py:8,138-143; models/cavp_model.py:16,123-128).  The stand-in restates the
     published timm 0.4.9 definition: fc1 -> act -> drop -> fc2 -> drop, attribute names fc1/act/fc2/drop.
     The reference holds no test pinning this boundary => the Mlp stand-in is "parity unpinned".
  2. `easydict.EasyDict` (attribute dict).
  3. an empty `ckpts/pretrained/resnet50.pth` in the CWD (models/visual/backbones/resnet.py:224 hard-codes it;
     utils/pyt_utils.py:57 loads with strict=False so random init survives).
"""
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


class _Mlp(nn.Module):
    """timm==0.4.9 timm/models/layers/mlp.py:Mlp (published definition, restated)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.drop_prob, "DropPath>0 is never used by the live path (models/attn.py:135)"
        return x


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("/root/reference is not mounted (GPU box?) - goldens are committed under tests/golden")
    sys.dont_write_bytecode = True  # never litter the read-only reference with __pycache__

    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_registry = types.ModuleType("timm.models.registry")
    timm_vit = types.ModuleType("timm.models.vision_transformer")
    timm_layers.Mlp = _Mlp
    timm_layers.DropPath = _DropPath
    timm_layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    timm_layers.trunc_normal_ = lambda t, std=1.0, **kw: nn.init.trunc_normal_(t, std=std)
    timm_registry.register_model = lambda f: f
    timm_vit._cfg = lambda **kw: dict(kw)
    timm.models = timm_models
    timm_models.layers = timm_layers
    timm_models.registry = timm_registry
    timm_models.vision_transformer = timm_vit
    for name, mod in [("timm", timm), ("timm.models", timm_models), ("timm.models.layers", timm_layers),
                      ("timm.models.registry", timm_registry), ("timm.models.vision_transformer", timm_vit)]:
        sys.modules.setdefault(name, mod)

    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules.setdefault("easydict", ed)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # shim 3: empty checkpoint relative to CWD
    work = tempfile.mkdtemp(prefix="cavp_ref_cwd_")
    os.makedirs(os.path.join(work, "ckpts", "pretrained"), exist_ok=True)
    torch.save({}, os.path.join(work, "ckpts", "pretrained", "resnet50.pth"))
    os.chdir(work)
    _installed = True


def make_args(num_classes, dilation, audio_backbone="vgg", batch_size=2):
    return EasyDict(
        seg_model="DeepLabV3Plus",
        last_three_dilation_stride=list(dilation),
        audio_backbone=audio_backbone,
        num_classes=num_classes,
        batch_size=batch_size,
        local_rank="cpu",
    )


def build_reference_cavp(num_classes, dilation, audio_backbone="vgg", in_plane=1, batch_size=2):
    """models/cavp_model.py:70-136 constructed exactly as main_vpo_mono.py:100-107 does."""
    cwd = os.getcwd()
    install_shims()
    try:
        if audio_backbone != "vgg":
            import torchvision
            import models.audio.audio_network as an

            an.resnet18 = lambda *a, **k: torchvision.models.resnet18(weights=None)
        from models.cavp_model import CAVP

        model = CAVP(50, None, num_classes=num_classes, ignore_index=255, audio_backbone_pretrain_path=None,
                     visual_backbone=50, args=make_args(num_classes, dilation, audio_backbone, batch_size),
                     in_plane=in_plane)
    finally:
        pass
    return model


def reference_losses(max_views=512):
    install_shims()
    from loss.contrastive_aud import ContrastLoss

    ce = nn.CrossEntropyLoss(ignore_index=255)  # loss/losser.py:53,60-62
    ctr = ContrastLoss(temperature=0.1, ignore_idx=255, max_views=max_views)  # trainer_cavp_vpo_mono.py:55
    return ce, ctr
