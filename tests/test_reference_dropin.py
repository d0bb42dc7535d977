"""Reference-side drop-in proof (CPU, build container only: needs /root/reference).

The cavp_b200 CAVP module is handed to the reference's OWN, unmodified host code:

  * `engine.utils.group_weight` (engine/utils.py:642-688) + `set_group_lr` (main_vpo_mono.py:45-65, executed from the
    reference source text - the entry script itself cannot be imported: its trainers need modules the repo lacks)
    must classify every parameter and produce the 12 SGD groups the trainers index positionally
    (trainer_cavp_vpo_mono.py:75-83), identical in sizes to what they produce for the reference model;
  * `torch.optim.SGD(param_lists_v, ...)` / `torch.optim.Adam(model_a.parameters())` (main_vpo_mono.py:118-125);
  * `nn.SyncBatchNorm.convert_sync_batchnorm` (main_vpo_mono.py:130,138) keeps names, parameters and buffers;
  * a released-checkpoint-style state dict (`module.` prefix, main_vpo_mono.py DDP save format) round-trips through
    `load_state_dict(strict=False)` after stripping the prefix, onto both the plain and the converted model.
"""
import ast
import sys
import types
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.reference


def _reference_group_fns():
    from oracle import ref_harness
    ref_harness.install_shims()
    for name in ("matplotlib", "matplotlib.pyplot", "terminaltables"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                m = types.ModuleType(name)
                m.AsciiTable = object
                sys.modules[name] = m
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from engine.utils import group_weight  # the reference's own function object
    src = open(ref_harness.REFERENCE_ROOT + "/main_vpo_mono.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "set_group_lr")
    ns = {"group_weight": group_weight, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "main_vpo_mono.py", "exec"), ns)
    return group_weight, ns["set_group_lr"]


def _ours(nc=22, audio="vgg", in_plane=1):
    from cavp_b200.models.cavp_model import CAVP
    args = SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=[False, True, True],
                           audio_backbone=audio, num_classes=nc, batch_size=2, local_rank="cpu")
    return CAVP(50, None, num_classes=nc, args=args, in_plane=in_plane)


def test_reference_set_group_lr_produces_the_same_12_groups():
    from oracle import ref_harness
    group_weight, set_group_lr = _reference_group_fns()
    hyp = SimpleNamespace(lr=1e-3, use_baseline=False)
    ours = _ours()
    ref = ref_harness.build_reference_cavp(22, [False, True, True], "vgg", 1)
    g_ours = set_group_lr(ours, hyp)  # group_weight's own assert checks every parameter was classified
    g_ref = set_group_lr(ref, hyp)
    assert len(g_ours) == len(g_ref) == 12
    for a, b in zip(g_ours, g_ref):
        pa, pb = list(a["params"]), list(b["params"])
        a["params"], b["params"] = pa, pb
        assert [tuple(p.shape) for p in pa] == [tuple(p.shape) for p in pb]
        assert a["lr"] == b["lr"] and a.get("weight_decay") == b.get("weight_decay")
    # trainer_cavp_vpo_mono.py:75-83: [:4] base LR, [4:] x10
    assert all(g["lr"] == hyp.lr for g in g_ours[:4]) and all(g["lr"] == hyp.lr * 10.0 for g in g_ours[4:])
    opt_v = torch.optim.SGD(g_ours, lr=hyp.lr, momentum=0.9, weight_decay=5e-4)  # main_vpo_mono.py:118-123
    opt_a = torch.optim.Adam(params=ours.audio_backbone.parameters(), lr=hyp.lr)  # :125
    n_v = sum(p.numel() for g in opt_v.param_groups for p in g["params"])
    n_a = sum(p.numel() for g in opt_a.param_groups for p in g["params"])
    assert n_v + n_a == sum(p.numel() for p in ours.parameters())


def test_convert_sync_batchnorm_and_module_prefixed_checkpoint_round_trip():
    from oracle import ref_harness
    ours = _ours()
    ref = ref_harness.build_reference_cavp(22, [False, True, True], "vgg", 1)
    # a checkpoint the way the reference's DDP run saves it: keys prefixed with "module."
    ckpt = {"module." + k: v.clone() for k, v in ref.state_dict().items()}
    stripped = {k[len("module."):]: v for k, v in ckpt.items()}
    missing, unexpected = ours.load_state_dict(stripped, strict=False)
    assert not missing and not unexpected
    for k, v in ref.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v), k
    # conv weights went back to channels_last storage after the load (what the kernels read)
    w = ours.backbone.backbone.layer1[0].conv2.weight
    assert w.permute(0, 2, 3, 1).is_contiguous()

    n_bn = sum(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in ours.modules())
    w_before, rm_before = ours.backbone.backbone.bn1.weight, ours.backbone.backbone.bn1.running_mean
    conv = nn.SyncBatchNorm.convert_sync_batchnorm(ours)  # main_vpo_mono.py:130
    assert sum(isinstance(m, nn.SyncBatchNorm) for m in conv.modules()) == n_bn > 50
    assert list(conv.state_dict().keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.equal(conv.state_dict()[k], v), k
    missing, unexpected = conv.load_state_dict(stripped, strict=False)
    assert not missing and not unexpected
    # the attribute tree main_vpo_mono.py / the trainers reach into survives the conversion
    assert len(conv.segment.business_layer) == 4 and conv.audio_backbone is not None and conv.memory is not None
    # (the reference builds its optimiser groups BEFORE the conversion - main_vpo_mono.py:116 vs :130 - and the
    # optimisers keep working because convert_sync_batchnorm re-uses the same Parameter objects)
    assert conv.backbone.backbone.bn1.weight is w_before and conv.backbone.backbone.bn1.running_mean is rm_before
    # and a DDP-style wrapper exposes .module as validation expects (trainer_cavp_vpo_mono.py:244-278)
    wrapped = nn.DataParallel(conv, device_ids=None) if torch.cuda.is_available() else SimpleNamespace(module=conv)
    assert wrapped.module is conv
