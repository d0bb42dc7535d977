"""Eval epilogue: argmax + confusion kernels and the MIoU / ForegroundDetect drop-ins against the
golden vectors produced by the unmodified reference classes (tests/golden/metrics.pt, oracle/make_golden_metrics.py)."""
import numpy
import pytest
import torch

from conftest import load_golden
from oracle import metrics_oracle as MO


def confusion_cpu(logits, target, C, ignore):
    """host restatement of what the kernel counts: conf[(C+1)][C]"""
    pred = logits.argmax(1).reshape(-1)
    t = target.reshape(-1)
    conf = numpy.zeros((C + 1, C), dtype=numpy.int64)
    for tt, pp in zip(t.tolist(), pred.tolist()):
        if tt == ignore or tt < 0:
            continue
        conf[tt if tt < C else C, pp] += 1
    return conf


def test_oracle_matches_reference_golden():
    gold = load_golden("metrics")
    for g in gold["cases"]:
        c = g["case"]
        logits, target = MO.metric_case(**c)
        sample = MO.miou_sample(logits, target, c["C"], 255)
        for a, b in zip(sample, g["sample"]):
            assert numpy.array_equal(numpy.asarray(a, dtype=numpy.float64), b.numpy())
        logits2, target2 = MO.metric_case(**{**c, "seed": c["seed"] + 100})
        cm = MO.foreground_confusion(logits, target, c["C"], 255) + MO.foreground_confusion(logits2, target2, c["C"], 255)
        assert numpy.array_equal(cm, g["confusion"].numpy())


def test_host_derivation_from_confusion_matches_golden():
    """cavp_b200.metrics derives every reference number from conf[label][pred]; checked here with a CPU-built conf."""
    from cavp_b200.metrics import MIoU, ForegroundDetect
    gold = load_golden("metrics")
    for g in gold["cases"]:
        c = g["case"]
        C = c["C"]
        m = MIoU(C, 255)
        fg = ForegroundDetect(num_classes=C)
        logits, target = MO.metric_case(**c)
        conf = confusion_cpu(logits, target, C, 255)
        sample = MIoU.sample_from_confusion(conf, C)
        for a, b in zip(sample, g["sample"]):
            assert numpy.array_equal(numpy.asarray(a, dtype=numpy.float64), b.numpy())
        r1 = m.update_from_confusion(conf)
        assert [float(x) for x in r1] == g["miou_after_1"]
        logits2, target2 = MO.metric_case(**{**c, "seed": c["seed"] + 100})
        conf2 = confusion_cpu(logits2, target2, C, 255)
        r2 = m.update_from_confusion(conf2)
        assert [float(x) for x in r2] == g["miou_after_2"]
        assert numpy.allclose(m.iou, g["iou"].numpy(), rtol=0, atol=0)
        fg.update_from_confusion(conf)
        fg.update_from_confusion(conf2)
        assert numpy.array_equal(fg.confusion_matrix_, g["confusion"].numpy())
        ours = fg.get_metric_results()
        ref = MO.foreground_scores(g["confusion"].numpy())
        assert all(numpy.array_equal(a, b) for a, b in zip(ours, ref))


@pytest.mark.gpu
def test_argmax_confusion_kernel_matches_golden():
    from cavp_b200.metrics import MIoU, ForegroundDetect, argmax_confusion
    gold = load_golden("metrics")
    for g in gold["cases"]:
        c = g["case"]
        C = c["C"]
        logits, target = MO.metric_case(**c)
        pred, conf = argmax_confusion(logits.cuda(), target.cuda(), 255, want_pred=True)
        assert torch.equal(pred.cpu(), logits.argmax(1))                       # bit-exact argmax indices
        assert numpy.array_equal(conf.cpu().numpy(), confusion_cpu(logits, target, C, 255))
        m = MIoU(C, 255)
        fg = ForegroundDetect(num_classes=C)
        r1 = m(logits.cuda(), target.cuda())
        assert [float(x) for x in r1] == g["miou_after_1"]
        logits2, target2 = MO.metric_case(**{**c, "seed": c["seed"] + 100})
        r2 = m(logits2.cuda(), target2.cuda())
        assert [float(x) for x in r2] == g["miou_after_2"]
        fg(logits.cuda(), target.cuda())
        fg(logits2.cuda(), target2.cuda())
        assert numpy.array_equal(fg.confusion_matrix_, g["confusion"].numpy())


@pytest.mark.gpu
def test_argmax_confusion_edge_cases():
    from cavp_b200.metrics import argmax_confusion
    # all pixels ignored, negative labels, labels >= C, ties (first maximum wins), one pixel, > 96 KB of counters
    logits = torch.zeros(1, 5, 3, 3)
    logits[0, 2] = 1.0
    logits[0, 4] = 1.0                                   # tie between class 2 and 4 -> 2
    target = torch.tensor([[[255, -1, 7], [2, 4, 0], [255, 255, 2]]])
    pred, conf = argmax_confusion(logits.cuda(), target.cuda(), 255, want_pred=True)
    assert torch.equal(pred.cpu(), torch.full((1, 3, 3), 2))
    exp = numpy.zeros((6, 5), dtype=numpy.int64)
    exp[5, 2] = 1; exp[2, 2] = 2; exp[4, 2] = 1; exp[0, 2] = 1
    assert numpy.array_equal(conf.cpu().numpy(), exp)
    _, conf = argmax_confusion(torch.randn(2, 3, 4, 4).cuda(), torch.full((2, 4, 4), 255).cuda(), 255)
    assert int(conf.sum()) == 0
    g = torch.Generator().manual_seed(0)
    big = torch.randn(1, 200, 8, 8, generator=g)         # (C+1)*C*4 B > 96 KB: global-atomic path
    tb = torch.randint(0, 200, (1, 8, 8), generator=g)
    _, conf = argmax_confusion(big.cuda(), tb.cuda(), 255)
    assert numpy.array_equal(conf.cpu().numpy(), confusion_cpu(big, tb, 200, 255))


@pytest.mark.gpu
def test_fused_upsample_argmax_equals_materialised_path():
    """CAVP.forward_eval_metrics (upsample + argmax + confusion in one kernel, no full-resolution logits) against
    forward_inference followed by the materialised kernel: identical predictions and counts."""
    from types import SimpleNamespace
    from cavp_b200.metrics import argmax_confusion
    from cavp_b200.models.cavp_model import CAVP
    from oracle import schema, seeded
    nc = 22
    args = SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=[False, True, True],
                           audio_backbone="vgg", num_classes=nc, batch_size=2, local_rank=0)
    model = CAVP(50, None, num_classes=nc, args=args, in_plane=1)
    model.load_state_dict(schema.seeded_state(nc, "vgg", 1, seed=0), strict=True)
    model = model.cuda().eval()
    batch = seeded.synthetic_batch(2, 64, 96, nc, seed=5)
    image, audio = batch["image"].cuda(), batch["audio"][:2].cuda()
    target = batch["pix_label"].cuda()
    with torch.no_grad():
        out_pred, _, _ = model(image, audio, eval_mode=True)
    pred_ref, conf_ref = argmax_confusion(out_pred, target, 255, want_pred=True)
    assert torch.equal(pred_ref.cpu(), out_pred.cpu().argmax(1))
    pred, conf = model.forward_eval_metrics(image, audio, target)
    # (a second forward pass: split-K layers accumulate with red.global.add, so its logits equal the first pass's to
    # fp32 rounding, not bitwise; the fused kernel itself is checked bit-exactly below on the SAME low-resolution logits)
    top2 = out_pred.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * float(out_pred.abs().max())
    assert torch.equal(pred[safe], pred_ref[safe])
    assert int((conf - conf_ref).abs().sum()) <= 2 * int((~safe).sum())
    # same logits, both kernels: upsample with cavp_bilinear_fwd then argmax  ==  fused upsample+argmax
    from cavp_b200 import _C
    low = torch.randn(2, 16, 24, 24, device="cuda")                      # NHWC [n, h, w, C=24 (22 used)]
    full = torch.empty(2, nc, 64, 96, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _C.call("cavp_bilinear_fwd", low.data_ptr(), 24, 16, 24, full.data_ptr(), 0, 64, 96, 2, nc, 0, 1, st)
    p_mat, c_mat = argmax_confusion(full, target, 255, want_pred=True)
    p_fused = torch.empty(2, 64, 96, dtype=torch.int64, device="cuda")
    c_fused = torch.zeros(nc + 1, nc, dtype=torch.int64, device="cuda")
    _C.call("cavp_upsample_argmax_confusion", low.data_ptr(), 24, 16, 24, 64, 96, 2, nc, target.data_ptr(), 255,
            p_fused.data_ptr(), c_fused.data_ptr(), st)
    assert torch.equal(p_fused, p_mat) and torch.equal(c_fused, c_mat)
