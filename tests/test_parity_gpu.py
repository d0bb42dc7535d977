"""Parity of the CUDA product (C-ABI kernels behind cavp_b200.models.CAVP / cavp_b200.trainer.train_step) against
the golden fixtures produced by the UNMODIFIED reference (tests/golden/*.pt, oracle/make_golden.py) and against the
oracle on the same seeded inputs.  Bar (BASELINE.json north_star): 1e-3 relative fp32 on logits; bit-exact argmax
wherever the top-2 margin exceeds the numerical noise."""
from types import SimpleNamespace

import pytest
import torch

from conftest import load_golden, rel_err
from oracle import schema, seeded
from oracle.make_golden import sample_idx

pytestmark = pytest.mark.gpu
TOL = 1e-3  # north_star tolerance
# End-to-end gradients of this random-weight, batch-stat-BN network are ill-conditioned in fp32: the reference
# arithmetic itself (oracle in fp32 vs the same oracle in fp64, tools/grad_sensitivity.py ->
# tests/golden/grad_sensitivity.json) differs by 3-4e-2 relative L2 per tensor and up to 9e-3 in norm, because tiny
# forward perturbations flip ReLU / max-pool decisions.  Two independent fp32 implementations can therefore only agree
# to a small multiple of that self-discrepancy: against the fp32 goldens we allow 2x the fixture's measured value (our
# error + the reference's own); elementwise gradient parity is asserted per op (2e-5) in tests/test_ops_gpu.py, and
# tests/test_parity_fullsize_oracle_gpu.py compares every gradient tensor with an fp64 oracle on the GPU, where the
# kernels must be at least as close to exact arithmetic as stock fp32 PyTorch is.
import json
import os

_SENS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad_sensitivity.json")))
GRAD_FACTOR = 2.0


def grad_tols(name):
    s = _SENS[name]
    return GRAD_FACTOR * s["worst_sample_l2"], max(GRAD_FACTOR * s["worst_norm"], 1e-2)


def build_model(cfg, prec=2):
    from cavp_b200.models.cavp_model import CAVP
    args = SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=list(cfg["dilation"]),
                           audio_backbone=cfg["audio"], num_classes=cfg["nc"], batch_size=cfg["B"], local_rank=0,
                           cavp_prec=prec)
    model = CAVP(50, None, num_classes=cfg["nc"], ignore_index=255, audio_backbone_pretrain_path=None,
                 visual_backbone=50, args=args, in_plane=cfg["in_plane"])
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.cuda()


def batch_for(cfg):
    return seeded.synthetic_batch(cfg["B"], cfg["H"], cfg["W"], cfg["nc"], seed=666, audio_frames=cfg["frames"],
                                  in_plane=cfg["in_plane"])


def test_state_dict_keys_match_reference_schema():
    cfg = load_golden("tiny_train")["config"]
    model = build_model(cfg)
    ref_keys = list(schema.cavp_schema(cfg["nc"], cfg["audio"], cfg["in_plane"]).shapes.keys())
    assert sorted(model.state_dict().keys()) == sorted(ref_keys)


def test_eval_cfgA_matches_reference_golden():
    g = load_golden("cfgA_eval_224")
    cfg = g["config"]
    model = build_model(cfg).eval()
    batch = batch_for(cfg)
    pred, fusion, pack = model(batch["image"].cuda(), batch["audio"][:cfg["B"]].cuda(), eval_mode=True)
    assert pred.shape == (1, 71, 224, 224) and fusion.shape == (1, 304, 56, 56)
    assert pack["attn_v"].shape == (1, 4, 3136, 1) and pack["audio"].shape == (1, 304, 1, 1)
    e_pred = rel_err(pred[:, :, ::4, ::4], g["pred_stride4"])
    e_fus = rel_err(fusion[:, :, ::2, ::2], g["fusion_stride2"])
    e_att = rel_err(pack["attn_v"], g["attn_v"])
    print("eval cfgA rel err: pred %.2e fusion %.2e attn %.2e" % (e_pred, e_fus, e_att))
    assert e_pred < TOL and e_fus < TOL and e_att < TOL
    am = pred.argmax(1).to(torch.uint8).cpu()
    safe = g["margin"].float() > 2e-3 * g["pred_summary"]["absmax"]
    assert torch.equal(am[safe], g["argmax"][safe])  # bit-exact argmax outside the noise margin
    mismatch = float((am != g["argmax"]).float().mean())
    print("argmax mismatch rate %.2e" % mismatch)
    assert mismatch < 1e-3


@pytest.mark.parametrize("name", ["tiny_train", "tiny_train_fff71", "tiny_train_r18_stereo", "cfgB_train_224"])
def test_train_step_matches_reference_golden(name):
    from cavp_b200.trainer import shuffled_labels, train_step
    g = load_golden(name)
    cfg = g["config"]
    model = build_model(cfg).train()
    batch = batch_for(cfg)
    B = cfg["B"]
    spl = shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    assert torch.equal(spl, seeded.shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"]))
    audio = batch["audio"][:B] if cfg["audio_func"] else batch["audio"]
    torch.manual_seed(1234)  # make_golden seeds the CPU generator right before ContrastLoss
    res = train_step(model, batch["image"].cuda(), audio.cuda(), batch["pix_label"], spl, max_views=cfg["max_views"],
                     shuffle_idx=batch["shuffle_idx"].cuda() if cfg["audio_func"] else None,
                     audio_func=cfg["audio_func"], keep_outputs=True)
    ps, fs, at = g["pred_stride"], g["fusion_stride"], g["attn_stride"]
    errs = dict(pred=rel_err(res.out_pred[:, :, ::ps, ::ps], g["pred"]),
                fusion=rel_err(res.out_fusion[:, :, ::fs, ::fs], g["fusion"]),
                attn=rel_err(res.attn_v[:, :, ::at], g["attn_v"]), audio=rel_err(res.audio, g["audio"]),
                l_ce=abs(float(res.l_ce) - g["l_ce"]) / abs(g["l_ce"]),
                l_ctr=abs(float(res.l_ctr) - g["l_ctr"]) / abs(g["l_ctr"]))
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < TOL, (k, v)
    sd = model.state_dict()
    for k, v in g["buffers"].items():
        assert rel_err(sd[k].float(), v.float()) < TOL, k
    worst, worst_n = (0.0, ""), (0.0, "")
    GRAD_SAMPLE_L2_TOL, GRAD_NORM_TOL = grad_tols(name)
    params = dict(model.named_parameters())
    for k, gs in g["grads"].items():
        p = params[k]
        if gs is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        gr = p.grad.detach().cpu().contiguous()
        got = gr.flatten()[sample_idx(gr.numel())].double()
        ref = gs["samples"].double()
        e = float((got - ref).norm()) / max(float(ref.norm()), 1e-3 * gs["absmax"] * ref.numel() ** 0.5, 1e-30)
        ne = abs(float(gr.double().norm()) - gs["norm"]) / (gs["norm"] + 1e-30)
        worst = max(worst, (e, k))
        worst_n = max(worst_n, (ne, k))
        assert e < GRAD_SAMPLE_L2_TOL and ne < GRAD_NORM_TOL, (k, e, ne)
    print(name, "worst grad sample-L2 %.2e (%s); worst norm err %.2e (%s)" % (worst + worst_n))


def test_module_autograd_path_equals_train_step():
    """nn.Module forward + torch-side slicing + cavp_b200 losses (the drop-in usage) == fused train_step."""
    from cavp_b200.loss import ContrastLoss, CrossEntropyLoss
    from cavp_b200.trainer import shuffled_labels, train_step
    cfg = load_golden("tiny_train")["config"]
    batch = batch_for(cfg)
    B = cfg["B"]
    spl = shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    GRAD_SAMPLE_L2_TOL, GRAD_NORM_TOL = grad_tols("tiny_train")
    m1 = build_model(cfg).train()
    torch.manual_seed(1234)
    r = train_step(m1, batch["image"].cuda(), batch["audio"].cuda(), batch["pix_label"], spl, max_views=cfg["max_views"])
    m2 = build_model(cfg).train()
    out_cat, ctr_cat, pack = m2(batch["image"].cuda(), batch["audio"].cuda(), None, False)
    assert out_cat.shape == (2 * B, cfg["nc"], cfg["H"], cfg["W"]) and pack["visual"].shape == ctr_cat.shape
    output = out_cat[:B] + out_cat[B:] * 0.0  # trainer_cavp_vpo_mono.py:171
    torch.manual_seed(1234)
    l_ctr = ContrastLoss(0.1, 255, cfg["max_views"])(ctr_cat[:B], batch["pix_label"].cuda(), ctr_cat[B:], spl.cuda())
    l_ce = CrossEntropyLoss(255)(output, batch["pix_label"].cuda())
    (l_ce + l_ctr).backward()
    assert abs(float(l_ce) - float(r.l_ce)) < 1e-5 * abs(float(r.l_ce))
    assert abs(float(l_ctr) - float(r.l_ctr)) < 1e-5 * abs(float(r.l_ctr))
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), k
        if p1.grad is not None:
            # same kernels; split-K / pooling atomics reorder fp32 sums run to run and the network amplifies that
            a, b = p2.grad.double().flatten(), p1.grad.double().flatten()
            assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < GRAD_SAMPLE_L2_TOL, k
            assert abs(float(a.norm() - b.norm())) / float(b.norm().clamp_min(1e-30)) < GRAD_NORM_TOL, k


# ---------------------------------------------------------------------------------------------------------------
# edge cases against the oracle on the same seeded inputs (the reference itself has no tests: SURVEY.md section 4)
# ---------------------------------------------------------------------------------------------------------------
def _oracle_train(cfg, batch, spl, seed, dtype=torch.float64):
    """The oracle in fp64 (CPU) is the arbiter for the tiny, badly conditioned edge shapes: against the fp32 oracle a
    comparison also contains the fp32 oracle's own rounding noise, which is as large as ours here."""
    from oracle import cavp_oracle as O
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0, requires_grad=True)
    sd = {k: (v.detach().to(dtype).requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    torch.manual_seed(seed)
    return O.train_step_losses(sd, b, spl, dilation_flags=cfg["dilation"], audio_kind=cfg["audio"],
                               max_views=cfg["max_views"])


@pytest.mark.parametrize("cfg", [
    # minimum legal train batch, binary AVSBench-object classes, non-square input
    dict(B=2, H=96, W=128, nc=2, dilation=(False, False, False), audio="vgg", in_plane=1, frames=96, max_views=16),
    # odd feature-map sizes: 72x104 -> stride-4 map 18x26, stride-16 maps 5x7 (ASPP dilations mostly in the padding)
    dict(B=3, H=72, W=104, nc=22, dilation=(False, True, True), audio="vgg", in_plane=1, frames=96, max_views=32),
])
def test_train_step_edge_shapes_match_oracle(cfg):
    from cavp_b200.trainer import shuffled_labels, train_step
    model = build_model(cfg).train()
    batch = batch_for(cfg)
    spl = shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    torch.manual_seed(77)
    res = train_step(model, batch["image"].cuda(), batch["audio"].cuda(), batch["pix_label"], spl,
                     max_views=cfg["max_views"], keep_outputs=True)
    l_ce, l_ctr, out_cat, ctr_cat, pack, newbuf = _oracle_train(cfg, batch, spl, 77)
    errs = dict(pred=rel_err(res.out_pred, out_cat), fusion=rel_err(res.out_fusion, ctr_cat),
                attn=rel_err(res.attn_v, pack["attn_v"]), l_ce=abs(float(res.l_ce) - float(l_ce)) / abs(float(l_ce)))
    if float(l_ctr) != 0.0:
        errs["l_ctr"] = abs(float(res.l_ctr) - float(l_ctr)) / abs(float(l_ctr))
    else:
        assert float(res.l_ctr) == 0.0  # no class reached max_views pixels: reference returns tensor([0.])
    print(cfg["H"], cfg["W"], {k: "%.2e" % v for k, v in errs.items()})
    # north-star tolerance against exact (fp64) arithmetic; the same shapes are checked at 1e-4 in eval mode below
    for k, v in errs.items():
        assert v < TOL, (k, v)
    sd = model.state_dict()
    for k, v in newbuf.items():
        assert rel_err(sd[k].float(), v.float()) < TOL, k


@pytest.mark.parametrize("cfg", [
    dict(B=2, H=136, W=168, nc=71, dilation=(False, False, False), audio="vgg", in_plane=1, frames=96),
    dict(B=3, H=72, W=104, nc=22, dilation=(False, True, True), audio="vgg", in_plane=1, frames=96),
    dict(B=1, H=96, W=128, nc=2, dilation=(False, False, False), audio="vgg", in_plane=1, frames=96),
])
def test_eval_batch_and_odd_size_match_oracle(cfg):
    from oracle import cavp_oracle as O
    model = build_model(cfg).eval()
    batch = batch_for(cfg)
    B = cfg["B"]
    pred, fusion, pack = model(batch["image"].cuda(), batch["audio"][:B].cuda(), eval_mode=True)
    sd = schema.seeded_state(cfg["nc"], "vgg", 1, seed=0)
    with torch.no_grad():
        rp, rf, rpack, _ = O.cavp_forward(sd, batch["image"], batch["audio"][:B], dilation_flags=cfg["dilation"],
                                          train=False)
    assert pred.shape == rp.shape and fusion.shape == rf.shape
    assert rel_err(pred, rp) < 1e-4 and rel_err(fusion, rf) < 1e-4 and rel_err(pack["attn_v"], rpack["attn_v"]) < 1e-4
    top2 = rp.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * float(rp.abs().max())
    assert torch.equal(pred.argmax(1).cpu()[safe], rp.argmax(1)[safe])


def test_forward_audio_path_with_sound_bank_overwrite():
    """trainer_cavp_vpo_stereo.py:211 usage: B audio rows + shuffle_info, ow_flag=True touches only the SoundBank."""
    cfg = load_golden("tiny_train_r18_stereo")["config"]
    model = build_model(cfg).train()
    batch = batch_for(cfg)
    B = cfg["B"]
    info = {"shuffle_idx": batch["shuffle_idx"].cuda(), "mod_idx_map": {0: 3}, "image_label": batch["img_label"].clone().cuda()}
    out_cat, ctr_cat, pack = model(batch["image"].cuda(), batch["audio"][:B].cuda(), info, True, audio_func=True)
    assert out_cat.shape == (2 * B, cfg["nc"], cfg["H"], cfg["W"]) and pack["audio"].shape == (2 * B, 304, 1, 1)
    assert torch.equal(pack["audio"][B:], pack["audio"][:B][batch["shuffle_idx"].cuda()])
    assert float(model.memory.bank_vault.abs().sum()) > 0  # update_bank queued single-label features
    out_cat.sum().backward()
    assert model.audio_backbone.backbone.fc.weight.grad is not None


def test_eval_forward_is_bitwise_deterministic():
    """Forward split-K stores per-split slabs and sums them in a fixed order (no atomics), BN / LN partials are reduced
    in a fixed order too: two passes over the same input give identical logits, hence identical argmax masks."""
    cfg = load_golden("tiny_train")["config"]
    model = build_model(cfg).eval()
    batch = batch_for(cfg)
    img, aud = batch["image"].cuda(), batch["audio"][:cfg["B"]].cuda()
    with torch.no_grad():
        p1, f1, _ = model(img, aud, eval_mode=True)
        p2, f2, _ = model(img, aud, eval_mode=True)
    assert torch.equal(p1, p2) and torch.equal(f1, f2)
