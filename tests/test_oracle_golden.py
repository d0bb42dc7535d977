"""Pins the restated oracle (oracle/cavp_oracle.py) to outputs of the unmodified reference
(tests/golden/*.pt, produced by oracle/make_golden.py in the build container)."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import cavp_oracle as O
from oracle import schema, seeded

TOL = 1e-4  # oracle vs reference on the same CPU: only op-ordering noise is allowed


def _setup(cfg, requires_grad):
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0, requires_grad=requires_grad)
    batch = seeded.synthetic_batch(cfg["B"], cfg["H"], cfg["W"], cfg["nc"], seed=666, audio_frames=cfg["frames"],
                                   in_plane=cfg["in_plane"])
    return sd, batch


def test_eval_cfgA_matches_reference():
    g = load_golden("cfgA_eval_224")
    cfg = g["config"]
    sd, batch = _setup(cfg, False)
    with torch.no_grad():
        pred, fusion, pack, _ = O.cavp_forward(sd, batch["image"], batch["audio"][: cfg["B"]],
                                               dilation_flags=cfg["dilation"], train=False)
    assert rel_err(pred[:, :, ::4, ::4], g["pred_stride4"]) < TOL
    assert rel_err(fusion[:, :, ::2, ::2], g["fusion_stride2"]) < TOL
    assert rel_err(pack["attn_v"], g["attn_v"]) < TOL
    am = pred.argmax(1).to(torch.uint8)
    safe = g["margin"].float() > 1e-3 * g["pred_summary"]["absmax"]
    assert torch.equal(am[safe], g["argmax"][safe])
    assert float((am != g["argmax"]).float().mean()) < 1e-3


@pytest.mark.parametrize("name", ["tiny_train", "tiny_train_fff71", "tiny_train_r18_stereo", "cfgB_train_224"])
def test_train_step_matches_reference(name):
    g = load_golden(name)
    cfg = g["config"]
    sd, batch = _setup(cfg, True)
    B = cfg["B"]
    spl = seeded.shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    if cfg["audio_func"]:
        out_cat, ctr_cat, pack, newbuf = O.cavp_forward(sd, batch["image"], batch["audio"][:B],
                                                        dilation_flags=cfg["dilation"], audio_kind=cfg["audio"],
                                                        train=True, shuffle_idx=batch["shuffle_idx"], audio_func=True)
    else:
        out_cat, ctr_cat, pack, newbuf = O.cavp_forward(sd, batch["image"], batch["audio"],
                                                        dilation_flags=cfg["dilation"], audio_kind=cfg["audio"],
                                                        train=True)
    output = out_cat[:B] + out_cat[B:] * 0.0
    torch.manual_seed(1234)
    l_ctr = O.contrast_loss(ctr_cat[:B], batch["pix_label"], ctr_cat[B:], spl, cfg["max_views"])
    l_ce = O.cross_entropy(output, batch["pix_label"])
    (l_ce + l_ctr).backward()

    ps, fs, at = g["pred_stride"], g["fusion_stride"], g["attn_stride"]
    assert rel_err(out_cat[:, :, ::ps, ::ps], g["pred"]) < TOL
    assert rel_err(ctr_cat[:, :, ::fs, ::fs], g["fusion"]) < TOL
    assert rel_err(pack["attn_v"][:, :, ::at], g["attn_v"]) < TOL
    assert rel_err(pack["audio"], g["audio"]) < TOL
    assert abs(float(l_ce) - g["l_ce"]) < TOL * abs(g["l_ce"])
    assert abs(float(l_ctr) - g["l_ctr"]) < TOL * abs(g["l_ctr"])
    for k, v in g["buffers"].items():
        assert rel_err(newbuf[k], v) < TOL, k
    from oracle.make_golden import sample_idx
    worst = 0.0
    for k, gs in g["grads"].items():
        p = sd[k]
        if gs is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        got = p.grad.flatten()[sample_idx(p.grad.numel())]
        e = float((got - gs["samples"]).abs().max()) / max(gs["absmax"], 1e-30)
        worst = max(worst, e)
        assert e < 5e-4, (k, e)
        assert abs(float(p.grad.double().norm()) - gs["norm"]) <= 5e-4 * gs["norm"] + 1e-12, k
    print(name, "worst grad rel err", worst)
