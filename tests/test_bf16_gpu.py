"""bf16-operand path (BASELINE.json configs[2] "bf16 training loop"; cavp_prec = 3): tcgen05 kind::f16 GEMMs with bf16
A/B and fp32 accumulation behind the same Graph / CAVP API.

The reference has no reduced-precision mode, so parity is defined against the oracle with bf16 OPERAND
ROUNDING (oracle/cavp_oracle.py:OPERAND_ROUND): every conv / Linear input and weight rounded to bf16, products and sums
exact - the arithmetic a bf16 tensor-core GEMM with wide accumulation performs.  Against that oracle:

  * op level (this file, fp64 reference on the rounded operands): forward and data-gradient GEMMs agree to fp32
    accumulation error (2e-5) - the bf16 rounding itself is bit-identical on both sides; weight gradients (plain TF32 in
    this mode) to TF32 rounding (2e-3);
  * end to end (eval forward): logits / embedding / attention as close to the fp64 rounded-operand oracle as an fp32
    evaluation of that same oracle is, and within BF16_E2E_TOL of exact arithmetic (measured values are printed);
    train mode: losses and gradient magnitudes (element-wise comparison is chaotic there, see the test).
"""
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import load_golden, rel_err
from test_ops_gpu import back, cl, seed, to_act

pytestmark = pytest.mark.gpu
TOL_ACC = 2e-5      # fp32 accumulation (both sides use identical bf16 operands)
TOL_TF32 = 2e-3     # weight gradients: TF32 products of dy and x
BF16_E2E_TOL = 5e-2   # stated tolerance of the bf16 row against fp32 / exact arithmetic (measured: 1.3e-2 .. 1.9e-2)


def bf(t):
    return t.to(torch.bfloat16).to(t.dtype)


def _g3():
    from cavp_b200.engine import Graph
    return Graph(torch.device("cuda"), prec=3, train=True)


@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,n,h,w", [
    (64, 256, 3, 1, 1, 1, 4, 30, 30),      # 256-column pair tiles, several 256-row pair tiles, K = 576 (9 k-blocks)
    (304, 256, 3, 1, 1, 1, 2, 28, 28),     # decoder conv: K = 2736 (42.75 k-blocks: K tail), taps straddle k-blocks
    (256, 304, 1, 1, 0, 1, 2, 40, 28),     # N = 304: 128-column tiles with an N tail
    (128, 48, 1, 1, 0, 1, 3, 19, 17),      # N = 48 < one tile; M = 969 (odd number of 128-row tiles, ragged last tile)
    (64, 128, 3, 2, 1, 1, 2, 33, 31),      # stride 2 (dgrad takes the strided-gather path)
    (256, 256, 3, 1, 6, 6, 2, 14, 14),     # ASPP dilation: most taps in the padding
    (512, 2048, 1, 1, 0, 1, 1, 7, 7),      # tiny M (49 rows): forward split-K slabs
    (8, 16, 3, 1, 1, 1, 1, 9, 9),          # smallest channel counts inside the envelope
    (96, 304, 3, 2, 1, 1, 3, 37, 29),      # bf16 wgrad: Cout = 304 (two 256-row pair tiles, ragged), stride 2, P % 64 != 0
    (256, 136, 1, 1, 0, 1, 2, 23, 25),     # bf16 wgrad: Cout just above one CTA's 128 rows, K = 256 (2 column tiles)
])
def test_bf16_conv_forward_dgrad_wgrad(cin, cout, k, stride, pad, dil, n, h, w):
    torch.manual_seed(cin * 7 + cout + k)
    conv = cl(nn.Conv2d(cin, cout, k, stride=stride, padding=pad, dilation=dil, bias=False)).cuda()
    x = torch.randn(n, cin, h, w)
    xr = bf(x.double()).requires_grad_(True)
    wr = bf(conv.weight.detach().double().cpu()).requires_grad_(True)
    y = F.conv2d(xr, wr, None, stride, pad, dil)
    dy = torch.randn_like(y)
    g = _g3()
    g.use_weight_cache(conv)
    assert g.wcache is not None and g.wcache.bf16
    xa = to_act(g, x)
    ya, _ = g.conv(xa, conv.weight, stride=stride, pad=pad, dil=dil)
    assert rel_err(back(ya), y) < TOL_ACC
    # backward: dgrad rounds dy and w to bf16 (exact products); wgrad multiplies fp32 dy and x on the TF32 pipe
    y.backward(bf(dy))
    dx_ref = xr.grad.clone()
    dw_ref_bf16 = wr.grad.clone()  # bf16(dy)^T * bf16(x), exact products: what the bf16 weight-gradient kernel computes
    xr.grad = None; wr.grad = None
    F.conv2d(x.double().requires_grad_(False), wr, None, stride, pad, dil).backward(dy)
    seed(g, ya, dy)
    g.backward()
    assert rel_err(back(g.grad_of(xa)), dx_ref) < TOL_ACC
    dw = g.param_grads[id(conv.weight)]
    if g.wgrad_bf16_ok(cin, cout):   # Cout > 128: tcgen05 kind::f16 with MN-major bf16 operands
        assert rel_err(dw.double().cpu(), dw_ref_bf16) < 1e-4
    else:                            # small layers: plain TF32 products of the fp32 dy and x
        assert rel_err(dw.double().cpu(), wr.grad) < TOL_TF32


def test_bf16_linear_bias_gelu_residual_and_stats():
    """Epilogue options ride along unchanged: bias + GELU (own pass), residual, BN statistics partials."""
    torch.manual_seed(3)
    lin1, lin2 = nn.Linear(304, 1216).cuda(), nn.Linear(1216, 304).cuda()
    mod = nn.Sequential(lin1, lin2)
    x = torch.randn(2, 304, 24, 20)
    from cavp_b200.engine import ACT_GELU
    g = _g3()
    g.use_weight_cache(mod)
    xa = to_act(g, x)
    h1, _ = g.conv(xa, lin1.weight, bias=lin1.bias, act=ACT_GELU, save_pre=True)
    out, _ = g.conv(h1, lin2.weight, bias=lin2.bias, res=xa)
    xt = x.double().permute(0, 2, 3, 1)
    w1, w2 = bf(lin1.weight.detach().double().cpu()), bf(lin2.weight.detach().double().cpu())
    r1 = F.gelu(F.linear(bf(xt), w1, lin1.bias.detach().double().cpu()))
    # h1 is stored in fp32 by the kernel and rounded to bf16 by the next GEMM's producers
    r2 = F.linear(bf(r1.float().double()), w2, lin2.bias.detach().double().cpu()) + xt
    assert rel_err(back(h1).permute(0, 2, 3, 1), r1) < TOL_ACC
    # one bf16 ulp flips where the fp32 h1 sits on a rounding boundary: compare with a bound that tolerates a few flips
    assert rel_err(back(out).permute(0, 2, 3, 1), r2) < 5e-4


def _oracle_bf16(fn):
    from oracle import cavp_oracle as O
    O.OPERAND_ROUND = torch.bfloat16
    try:
        return fn(O)
    finally:
        O.OPERAND_ROUND = None


def _state64(cfg):
    from oracle import schema
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0, requires_grad=True)
    return {k: (v.detach().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


@pytest.mark.parametrize("name", ["cfgA_eval_224", "tiny_train"])
def test_bf16_eval_forward_matches_rounded_operand_oracle(name):
    """Eval mode (running statistics), whole forward graph in bf16 mode against the rounded-operand oracle in fp64.

    bf16 rounding is discontinuous: wherever an activation sits next to a rounding boundary, two evaluations that differ
    in the last fp32 bits round to different bf16 neighbours (a 2^-9 jump), and ~50 layers of that add up.  The
    yardstick is therefore the oracle ITSELF: the same rounded-operand oracle evaluated in fp32 instead of fp64 differs
    from the fp64 one by ~1e-2 (max-norm, measured here on every run) - as much as bf16 differs from exact arithmetic.
    Our kernels must be as close to the fp64 rounded-operand oracle as that fp32 evaluation of the same model is."""
    from oracle import cavp_oracle as O
    from oracle import schema
    from test_parity_gpu import batch_for, build_model
    cfg = load_golden(name)["config"]
    model = build_model(cfg, prec=3).eval()
    batch = batch_for(cfg)
    B = cfg["B"]
    pred, fusion, pack = model(batch["image"].cuda(), batch["audio"][:B].cuda(), eval_mode=True)

    def oracle(dtype, rounded):
        sd = {k: (v.to(dtype) if v.is_floating_point() else v)
              for k, v in schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0).items()}
        O.OPERAND_ROUND = torch.bfloat16 if rounded else None
        try:
            with torch.no_grad():
                p_, f_, pk, _ = O.cavp_forward(sd, batch["image"].to(dtype), batch["audio"][:B].to(dtype),
                                               dilation_flags=cfg["dilation"], train=False)
            return dict(pred=p_, fusion=f_, attn=pk["attn_v"])
        finally:
            O.OPERAND_ROUND = None
    r64, r32, exact = oracle(torch.float64, True), oracle(torch.float32, True), oracle(torch.float64, False)
    ours = dict(pred=pred, fusion=fusion, attn=pack["attn_v"])
    errs = {k: rel_err(ours[k], r64[k]) for k in ours}
    yard = {k: rel_err(r32[k], r64[k]) for k in ours}
    cost = {k: rel_err(r64[k], exact[k]) for k in ours}
    print(name, "bf16 eval forward vs rounded-operand fp64 oracle", {k: "%.2e" % v for k, v in errs.items()},
          "| fp32 evaluation of the same oracle vs fp64", {k: "%.2e" % v for k, v in yard.items()},
          "| rounded-operand oracle vs exact fp64 (what bf16 operands cost)", {k: "%.2e" % v for k, v in cost.items()})
    for k, v in errs.items():
        assert v < 2.0 * yard[k] + 1e-3, (k, v, yard[k])
        assert v < BF16_E2E_TOL, (k, v)
    # against exact arithmetic the bf16 path stays within the stated bf16 tolerance as well
    for k in ours:
        assert rel_err(ours[k], exact[k]) < BF16_E2E_TOL, k


@pytest.mark.parametrize("name", ["tiny_train", "tiny_train_fff71"])
def test_bf16_train_step_runs_and_tracks_the_oracle_losses(name):
    """Train mode: batch-statistics BatchNorm on random weights amplifies a single one-ulp bf16 flip (2^-9) into O(0.1)
    differences of the logits (the same mechanism that forces 3xTF32 in the fp32 row), so tensors cannot be
    compared element-wise with ANY other bf16 implementation; the op-level tests above pin the arithmetic.  Here: the
    step runs through the public API, losses stay within a few percent of the rounded-operand oracle, every gradient
    is finite and has the oracle's magnitude."""
    from cavp_b200.trainer import shuffled_labels, train_step
    from test_parity_gpu import batch_for, build_model
    cfg = load_golden(name)["config"]
    model = build_model(cfg, prec=3).train()
    batch = batch_for(cfg)
    spl = shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    torch.manual_seed(99)
    res = train_step(model, batch["image"].cuda(), batch["audio"].cuda(), batch["pix_label"], spl,
                     max_views=cfg["max_views"], keep_outputs=True)
    sd = _state64(cfg)
    b = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
    torch.manual_seed(99)
    l_ce, l_ctr, out_cat, ctr_cat, pack, newbuf = _oracle_bf16(lambda O: O.train_step_losses(
        sd, b, spl, dilation_flags=cfg["dilation"], audio_kind=cfg["audio"], max_views=cfg["max_views"]))
    e_ce = abs(float(res.l_ce) - float(l_ce)) / abs(float(l_ce))
    e_ctr = abs(float(res.l_ctr) - float(l_ctr.sum())) / max(abs(float(l_ctr.sum())), 1e-12)
    print(name, "bf16 train step vs rounded-operand oracle: l_ce %.2e l_ctr %.2e; logits rel err %.2e (chaotic, not "
          "asserted)" % (e_ce, e_ctr, rel_err(res.out_pred, out_cat)))
    assert e_ce < 5e-2 and (float(l_ctr.sum()) == 0.0 or e_ctr < 5e-2)
    (l_ce + l_ctr.sum()).backward()
    ratios = []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        assert torch.isfinite(p.grad).all(), k
        ratios.append(float(p.grad.double().norm().cpu() / sd[k].grad.norm().clamp_min(1e-30)))
    ratios.sort()
    print(name, "gradient norm ratio ours / oracle: min %.3f median %.3f max %.3f" % (ratios[0], ratios[len(ratios) // 2], ratios[-1]))
    assert 0.5 < ratios[len(ratios) // 2] < 2.0
