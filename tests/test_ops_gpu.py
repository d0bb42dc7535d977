"""Op-level parity of the C-ABI kernels (through cavp_b200.engine.Graph) against torch CPU fp64 restatements of the
same reference ops.  Tolerance: 1e-3 relative (north_star) - in practice these land around 1e-6."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _g(train=True, prec=2):
    from cavp_b200.engine import Graph
    return Graph(torch.device("cuda"), prec=prec, train=train)


def to_act(g, t, needs_grad=True):
    from cavp_b200.engine import new_act
    n, c, h, w = t.shape
    a = new_act(n, h, w, c, g.device, needs_grad=needs_grad)
    tc = t.float().contiguous().cuda()
    g.call("cavp_nchw_to_nhwc", tc.data_ptr(), a.ptr, n, c, h * w, c)
    return a


def seed(g, a, dy):
    d, acc = g.grad_target(a)
    assert not acc
    tc = dy.float().contiguous().cuda()
    if d.ld == a.c:
        g.call("cavp_nchw_to_nhwc", tc.data_ptr(), d.ptr, a.n, a.c, a.h * a.w, a.c)
    else:
        from cavp_b200.engine import new_act
        tmp = new_act(a.n, a.h, a.w, a.c, g.device)
        g.call("cavp_nchw_to_nhwc", tc.data_ptr(), tmp.ptr, a.n, a.c, a.h * a.w, a.c)
        g.copy_act(d, tmp)


def back(a):
    return a.nchw().double().cpu()


def cl(m):
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d) and mod.weight.shape[1] % 4 == 0:
            mod.weight.data = mod.weight.data.contiguous(memory_format=torch.channels_last)
    return m


@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,n,h,w,use_res", [
    (64, 64, 3, 1, 1, 1, 2, 14, 14, True),
    (32, 96, 3, 2, 1, 1, 3, 15, 13, False),
    (128, 256, 1, 1, 0, 1, 2, 8, 8, True),
    (64, 128, 1, 2, 0, 1, 2, 12, 12, False),
    (256, 48, 3, 1, 4, 4, 1, 10, 10, False),
    (3, 64, 3, 2, 1, 1, 2, 16, 16, False),
    (64, 256, 3, 1, 1, 1, 4, 30, 30, False),   # N = 256: eligible for the 256-column pair tile (forced-schedule runs)
    (304, 256, 3, 1, 1, 1, 2, 28, 28, False),  # decoder conv: K = 2736 (K tail inside a promotion unit)
    (96, 512, 3, 2, 1, 1, 3, 17, 15, True),    # two 256-column tiles, stride 2, ragged last row tile, residual in BN
])
def test_conv_bn_relu_train(cin, cout, k, stride, pad, dil, n, h, w, use_res):
    torch.manual_seed(0)
    conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=False).double()
    bn = nn.BatchNorm2d(cout).double()
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    x = torch.randn(n, cin, h, w, dtype=torch.double, requires_grad=cin % 4 == 0)
    ref_y = conv(x)
    res = torch.randn_like(ref_y, requires_grad=True) if use_res else None
    bn.train()
    z = F.relu(bn(ref_y) + (res if use_res else 0))
    dz = torch.randn_like(z)
    z.backward(dz)

    conv32, bn32 = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=False).cuda(), nn.BatchNorm2d(cout).cuda()
    conv32.weight.data.copy_(conv.weight.data.float()); cl(conv32)
    bn32.weight.data.copy_(bn.weight.data.float()); bn32.bias.data.copy_(bn.bias.data.float())
    bn32.running_mean.copy_(rm0.float()); bn32.running_var.copy_(rv0.float())

    g = _g()
    from cavp_b200.engine import pad4
    if cin % 4:
        xp = torch.zeros(n, pad4(cin), h, w); xp[:, :cin] = x.detach().float()
        xa = to_act(g, xp, needs_grad=False)
    else:
        xa = to_act(g, x.detach())
    ra = to_act(g, res.detach()) if use_res else None
    za = g.conv_bn(xa, conv32.weight, bn32, stride=stride, pad=pad, dil=dil, res=ra)
    assert rel_err(back(za), z) < TOL
    assert rel_err(bn32.running_mean, bn.running_mean) < TOL
    assert rel_err(bn32.running_var, bn.running_var) < TOL
    seed(g, za, dz)
    g.backward()
    assert rel_err(g.param_grads[id(conv32.weight)], conv.weight.grad) < TOL
    assert rel_err(g.param_grads[id(bn32.weight)], bn.weight.grad) < TOL
    assert rel_err(g.param_grads[id(bn32.bias)], bn.bias.grad) < TOL
    if cin % 4 == 0:
        assert rel_err(back(g.grad_of(xa)), x.grad) < TOL
    if use_res:
        assert rel_err(back(g.grad_of(ra)), res.grad) < TOL


def test_conv_bn_eval_fused():
    torch.manual_seed(1)
    conv = cl(nn.Conv2d(64, 96, 3, 1, 2, 2, bias=False)).cuda()
    bn = nn.BatchNorm2d(96).cuda()
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, 64, 9, 11)
    g = _g(train=False)
    za = g.conv_bn(to_act(g, x), conv.weight, bn, stride=1, pad=2, dil=2)
    ref = F.relu(F.batch_norm(F.conv2d(x.double(), conv.weight.double().cpu(), None, 1, 2, 2),
                              bn.running_mean.double().cpu(), bn.running_var.double().cpu(), bn.weight.double().cpu(),
                              bn.bias.double().cpu(), False, 0.1, bn.eps))
    assert rel_err(back(za), ref) < TOL


def test_conv_n256_eval_epilogue_dgrad_accumulate_and_splitk():
    """Shapes with N % 256 == 0 on both sides (eligible for the 256-column pair tile under CAVP_IGEMM_BN256=1, see
    tests/test_forced_schedules_gpu.py): eval-mode folded BN + LeakyReLU epilogue, a data gradient that accumulates in
    place into an existing gradient (two consumers of one tensor), and the deterministic split-K slabs of a tiny map."""
    from cavp_b200.engine import ACT_LEAKY
    torch.manual_seed(11)
    conv = cl(nn.Conv2d(256, 256, 3, 1, 2, 2, bias=False)).cuda()
    bn = nn.BatchNorm2d(256).cuda()
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(3, 256, 13, 11)
    g = _g(train=False)
    za = g.conv_bn(to_act(g, x), conv.weight, bn, stride=1, pad=2, dil=2, act=ACT_LEAKY)
    ref = F.leaky_relu(F.batch_norm(F.conv2d(x.double(), conv.weight.double().cpu(), None, 1, 2, 2),
                                    bn.running_mean.double().cpu(), bn.running_var.double().cpu(),
                                    bn.weight.double().cpu(), bn.bias.double().cpu(), False, 0.1, bn.eps), 0.01)
    assert rel_err(back(za), ref) < TOL
    # two convs read the same input: the second data gradient accumulates in place (igemm residual == output)
    c1, c2 = cl(nn.Conv2d(256, 64, 3, 1, 1, 1, bias=False)).cuda(), cl(nn.Conv2d(256, 512, 1, 1, 0, 1, bias=False)).cuda()
    xr = x.double().requires_grad_(True)
    y1 = F.conv2d(xr, c1.weight.double().cpu(), None, 1, 1, 1)
    y2 = F.conv2d(xr, c2.weight.double().cpu())
    d1, d2 = torch.randn_like(y1), torch.randn_like(y2)
    (y1 * d1).sum().backward(retain_graph=True)
    (y2 * d2).sum().backward()
    g = _g()
    xa = to_act(g, x)
    a1, _ = g.conv(xa, c1.weight, pad=1)
    a2, _ = g.conv(xa, c2.weight)
    assert rel_err(back(a1), y1) < TOL and rel_err(back(a2), y2) < TOL
    seed(g, a1, d1)
    seed(g, a2, d2)
    g.backward()
    assert rel_err(back(g.grad_of(xa)), xr.grad) < TOL
    # tiny map, long K: forward split-K (private slabs, fixed-order sum) and dgrad split-K (red.add)
    c3 = cl(nn.Conv2d(1024, 256, 3, 1, 1, 1, bias=False)).cuda()
    x3 = torch.randn(2, 1024, 6, 5)
    x3r = x3.double().requires_grad_(True)
    y3 = F.conv2d(x3r, c3.weight.double().cpu(), None, 1, 1, 1)
    d3 = torch.randn_like(y3)
    y3.backward(d3)
    g = _g()
    x3a = to_act(g, x3)
    a3, _ = g.conv(x3a, c3.weight, pad=1)
    assert rel_err(back(a3), y3) < TOL
    seed(g, a3, d3)
    g.backward()
    assert rel_err(back(g.grad_of(x3a)), x3r.grad) < TOL
    # dW[o][i][ky][kx] = sum_{n,y,x} d3[n][o][y][x] * x3[n][i][y+ky-1][x+kx-1]  ==  a correlation of x3 with d3
    wref = F.conv2d(x3.double().transpose(0, 1), d3.transpose(0, 1), None, 1, 1, 1).transpose(0, 1)
    assert rel_err(g.param_grads[id(c3.weight)], wref) < TOL


def test_linear_bias_gelu_and_residual():
    torch.manual_seed(2)
    lin1, lin2 = nn.Linear(304, 256).double(), nn.Linear(256, 304).double()
    x = torch.randn(2, 304, 5, 7, dtype=torch.double, requires_grad=True)
    tok = x.flatten(2).transpose(1, 2)
    out = lin2(F.gelu(lin1(tok))) + tok
    dy = torch.randn_like(out)
    out.backward(dy)
    l1, l2 = nn.Linear(304, 256).cuda(), nn.Linear(256, 304).cuda()
    l1.load_state_dict({k: v.float() for k, v in lin1.state_dict().items()})
    l2.load_state_dict({k: v.float() for k, v in lin2.state_dict().items()})
    g = _g()
    from cavp_b200.engine import ACT_GELU
    xa = to_act(g, x.detach())
    h, _ = g.conv(xa, l1.weight, bias=l1.bias, act=ACT_GELU, save_pre=True)
    y, _ = g.conv(h, l2.weight, bias=l2.bias, res=xa)
    ref = out.transpose(1, 2).reshape(2, 304, 5, 7)
    assert rel_err(back(y), ref) < TOL
    seed(g, y, dy.transpose(1, 2).reshape(2, 304, 5, 7))
    g.backward()
    assert rel_err(back(g.grad_of(xa)), x.grad) < TOL
    for a, b in ((l1, lin1), (l2, lin2)):
        assert rel_err(g.param_grads[id(a.weight)], b.weight.grad) < TOL
        assert rel_err(g.param_grads[id(a.bias)], b.bias.grad) < TOL


def test_vgg_tail_conv_bias_relu_pool_flatten_fc_splitk():
    torch.manual_seed(3)
    conv = nn.Conv2d(1, 64, 3, padding=1).double()
    conv2 = nn.Conv2d(64, 128, 3, padding=1).double()
    fc = nn.Linear(128 * 6 * 4, 64).double()
    x = torch.randn(4, 1, 24, 16, dtype=torch.double)
    y = F.max_pool2d(F.relu(conv(x)), 2, 2)
    y = F.max_pool2d(F.relu(conv2(y)), 2, 2)
    flat = y.permute(0, 2, 3, 1).contiguous().view(4, -1)
    out = F.relu(fc(flat))
    dy = torch.randn_like(out)
    out.backward(dy)
    c1, c2, f1 = cl(nn.Conv2d(1, 64, 3, padding=1)).cuda(), cl(nn.Conv2d(64, 128, 3, padding=1)).cuda(), nn.Linear(3072, 64).cuda()
    for a, b in ((c1, conv), (c2, conv2), (f1, fc)):
        a.weight.data.copy_(b.weight.data.float()); a.bias.data.copy_(b.bias.data.float())
    cl(c2)
    g = _g()
    from cavp_b200.engine import ACT_RELU
    xp = torch.zeros(4, 4, 24, 16); xp[:, :1] = x.float()
    xa = to_act(g, xp, needs_grad=False)
    a, _ = g.conv(xa, c1.weight, bias=c1.bias, act=ACT_RELU, pad=1)
    a = g.maxpool(a, 2, 2, 0)
    a, _ = g.conv(a, c2.weight, bias=c2.bias, act=ACT_RELU, pad=1)
    a = g.maxpool(a, 2, 2, 0)
    a = g.flatten(a)
    o, _ = g.conv(a, f1.weight, bias=f1.bias, act=ACT_RELU)
    assert rel_err(o.dense().reshape(4, 64).double().cpu(), out) < TOL
    d, _ = g.grad_target(o)
    d.buf.copy_(dy.float().cuda())
    g.backward()
    for a_, b_ in ((c1, conv), (c2, conv2), (f1, fc)):
        assert rel_err(g.param_grads[id(a_.weight)], b_.weight.grad) < TOL, type(a_)
        assert rel_err(g.param_grads[id(a_.bias)], b_.bias.grad) < TOL, type(a_)


@pytest.mark.parametrize("align", [True, False])
def test_bilinear_and_pools(align):
    torch.manual_seed(4)
    x = torch.randn(2, 8, 7, 5, dtype=torch.double, requires_grad=True)
    up = F.interpolate(x, size=(28, 20), mode="bilinear", align_corners=align)
    mp = F.max_pool2d(up, 3, 2, 1)
    gp = up.mean(dim=(2, 3), keepdim=True)
    dmp, dgp = torch.randn_like(mp), torch.randn_like(gp)
    (mp * dmp).sum().backward(retain_graph=True)
    g1 = x.grad.clone(); x.grad = None
    (gp * dgp).sum().backward()
    g2 = x.grad.clone()
    g = _g()
    xa = to_act(g, x.detach())
    ua = g.bilinear(xa, 28, 20, align)
    ma = g.maxpool(ua, 3, 2, 1)
    ga = g.global_avgpool(ua)
    assert rel_err(back(ua), up) < TOL
    assert rel_err(back(ma), mp) < TOL
    assert rel_err(back(ga), gp) < TOL
    seed(g, ma, dmp)
    seed(g, ga, dgp)
    g.backward()
    assert rel_err(back(g.grad_of(xa)), g1 + g2) < TOL


def test_upsample_to_nchw_and_backward():
    torch.manual_seed(5)
    nc = 22
    x = torch.randn(4, 24, 14, 14, dtype=torch.double); x[:, nc:] = 0
    xr = x[:, :nc].clone().requires_grad_(True)
    pred = F.interpolate(xr, size=(56, 56), mode="bilinear", align_corners=False)
    d = torch.randn_like(pred); d[2:] = 0
    pred.backward(d)
    g = _g()
    xa = to_act(g, x)
    p = g.upsample_to_nchw(xa, nc, 56, 56)
    assert rel_err(p, pred) < TOL
    g.upsample_to_nchw_backward(xa, nc, d[:2].float().cuda().contiguous(), n_valid=2)
    got = back(g.grad_of(xa))
    assert rel_err(got[:, :nc], xr.grad) < TOL
    assert float(got[:, nc:].abs().max()) == 0.0


def test_layernorm_gate_block():
    """attn.py:146-162 live branch on tiny shapes, against the oracle restatement."""
    from oracle import cavp_oracle as O
    torch.manual_seed(6)
    B, rows, hh, ww, C = 2, 4, 5, 6, 304
    sd = {}
    for name, shape in [("norm1.weight", (C,)), ("norm1.bias", (C,)), ("attn.q.weight", (C, C)), ("attn.k.weight", (C, C)),
                        ("attn.v.weight", (C, C)), ("attn.proj.weight", (C, C)), ("attn.proj.bias", (C,))]:
        t = torch.randn(shape, dtype=torch.double) * (0.05 if len(shape) == 2 else 0.3)
        if name == "norm1.weight":
            t = t + 1
        sd[name] = t.requires_grad_(True)
    fv = torch.randn(B, hh * ww, C, dtype=torch.double, requires_grad=True)
    fa = torch.randn(rows, 1, C, dtype=torch.double, requires_grad=True)
    st = O.State(sd, True)
    fvn = O.layer_norm(st.sub("norm1"), fv)
    fan = O.layer_norm(st.sub("norm1"), fa)
    o, attn = O.attention(st.sub("attn"), torch.cat((fvn, fvn)), fan, fan)
    f1 = torch.cat((fvn, fvn)) + o
    dy = torch.randn_like(f1)
    f1.backward(dy)

    g = _g()
    ln = nn.LayerNorm(C).cuda()
    ln.weight.data.copy_(sd["norm1.weight"].detach().float()); ln.bias.data.copy_(sd["norm1.bias"].detach().float())
    lins = {}
    for n_ in ("q", "k", "v", "proj"):
        lin = nn.Linear(C, C, bias=(n_ == "proj")).cuda()
        lin.weight.data.copy_(sd[f"attn.{n_}.weight"].detach().float())
        if n_ == "proj":
            lin.bias.data.copy_(sd["attn.proj.bias"].detach().float())
        lins[n_] = lin
    fva = to_act(g, fv.detach().transpose(1, 2).reshape(B, C, hh, ww))
    faa = to_act(g, fa.detach().transpose(1, 2).reshape(rows, C, 1, 1))
    a_fvn = g.layernorm(fva, ln)
    a_fan = g.layernorm(faa, ln)
    q, _ = g.conv(a_fvn, lins["q"].weight)
    k, _ = g.conv(a_fan, lins["k"].weight)
    v, _ = g.conv(a_fan, lins["v"].weight)
    x, at = g.gate(q, k, v, heads=4)
    y, _ = g.conv(x, lins["proj"].weight, bias=lins["proj"].bias, res=a_fvn, res_mod=a_fvn.rows)
    ref = f1.transpose(1, 2).reshape(rows, C, hh, ww)
    assert rel_err(back(y), ref) < TOL
    assert rel_err(at.unsqueeze(-1), attn) < TOL
    seed(g, y, dy.transpose(1, 2).reshape(rows, C, hh, ww))
    g.backward()
    assert rel_err(back(g.grad_of(fva)), fv.grad.transpose(1, 2).reshape(B, C, hh, ww)) < TOL
    assert rel_err(back(g.grad_of(faa)), fa.grad.transpose(1, 2).reshape(rows, C, 1, 1)) < TOL
    assert rel_err(g.param_grads[id(ln.weight)], sd["norm1.weight"].grad) < TOL
    assert rel_err(g.param_grads[id(ln.bias)], sd["norm1.bias"].grad) < TOL
    for n_ in ("q", "k", "v", "proj"):
        assert rel_err(g.param_grads[id(lins[n_].weight)], sd[f"attn.{n_}.weight"].grad) < TOL, n_
    assert rel_err(g.param_grads[id(lins["proj"].bias)], sd["attn.proj.bias"].grad) < TOL


def test_cross_entropy_module():
    from cavp_b200.loss import CrossEntropyLoss
    torch.manual_seed(7)
    logits = torch.randn(3, 22, 17, 19, dtype=torch.double, requires_grad=True)
    labels = torch.randint(0, 22, (3, 17, 19)); labels[0, :3, :4] = 255
    ref = F.cross_entropy(logits, labels, ignore_index=255)
    (ref * 1.7).backward()
    lg = logits.detach().float().cuda().requires_grad_(True)
    loss = CrossEntropyLoss(255)(lg, labels.cuda())
    (loss * 1.7).backward()
    assert abs(float(loss) - float(ref)) < TOL * abs(float(ref))
    assert rel_err(lg.grad, logits.grad) < TOL


def test_contrast_loss_module_matches_oracle():
    from cavp_b200.loss import ContrastLoss
    from oracle import cavp_oracle as O
    torch.manual_seed(8)
    B, C, h, w = 3, 304, 12, 12
    em = torch.randn(B, C, h, w, dtype=torch.double, requires_grad=True)
    es = torch.randn(B, C, h, w, dtype=torch.double, requires_grad=True)
    gt = torch.zeros(B, 48, 48, dtype=torch.int64)
    gt[0, 8:40, 8:40] = 3; gt[1, 4:44, 10:30] = 5; gt[2, 10:30, 4:44] = 3; gt[:, :2, :2] = 255
    gsh = gt.clone(); gsh[1] = 0
    torch.manual_seed(99)
    ref = O.contrast_loss(em, gt, es, gsh, max_views=32)
    ref.backward()
    em32 = em.detach().float().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    es32 = es.detach().float().cuda().requires_grad_(True)
    torch.manual_seed(99)
    loss = ContrastLoss(0.1, 255, 32)(em32, gt.cuda(), es32, gsh.cuda())
    loss.backward()
    assert abs(float(loss) - float(ref)) < TOL * abs(float(ref))
    assert rel_err(em32.grad, em.grad) < 1e-4
    assert rel_err(es32.grad, es.grad) < 1e-4


def test_contrast_loss_empty_selection_returns_zero():
    from cavp_b200.loss import ContrastLoss
    em = torch.randn(2, 304, 8, 8, device="cuda")
    gt = torch.zeros(2, 32, 32, dtype=torch.int64, device="cuda")
    out = ContrastLoss(0.1, 255, 512)(em, gt, em, gt)
    assert out.shape == (1,) and float(out) == 0.0


@pytest.mark.parametrize("nc,hin,win,hout,wout", [(22, 14, 14, 56, 56), (71, 9, 13, 36, 52), (2, 7, 5, 23, 17)])
def test_fused_upsample_cross_entropy_matches_torch(nc, hin, win, hout, wout):
    """cavp_upsample_ce_fwd / bwd (the train-step path that never writes the full-resolution logits) against
    F.interpolate(align_corners=False) + F.cross_entropy(ignore_index=255) in fp64, on the first B of 2B rows
    (trainer_cavp_vpo_mono.py:171: the shuffled half is weighted by zero)."""
    from cavp_b200 import _C
    from cavp_b200.engine import pad4
    torch.manual_seed(9)
    B, rows, cp = 3, 6, pad4(nc)
    x = torch.randn(rows, cp, hin, win, dtype=torch.double); x[:, nc:] = 0
    xr = x[:B, :nc].clone().requires_grad_(True)
    labels = torch.randint(0, nc, (B, hout, wout)); labels[0, :3, :4] = 255; labels[1, 5, 5] = 255
    ref = F.cross_entropy(F.interpolate(xr, size=(hout, wout), mode="bilinear", align_corners=False), labels,
                          ignore_index=255)
    ref.backward()
    g = _g()
    xa = to_act(g, x)
    lab = labels.cuda()
    lse = g.empty(B * hout * wout)
    partials = g.empty(_C.query("cavp_ce_nblocks", B, hout * wout), 2)
    out = g.empty(2)
    g.call("cavp_upsample_ce_fwd", xa.ptr, xa.ld, hin, win, hout, wout, B, nc, lab.data_ptr(), 255, lse.data_ptr(),
           partials.data_ptr(), out.data_ptr())
    assert abs(float(out[0]) - float(ref)) < TOL * abs(float(ref))
    assert int(out[1]) == int((labels != 255).sum())
    dx, acc = g.grad_target(xa)
    dx.buf.fill_(float("nan"))  # the kernel must write every element (pad channels and rows >= B as zeros)
    g.call("cavp_upsample_ce_bwd", xa.ptr, xa.ld, hin, win, hout, wout, rows, B, nc, cp, lab.data_ptr(), 255,
           lse.data_ptr(), out.data_ptr(), 0, dx.ptr, dx.ld)
    got = back(dx)
    assert rel_err(got[:B, :nc], xr.grad) < TOL
    assert float(got[:B, nc:].abs().max() if cp > nc else 0.0) == 0.0 and float(got[B:].abs().max()) == 0.0
    # same loss value as the materialised path (bilinear kernel + CE kernel) to fp32 rounding of the reduction order
    p = g.upsample_to_nchw(xa, nc, hout, wout)
    from cavp_b200.loss import ce_forward
    ce2 = ce_forward(g, p.data_ptr(), lab, B, nc, hout * wout, 255)
    assert abs(float(ce2[0]) - float(out[0])) < 1e-6 * abs(float(out[0]))


def test_cross_entropy_out_of_range_labels_are_ignored_not_dereferenced():
    from cavp_b200.loss import CrossEntropyLoss
    torch.manual_seed(8)
    logits = torch.randn(2, 5, 9, 9, dtype=torch.double, requires_grad=True)
    labels = torch.randint(0, 5, (2, 9, 9))
    bad = labels.clone(); bad[0, 0, :4] = 254; bad[1, 2, 2] = -3
    ok = labels.clone(); ok[0, 0, :4] = 255; ok[1, 2, 2] = 255
    ref = F.cross_entropy(logits, ok, ignore_index=255)
    ref.backward()
    lg = logits.detach().float().cuda().requires_grad_(True)
    loss = CrossEntropyLoss(255)(lg, bad.to(torch.int32).cuda())  # int32 labels are widened, not reinterpreted
    loss.backward()
    assert abs(float(loss) - float(ref)) < TOL * abs(float(ref))
    assert rel_err(lg.grad, logits.grad) < TOL
