"""Full-size (BASELINE.json configs[1], bs32) checks of the tensor-core path through size-independent properties - the
CPU oracle cannot run these shapes in seconds.

1. Adjoint identities.  Forward, data gradient and weight gradient are three different kernels / schedules; for a linear
   map y = conv(x, w) they must satisfy   <conv(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(dy, x)>   (fp64 dot products
   of fp32 tensors).  A tile, stride, dilation, padding or split-K mistake in any one of them breaks the equality.
2. Schedule independence.  The one-tile-per-CTA kernel, the persistent kernel and the CTA-pair kernel must agree on the
   same inputs to fp32 rounding (they accumulate in the same 64-wide K units).
Tolerance: 1e-5 relative (the products are fp32-grade: 3xTF32 + promotion); the north-star tolerance is 1e-3.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (name, nimg, h, cin, cout, k, stride, pad, dil): decoder conv, ASPP dilated conv, stride-2 bottleneck conv, MLP linear
SHAPES = [
    ("decoder 3x3 304->256 @64x56^2", 64, 56, 304, 256, 3, 1, 1, 1),
    ("aspp 3x3 d12 2048->256 @32x28^2", 32, 28, 2048, 256, 3, 1, 12, 12),
    ("layer2 3x3 s2 128->128 @32x56^2", 32, 56, 128, 128, 3, 2, 1, 1),
    ("mlp fc1 304->1216 @64x3136", 64, 56, 304, 1216, 1, 1, 0, 1),
    ("stem 3x3 64->128 @32x112^2", 32, 112, 64, 128, 3, 1, 1, 1),
]


def _run(g, x, w, dy, stride, pad, dil):
    from cavp_b200.engine import Act
    y, _ = g.conv(x, w, stride=stride, pad=pad, dil=dil)
    d, acc = g.grad_target(y)
    assert not acc
    d.buf.copy_(dy)
    g.backward()
    return y, g.grad_of(x), g.param_grads[id(w)]


@pytest.mark.parametrize("prec,tol", [(2, 1e-5), (1, 3e-3)])
@pytest.mark.parametrize("name,nimg,h,cin,cout,k,stride,pad,dil", SHAPES)
def test_adjoint_identities_at_full_size(name, nimg, h, cin, cout, k, stride, pad, dil, prec, tol):
    """prec 2 = the fp32-parity mode (3xTF32 + promotion); prec 1 = plain TF32 (`bench.py --prec 1`), same kernels with
    one MMA per product - covered here so that the reduced-precision instantiations stay correct too."""
    from cavp_b200.engine import Graph, new_act
    dev = torch.device("cuda")
    torch.manual_seed(0)
    g = Graph(dev, prec=prec, train=True)
    x = new_act(nimg, h, h, cin, dev)
    x.buf.normal_()
    w = torch.nn.Parameter((torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
                           .contiguous(memory_format=torch.channels_last))
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn(nimg * ho * ho, cout, device=dev)
    y, dx, dw = _run(g, x, w, dy, stride, pad, dil)
    torch.cuda.synchronize()
    a = torch.dot(y.buf.double().flatten(), dy.double().flatten())
    b = torch.dot(x.buf.double().flatten(), dx.buf.double().flatten())
    c = torch.dot(w.detach().double().flatten(), dw.double().flatten())
    scale = (y.buf.double().norm() * dy.double().norm()).item()
    assert abs(a - b).item() < tol * scale, (name, float(a), float(b))
    assert abs(a - c).item() < tol * scale, (name, float(a), float(c))


def test_schedules_agree_at_full_size():
    """tile / persistent / CTA-pair kernels on the decoder conv and the MLP linear (separate processes: the schedule
    override CAVP_IGEMM_WS is read once per process)."""
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from cavp_b200.engine import Graph, new_act
dev = torch.device("cuda")
out = {}
for name, nimg, h, cin, cout, k, pad in (("dec", 64, 56, 304, 256, 3, 1), ("fc", 64, 56, 304, 1216, 1, 0)):
    torch.manual_seed(1)
    g = Graph(dev, prec=2, train=False)
    x = new_act(nimg, h, h, cin, dev); x.buf.normal_()
    w = torch.nn.Parameter((torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
                           .contiguous(memory_format=torch.channels_last))
    y, _ = g.conv(x, w, pad=pad)
    torch.cuda.synchronize()
    out[name] = y.buf.cpu()
torch.save(out, sys.argv[1])
''' % ROOT
    outs = []
    for sched in ("0", "1", "2"):
        path = f"/tmp/cavp_sched_{sched}.pt"
        env = dict(os.environ, CAVP_IGEMM_WS=sched)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        outs.append(torch.load(path))
    for name in ("dec", "fc"):
        ref = outs[0][name].double()
        for o in outs[1:]:
            err = (o[name].double() - ref).abs().max() / ref.abs().max()
            assert err < 2e-6, (name, float(err))
