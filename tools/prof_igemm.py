"""Single-shape driver for ncu captures of the tcgen05 tile kernel (fwd + wgrad), prec 2 by default."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cavp_b200 import _C
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
nimg, h, c, cout, r, pad = 64, 56, 256, 256, 3, 1
xh = torch.randn(nimg, h, h, c, device=dev); wh = torch.randn(cout, r * r * c, device=dev)
M = nimg * h * h
y = torch.empty(M, cout, device=dev)
dw = torch.zeros(cout, r * r * c, device=dev)
sp = torch.empty(2, cout, r * r * c, device=dev)
_C.call("cavp_split_tf32", _C.ptr(wh), _C.ptr(sp[0]), _C.ptr(sp[1]), wh.numel(), _C.stream())
LO_OFF = sp[0].numel()
for _ in range(4):
    _C.call("cavp_igemm", _C.ptr(xh), _C.ptr(sp[0]), _C.ptr(y), 0, 0, 0, 0, 0, nimg, h, h, c, c, h, h, r, r, 1, pad, 1, 0,
            cout, r * r * c, cout, 0, 0, 0, cout, 0, 0.0, 1, prec, LO_OFF, _C.stream())
    _C.call("cavp_igemm_wgrad", _C.ptr(y), _C.ptr(xh), _C.ptr(dw), nimg, h, h, c, c, h, h, r, r, 1, pad, 1, cout, cout, 9,
            prec, _C.stream())
torch.cuda.synchronize()
print("done")
