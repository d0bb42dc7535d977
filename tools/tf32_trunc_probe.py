"""Does tcgen05.mma kind::tf32 truncate or round the low 13 mantissa bits of an fp32 operand?  (diagnostic tool)
B is handed to the kernel raw (TMA, no split kernel) at PREC=1; A holds exact small integers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cavp_b200 import _C
dev = "cuda"
M, N, K = 128, 128, 32
a = torch.zeros(M, K, device=dev); a[:, 0] = 1.0
for name, val in [("1+2^-11+2^-12 (RN up, trunc down)", 1 + 2**-11 + 2**-12), ("1+2^-12 (both down)", 1 + 2**-12),
                  ("1+2^-10+2^-11+2^-13 (RN up)", 1 + 2**-10 + 2**-11 + 2**-13), ("-(1+2^-11+2^-12)", -(1 + 2**-11 + 2**-12))]:
    b = torch.zeros(N, K, device=dev); b[:, 0] = val
    lo = torch.zeros(N, K, device=dev)
    both = torch.stack([b, lo]).contiguous()
    y = torch.zeros(M, N, device=dev)
    _C.call("cavp_igemm", _C.ptr(a), _C.ptr(both[0]), _C.ptr(y), 0, 0, 0, 0, 0, M, 1, 1, K, K, 1, 1, 1, 1, 1, 0, 1, 0, N, K, N,
            0, 0, 0, N, 0, 0.0, 1, 1, both[0].numel(), _C.stream())
    torch.cuda.synchronize()
    got = float(y[0, 0])
    tr = torch.tensor(val, dtype=torch.float32).view(torch.int32)
    trunc = float((tr & ~0x1FFF).view(torch.float32))
    rn = float(((tr + 0x1000) & ~0x1FFF).view(torch.float32))
    print(f"{name:40s} got {got!r}  trunc {trunc!r}  rn {rn!r}  ->", "TRUNC" if got == trunc else ("RN" if got == rn else "OTHER"))
