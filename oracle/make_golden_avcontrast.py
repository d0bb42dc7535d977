"""Generate tests/golden/avcontrast.pt from the UNMODIFIED reference loss/av_contrast.py:AVContrast on CPU.

The class hard-codes `.cuda(self.local_rank)` (:51,104); for the duration of the call torch.Tensor.cuda is patched to
the identity so that the reference code itself runs here (no GPU in the build container).
Build container only:   python oracle/make_golden_avcontrast.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import avcontrast_oracle as AO  # noqa: E402

CASES = [dict(seed=21, b=4, c=304, H=64, W=96), dict(seed=22, b=6, c=64, H=128, W=128, empty=(2,)),
         dict(seed=23, b=3, c=304, H=224, W=224, empty=(0, 1, 2))]


def main():
    sys.path.insert(0, "/root/reference")
    from loss.av_contrast import AVContrast
    out = []
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for c in CASES:
            f_v, f_a, labels = AO.case(**c)
            f_v = f_v.double().requires_grad_(True)
            f_a = f_a.double().requires_grad_(True)
            loss = AVContrast(0.1, "cpu")(f_v, f_a, labels)
            if loss.requires_grad:
                loss.backward()
            gv = f_v.grad if f_v.grad is not None else torch.zeros_like(f_v)
            ga = f_a.grad if f_a.grad is not None else torch.zeros_like(f_a)
            idx = (torch.arange(2048, dtype=torch.int64) * (gv.numel() - 1)) // 2047
            out.append(dict(case=c, loss=float(loss), grad_fa=ga.float(), grad_fv_idx=idx,
                            grad_fv_samples=gv.flatten()[idx].float(), grad_fv_norm=float(gv.norm())))
    finally:
        torch.Tensor.cuda = orig
    path = os.path.join(ROOT, "tests", "golden", "avcontrast.pt")
    torch.save(dict(torch_version=torch.__version__, cases=out), path)
    print("wrote", path, os.path.getsize(path), [o["loss"] for o in out])


if __name__ == "__main__":
    main()
