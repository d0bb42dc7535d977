"""Generate tests/golden/metrics.pt by running the UNMODIFIED reference metric classes (utils/eval_utils.py) on CPU.

Build container only:   python oracle/make_golden_metrics.py
torchmetrics (imported at the top of eval_utils.py, unused by MIoU / ForegroundDetect) is absent here and stubbed.
ForegroundDetect.get_metric_results moves the matrix to .cuda(): only its CPU parts (__call__, _fast_hist) are run.
"""
import os
import sys
import types

import numpy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import metrics_oracle as MO  # noqa: E402

CASES = [dict(seed=1, B=2, C=22, H=24, W=40), dict(seed=2, B=1, C=71, H=32, W=32), dict(seed=3, B=3, C=2, H=16, W=16),
         dict(seed=4, B=2, C=24, H=20, W=28, ignore_frac=0.0)]


def main():
    if "torchmetrics" not in sys.modules:
        try:
            import torchmetrics  # noqa: F401
        except Exception:
            sys.modules["torchmetrics"] = types.ModuleType("torchmetrics")
    sys.path.insert(0, "/root/reference")
    from utils.eval_utils import MIoU, ForegroundDetect
    out = []
    for c in CASES:
        logits, target = MO.metric_case(**c)
        C = c["C"]
        m = MIoU(C, 255, "cpu")
        res1 = m(logits.clone(), target.clone())
        logits2, target2 = MO.metric_case(**{**c, "seed": c["seed"] + 100})
        res2 = m(logits2.clone(), target2.clone())          # accumulation over two calls
        sample = MIoU(C, 255, "cpu").calculate_current_sample(logits.clone(), target.clone())
        fg = ForegroundDetect(num_classes=C, local_rank="cpu")
        fg(logits.clone(), target.clone())
        fg(logits2.clone(), target2.clone())
        out.append(dict(case=c, miou_after_1=[float(x) for x in res1], miou_after_2=[float(x) for x in res2],
                        sample=[torch.as_tensor(numpy.asarray(x, dtype=numpy.float64)) for x in sample],
                        iou=torch.as_tensor(numpy.asarray(m.iou, dtype=numpy.float64)),
                        confusion=torch.as_tensor(numpy.asarray(fg.confusion_matrix_, dtype=numpy.float64))))
    path = os.path.join(ROOT, "tests", "golden", "metrics.pt")
    torch.save(dict(torch_version=torch.__version__, cases=out), path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
