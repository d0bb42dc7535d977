"""fp32-vs-fp64 self-discrepancy of the reference arithmetic (oracle) on the golden train fixtures.

End-to-end gradients of the random-weight, batch-stat-BN network are ill-conditioned: running the SAME oracle in fp64
instead of fp32 changes per-tensor gradients by a few percent (ReLU / max-pool decisions flip).  This script measures
that and writes tests/golden/grad_sensitivity.json; tests/test_parity_gpu.py scales its end-to-end gradient tolerances
from it (elementwise gradient parity is asserted per op in tests/test_ops_gpu.py).
Usage: python tools/grad_sensitivity.py [fixture ...]   (build container, CPU, ~5 min for all four)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_err  # noqa: E402
from oracle import cavp_oracle as O, schema, seeded  # noqa: E402
from oracle.make_golden import sample_idx  # noqa: E402


def run(cfg, dtype):
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0)
    sd = {k: (v.to(dtype).requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    batch = seeded.synthetic_batch(cfg["B"], cfg["H"], cfg["W"], cfg["nc"], seed=666, audio_frames=cfg["frames"],
                                   in_plane=cfg["in_plane"])
    B = cfg["B"]
    spl = seeded.shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    kw = dict(dilation_flags=cfg["dilation"], audio_kind=cfg["audio"], train=True)
    if cfg["audio_func"]:
        out_cat, ctr_cat, pack, _ = O.cavp_forward(sd, batch["image"].to(dtype), batch["audio"][:B].to(dtype),
                                                   shuffle_idx=batch["shuffle_idx"], audio_func=True, **kw)
    else:
        out_cat, ctr_cat, pack, _ = O.cavp_forward(sd, batch["image"].to(dtype), batch["audio"].to(dtype), **kw)
    output = out_cat[:B] + out_cat[B:] * 0.0
    torch.manual_seed(1234)
    l_ctr = O.contrast_loss(ctr_cat[:B], batch["pix_label"], ctr_cat[B:], spl, cfg["max_views"])
    l_ce = O.cross_entropy(output, batch["pix_label"])
    (l_ce + l_ctr).backward()
    return out_cat, ctr_cat, sd


def main(names):
    path = os.path.join(ROOT, "tests", "golden", "grad_sensitivity.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        cfg = load_golden(name)["config"]
        o64, f64, sd64 = run(cfg, torch.float64)
        o32, f32, sd32 = run(cfg, torch.float32)
        worst_l2, worst_norm = (0.0, ""), (0.0, "")
        for k in sd64:
            if not sd64[k].is_floating_point() or sd64[k].grad is None:
                continue
            a, b = sd32[k].grad.double().flatten(), sd64[k].grad.flatten()
            idx = sample_idx(a.numel())
            l2s = float((a[idx] - b[idx]).norm() / b[idx].norm().clamp_min(1e-30))
            nd = abs(float(a.norm() - b.norm())) / float(b.norm())
            worst_l2, worst_norm = max(worst_l2, (l2s, k)), max(worst_norm, (nd, k))
        out[name] = {"pred": rel_err(o32, o64), "fusion": rel_err(f32, f64), "worst_sample_l2": worst_l2[0],
                     "worst_sample_l2_key": worst_l2[1], "worst_norm": worst_norm[0], "worst_norm_key": worst_norm[1]}
        print(name, out[name], flush=True)
        json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny_train", "tiny_train_fff71", "tiny_train_r18_stereo", "cfgB_train_224"])
