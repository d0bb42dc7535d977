"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python oracle/make_golden.py
The reference has no tests/golden vectors of its own, so these fixtures are the pin:
the restated oracle (oracle/cavp_oracle.py) and the CUDA product are both checked against them.
Fixtures hold no parameters: state and inputs are regenerated from oracle/seeded.py (per-key seeds).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

from oracle import ref_harness, seeded  # noqa: E402

CONFIGS = {
    # name: dict(B, H, W, nc, dilation, audio, in_plane, frames, train, max_views, audio_func)
    "tiny_train": dict(B=4, H=64, W=64, nc=22, dilation=(False, True, True), audio="vgg", in_plane=1, frames=96,
                       train=True, max_views=64, audio_func=False),
    "tiny_train_fff71": dict(B=3, H=64, W=96, nc=71, dilation=(False, False, False), audio="vgg", in_plane=1,
                             frames=96, train=True, max_views=64, audio_func=False),
    "tiny_train_r18_stereo": dict(B=4, H=64, W=64, nc=22, dilation=(False, True, True), audio="18", in_plane=2,
                                  frames=96, train=True, max_views=64, audio_func=True),
    "cfgA_eval_224": dict(B=1, H=224, W=224, nc=71, dilation=(False, False, False), audio="vgg", in_plane=1,
                          frames=96, train=False, max_views=512, audio_func=False),
    "cfgB_train_224": dict(B=3, H=224, W=224, nc=22, dilation=(False, True, True), audio="vgg", in_plane=1,
                           frames=96, train=True, max_views=512, audio_func=False),
}

SAMPLE_POINTS = 48


def sample_idx(numel, n=SAMPLE_POINTS):
    if numel <= n:
        return torch.arange(numel)
    return (torch.arange(n, dtype=torch.int64) * (numel - 1)) // (n - 1)


def summarize(t):
    f = t.detach().flatten().double()
    return dict(norm=float(f.norm()), sum=float(f.sum()), absmax=float(f.abs().max()),
                samples=t.detach().flatten()[sample_idx(t.numel())].clone(), shape=tuple(t.shape))


def run(name, cfg):
    torch.manual_seed(0)
    model = ref_harness.build_reference_cavp(cfg["nc"], cfg["dilation"], cfg["audio"], cfg["in_plane"],
                                             batch_size=cfg["B"])
    seeded.fill_module_(model, seed=0)
    batch = seeded.synthetic_batch(cfg["B"], cfg["H"], cfg["W"], cfg["nc"], seed=666,
                                   audio_frames=cfg["frames"], in_plane=cfg["in_plane"])
    out = dict(config=cfg, torch_version=torch.__version__)
    B = cfg["B"]
    if not cfg["train"]:
        model.eval()
        with torch.no_grad():
            pred, fusion, pack = model(batch["image"], batch["audio"][:B], eval_mode=True)
        top2 = pred.topk(2, dim=1).values
        out.update(
            pred_stride4=pred[:, :, ::4, ::4].clone(),
            argmax=pred.argmax(1).to(torch.uint8),
            margin=(top2[:, 0] - top2[:, 1]).half(),
            fusion_stride2=fusion[:, :, ::2, ::2].clone(),
            attn_v=pack["attn_v"].clone(),
            pred_summary=summarize(pred), fusion_summary=summarize(fusion),
        )
        return out

    model.train()
    ce, ctr = ref_harness.reference_losses(cfg["max_views"])
    shuffle_pix_label = seeded.shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    if cfg["audio_func"]:
        # trainer_cavp_vpo_stereo.py:211: audio batch is B, shuffle at feature level
        info = {"shuffle_idx": batch["shuffle_idx"], "mod_idx_map": None, "image_label": batch["img_label"].clone()}
        pred, fusion, pack = model(batch["image"], batch["audio"][:B], info, False, audio_func=True)
    else:
        pred, fusion, pack = model(batch["image"], batch["audio"], None, False)
    output = pred[:B] + pred[B:] * 0.0
    torch.manual_seed(1234)  # pins the randperm stream of ContrastLoss
    l_ctr = ctr(fusion[:B], batch["pix_label"], fusion[B:], shuffle_pix_label)
    l_ce = ce(output, batch["pix_label"])
    (l_ce + l_ctr).backward()

    ps = 4 if pred.numel() > 500_000 else 1  # spatial strides of the stored slices (kept in the fixture)
    fs = 4 if fusion.numel() > 3_000_000 else (2 if fusion.numel() > 400_000 else 1)
    at = 4 if pack["attn_v"].numel() > 100_000 else 1
    out.update(
        l_ce=float(l_ce.detach()), l_ctr=float(l_ctr.detach().reshape(-1)[0]),
        pred_stride=ps, fusion_stride=fs, attn_stride=at,
        pred=pred[:, :, ::ps, ::ps].detach().clone(),
        fusion=fusion[:, :, ::fs, ::fs].detach().clone(),
        attn_v=pack["attn_v"][:, :, ::at].detach().clone(),
        visual=summarize(pack["visual"]), audio=pack["audio"].detach().clone(),
        pred_summary=summarize(pred), fusion_summary=summarize(fusion),
        grads={k: (None if p.grad is None else summarize(p.grad)) for k, p in model.named_parameters()},
        buffers={k: v.detach().clone() for k, v in model.state_dict().items()
                 if k.endswith("running_mean") or k.endswith("running_var")},
    )
    return out


def main():
    only = sys.argv[1:]
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        res = run(name, cfg)
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(res, path)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB",
              {k: res[k] for k in ("l_ce", "l_ctr") if k in res})


if __name__ == "__main__":
    main()
