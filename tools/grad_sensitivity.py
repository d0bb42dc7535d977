import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import load_golden, rel_err
from oracle import cavp_oracle as O, schema, seeded
from oracle.make_golden import sample_idx
name = sys.argv[1]
g = load_golden(name); cfg = g["config"]
def run(dtype):
    sd = schema.seeded_state(cfg["nc"], cfg["audio"], cfg["in_plane"], seed=0)
    sd = {k: (v.to(dtype).requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    batch = seeded.synthetic_batch(cfg["B"], cfg["H"], cfg["W"], cfg["nc"], seed=666, audio_frames=cfg["frames"], in_plane=cfg["in_plane"])
    B = cfg["B"]
    spl = seeded.shuffled_labels(batch["pix_label"], batch["img_label"], batch["shuffle_idx"])
    out_cat, ctr_cat, pack, newbuf = O.cavp_forward(sd, batch["image"].to(dtype), batch["audio"].to(dtype), dilation_flags=cfg["dilation"], audio_kind=cfg["audio"], train=True)
    output = out_cat[:B] + out_cat[B:] * 0.0
    torch.manual_seed(1234)
    l_ctr = O.contrast_loss(ctr_cat[:B], batch["pix_label"], ctr_cat[B:], spl, cfg["max_views"])
    l_ce = O.cross_entropy(output, batch["pix_label"])
    (l_ce + l_ctr).backward()
    return out_cat, ctr_cat, sd
o64, f64, sd64 = run(torch.float64)
o32, f32, sd32 = run(torch.float32)
print("pred", rel_err(o32, o64), "fusion", rel_err(f32, f64))
errs = []
for k in sd64:
    if not sd64[k].is_floating_point() or sd64[k].grad is None: continue
    a, b = sd32[k].grad.double(), sd64[k].grad
    errs.append((float((a-b).abs().max()/b.abs().max()), abs(float(a.norm()-b.norm()))/float(b.norm()), k))
errs.sort(reverse=True)
for e in errs[:8]: print("%.2e %.2e %s" % e)
import statistics; print("median", statistics.median(e[0] for e in errs))
print("---- robust metrics (fp32 vs fp64 oracle)")
rows = []
for k in sd64:
    if not sd64[k].is_floating_point() or sd64[k].grad is None: continue
    a, b = sd32[k].grad.double().flatten(), sd64[k].grad.flatten()
    idx = sample_idx(a.numel())
    l2 = float((a-b).norm()/b.norm()); l2s = float((a[idx]-b[idx]).norm()/b[idx].norm())
    cos = float(torch.dot(a,b)/(a.norm()*b.norm()))
    rows.append((l2, l2s, 1-cos, k))
rows.sort(reverse=True)
for r in rows[:10]: print("relL2 %.2e  sampleL2 %.2e  1-cos %.2e  %s" % r)
print("worst sampleL2", max(r[1] for r in rows), "median relL2", statistics.median(r[0] for r in rows))
print("worst norm discrepancy", max((abs(float(sd32[k].grad.double().norm() - sd64[k].grad.norm())) / float(sd64[k].grad.norm()), k)
      for k in sd64 if sd64[k].is_floating_point() and sd64[k].grad is not None))
