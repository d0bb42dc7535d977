"""Host-side view of train_step + optimisers (kernel launches are async): issue time, implicit syncs, cProfile."""
import cProfile, pstats, os, sys, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cavp_b200.models.cavp_model import CAVP
from cavp_b200.trainer import train_step

dev = torch.device("cuda", 0)
B = 32
model = CAVP(50, None, num_classes=bench.CFG["nc"], args=bench.make_args(B, 0, 2), in_plane=1).to(dev).train()
audio_params = list(model.audio_backbone.backbone.parameters())
ids = {id(p) for p in audio_params}
from cavp_b200.optim import SGD, Adam
opt_v = SGD([p for p in model.parameters() if id(p) not in ids], lr=1e-3, momentum=0.9, weight_decay=5e-4)
opt_a = Adam(audio_params, lr=1e-4)
image, audio, pix, spl = bench.synthetic_batch(B, 666)
image, audio, pixd = image.to(dev), audio.to(dev), pix.to(dev)

def step():
    opt_v.zero_grad(set_to_none=True); opt_a.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    train_step(model, image, audio, pix, spl, labels_dev=pixd)
    t1 = time.perf_counter()
    opt_v.step(); opt_a.step()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1

for _ in range(3):
    step()
torch.cuda.synchronize()
# pure host cost: start every step with an empty launch queue
pure = []
for _ in range(3):
    torch.cuda.synchronize()
    pure.append(step())
print("issue ms per step with an EMPTY queue: train_step", [round(1e3 * a, 1) for a, b in pure], "optim",
      [round(1e3 * b, 1) for a, b in pure])
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
ts = [step() for _ in range(4)]
torch.cuda.set_sync_debug_mode("default")
t0 = time.perf_counter(); torch.cuda.synchronize(); drain = time.perf_counter() - t0
print("issue ms per step: train_step", [round(1e3 * a, 1) for a, b in ts], "optim", [round(1e3 * b, 1) for a, b in ts],
      "drain after 4 steps ms", round(1e3 * drain, 1))
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
