"""GPU probe for the tcgen05 implicit-GEMM kernels: prints error tables against torch CPU fp64 (diagnostic tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cavp_b200 import _C

dev = "cuda"
results = []

def relerr(got, ref):
    ref = ref.double(); got = got.double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))

TMA = [False]

def presplit(w):
    sp = torch.empty(2, *w.shape, device=dev)
    _C.call("cavp_split_tf32", _C.ptr(w), _C.ptr(sp[0]), _C.ptr(sp[1]), w.numel(), _C.stream())
    return sp

def igemm(x, w, nimg, hs, ws, c, ho, wo, r, s, stride, pad, dil, dgrad, ncols, prec, splits=1, scale=None, shift=None,
          res=None, act=0, stats=False, y_pre=False, ldy=None):
    lo_off = 0
    ldw_ = w.shape[-1] if w.dim() == 2 else r * s * c
    if TMA[0]:
        sp = presplit(w.contiguous())
        w, lo_off = sp[0], sp[0].numel()
    M = nimg * ho * wo
    ldy = ldy or ncols
    y = torch.zeros(M, ldy, device=dev)
    yp = torch.zeros(M, ldy, device=dev) if y_pre else None
    st = torch.zeros(((M + 127) // 128) * 4, 2, ncols, device=dev) if stats else None
    _C.call("cavp_igemm", _C.ptr(x), _C.ptr(w), _C.ptr(y), _C.ptr(yp), _C.ptr(scale), _C.ptr(shift), _C.ptr(res),
            _C.ptr(st), nimg, hs, ws, c, x.shape[-1], ho, wo, r, s, stride, pad, dil, dgrad, ncols, ldw_,
            ldy, res.shape[-1] if res is not None else 0, 0, 0, ncols, act, 0.01, splits, prec, lo_off, _C.stream())
    torch.cuda.synchronize()
    return y, yp, st

def report(name, err, tol):
    ok = err < tol
    results.append(dict(name=name, err=err, tol=tol, ok=ok))
    print(f"{'OK  ' if ok else 'FAIL'} {name:60s} err={err:.3e} tol={tol:.1e}", flush=True)

for _tma in (False, True):
  TMA[0] = _tma
  print('==== B operand via', 'TMA (pre-split)' if _tma else 'producer warps', flush=True)
  torch.manual_seed(0)
  # 1. plain GEMM
  for (M, N, K) in [(128, 128, 32), (128, 128, 64), (300, 200, 304), (64, 48, 100), (1000, 22 + 2, 256)]:
      a = torch.randn(M, K); b = torch.randn(N, K)
      ref = a.double() @ b.double().t()
      for prec in (1, 2):
          y, _, _ = igemm(a.to(dev), b.to(dev), M, 1, 1, K, 1, 1, 1, 1, 1, 0, 1, 0, N, prec)
          e = relerr(y, ref)
          report(f"gemm M{M} N{N} K{K} prec{prec}", e, 3e-3 if prec == 1 else 2e-6)
          if e > 1e-2 and M == 128 and K == 32:
              print("got", y[:4, :8].cpu()); print("ref", ref[:4, :8])
  # split-K
  a = torch.randn(64, 4096); b = torch.randn(300, 4096)
  ref = a.double() @ b.double().t()
  y, _, _ = igemm(a.to(dev), b.to(dev), 64, 1, 1, 4096, 1, 1, 1, 1, 1, 0, 1, 0, 300, 2, splits=8)
  report("gemm splitK8 M64 N300 K4096 prec2", relerr(y, ref), 2e-6)

  # 2. convs (NHWC)
  def conv_case(nimg, h, w_, c, cout, r, stride, pad, dil, prec, tag=""):
      x = torch.randn(nimg, c, h, w_); wt = torch.randn(cout, c, r, r) / (c * r * r) ** 0.5
      ref = F.conv2d(x.double(), wt.double(), None, stride, pad, dil)
      ho, wo = ref.shape[-2:]
      xh = x.permute(0, 2, 3, 1).contiguous().to(dev)
      wh = wt.permute(0, 2, 3, 1).contiguous().reshape(cout, -1).to(dev)
      y, _, _ = igemm(xh, wh, nimg, h, w_, c, ho, wo, r, r, stride, pad, dil, 0, cout, prec)
      got = y.view(nimg, ho, wo, cout).permute(0, 3, 1, 2)
      report(f"conv{tag} n{nimg} {h}x{w_} c{c}->{cout} k{r} s{stride} p{pad} d{dil} prec{prec}", relerr(got, ref),
             3e-3 if prec == 1 else 2e-6)
      return x, wt, ref

  for prec in (1, 2):
      conv_case(2, 14, 14, 64, 96, 3, 1, 1, 1, prec)
      conv_case(2, 15, 13, 32, 40, 3, 2, 1, 1, prec)
      conv_case(1, 14, 14, 128, 256, 3, 1, 6, 6, prec)
      conv_case(3, 9, 9, 4, 64, 3, 2, 1, 1, prec)
      conv_case(2, 20, 12, 304, 48, 1, 1, 0, 1, prec)
      conv_case(2, 10, 10, 8, 16, 7, 2, 3, 1, prec)

  # 3. dgrad / wgrad against autograd
  def grad_case(nimg, h, w_, c, cout, r, stride, pad, dil, prec):
      x = torch.randn(nimg, c, h, w_, dtype=torch.double, requires_grad=True)
      wt = (torch.randn(cout, c, r, r, dtype=torch.double) / (c * r * r) ** 0.5).requires_grad_(True)
      y = F.conv2d(x, wt, None, stride, pad, dil)
      ho, wo = y.shape[-2:]
      dy = torch.randn_like(y)
      y.backward(dy)
      dyh = dy.float().permute(0, 2, 3, 1).contiguous().to(dev)          # [n,ho,wo,cout]
      xh = x.detach().float().permute(0, 2, 3, 1).contiguous().to(dev)
      # dgrad: rows = input pixels, source = dy, weights transposed to [cin][r][s][cout]
      wt_t = wt.detach().float().permute(1, 2, 3, 0).contiguous().reshape(c, -1).to(dev)
      dx, _, _ = igemm(dyh, wt_t, nimg, ho, wo, cout, h, w_, r, r, stride, pad, dil, 1, c, prec)
      got = dx.view(nimg, h, w_, c).permute(0, 3, 1, 2)
      report(f"dgrad n{nimg} {h}x{w_} c{c}->{cout} k{r} s{stride} p{pad} d{dil} prec{prec}", relerr(got, x.grad),
             3e-3 if prec == 1 else 2e-6)
      for splits in (1, 3):
          dw = torch.zeros(cout, r * r * c, device=dev)
          _C.call("cavp_igemm_wgrad", _C.ptr(dyh), _C.ptr(xh), _C.ptr(dw), nimg, h, w_, c, c, ho, wo, r, r, stride, pad,
                  dil, cout, cout, splits, prec, _C.stream())
          torch.cuda.synchronize()
          gotw = dw.view(cout, r, r, c).permute(0, 3, 1, 2)
          report(f"wgrad(splits{splits}) n{nimg} {h}x{w_} c{c}->{cout} k{r} s{stride} p{pad} d{dil} prec{prec}",
                 relerr(gotw, wt.grad), 3e-3 if prec == 1 else 2e-6)

  for prec in (1, 2):
      grad_case(2, 14, 14, 64, 96, 3, 1, 1, 1, prec)
      grad_case(2, 15, 13, 32, 40, 3, 2, 1, 1, prec)
      grad_case(1, 14, 14, 128, 256, 3, 1, 6, 6, prec)
      grad_case(2, 20, 12, 304, 48, 1, 1, 0, 1, prec)
      grad_case(2, 12, 12, 64, 128, 1, 2, 0, 1, prec)

  # 4. epilogue
  M, N, K = 300, 200, 96
  a = torch.randn(M, K); b = torch.randn(N, K) / K ** 0.5
  sc = torch.rand(N) + 0.5; sh = torch.randn(N); rs = torch.randn(M, N)
  pre = (a.double() @ b.double().t()) * sc.double() + sh.double() + rs.double()
  y, yp, st = igemm(a.to(dev), b.to(dev), M, 1, 1, K, 1, 1, 1, 1, 1, 0, 1, 0, N, 2, scale=sc.to(dev), shift=sh.to(dev),
                    res=rs.to(dev), act=1, stats=True, y_pre=True)
  report("epilogue relu out", relerr(y, pre.clamp_min(0)), 2e-6)
  report("epilogue pre-act", relerr(yp, pre), 2e-6)
  relu = pre.clamp_min(0)
  report("epilogue stats sum", relerr(st[:, 0].sum(0), relu.sum(0)), 1e-5)
  report("epilogue stats sumsq", relerr(st[:, 1].sum(0), (relu * relu).sum(0)), 1e-5)
  y, _, _ = igemm(a.to(dev), b.to(dev), M, 1, 1, K, 1, 1, 1, 1, 1, 0, 1, 0, N, 2, shift=sh.to(dev), act=3)
  report("epilogue gelu", relerr(y, F.gelu(a.double() @ b.double().t() + sh.double())), 2e-6)


TMA[0] = False
# 5. speed probe
def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

speeds = {}
for (nimg, h, c, cout, r, pad) in [(64, 56, 256, 256, 3, 1), (64, 56, 304, 256, 3, 1), (32, 56, 64, 256, 1, 0), (32, 28, 512, 512, 3, 2)]:
    xh = torch.randn(nimg, h, h, c, device=dev); wh = torch.randn(cout, r * r * c, device=dev)
    M = nimg * h * h
    y = torch.empty(M, cout, device=dev)
    sp = presplit(wh)
    for prec in (1, 2):
      for tma in (False, True):
        wptr, LO_OFF = (sp[0], sp[0].numel()) if tma else (wh, 0)
        def run():
            _C.call("cavp_igemm", _C.ptr(xh), _C.ptr(wptr), _C.ptr(y), 0, 0, 0, 0, 0, nimg, h, h, c, c, h, h, r, r, 1, pad, pad if r == 3 else 1, 0,
                    cout, r * r * c, cout, 0, 0, 0, cout, 0, 0.0, 1, prec, LO_OFF, _C.stream())
        ms = bench(run)
        fl = 2.0 * M * cout * r * r * c
        speeds[f"fwd n{nimg} {h}^2 c{c}->{cout} k{r} prec{prec} tma{int(tma)}"] = dict(ms=ms, tflops=fl / ms / 1e9)
        print(f"fwd n{nimg} {h}^2 c{c}->{cout} k{r} prec{prec} tma{int(tma)}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    dyh = torch.randn(M, cout, device=dev); dw = torch.zeros(cout, r * r * c, device=dev)
    for prec in (1, 2):
        tiles = ((cout + 127) // 128) * ((r * r * c + 127) // 128)
        splits = max(1, min(M // 32 // 8, (296 + tiles - 1) // tiles))
        def runw():
            _C.call("cavp_igemm_wgrad", _C.ptr(dyh), _C.ptr(xh), _C.ptr(dw), nimg, h, h, c, c, h, h, r, r, 1, pad, pad if r == 3 else 1,
                    cout, cout, splits, prec, _C.stream())
        ms = bench(runw)
        fl = 2.0 * M * cout * r * r * c
        print(f"wgrad n{nimg} {h}^2 c{c}->{cout} k{r} prec{prec} splits{splits}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        speeds[f"wgrad n{nimg} {h}^2 c{c}->{cout} k{r} prec{prec}"] = dict(ms=ms, tflops=fl / ms / 1e9, splits=splits)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(results=results, speeds=speeds), open("gpurun_out/probe_igemm.json", "w"), indent=1)
print("FAILED:", [r["name"] for r in results if not r["ok"]])
