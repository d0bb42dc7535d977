"""Per-shape timing of the tcgen05 tile kernels over the GEMM shapes of one bs32 train step (diagnostic tool).

  python tools/shape_sweep.py [--prec 2] [--only SUBSTR] [--iters 10] [--json out.json]

Every case goes through the C-ABI (cavp_igemm / cavp_igemm_wgrad) exactly as engine.Graph issues it: weights pre-split
and fetched by TMA, epilogue flags as in the step.  Schedule selection is whatever the library decides (or what the
CAVP_IGEMM_WS / CAVP_IGEMM_SCHED environment forces), so running it under different settings compares schedules.
Timing: CUDA events on the launching stream, 3 warm-up launches, inputs far larger than L2 for the big cases.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cavp_b200 import _C  # noqa: E402

dev = "cuda"

# name, kind, nimg, h(=w) of the SOURCE, c, ncols, r, stride, pad, dil, flags
# kind 'row' = forward / dgrad through cavp_igemm; 'wgrad' through cavp_igemm_wgrad (ncols = cout, splits in flags)
CASES = [
    ("fc1 fwd gelu+pre  M200704 N1216 K304", "row", 64, 56, 304, 1216, 1, 1, 0, 1, dict(shift=1, act=3, pre=1)),
    ("fc2 dgrad         M200704 N1216 K304", "row", 64, 56, 304, 1216, 1, 1, 0, 1, dict()),
    ("fc2 fwd bias+res  M200704 N304 K1216", "row", 64, 56, 1216, 304, 1, 1, 0, 1, dict(shift=1, res=1)),
    ("fc1 dgrad         M200704 N304 K1216", "row", 64, 56, 1216, 304, 1, 1, 0, 1, dict()),
    ("fc1 dgrad +acc    M200704 N304 K1216", "row", 64, 56, 1216, 304, 1, 1, 0, 1, dict(res=1)),
    ("proj fwd bias     M200704 N304 K304 ", "row", 64, 56, 304, 304, 1, 1, 0, 1, dict(shift=1)),
    ("proj dgrad        M200704 N304 K304 ", "row", 64, 56, 304, 304, 1, 1, 0, 1, dict()),
    ("pe fwd bias       M100352 N304 K304 ", "row", 32, 56, 304, 304, 1, 1, 0, 1, dict(shift=1)),
    ("dec1 fwd stats    M200704 N256 K2736", "row", 64, 56, 304, 256, 3, 1, 1, 1, dict(stats=1)),
    ("dec2 fwd stats    M200704 N256 K2304", "row", 64, 56, 256, 256, 3, 1, 1, 1, dict(stats=1)),
    ("dec1 dgrad        M200704 N304 K2304", "row", 64, 56, 256, 304, 3, 1, 1, 1, dict(dgrad=1)),
    ("aspp fwd d6 stats M25088 N256 K18432", "row", 32, 28, 2048, 256, 3, 1, 6, 6, dict(stats=1)),
    ("aspp dgrad d12    M25088 N2048 K2304", "row", 32, 28, 256, 2048, 3, 1, 12, 12, dict(dgrad=1, res=1)),
    ("l4 1x1 fwd stats  M25088 N2048 K512 ", "row", 32, 28, 512, 2048, 1, 1, 0, 1, dict(stats=1)),
    ("l4 1x1 fwd stats  M25088 N512 K2048 ", "row", 32, 28, 2048, 512, 1, 1, 0, 1, dict(stats=1)),
    ("l4 3x3 fwd d2     M25088 N512 K4608 ", "row", 32, 28, 512, 512, 3, 1, 2, 2, dict(stats=1)),
    ("l3 1x1 fwd stats  M25088 N1024 K256 ", "row", 32, 28, 256, 1024, 1, 1, 0, 1, dict(stats=1)),
    ("l3 3x3 fwd d1     M25088 N256 K2304 ", "row", 32, 28, 256, 256, 3, 1, 1, 1, dict(stats=1)),
    ("l2 3x3 fwd stats  M25088 N128 K1152 ", "row", 32, 28, 128, 128, 3, 1, 1, 1, dict(stats=1)),
    ("l2 1x1 fwd stats  M25088 N128 K512  ", "row", 32, 28, 512, 128, 1, 1, 0, 1, dict(stats=1)),
    ("l2 1x1 dgrad      M25088 N512 K128  ", "row", 32, 28, 128, 512, 1, 1, 0, 1, dict(dgrad=1)),
    ("vgg c3 fwd relu   M24576 N256 K2304 ", "row", 64, 24, 256, 256, 3, 1, 1, 1, dict(shift=1, act=1)),
    ("vgg c5 fwd relu   M6144 N512 K4608  ", "row", 64, 12, 512, 512, 3, 1, 1, 1, dict(shift=1, act=1)),
    ("l1 1x1 fwd stats  M100352 N256 K64  ", "row", 32, 56, 64, 256, 1, 1, 0, 1, dict(stats=1)),
    ("l1 3x3 fwd stats  M100352 N64 K576  ", "row", 32, 56, 64, 64, 3, 1, 1, 1, dict(stats=1)),
    ("stem3 fwd stats   M401408 N128 K576 ", "row", 32, 112, 64, 128, 3, 1, 1, 1, dict(stats=1)),
    ("stem3 dgrad       M401408 N64 K1152 ", "row", 32, 112, 128, 64, 3, 1, 1, 1, dict(dgrad=1)),
    ("vgg fc1           M64 N4096 K12288  ", "row", 64, 1, 12288, 4096, 1, 1, 0, 1, dict(shift=1, act=1, splits=12)),
    ("wgrad dec1  P200704 Cout256 K2736 auto   ", "wgrad", 64, 56, 304, 256, 3, 1, 1, 1, dict()),
    ("wgrad fc1   P200704 Cout1216 K304 auto   ", "wgrad", 64, 56, 304, 1216, 1, 1, 0, 1, dict()),
    ("wgrad fc2   P200704 Cout304 K1216 auto   ", "wgrad", 64, 56, 1216, 304, 1, 1, 0, 1, dict()),
    ("wgrad aspp  P25088 Cout256 K18432 auto   ", "wgrad", 32, 28, 2048, 256, 3, 1, 6, 6, dict()),
    ("wgrad l4    P25088 Cout2048 K512 auto   ", "wgrad", 32, 28, 512, 2048, 1, 1, 0, 1, dict()),
    ("wgrad l4    P25088 Cout512 K4608 auto   ", "wgrad", 32, 28, 512, 512, 3, 1, 2, 2, dict()),
    ("wgrad stem3 P401408 Cout128 K576 auto   ", "wgrad", 32, 112, 64, 128, 3, 1, 1, 1, dict()),
    ("wgrad proj  P200704 Cout304 K304 auto   ", "wgrad", 64, 56, 304, 304, 1, 1, 0, 1, dict()),
]


def time_it(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


LDX0 = [False]  # --ldx0: every A row reads the same address (L1 hits): the no-global-latency ceiling of the main loop


def run_case(case, prec, iters):
    name, kind, nimg, h, c, ncols, r, stride, pad, dil, fl = case
    K = r * r * c
    if kind == "row":
        dgrad = fl.get("dgrad", 0)
        ho = h if (dgrad or stride == 1) else (h + 2 * pad - dil * (r - 1) - 1) // stride + 1
        M = nimg * ho * ho
        x = torch.randn(nimg * h * h, c, device=dev)
        w = torch.randn(ncols, K, device=dev) / K ** 0.5
        sp = torch.empty(2, ncols, K, device=dev)
        _C.call("cavp_split_tf32", _C.ptr(w), _C.ptr(sp[0]), _C.ptr(sp[1]), w.numel(), _C.stream())
        splits = fl.get("splits", 1)
        y = torch.zeros(M, ncols, device=dev)
        plain = splits > 1
        ypre = torch.empty(M, ncols, device=dev) if fl.get("pre") and not plain else None
        shift = torch.randn(ncols, device=dev) if fl.get("shift") and not plain else None
        res = (y if fl.get("res") else None) if not plain else None
        nparts = ((M + 127) // 128) * 4
        st = torch.empty(nparts, 2, ncols, device=dev) if fl.get("stats") and not plain else None
        act = 0 if plain else fl.get("act", 0)

        w16 = w.to(torch.bfloat16) if prec == 3 else None  # --prec 3: the bf16-operand kernel (cavp_igemm_bf16)

        def fn():
            if w16 is not None:
                _C.call("cavp_igemm_bf16", _C.ptr(x), _C.ptr(w), _C.ptr(w16), _C.ptr(y), _C.ptr(ypre), 0, _C.ptr(shift),
                        _C.ptr(res), _C.ptr(st), nimg, h, h, c, 0 if LDX0[0] else c, ho, ho, r, r, stride, pad, dil, dgrad,
                        ncols, K, ncols, ncols if res is not None else 0, 0, 0, ncols, act, 0.01, splits, 0, _C.stream())
                return
            _C.call("cavp_igemm", _C.ptr(x), _C.ptr(sp[0]), _C.ptr(y), _C.ptr(ypre), 0, _C.ptr(shift), _C.ptr(res),
                    _C.ptr(st), nimg, h, h, c, 0 if LDX0[0] else c, ho, ho, r, r, stride, pad, dil, dgrad, ncols, K, ncols,
                    ncols if res is not None else 0, 0, 0, ncols, act, 0.01, splits, prec, sp[0].numel(), _C.stream())
        flops = 2.0 * M * ncols * K
    else:
        ho = (h + 2 * pad - dil * (r - 1) - 1) // stride + 1
        P = nimg * ho * ho
        x = torch.randn(nimg * h * h, c, device=dev)
        dy = torch.randn(P, ncols, device=dev)
        dw = torch.zeros(ncols, K, device=dev)

        from cavp_b200.engine import Graph
        wsplits = fl.get("splits") or Graph.wgrad_splits(P, ncols, K)  # the engine's choice unless forced

        if Graph.wgrad_via_tma(P, ncols, K):  # as the engine: split dY (timed) + TMA-fed kernel
            gsp = torch.empty(2, P, ncols, device=dev)

            def fn():
                _C.call("cavp_split_tf32_2d", _C.ptr(dy), ncols, P, ncols, _C.ptr(gsp[0]), _C.ptr(gsp[1]), _C.stream())
                _C.call("cavp_igemm_wgrad_tma", _C.ptr(gsp[0]), gsp[0].numel(), _C.ptr(x), _C.ptr(dw), nimg, h, h, c, c,
                        ho, ho, r, r, stride, pad, dil, ncols, wsplits, prec, _C.stream())
        else:
            def fn():
                _C.call("cavp_igemm_wgrad", _C.ptr(dy), _C.ptr(x), _C.ptr(dw), nimg, h, h, c, c, ho, ho, r, r, stride,
                        pad, dil, ncols, ncols, wsplits, prec, _C.stream())
        flops = 2.0 * P * ncols * K
    ms = time_it(fn, iters)
    return ms, flops / ms / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", type=int, default=2)
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--ldx0", action="store_true")
    args = ap.parse_args()
    LDX0[0] = args.ldx0
    out = {}
    env = {k: v for k, v in os.environ.items() if k.startswith("CAVP_")}
    print(f"# shape sweep prec={args.prec} env={env}", flush=True)
    for case in CASES:
        if args.only and args.only not in case[0]:
            continue
        if args.prec == 3 and case[1] != "row":
            continue  # weight gradients run the TF32 kernels in bf16 mode
        ms, tf = run_case(case, 1 if (args.prec == 3 and case[1] != "row") else args.prec, args.iters)
        out[case[0].strip()] = {"ms": ms, "tflops": tf}
        print(f"{ms:8.3f} ms  {tf:7.1f} TF  {case[0]}", flush=True)
        torch.cuda.empty_cache()
    if args.json:
        json.dump({"env": env, "prec": args.prec, "cases": out}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
