#!/usr/bin/env python
"""bench.py - AVS train-step images/sec @224^2 bs32/GPU on 1/2/4/8 B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference arithmetic on the host CPU cores, see below)

A "step" is one pass of the hot path over one synthetic batch: CAVP forward (ResNet-50 + VGG audio + cross-modal
fusion + decoder) -> CE + ContrastLoss -> backward -> SGD(visual) + Adam(audio), i.e. the loop body of
trainer/trainer_cavp_vpo_mono.py:142-193.  Workload = BASELINE.json configs[1]: VPO-SS shapes (22 classes, dilation
[F,T,T], VGG audio 96x64), fp32, 32 images per GPU.  Weak scaling: every rank processes its own 32 images, then ONE
NCCL all-reduce averages the flat gradient buffer.

`value`  : images/s with the batch already resident in HBM.
`e2e`    : same through the public call (cavp_b200.trainer.train_step) from pinned HOST buffers: H2D of image / audio /
           labels and a D2H read of the two losses inside the timed region, every step.
`roofline`: the dominant kernel (the tcgen05 implicit-GEMM tile kernel igemm_kernel, all its launches of one step),
           timed with CUDA events on the launching stream in an extra instrumented step.
`cpu_baseline` / `--impl reference`: the oracle port of the reference (oracle/cavp_oracle.py, torch CPU fp32 - the
           reference itself is Python/PyTorch and cannot travel to the GPU box) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs[1] (the headline: fp32, VPO-SS shapes) and configs[2] (AVSBench-Semantics shapes, bf16 loop)
CONFIGS = {
    1: dict(cfg=dict(nc=22, dilation=(False, True, True), audio="vgg", in_plane=1, H=224, W=224, frames=96,
                     max_views=512),
            workload="configs[1]: ResNet-50 + VGGish CAVP fwd+bwd bs32/GPU, synthetic VPO-SS shapes (22 cls, dilation "
                     "FTT), fp32",
            flops=298.67e9),  # reference-equivalent fwd+bwd conv+GEMM FLOPs per image
    2: dict(cfg=dict(nc=71, dilation=(False, False, False), audio="vgg", in_plane=1, H=224, W=224, frames=96,
                     max_views=512),
            workload="configs[2]: AVSBench-Semantics config (71 cls, dilation FFF) bf16 training loop, bs32/GPU, NCCL "
                     "grad allreduce",
            flops=169.20e9),
}
CFG = dict(CONFIGS[1]["cfg"])
WORKLOAD = CONFIGS[1]["workload"]
METRIC = "AVS train-step images/sec @224^2 bs32/GPU"
NCCL_MAX_CTAS_DEFAULT = 0  # chosen by measurement at 8 GPUs (profiles/README.md)
FLOPS_PER_IMAGE = CONFIGS[1]["flops"]


def select_config(idx):
    global WORKLOAD, FLOPS_PER_IMAGE
    CFG.clear()
    CFG.update(CONFIGS[idx]["cfg"])
    WORKLOAD = CONFIGS[idx]["workload"]
    FLOPS_PER_IMAGE = CONFIGS[idx]["flops"]


def synthetic_batch(B, seed):
    """SURVEY.md 8(d) cfg 2: randn image / log-mel, one centred rectangle of a random foreground class per image,
    an 8x8 corner of ignore (255); the shuffled half follows trainer_cavp_vpo_mono.py:148-151,178-180."""
    g = torch.Generator().manual_seed(seed)
    H, W, nc = CFG["H"], CFG["W"], CFG["nc"]
    image = torch.randn(B, 3, H, W, generator=g)
    audio_m = torch.randn(B, 1, CFG["frames"], 64, generator=g)
    pix = torch.zeros(B, H, W, dtype=torch.int64)
    img_label = torch.zeros(B, nc, dtype=torch.int64)
    cls = torch.randint(1, nc, (B,), generator=g)
    for b in range(B):
        pix[b, H // 8:H - H // 8, W // 6:W - W // 6] = cls[b]
        pix[b, :8, :8] = 255
        img_label[b, cls[b]] = 1
    img_label[:, 0] = 1
    shuffle_idx = torch.randperm(B, generator=g)
    audio = torch.cat((audio_m, audio_m[shuffle_idx]), 0)
    from cavp_b200.trainer import shuffled_labels
    spl = shuffled_labels(pix, img_label, shuffle_idx)
    return image, audio, pix, spl


def make_args(B, device, prec):
    from types import SimpleNamespace
    return SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=list(CFG["dilation"]),
                           audio_backbone=CFG["audio"], num_classes=CFG["nc"], batch_size=B, local_rank=device,
                           cavp_prec=prec)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe): ONE `nvidia-smi -lms`
    process started before the region and killed after it; only the lines that arrive while the region is armed are
    kept.  (A fresh nvidia-smi per sample re-initialises NVML every time, which stalls this process' kernel launches
    for tens of ms - visible as a GPU bubble in the first timed step, when the host has no lead over the device.)
    If the looping process produces no line within 3 s the sampler falls back to one nvidia-smi per sample."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 100

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.armed, self.proc, self.first_line, self.oneshot = False, None, threading.Event(), False

    def _cmd(self, loop):
        return (["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)]
                + (["-lms", str(self.PERIOD_MS)] if loop else []))

    def run(self):
        try:
            self.proc = subprocess.Popen(self._cmd(True), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                         bufsize=1)
            for line in self.proc.stdout:
                self.first_line.set()
                if self.armed and line.strip():
                    self.samples.append([x.strip() for x in line.split(",")])
                if self.stop_flag or self.oneshot:
                    break
        except Exception:
            pass

    def _oneshot_loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(self._cmd(False), capture_output=True, text=True, timeout=5).stdout.strip()
                if out and self.armed:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def arm(self):
        """Call right before the timed region: waits until the looping nvidia-smi is up, then starts keeping samples."""
        if not self.first_line.wait(timeout=3.0):
            self.oneshot = True
            if self.proc is not None:
                try:
                    self.proc.kill()
                except Exception:
                    pass
            threading.Thread(target=self._oneshot_loop, daemon=True).start()
        self.armed = True

    def stop(self):
        self.stop_flag = True
        self.armed = False
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ======================================================================================================================
def host_threads():
    """Host cores this process may use (affinity-aware).  torch.distributed.run exports OMP_NUM_THREADS=1 to its
    children; the CPU arms override that explicitly so they always use every core the box gives us."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _oracle_state(device, dtype=torch.float32):
    from oracle import schema
    sd = schema.seeded_state(CFG["nc"], CFG["audio"], CFG["in_plane"], seed=0, requires_grad=False)
    out = {}
    for k, v in sd.items():
        v = v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device)
        out[k] = v.requires_grad_(True) if (v.is_floating_point() and "running_" not in k) else v
    return out


def _oracle_stepper(sd, batch, device):
    """-> step() running the restated trainer body (trainer_cavp_vpo_mono.py:142-193) with torch.optim SGD + Adam."""
    from oracle import cavp_oracle as O
    leaves = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    audio_params = [v for k, v in sd.items() if k.startswith("audio_backbone.backbone") and v.is_floating_point()
                    and v.requires_grad]
    audio_ids = {id(v) for v in audio_params}
    opt_v = torch.optim.SGD([v for v in leaves if id(v) not in audio_ids], lr=1e-3, momentum=0.9, weight_decay=5e-4)
    opt_a = torch.optim.Adam(audio_params, lr=1e-4)
    image, audio, pix, spl = synthetic_batch(batch, 666)
    b = {"image": image.to(device), "audio": audio.to(device), "pix_label": pix.to(device)}
    spl = spl.to(device)

    def step():
        opt_v.zero_grad(set_to_none=True)
        opt_a.zero_grad(set_to_none=True)
        l_ce, l_ctr, *_rest, newbuf = O.train_step_losses(sd, b, spl, dilation_flags=CFG["dilation"],
                                                          audio_kind=CFG["audio"], max_views=CFG["max_views"])
        (l_ce + l_ctr).backward()
        opt_v.step()
        opt_a.step()
        for k, v in newbuf.items():
            sd[k] = v
        return l_ce, l_ctr
    return step


def cpu_reference_arm(steps, warmup, batch, threads=None):
    """The reference arithmetic on the host cores: oracle port (torch CPU fp32) of the same train step + optimisers."""
    torch.set_num_threads(threads or host_threads())
    torch.manual_seed(666)
    step = _oracle_stepper(_oracle_state("cpu"), batch, "cpu")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return batch / (ms / 1e3), ms, torch.get_num_threads()


def gpu_stock_baseline(dev, batch, steps=5, warmup=2):
    """SURVEY.md 8(d) / BASELINE.md 5.3: the reference arithmetic (oracle port = the same torch ops the reference
    modules call) under STOCK PyTorch / cuDNN / cuBLAS on this B200, same workload, same step body + torch.optim,
    timed with CUDA events.  Settings of the reference entry point (main_vpo_mono.py:33-42): cudnn.benchmark = True and
    torch's defaults, i.e. TF32 allowed for cuDNN convolutions, not for cuBLAS matmuls ("reference_default")."""
    modes = [("fp32_no_tf32", False, False, None), ("reference_default", True, False, None),
             ("tf32_all", True, True, None), ("bf16_autocast", True, True, torch.bfloat16)]
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out = {"what": "oracle port (same torch ops as the reference modules) on cuda, stock PyTorch %s / cuDNN %s, "
                   "cudnn.benchmark=True, bs%d, %d warm-up + %d timed steps, CUDA events"
                   % (torch.__version__, torch.backends.cudnn.version(), batch, warmup, steps)}
    try:
        torch.backends.cudnn.benchmark = True
        for name, conv_tf32, mm_tf32, amp in modes:
            try:
                torch.backends.cudnn.allow_tf32 = conv_tf32
                torch.backends.cuda.matmul.allow_tf32 = mm_tf32
                torch.manual_seed(666)
                step = _oracle_stepper(_oracle_state(dev), batch, dev)
                ctx = torch.autocast("cuda", dtype=amp) if amp is not None else None

                def run():
                    if ctx is None:
                        return step()
                    with ctx:
                        return step()
                for _ in range(warmup):
                    run()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    l_ce, l_ctr = run()
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"images_per_s": batch / (ms / 1e3), "ms_per_step": ms,
                             "cudnn_allow_tf32": conv_tf32, "matmul_allow_tf32": mm_tf32,
                             "autocast": None if amp is None else "bf16", "l_ce": float(l_ce.detach()),
                             "l_ctr": float(l_ctr.detach().sum())}
                del step
            except Exception as e:  # a stock-library failure must not take the bench line down
                out[name] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


def measure_tf32_peak(dev, seconds=2.0):
    """TF32 dense matmul peak measured like MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with allow_tf32,
    best of 10 (burst) and back to back for `seconds` (sustained).  3xTF32 (fp32-parity mode) can reach a third of it."""
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize(dev)
        sus = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"tf32_tflops": fl / (best / 1e3) / 1e12, "tf32_tflops_sustained": fl / (sus / 1e3) / 1e12,
                "how": "torch.matmul fp32 8192^3, allow_tf32=True: best of 10 and back to back for %.0f s" % seconds}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    batch = args.cpu_batch or 32  # the same 32 images per step as the GPU arm (configs[1])
    value, ms, cores = cpu_reference_arm(args.steps, args.warmup, batch, threads=host_threads())
    sample = (f"{batch} images per step" + ("" if batch == 32 else " (bounded sample of the bs32 workload)") +
              f", {args.warmup} warm-up + {args.steps} timed steps, {cores} threads "
              f"(os.cpu_count()={os.cpu_count()}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} overridden)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": 32, "images_per_step": batch,
                       "note": "reference arithmetic on the host CPU (oracle port: the reference is Python/PyTorch and "
                               "cannot travel to the GPU box); one CPU replica regardless of --gpus"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ======================================================================================================================
def run_ours(args):
    import torch.distributed as dist
    from cavp_b200.models.cavp_model import CAVP
    from cavp_b200.parallel import FlatGradBuffer, cavp_buckets
    from cavp_b200.trainer import train_step

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's INIT lines (communicator size, NVLS / ring choice) go to stderr: the driver reads the rank count from
        # them, and stdout stays the one JSON line.  A level set by the caller wins.
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        nccl_log = None
        if "NCCL_DEBUG_FILE" not in os.environ:  # NCCL writes to a per-rank file; it is copied to stderr below
            nccl_log = f"/tmp/cavp_nccl_rank{rank}_{os.getpid()}.log"
            os.environ["NCCL_DEBUG_FILE"] = nccl_log
        # NCCL's kernels and ours compete for SMs while a bucket's all-reduce overlaps the backward pass (our GEMM kernels
        # are persistent, one CTA or CTA pair per SM / TPC): cap the collective's CTAs (CAVP_NCCL_MAX_CTAS, 0 = NCCL's choice)
        max_ctas = int(os.environ.get("CAVP_NCCL_MAX_CTAS", str(NCCL_MAX_CTAS_DEFAULT)))
        pg_opts = None
        if max_ctas > 0:
            try:
                pg_opts = dist.ProcessGroupNCCL.Options()
                pg_opts.config.max_ctas = max_ctas
                pg_opts.config.min_ctas = min(max_ctas, 4)
            except Exception:
                pg_opts = None
        if pg_opts is not None:
            dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
        else:
            dist.init_process_group("nccl", device_id=dev)
        probe = torch.ones(1, device=dev)
        dist.all_reduce(probe)  # forces communicator creation so that its INIT lines exist before the timed region
        torch.cuda.synchronize()
        sys.stderr.write(f"[cavp_b200] rank {rank}: torch.distributed NCCL communicator up, nranks {world} "
                         f"(all-reduce probe = {int(probe.item())}), NCCL {'.'.join(map(str, torch.cuda.nccl.version()))}\n")
        if nccl_log and os.path.exists(nccl_log):
            for ln in open(nccl_log, errors="replace"):
                if any(k in ln for k in ("nranks", "Init COMPLETE", "NVLS", "Connected", "NCCL version", "Channel 00")):
                    sys.stderr.write(ln)
        sys.stderr.flush()
    B = args.batch
    torch.manual_seed(666 + rank)  # main_*.py: seed_it(seed + local_rank), seed 666
    model = CAVP(50, None, num_classes=CFG["nc"], ignore_index=255, audio_backbone_pretrain_path=None,
                 visual_backbone=50, args=make_args(B, local_rank, args.prec), in_plane=CFG["in_plane"]).to(dev).train()
    sync_bn = bool(args.sync_bn) and world > 1
    if sync_bn:  # main_vpo_mono.py:130: the reference's multi-GPU launch converts every BatchNorm
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    if world > 1:  # identical initial weights on every rank (DDP broadcasts rank 0's)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    audio_params = list(model.audio_backbone.backbone.parameters())
    audio_ids = {id(p) for p in audio_params}
    visual_params = [p for p in model.parameters() if id(p) not in audio_ids]
    # main_vpo_mono.py:118-125: SGD(momentum 0.9, wd 5e-4) on the visual groups (lr *= gpus), Adam on the audio backbone
    # (fused drop-ins of torch.optim.SGD / Adam: cavp_b200/optim.py, one kernel launch per optimiser)
    from cavp_b200.optim import SGD, Adam
    opt_v = SGD(visual_params, lr=1e-3 * world, momentum=0.9, weight_decay=5e-4)
    opt_a = Adam(audio_params, lr=1e-4 * world)
    # gradients live in one flat, bucketed buffer (zeroed by ONE memset per step, written in place by the weight-gradient
    # kernels); with more than one rank each bucket is all-reduced as soon as the backward pass has completed it
    flat = FlatGradBuffer(cavp_buckets(model), dev)

    image_h, audio_h, pix_h, spl_h = synthetic_batch(B, 666 + rank)
    pinned = [t.pin_memory() for t in (image_h, audio_h, pix_h)]
    image_d, audio_d, pix_d = (t.to(dev) for t in pinned)
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned)
    launches = [0]
    host_ms = [0.0]

    # e2e input path: pinned host tensors -> device on a copy stream, double-buffered like a DataLoader(pin_memory=True)
    # prefetcher: the H2D copy of step i+1 overlaps the kernels of step i; every step still copies all of its inputs
    # inside the timed region (h2d_bytes_per_step) and reads its two losses back (8 bytes, one transfer)
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {"bufs": None, "event": None}

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            bufs = [t.to(dev, non_blocking=True) for t in pinned]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["bufs"], staged["event"] = bufs, ev

    def step(resident, profile=None):
        if resident:
            im, au, px = image_d, audio_d, pix_d
        else:
            if staged["bufs"] is None:
                stage_inputs()
            torch.cuda.current_stream(dev).wait_event(staged["event"])
            im, au, px = staged["bufs"]
            for t in (im, au, px):
                t.record_stream(torch.cuda.current_stream(dev))
            stage_inputs()  # next step's copy, overlapping this step's kernels
        opt_v.zero_grad(set_to_none=True)
        opt_a.zero_grad(set_to_none=True)
        res = train_step(model, im, au, pix_h, spl_h, max_views=CFG["max_views"], labels_dev=px, profile=profile,
                         grad_sink=flat)  # N>1: bucketed all-reduce overlapped with the backward pass
        opt_v.step()
        opt_a.step()
        launches[0] += res.launches + 2  # + the two fused optimiser kernels
        if not resident:
            # D2H read of the step's result, every step: the two losses go to a pinned slot with a non-blocking copy and
            # are consumed on the host one step later (after the next step has been issued), the way a training loop
            # logs its losses without draining the launch queue; the last step's losses are read inside the timed region
            slot = readback["ring"][readback["n"] % 2]
            slot["buf"].copy_(torch.stack((res.l_ce.reshape(()), res.l_ctr.reshape(()))), non_blocking=True)
            slot["ev"].record(torch.cuda.current_stream(dev))
            if readback["n"] > 0:
                prev = readback["ring"][(readback["n"] - 1) % 2]
                prev["ev"].synchronize()
                readback["last"] = prev["buf"].tolist()
            readback["n"] += 1
            return None
        return res

    readback = {"ring": [{"buf": torch.empty(2, dtype=torch.float32).pin_memory(), "ev": torch.cuda.Event()}
                         for _ in range(2)], "n": 0, "last": None}

    def drain_readback():
        if readback["n"] > 0:
            cur = readback["ring"][(readback["n"] - 1) % 2]
            cur["ev"].synchronize()
            readback["last"] = cur["buf"].tolist()

    def timed(resident):
        for _ in range(args.warmup):
            step(resident)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        sampler.start()
        sampler.arm()
        e0.record()
        t_host = time.perf_counter()
        marks = []
        if args.trace_steps:
            import gc
            gc_log, gc_t0 = [], [0.0]

            def on_gc(phase, info):
                if phase == "start":
                    gc_t0[0] = time.perf_counter()
                else:
                    gc_log.append((info["generation"], round(1e3 * (time.perf_counter() - gc_t0[0]), 2),
                                   info["collected"], len(marks)))
            gc.callbacks.append(on_gc)
        for _ in range(args.steps):
            step(resident)
            if args.trace_steps:  # per-step device time and host time (diagnostic; adds one event record per step)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((ev, time.perf_counter()))
        if not resident:
            drain_readback()  # the last step's losses, still inside the timed region
        host_ms[0] = 1e3 * (time.perf_counter() - t_host) / args.steps  # time to ISSUE a step (no device sync)
        e1.record()
        torch.cuda.synchronize()
        sampler.stop()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sampler.join(timeout=2)
        if args.trace_steps:
            gc.callbacks.remove(on_gc)
            slow = [x for x in gc_log if x[1] > 1.0]
            print(f"[trace] python gc: {len(gc_log)} collections, (generation, ms, collected, during step) of those over "
                  f"1 ms: {slow}", file=sys.stderr, flush=True)
        if marks and rank == 0:
            prev_e, prev_t, dev_ms, issue_ms = e0, t_host, [], []
            for ev, t in marks:
                dev_ms.append(round(prev_e.elapsed_time(ev), 2))
                issue_ms.append(round(1e3 * (t - prev_t), 2))
                prev_e, prev_t = ev, t
            print(f"[trace] {'resident' if resident else 'e2e'} loop: device ms per step {dev_ms}; host ms to issue "
                  f"each step {issue_ms}", file=sys.stderr, flush=True)
        return float(ms) / args.steps, sampler.summary(), launches[0] // args.steps

    # allocator priming (part of set-up, like building the model): the first steps of a process still grow the caching
    # allocator's pools (~12 GB of activations per step) with synchronous cudaMallocs; the W warm-up steps then run in
    # the steady state every later step of a training job sees
    for i in range(3):
        step(True)
        if args.mem_trace:
            torch.cuda.synchronize()
            print(f"[mem] after set-up step {i}: {torch.cuda.memory_allocated(dev) / 2 ** 30:.2f} GiB live, "
                  f"{torch.cuda.max_memory_allocated(dev) / 2 ** 30:.2f} GiB peak, "
                  f"{torch.cuda.memory_reserved(dev) / 2 ** 30:.2f} GiB reserved", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    ms_res, clocks, launches_per_step = timed(True)
    host_issue_ms = host_ms[0]
    ms_e2e, clocks_e2e, _ = timed(False)
    hbm_peak_gib = torch.cuda.max_memory_allocated(dev) / 2 ** 30  # live tensors at the peak of a train step
    value = world * B / (ms_res / 1e3)
    e2e = world * B / (ms_e2e / 1e3)

    # ---- instrumented step: CUDA events around every kernel launch (rank 0)
    roofline, attn_roof, breakdown = None, None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    prof = []
    step(True, profile=prof)  # every rank runs it (the step contains the gradient all-reduce)
    torch.cuda.synchronize()
    if rank == 0:
        agg = {}
        rows_ = []
        for name, a, b, fl, nb, tag in prof:
            rows_.append((a.elapsed_time(b), name, tag, fl))
            d = agg.setdefault(name, [0.0, 0, 0.0, 0.0])
            d[0] += a.elapsed_time(b); d[1] += 1; d[2] += fl; d[3] += nb
        total_ms = sum(v[0] for v in agg.values())
        if args.dump_profile:
            rows_.sort(reverse=True)
            with open(args.dump_profile, "w") as f:
                for ms_, name, tag, fl in rows_[:150]:
                    f.write(f"{ms_:8.3f} ms  {fl / ms_ / 1e9 if ms_ > 0 else 0:7.1f} TF  {name}  {tag}\n")
        ig = [agg.get(k, [0, 0, 0, 0]) for k in ("cavp_igemm", "cavp_igemm_wgrad", "cavp_igemm_wgrad_tma",
                                                 "cavp_igemm_bf16", "cavp_igemm_wgrad_bf16")]
        ig_ms, ig_n, ig_fl = sum(v[0] for v in ig), sum(v[1] for v in ig), sum(v[2] for v in ig)
        # DRAM traffic per launch of the same kernels, from the committed ncu launch list of this command
        # (profiles/r01_ncu_step_summary.json, written by tools/ncu_summarize.py; ncu counters cannot be read live)
        traffic, traffic_src = None, None
        for prof_name in ("r02_ncu_step_summary.json", "r01_ncu_step_summary.json"):
            if args.prec != 2:
                break  # the committed launch list is of the default (fp32-parity) command
            try:
                summ = json.load(open(os.path.join(ROOT, "profiles", prof_name)))["kernels"]
                igk = [v for k, v in summ.items() if "igemm" in k]
                n_l = sum(v["launches"] for v in igk)
                if n_l:
                    traffic = 1e6 * sum(v["dram_read_MB"] + v["dram_write_MB"] for v in igk) / n_l
                    traffic_src = (f"profiles/{prof_name}: dram__bytes_read.sum + dram__bytes_write.sum over "
                                   f"{n_l} igemm launches of the bench command, bytes per launch")
                    break
            except Exception:
                continue
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = ig_fl / (ig_ms / 1e3) / 1e12 if ig_ms else 0.0
        roofline = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM tile kernels igemm_ws / igemm_ws2 / igemm_kernel (cavp_igemm + cavp_igemm_wgrad[_tma])",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": traffic, "traffic_source": traffic_src, "launches_per_step": ig_n, "avg_launch_ms": ig_ms / max(ig_n, 1),
                    "flop_per_step_executed": ig_fl, "share_of_kernel_time": ig_ms / total_ms if total_ms else None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback",
                    "note": {2: "fp32-parity mode issues 3 TF32 MMAs per product (<= 1/6 of the bf16 peak by construction)",
                             1: "plain TF32: one TF32 MMA per product (<= 1/2 of the bf16 peak by construction)",
                             3: "bf16 operands for forward / dgrad / large weight gradients (kind::f16), TF32 for the "
                                "small weight gradients; fp32 activations are converted by the producer warps"}[args.prec]}
        gate = agg.get("cavp_gate_fwd")
        if gate and gate[0] > 0:
            hbm = peaks.get("hbm_gbs", 6650.0)
            gbs = gate[3] / (gate[0] / 1e3) / 1e9
            attn_roof = {"bound": "hbm", "kernel": "gate_fwd_kernel (cross-attention core, models/attn.py:73-106)",
                         "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "bytes_per_launch": gate[3],
                         "ms": gate[0]}
        breakdown = {k: round(v[0], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]}

    cpu, stock, tf32_peak = None, None, None
    if rank == 0 and world == 1 and not args.no_stock_baseline:
        torch.cuda.empty_cache()
        tf32_peak = measure_tf32_peak(dev)
        stock = gpu_stock_baseline(dev, B)
        if roofline is not None and tf32_peak:
            roofline["tf32_peak_measured"] = tf32_peak
        if roofline is not None and tf32_peak and args.prec == 2:
            ceil3 = tf32_peak["tf32_tflops_sustained"] / 3.0
            roofline["frac_of_3xtf32_ceiling"] = roofline["achieved"] / ceil3
            roofline["note"] += ("; against the TF32 matmul peak measured in this run (sustained / 3 = %.0f TFLOP/s) "
                                 "the tile kernels reach frac_of_3xtf32_ceiling" % ceil3)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or 8
        v, ms, cores = cpu_reference_arm(2, 1, cb, threads=host_threads())
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"oracle port, {cb} images per step (bounded sample of the bs32 workload; --impl reference runs "
                         f"all 32), 1 warm-up + 2 timed steps ({ms:.0f} ms/step), {cores} threads"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {1: "tf32", 2: "f32", 3: "bf16"}[args.prec], "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world,
                           "parallelism": f"dp{world}",
                           "precision": {2: "fp32 I/O, 3xTF32 tcgen05 products + fp32 promotion",
                                         1: "fp32 I/O, TF32 tcgen05 products",
                                         3: "bf16 operands (tcgen05 kind::f16, fp32 accumulate in TMEM) for forward / "
                                            "dgrad GEMMs, TF32 weight gradients, fp32 activations / BN statistics / "
                                            "LayerNorm / losses; stems and classifier fp32-grade"}[args.prec],
                           "l2": "working set per step (>10 GB of activations) far exceeds the 126 MB L2; no flush needed",
                           "bn": ("SyncBatchNorm: fp64 [sum, sum^2, count] of every BN layer all-reduced on the device in "
                                  "forward and backward (--sync-bn)") if sync_bn
                           else "local (per-rank) BatchNorm statistics",
                           "allreduce": (f"5 gradient buckets produced in place, NCCL all-reduce(AVG) of each bucket launched "
                                         f"asynchronously from a backward-tape marker (overlaps the rest of the backward); "
                                         f"NCCL max_ctas={os.environ.get('CAVP_NCCL_MAX_CTAS', NCCL_MAX_CTAS_DEFAULT)}")
                           if world > 1 else None},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 8,
                        "how": "cavp_b200.trainer.train_step from pinned host buffers: H2D of image / audio / labels every "
                               "step (prefetched on a copy stream), both losses copied D2H every step and read on the host "
                               "one step later; last losses read: %s" % (readback["last"],)},
                "gpu_launches": launches_per_step * args.steps,
                "gpu_launches_per_step": launches_per_step, "host_issue_ms_per_step": host_issue_ms,
                "hbm_peak_alloc_gib": round(hbm_peak_gib, 2),
                "roofline": roofline, "attn_roofline": attn_roof, "kernel_ms_top": breakdown,
                "reference_equiv_tflops": value * FLOPS_PER_IMAGE / 1e12, "cpu_baseline": cpu,
                "gpu_stock_baseline": stock}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--prec", type=int, default=2, choices=[1, 2, 3],
                    help="2 = fp32-parity (3xTF32 + promotion, the headline), 1 = plain TF32, 3 = bf16 operands")
    ap.add_argument("--config", type=int, default=None, choices=[1, 2],
                    help="BASELINE.json configs index: 1 = VPO-SS fp32 (default), 2 = AVSS 71 cls FFF (default for --prec 3)")
    ap.add_argument("--cpu-batch", type=int, default=None,
                    help="images per CPU reference step (default: 32 for --impl reference, 8 for the inline cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-bn", action="store_true",
                    help="N > 1: convert the BatchNorms to SyncBatchNorm as the reference's multi-GPU launch does "
                         "(main_vpo_mono.py:130); default: per-rank statistics")
    ap.add_argument("--trace-steps", action="store_true", help="print device and host time of every timed step")
    ap.add_argument("--mem-trace", action="store_true", help="print live / peak / reserved device memory after each set-up step")
    ap.add_argument("--no-stock-baseline", action="store_true",
                    help="skip the stock-PyTorch-on-this-GPU arm (gpu_stock_baseline) and the TF32 peak measurement")
    ap.add_argument("--dump-profile", default=None, help="write the per-launch CUDA-event profile of one step here")
    args = ap.parse_args()
    if args.cpu_batch is not None and args.cpu_batch < 2:
        ap.error("--cpu-batch must be >= 2 (train-mode BatchNorm of the pooled ASPP branch needs two images, as in the reference)")
    select_config(args.config if args.config is not None else (2 if args.prec == 3 else 1))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
