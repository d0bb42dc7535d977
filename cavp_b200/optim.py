"""Fused optimiser steps for the CAVP train loop.

Drop-ins for the two optimisers the reference entry scripts build (main_vpo_mono.py:118-125):

    optimizer_v = torch.optim.SGD(param_lists_v, lr=..., momentum=..., weight_decay=...)   -> cavp_b200.optim.SGD
    optimizer_a = torch.optim.Adam(params=model_a.parameters(), lr=...)                   -> cavp_b200.optim.Adam

Same constructor arguments, `param_groups` (the trainers rewrite `param_groups[i]["lr"]` every iteration,
trainer_cavp_vpo_mono.py:73-83), `zero_grad`, `state_dict` keys (`momentum_buffer`, `exp_avg`, `exp_avg_sq`, `step`: the
format engine/engine.py:72-100 checkpoints).  `step()` is ONE kernel launch per optimiser (csrc/optim.cu) over a table of
all parameters instead of torch's per-dtype/per-group foreach launches; there is no CPU fallback.
"""
import torch

from . import _C


def build_work(sizes, chunk):
    """(table row, chunk index) pairs covering tensors of the given sizes -> int32 tensor [nwork, 2]."""
    rows = []
    for i, n in enumerate(sizes):
        for c in range((n + chunk - 1) // chunk):
            rows.append((i, c))
    return torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)


def same_layout(a, b):
    return a.shape == b.shape and a.stride() == b.stride()


class _FusedBase(torch.optim.Optimizer):
    """Table bookkeeping shared by SGD and Adam.  Parameters must be dense fp32 CUDA tensors (any memory format: the
    kernels walk raw storage, and state / gradients are kept in the parameter's own layout)."""

    _state_keys = ()

    def _prepare(self):
        params, lrs, wds = [], [], []
        for grp in self.param_groups:
            for p in grp["params"]:
                params.append(p)
                lrs.append(float(grp["lr"]))
                wds.append(float(grp["weight_decay"]))
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("cavp_b200.optim runs on CUDA (sm_100a) only; there is no CPU fallback")
        cache = self.__dict__.get("_cavp_cache")
        sizes = tuple(p.numel() for p in params)
        if cache is None or cache["sizes"] != sizes:
            chunk = _C.query("cavp_opt_chunk_elems")
            work = build_work(sizes, chunk)
            cache = {"sizes": sizes, "work": work.to(dev), "nwork": work.shape[0],
                     # the pinned staging rows are rotated: the async upload of step i may still be queued behind the
                     # GPU work of step i-1 when the host starts filling the table of step i+1
                     "host": [torch.empty(len(params), 6, dtype=torch.int64).pin_memory() for _ in range(3)],
                     "events": [None, None, None], "turn": 0,
                     "dev": torch.empty(len(params), 6, dtype=torch.int64, device=dev)}
            self.__dict__["_cavp_cache"] = cache
        return params, lrs, wds, cache

    def _dense_like(self, p):
        return torch.zeros_like(p, memory_format=torch.preserve_format)

    def _fill_table(self, params, lrs, wds, cache, state_names):
        turn = cache["turn"]
        cache["turn"] = (turn + 1) % len(cache["host"])
        host = cache["host"][turn]
        if cache["events"][turn] is not None:
            cache["events"][turn].synchronize()
        fl = torch.empty(len(params), 2, dtype=torch.float32)
        rows = []
        keep = []  # gradients re-laid out for the kernel must outlive the launch
        for i, p in enumerate(params):
            g = p.grad
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("cavp_b200.optim: parameters must be fp32 CUDA tensors")
            st = self.state[p]
            # state loaded from a torch.optim checkpoint (the reference saves both optimisers' state_dicts,
            # engine/engine.py:93-94): `step` arrives as a per-parameter tensor and the moment buffers in the
            # checkpoint's NCHW-contiguous strides - normalise both once, on first use
            if "step" in st and not isinstance(st["step"], int):
                st["step"] = int(float(st["step"]))
            for name in state_names:
                if name not in st:
                    st[name] = self._dense_like(p)
                elif not same_layout(st[name], p) or st[name].dtype != p.dtype or st[name].device != p.device:
                    fixed = torch.empty_like(p, memory_format=torch.preserve_format)
                    fixed.copy_(st[name])  # plumbing copy into the parameter's own layout
                    st[name] = fixed
            gp = 0
            if g is not None:
                if g.is_sparse:
                    raise RuntimeError("cavp_b200.optim does not support sparse gradients")
                if not same_layout(g, p):
                    gl = torch.empty_like(p, memory_format=torch.preserve_format)
                    gl.copy_(g)  # plumbing copy into the parameter's layout
                    keep.append(gl)
                    g = gl
                gp = g.data_ptr()
                st["step"] = st.get("step", 0) + 1
            s = [st[name] for name in state_names]
            rows.append((p.data_ptr(), gp, s[0].data_ptr(), s[1].data_ptr() if len(s) > 1 else 0, p.numel()))
            fl[i, 0], fl[i, 1] = lrs[i], wds[i]
        host[:, :5] = torch.tensor(rows, dtype=torch.int64)
        host[:, 5] = fl.view(torch.int64).view(-1)
        cache["dev"].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        cache["events"][turn] = ev
        return keep

    @staticmethod
    def _stream(dev):
        return torch.cuda.current_stream(dev).cuda_stream


class SGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum, dampening=0, weight_decay, nesterov=False) on one fused kernel."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if dampening != 0.0 or nesterov:
            raise NotImplementedError("the reference uses plain momentum SGD (main_vpo_mono.py:118-123)")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))
        moms = {g["momentum"] for g in self.param_groups}
        if len(moms) > 1:
            raise NotImplementedError("per-group momentum")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        prep = self._prepare()
        if prep is None:
            return loss
        params, lrs, wds, cache = prep
        keep = self._fill_table(params, lrs, wds, cache, ("momentum_buffer",))
        _C.call("cavp_sgd_multi", cache["dev"].data_ptr(), cache["work"].data_ptr(), cache["nwork"],
                float(self.param_groups[0]["momentum"]), self._stream(params[0].device))
        del keep
        return loss


class Adam(_FusedBase):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) (no amsgrad) on one fused kernel.  All parameters that
    receive a gradient share the step counter (they do in the reference: the whole audio backbone steps together)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        prep = self._prepare()
        if prep is None:
            return loss
        params, lrs, wds, cache = prep
        keep = self._fill_table(params, lrs, wds, cache, ("exp_avg", "exp_avg_sq"))
        steps = {self.state[p]["step"] for p in params if p.grad is not None}
        if not steps:
            return loss
        if len(steps) > 1:
            raise NotImplementedError("parameters with different Adam step counts in one optimizer")
        t = steps.pop()
        grp = self.param_groups[0]
        b1, b2 = grp["betas"]
        _C.call("cavp_adam_multi", cache["dev"].data_ptr(), cache["work"].data_ptr(), cache["nwork"], float(b1),
                float(b2), float(grp["eps"]), 1.0 - b1 ** t, 1.0 - b2 ** t, self._stream(params[0].device))
        del keep
        return loss
