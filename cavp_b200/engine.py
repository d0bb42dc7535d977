"""Host-side graph bookkeeping for the CAVP hot path.

A `Graph` records, per forward op, a closure that computes the op's gradients; `backward()` replays them in reverse.
Every arithmetic step is a C-ABI kernel from libcavp_b200.so (cavp_b200/_C.py); torch only owns device memory and
streams.  Gradient fan-in is resolved inside kernels (igemm residual epilogue / in-place adds), never by torch ops.

Activations are `Act`s: NHWC fp32 storage `[n*h*w, ld]` with a channel window `[off, off+c)`, so concat buffers and
their slices need no copies.  Token tensors `[rows, N, C]` are the same thing with (h, w) = token grid.
"""
import os

import torch

from . import _C

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
NUM_SMS = 148
LEAKY_SLOPE = 0.01  # nn.LeakyReLU default (models/visual/deeplabv3/encoder_decoder.py:135)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pad4(c):
    return (c + 3) // 4 * 4


class Act:
    """NHWC activation view: storage `buf` [n*h*w, ld] (contiguous rows), channels [off, off+c)."""

    __slots__ = ("buf", "n", "h", "w", "c", "off", "grad", "parent", "needs_grad", "want_split", "split")

    def __init__(self, buf, n, h, w, c, off=0, parent=None, needs_grad=True):
        assert buf.dim() == 2 and buf.shape[0] == n * h * w, (tuple(buf.shape), n, h, w)
        self.buf, self.n, self.h, self.w, self.c, self.off = buf, n, h, w, c, off
        self.grad = None
        self.parent = parent
        self.needs_grad = needs_grad
        self.want_split = False  # a consumer's weight-gradient kernel wants d(this) as a dense TF32 hi | lo split
        self.split = None        # (on a gradient Act) that split, [2, rows, c], written by the kernel that produced it

    @property
    def ld(self):
        return self.buf.stride(0)

    @property
    def rows(self):
        return self.n * self.h * self.w

    @property
    def ptr(self):
        return self.buf.data_ptr() + 4 * self.off

    @property
    def root(self):
        return self if self.parent is None else self.parent

    def slice(self, off, c):
        return Act(self.buf, self.n, self.h, self.w, c, self.off + off, parent=self.root, needs_grad=self.needs_grad)

    def dense(self):
        """torch view [n, h, w, c] (no copy)."""
        return self.buf[:, self.off:self.off + self.c].unflatten(0, (self.n, self.h, self.w))

    def nchw(self):
        """logical NCHW view with channels_last strides (no copy) - what the reference-facing API returns."""
        return self.dense().permute(0, 3, 1, 2)


def new_act(n, h, w, c, device, needs_grad=True):
    return Act(torch.empty(n * h * w, c, device=device, dtype=torch.float32), n, h, w, c, needs_grad=needs_grad)


class WeightRef:
    """K-major [Cout_eff, K] operand of a Conv2d / Linear parameter plus the way its gradient goes back.

    kinds: 'linear' [out, in]; 'cl' conv stored channels_last (zero-copy OHWI view); 'packed' conv with Cin % 4 != 0
    (re-packed to pad4(Cin) every step); any kind may be row-padded to `pad_cout` (classifier: nc -> pad4(nc))."""

    def __init__(self, g, param, pad_cout=None):
        self.g, self.param = g, param
        if param.dim() == 2:
            self.kind, self.r, self.s = "linear", 1, 1
            self.cout, self.cin = param.shape
            wk = param.detach()
            assert wk.is_contiguous()
        else:
            self.cout, ci, self.r, self.s = param.shape
            if ci % 4 != 0:
                self.kind, self.cin = "packed", pad4(ci)
                wc = param.detach().contiguous()
                wk = g.empty(self.cout, self.r * self.s * self.cin)
                g.call("cavp_nchw_to_nhwc", wc.data_ptr(), wk.data_ptr(), self.cout, ci, self.r * self.s, self.cin)
            else:
                self.kind, self.cin = "cl", ci
                v = param.detach().permute(0, 2, 3, 1)
                assert v.is_contiguous(), "conv weights must be stored channels_last (cavp_b200.models converts them)"
                wk = v.reshape(self.cout, self.r * self.s * ci)
        self.K = self.r * self.s * self.cin
        self.cout_eff = self.cout if pad_cout is None else pad_cout
        if self.cout_eff != self.cout:
            wp = g.zeros(self.cout_eff, self.K)
            wp[:self.cout].copy_(wk)  # plumbing copy
            wk = wp
        self.wk = wk
        self._wt = None
        self._split = None
        self._wt_split = None

    def _presplit(self, w):
        """[hi | lo] TF32 split of a K-major operand, done once per step; the igemm kernel then fetches it by TMA."""
        sp = self.g.empty(2, w.shape[0], w.shape[1])
        self.g.call("cavp_split_tf32", w.data_ptr(), sp[0].data_ptr(), sp[1].data_ptr(), w.numel())
        return sp

    def operand(self):
        """(tensor, lo offset in elements) for cavp_igemm's B operand."""
        if self._split is None:
            cache = self.g.wcache
            if cache is not None and self.kind in ("linear", "cl") and self.cout_eff == self.cout:
                if cache.bf16:
                    hit = cache.operand_bf16(self.param)
                    if hit is not None:  # (fp32 operand for the ABI's fallback slot, no split, bf16 copy)
                        self._split = (self.wk, 0, hit)
                        return self._split
                else:
                    hit = cache.operand(self.param)
                    if hit is not None:
                        self._split = hit
                        return self._split
            self._split = (self._presplit(self.wk), self.wk.numel())
        return self._split

    def operand_t(self):
        """[hi | lo] split of the transposed operand [Cin][taps][Cout_eff] for the data-gradient GEMM (one kernel)."""
        if self._wt_split is None:
            cache = self.g.wcache
            if cache is not None and self.kind in ("linear", "cl") and self.cout_eff == self.cout:
                if cache.bf16:
                    hit = cache.operand_t_bf16(self.param)
                    if hit is not None:
                        return (self.wk, 0, hit)
                else:
                    hit = cache.operand_t(self.param)
                    if hit is not None:
                        self._wt_split = hit[0]
                        self._wt_lo_off = hit[1]
                        return self._wt_split, self._wt_lo_off
            taps = self.r * self.s
            sp = self.g.empty(2, self.cin, taps * self.cout_eff)
            self.g.call("cavp_transpose_split", self.wk.data_ptr(), sp[0].data_ptr(), sp[1].data_ptr(), self.cout_eff,
                        self.cin, taps * self.cin, taps * self.cout_eff, taps, self.cin, self.cout_eff)
            self._wt_split = sp[0]
            self._wt_lo_off = sp[0].numel()
        return self._wt_split, self._wt_lo_off

    def transposed(self):
        """[Cin][taps][Cout_eff] for the data-gradient GEMM."""
        if self._wt is None:
            taps = self.r * self.s
            wt = self.g.empty(self.cin, taps * self.cout_eff)
            self.g.call("cavp_transpose", self.wk.data_ptr(), wt.data_ptr(), self.cout_eff, self.cin, taps * self.cin,
                        taps * self.cout_eff, taps, self.cin, self.cout_eff)
            self._wt = wt
        return self._wt

    def deliver_grad(self, dwk):
        """dwk: [Cout_eff, K] -> gradient with the parameter's shape / memory format."""
        p, g = self.param, self.g
        dwk = dwk[:self.cout]
        if self.kind == "linear":
            grad = dwk
        elif self.kind == "cl":
            grad = dwk.view(self.cout, self.r, self.s, self.cin).permute(0, 3, 1, 2)
        else:
            ci = p.shape[1]
            grad = g.empty(self.cout, ci, self.r, self.s)
            g.call("cavp_nhwc_to_nchw", dwk.data_ptr(), grad.data_ptr(), self.cout, ci, self.r * self.s, self.cin)
        g.add_param_grad(p, grad)


class WeightSplitCache:
    """Persistent [hi | lo] TF32 splits of every K-major weight operand of a model (channels_last conv weights with
    Cin % 4 == 0 and Linear weights), refreshed by ONE kernel launch per step (cavp_split_tf32_multi).  The buffers and
    the device table are built once; they are rebuilt if a parameter's storage moves."""

    def __init__(self, params, device, bf16=False):
        self.params = list(params)
        self.device = device
        self.bf16 = bf16  # prec 3: bf16 copies (K-major and transposed) instead of the TF32 hi | lo splits
        if bf16:
            self._init_bf16()
            return
        self.ptrs = [p.data_ptr() for p in self.params]
        total = sum(p.numel() for p in self.params)
        self.buf = torch.empty(2, total, device=device, dtype=torch.float32)
        self.views, rows, sizes, off = {}, [], [], 0
        for p in self.params:
            n = p.numel()
            hi, lo = self.buf[0, off:off + n], self.buf[1, off:off + n]
            self.views[id(p)] = (hi, lo)
            rows.append((p.data_ptr(), hi.data_ptr(), lo.data_ptr(), n))
            sizes.append(n)
            off += n
        chunk = _C.query("cavp_opt_chunk_elems")
        work = [(i, c) for i, n in enumerate(sizes) for c in range((n + chunk - 1) // chunk)]
        self.table = torch.tensor(rows, dtype=torch.int64).to(device)
        self.work = torch.tensor(work, dtype=torch.int32).reshape(-1, 2).to(device)
        self.lo_off = total  # lo = hi + total elements for every weight
        self._t = None  # transposed operands of the data-gradient GEMMs: built on the first training step

    def _build_transposed(self):
        """[Cin][taps][Cout] hi | lo splits of every weight (the dgrad B operand), one persistent buffer + one table."""
        total = self.lo_off
        buf = torch.empty(2, total, device=self.device, dtype=torch.float32)
        views, rows, work, off = {}, [], [], 0
        for i, p in enumerate(self.params):
            n = p.numel()
            cout = p.shape[0]
            cin = p.shape[1]
            taps = n // (cout * cin)
            hi, lo = buf[0, off:off + n], buf[1, off:off + n]
            views[id(p)] = hi.view(cin, taps * cout)
            tiles_c, tiles_r = (cin + 31) // 32, (cout + 31) // 32
            rows.append((p.data_ptr(), hi.data_ptr(), lo.data_ptr(), cout | (cin << 32), taps * cin, taps * cout, cin,
                         cout, tiles_c | (tiles_r << 32)))
            work.extend((i, t) for t in range(taps * tiles_c * tiles_r))
            off += n
        self._t = {"buf": buf, "views": views,
                   "table": torch.tensor(rows, dtype=torch.int64).to(self.device),
                   "work": torch.tensor(work, dtype=torch.int32).reshape(-1, 2).to(self.device)}

    def refresh_transposed(self, g):
        if self._t is None:
            self._build_transposed()
        g.call("cavp_transpose_split_multi", self._t["table"].data_ptr(), self._t["work"].data_ptr(),
               self._t["work"].shape[0], 0)

    def operand_t(self, param):
        """(hi tensor [Cin, taps*Cout], lo offset in elements) or None"""
        if self._t is None:
            return None
        v = self._t["views"].get(id(param))
        return None if v is None else (v, self.lo_off)

    def _init_bf16(self):
        """bf16 operands of the configs[2] path: [cout][K] copies for forward and [Cin][taps][Cout] transposes for the
        data-gradient GEMMs; each refreshed by one launch per step (cavp_cvt_bf16_multi / cavp_transpose_split_multi)."""
        self.ptrs = [p.data_ptr() for p in self.params]
        total = sum((p.numel() + 7) // 8 * 8 for p in self.params)  # every operand starts 16-byte aligned
        self.buf = torch.empty(2, total, device=self.device, dtype=torch.bfloat16)
        chunk = _C.query("cavp_opt_chunk_elems")
        self.views, self.t_views, rows, trows, work, twork, off = {}, {}, [], [], [], [], 0
        for i, p in enumerate(self.params):
            n = p.numel()
            cout, cin = p.shape[0], p.shape[1]
            taps = n // (cout * cin)
            fw, tr = self.buf[0, off:off + n], self.buf[1, off:off + n]
            self.views[id(p)] = fw.view(cout, taps * cin)
            self.t_views[id(p)] = tr.view(cin, taps * cout)
            rows.append((p.data_ptr(), fw.data_ptr(), n))
            work.extend((i, c) for c in range((n + chunk - 1) // chunk))
            tiles_c, tiles_r = (cin + 31) // 32, (cout + 31) // 32
            trows.append((p.data_ptr(), tr.data_ptr(), 0, cout | (cin << 32), taps * cin, taps * cout, cin, cout,
                          tiles_c | (tiles_r << 32)))
            twork.extend((i, t) for t in range(taps * tiles_c * tiles_r))
            off += (n + 7) // 8 * 8
        dev = self.device
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.work = torch.tensor(work, dtype=torch.int32).reshape(-1, 2).to(dev)
        self.t_table = torch.tensor(trows, dtype=torch.int64).to(dev)
        self.t_work = torch.tensor(twork, dtype=torch.int32).reshape(-1, 2).to(dev)

    def refresh_bf16(self, g, train):
        g.call("cavp_cvt_bf16_multi", self.table.data_ptr(), self.work.data_ptr(), self.work.shape[0])
        if train:
            g.call("cavp_transpose_split_multi", self.t_table.data_ptr(), self.t_work.data_ptr(), self.t_work.shape[0], 1)

    def operand_bf16(self, param):
        return self.views.get(id(param))

    def operand_t_bf16(self, param):
        return self.t_views.get(id(param))

    @staticmethod
    def eligible(model, bf16=False):
        out = []
        for mod in model.modules():
            w = getattr(mod, "weight", None)
            if w is None or not isinstance(mod, (torch.nn.Linear, torch.nn.Conv2d)):
                continue
            if bf16 and (w.shape[0] % 8 or w.shape[1] % 8):
                continue  # 8-element bf16 chunks / 16-byte TMA strides; such layers run the TF32 kernels
            if isinstance(mod, torch.nn.Linear) and w.is_contiguous():
                out.append(w)
            elif isinstance(mod, torch.nn.Conv2d) and w.shape[1] % 4 == 0 and w.permute(0, 2, 3, 1).is_contiguous():
                out.append(w)
        return out

    def valid(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self.ptrs))

    def operand(self, param):
        """(hi tensor viewed [cout, K], lo offset in elements) or None"""
        v = self.views.get(id(param))
        if v is None:
            return None
        cout = param.shape[0]
        return v[0].view(cout, -1), self.lo_off


class Graph:
    def __init__(self, device, prec=2, train=True, sync_bn_group=None, grad_sink=None):
        self.device = device
        # grad_sink: a cavp_b200.parallel.FlatGradBuffer - weight-gradient kernels write straight into its views
        self.grad_sink = grad_sink
        self._dw_prezeroed = False
        self.callbacks = {}  # name -> callable, fired by the matching tape marker during backward()
        self.prec = prec  # 2 = fp32 parity (3xTF32 + promotion), 1 = plain TF32, 3 = bf16 operands (configs[2])
        # bf16 mode: forward / dgrad GEMMs outside the bf16 kernel's envelope (3/4-channel stems, the padded classifier,
        # the InfoNCE similarity) keep fp32-grade products - a TF32-rounded stem would be amplified ~1000x by the
        # batch-stat BN chain; weight gradients run plain TF32 (one MMA per product)
        self.prec_tf = 2 if prec == 3 else prec
        self.prec_wg = 1 if prec == 3 else prec
        self.train = train
        self.tape = []
        self.param_grads = {}  # id(param) -> gradient tensor
        self.sync_bn_group = sync_bn_group
        self._const = {}
        self.launches = 0
        self.profile = None  # set to [] to record (name, start_event, end_event, flops, bytes) per kernel launch
        self._stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
        _C.lib()
        self._fns = _C._fns
        self._work = (0.0, 0.0, "")
        self.wcache = None

    def use_weight_cache(self, model):
        """Refresh (one launch) the persistent TF32 splits of the model's weight operands for this step."""
        bf16 = self.prec == 3
        cache = model.__dict__.get("_cavp_wsplit")
        if cache is None or cache.device != self.device or cache.bf16 != bf16 or not cache.valid():
            params = WeightSplitCache.eligible(model, bf16)
            if not params:
                return
            cache = WeightSplitCache(params, self.device, bf16)
            model.__dict__["_cavp_wsplit"] = cache
        if bf16:
            cache.refresh_bf16(self, self.train)
        else:
            self.call("cavp_split_tf32_multi", cache.table.data_ptr(), cache.work.data_ptr(), cache.work.shape[0])
            if self.train:
                cache.refresh_transposed(self)  # the dgrad operands of every layer, one launch
        self.wcache = cache

    def all_reduce_stats(self, t):
        """In-place sum of a small SyncBatchNorm statistics vector over `sync_bn_group`: one peer-memory kernel
        (csrc/peer.cu) when the ranks of the group can map each other's memory, else torch.distributed (NCCL)."""
        from .parallel import peer_reduce_for
        red = peer_reduce_for(self.sync_bn_group, self.device)
        if red is not None and t.numel() * t.element_size() <= red.SLOT_BYTES:
            self.call("cavp_peer_allreduce", *red.args(t))
        else:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.sync_bn_group)

    # ------------------------------------------------------------------ small helpers
    def call(self, name, *args):
        self.launches += 1
        if self.profile is None:
            rc = self._fns[name](*args, self._stream)
            if rc != 0:
                raise _C.CavpError(f"{name} failed with status {rc}")
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.call(name, *args, self._stream)
        e1.record()
        self.profile.append((name, e0, e1) + self._work)
        self._work = (0.0, 0.0, "")

    def work(self, flops=0.0, nbytes=0.0, tag=""):
        """algorithmic work of the NEXT kernel launch (only used when profiling)."""
        if self.profile is not None:
            self._work = (float(flops), float(nbytes), tag)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, device=self.device, dtype=dtype)

    def zeros(self, *shape, dtype=torch.float32):
        t = torch.empty(*shape, device=self.device, dtype=dtype)
        self.call("cavp_zero", t.data_ptr(), t.numel() * t.element_size())
        return t

    def const_vec(self, n, value):
        key = (n, value)
        if key not in self._const:
            self._const[key] = torch.full((n,), float(value), device=self.device)
        return self._const[key]

    def zero_act(self, a):
        if a.off == 0 and a.c == a.ld:
            self.call("cavp_zero", a.buf.data_ptr(), a.rows * a.ld * 4)
        else:
            self.call("cavp_fill_strided", a.ptr, a.ld, a.rows, a.c, 0.0)

    def copy_act(self, dst, src):
        self.call("cavp_bn_apply", src.ptr, src.ld, self.const_vec(src.c, 1).data_ptr(),
                  self.const_vec(src.c, 0).data_ptr(), 0, 0, dst.ptr, dst.ld, src.rows, src.c, ACT_NONE, 0.0)

    def add_act(self, dst, src):
        """dst += src"""
        self.call("cavp_bn_apply", src.ptr, src.ld, self.const_vec(src.c, 1).data_ptr(),
                  self.const_vec(src.c, 0).data_ptr(), dst.ptr, dst.ld, dst.ptr, dst.ld, src.rows, src.c, ACT_NONE, 0.0)

    def grad_of(self, a):
        """Gradient Act of `a` (None if nobody produced one)."""
        root = a.root
        if root.grad is None:
            return None
        g = root.grad
        if a is root:
            return g
        return Act(g.buf, a.n, a.h, a.w, a.c, a.off - root.off + g.off)

    def grad_target(self, a):
        """-> (gradient Act, accumulate?)  allocating the gradient buffer on first use."""
        root = a.root
        fresh = root.grad is None
        if fresh:
            root.grad = Act(torch.empty(root.rows, root.ld, device=self.device), root.n, root.h, root.w, root.c, root.off)
            if a is not root and a.c != root.c:  # a partial writer comes first: start the whole buffer from zero
                self.call("cavp_zero", root.grad.buf.data_ptr(), root.grad.buf.numel() * 4)
                fresh = False
        return self.grad_of(a), (not fresh)

    def mark(self, name):
        """Tape marker: during backward(), `callbacks[name]` runs once every op recorded AFTER this point has finished
        its backward (used to launch a gradient bucket's all-reduce while the rest of the backward still runs)."""
        if self.train:
            def fire():
                cb = self.callbacks.get(name)
                if cb is not None:
                    cb()
            self.tape.append(fire)

    def weight_grad_buffer(self, wr, co, K):
        """[co, K] K-major buffer for a weight gradient: the parameter's own view inside the flat all-reduce buffer when
        its memory format is exactly what the kernel writes (Linear, channels_last conv), else a fresh tensor."""
        sink = self.grad_sink
        if (sink is not None and wr.kind in ("linear", "cl") and wr.cout_eff == wr.cout
                and id(wr.param) not in self.param_grads):
            v = sink.view_of(wr.param)
            if v is not None:
                self._dw_prezeroed = bool(getattr(sink, "prezeroed", False))
                return (v if wr.kind == "linear" else v.permute(0, 2, 3, 1)).reshape(co, K)
        self._dw_prezeroed = False
        return self.empty(co, K)

    def add_param_grad(self, p, g):
        key = id(p)
        if key in self.param_grads:
            old = self.param_grads[key]
            if not (old.is_contiguous() and g.is_contiguous()):
                old = old.contiguous(memory_format=torch.channels_last) if old.dim() == 4 else old.contiguous()
                g = g.contiguous(memory_format=torch.channels_last) if g.dim() == 4 else g.contiguous()
                self.param_grads[key] = old
            self.call("cavp_add_inplace", old.data_ptr(), g.data_ptr(), g.numel(), 1.0)
        else:
            self.param_grads[key] = g

    def accumulate_grad(self, a, g, res_mod=0, res_div=0):
        """a.grad += g where g has the row layout of a consumer that read `a` through an igemm residual epilogue:
        consumer row r reads a-row (r / res_div) % res_mod."""
        da, accumulate = self.grad_target(a)
        if res_div > 1:  # broadcast over the pixels of an image: sum over pixels
            assert res_mod == 0
            tmp = da if (not accumulate and da.ld == da.c) else new_act(a.n, a.h, a.w, a.c, self.device)
            self.call("cavp_pixel_sum", g.ptr, g.ld, tmp.ptr, a.rows, res_div, a.c, 1.0)
            if tmp is not da:
                (self.add_act if accumulate else self.copy_act)(da, tmp)
            return
        reps = 1 if res_mod == 0 else g.rows // res_mod
        nrows = g.rows // reps
        for i in range(reps):
            part = Act(g.buf[i * nrows:(i + 1) * nrows], a.n, a.h, a.w, a.c, g.off)
            if i == 0 and not accumulate:
                self.copy_act(da, part)
            else:
                self.add_act(da, part)

    # ------------------------------------------------------------------ igemm wrappers
    def _igemm(self, x, wk, ncols, ldw, y, *, geom, src_hw=None, dgrad=0, y_pre=None, scale=None, shift=None, res=None,
               res_mod=0, res_div=0, stats_ptr=0, ldstat=0, act=ACT_NONE, splits=1):
        """geom = (ho, wo, r, s, stride, pad, dil).  x: A-operand source Act; y: output Act (rows = x.n*ho*wo)."""
        ho, wo, r, s, stride, pad, dil = geom
        hs, ws = (x.h, x.w) if src_hw is None else src_hw
        b_lo_off, w16 = 0, None
        if isinstance(wk, tuple):  # pre-split weight operand -> TMA path; (fp32, 0, bf16 copy) -> bf16 kernel
            if len(wk) == 3:
                wk, b_lo_off, w16 = wk
            else:
                wk, b_lo_off = wk
        self.work(flops=2.0 * x.n * ho * wo * ncols * r * s * x.c,
                  tag=f"{'dgrad' if dgrad else 'fwd'} M{x.n * ho * wo} N{ncols} K{r * s * x.c} k{r} s{stride} d{dil} splits{splits}")
        if w16 is not None:
            self.call("cavp_igemm_bf16", x.ptr, wk.data_ptr(), w16.data_ptr(), y.ptr, 0 if y_pre is None else y_pre.ptr,
                      _C.ptr(scale), _C.ptr(shift), 0 if res is None else res.ptr, stats_ptr, x.n, hs, ws, x.c, x.ld, ho,
                      wo, r, s, stride, pad, dil, dgrad, ncols, ldw, y.ld, 0 if res is None else res.ld, res_mod,
                      res_div, ldstat, act, LEAKY_SLOPE, splits, b_lo_off)
            return
        self.call("cavp_igemm", x.ptr, wk.data_ptr(), y.ptr, 0 if y_pre is None else y_pre.ptr, _C.ptr(scale),
                  _C.ptr(shift), 0 if res is None else res.ptr, stats_ptr, x.n, hs, ws, x.c, x.ld, ho, wo, r, s, stride,
                  pad, dil, dgrad, ncols, ldw, y.ld, 0 if res is None else res.ld, res_mod, res_div, ldstat, act,
                  LEAKY_SLOPE, splits, self.prec_tf, b_lo_off)

    @staticmethod
    def fwd_splits(M, ncols, K):
        bn = 128 if ncols > 64 else 64
        tiles = ((M + 127) // 128) * ((ncols + bn - 1) // bn)
        num_kb = (K + 31) // 32
        if ncols % 256 == 0 and num_kb >= 128:
            # long-K layers whose 256-column pair tiles (csrc/igemm_ws2x.cuh) would fill the 74 CTA pairs unevenly
            # (ASPP 3x3 convs: 98 items): a deterministic split-K of 2-3 restores the fill at the price of one small
            # slab-sum pass, and keeps them on the 256-column kernel (85 % instead of 66 % tensor-pipe active)
            items = ((((M + 127) // 128) + 1) // 2) * (ncols // 256)
            fill = lambda n: n / (((n + 73) // 74) * 74)
            if 74 <= items < 74 * 6 and fill(items) < 0.85:
                for sk in (2, 3):
                    if fill(items * sk) >= 0.85:
                        return sk
        if tiles >= 96 or num_kb < 16:
            return 1
        return max(1, min(num_kb // 8, (NUM_SMS + tiles - 1) // tiles))

    @staticmethod
    def wgrad_splits(P, cout, K):
        """Split count of the pixel reduction: CTAs = tiles * splits run in waves of one CTA per SM, so the cost is
        waves * (k-blocks per split + a fixed per-CTA overhead of ~12 k-blocks: setup, pipeline fill, red.add epilogue).
        A count that lands just past a multiple of 148 CTAs wastes most of a wave (measured: 300 CTAs took 3 waves)."""
        bn = 128 if K > 64 else 64
        tiles = ((cout + 127) // 128) * ((K + bn - 1) // bn)
        num_kb = (P + 31) // 32
        overhead = int(os.environ.get("CAVP_WGRAD_OVERHEAD", "12"))
        best, best_cost = 1, None
        for s in range(1, max(1, min(num_kb // 4, 96)) + 1):
            waves = (tiles * s + NUM_SMS - 1) // NUM_SMS
            cost = waves * ((num_kb + s - 1) // s + overhead)
            if best_cost is None or cost < best_cost:
                best, best_cost = s, cost
        return best

    def wgrad_bf16_ok(self, cin, cout):
        """bf16 row: weight gradients with bf16 operands (csrc/igemm_wgrad_bf16.cuh) inside that kernel's envelope."""
        return (self.prec == 3 and cin % 8 == 0 and cout % 8 == 0 and cout > 128
                and os.environ.get("CAVP_WGRAD_BF16", "1") != "0")

    @staticmethod
    def wgrad_via_tma(P, cout, K):
        """The split pass over dY costs 12 B per element; it pays when enough column tiles (K / 128) re-read dY."""
        forced = os.environ.get("CAVP_WGRAD_TMA")
        if forced is not None:
            return forced != "0"
        return cout >= 32 and (K >= 1024 or (K >= 512 and cout >= 512))

    @staticmethod
    def colreduce_blocks(M, C):
        c4 = C // 4
        lc = max((64, 32, 16), key=lambda cand: (1000 * c4 // (((c4 + cand - 1) // cand) * cand), cand))
        # ^ column lanes per block, the same choice as csrc/elementwise.cu:cavp_colreduce
        return max(1, min((M + 31) // 32, 4 * NUM_SMS // max(1, (c4 + lc - 1) // lc)))

    def stats_from_tensor(self, y, stats):
        """BN statistics partials (sum, sum of squares) of an already materialised tensor (split-K path)."""
        stats_t, stats_ptr, nparts, ldstat = stats
        window = Act(stats_t.view(nparts * 2, ldstat), 1, 1, nparts * 2, y.c, (stats_ptr - stats_t.data_ptr()) // 4)
        self.zero_act(window)
        nblk = min(nparts, self.colreduce_blocks(y.rows, y.c))
        self.call("cavp_colreduce", y.ptr, y.ld, 0, 0, y.ptr, y.ld, self.const_vec(y.c, 0).data_ptr(),
                  self.const_vec(y.c, 1).data_ptr(), y.rows, y.c, ACT_NONE, 0.0, 0, 0, stats_ptr, ldstat, nblk, 0, 0)

    # ------------------------------------------------------------------ conv / linear
    def conv(self, x, w, *, stride=1, pad=0, dil=1, bias=None, act=ACT_NONE, want_stats=False, out=None, res=None,
             res_mod=0, res_div=0, save_pre=False, stats_buf=None, pad_cout=None, name=""):
        """y = act(conv(x, w) + bias + res)  (Linear = 1x1 conv over tokens).  `w`, `bias` are nn.Parameters.
        want_stats: also produce the BatchNorm sum / sum-of-squares partials of y.  Returns (y, stats)."""
        wr = w if isinstance(w, WeightRef) else WeightRef(self, w, pad_cout)
        co, r, s, K = wr.cout_eff, wr.r, wr.s, wr.K
        assert wr.cin == x.c, (name, wr.cin, x.c)
        bias_t = None
        if bias is not None:
            bias_t = bias.detach()
            if co != wr.cout:
                bp = self.zeros(co)
                bp[:wr.cout].copy_(bias_t)
                bias_t = bp
        ho = (x.h + 2 * pad - dil * (r - 1) - 1) // stride + 1
        wo = (x.w + 2 * pad - dil * (s - 1) - 1) // stride + 1
        y = out if out is not None else new_act(x.n, ho, wo, co, self.device)
        assert y.rows == x.n * ho * wo and y.c == co, (name, y.rows, x.n, ho, wo, y.c, co)
        M = y.rows
        geom = (ho, wo, r, s, stride, pad, dil)
        pre = new_act(x.n, ho, wo, co, self.device) if save_pre else None
        stats = None
        if want_stats:
            nparts = ((M + 127) // 128) * 4
            if stats_buf is None:
                stats_t = self.empty(nparts, 2, co)
                stats = (stats_t, stats_t.data_ptr(), nparts, co)
            else:
                stats = stats_buf  # (tensor [nparts, 2, ldstat], pointer offset to our columns, nparts, ldstat)
                assert stats[2] == nparts
        # A conv whose output feeds a BatchNorm (want_stats) gets d(y) from bn_bwd_apply, which can emit the TF32 split /
        # bf16 copy on its way out: the TMA-fed weight-gradient kernels then cost no extra pass, so they are used from
        # K >= 256 when CAVP_WGRAD_FREE_SPLIT=1 (experimental; default: only where the split pass would pay for itself)
        free_split = want_stats and os.environ.get("CAVP_WGRAD_FREE_SPLIT", "0") != "0" and co >= 64 and K >= 256
        if self.train and wr.param.requires_grad and y.off == 0 and y.ld == co and (
                self.wgrad_bf16_ok(x.c, co) or self.wgrad_via_tma(M, co, K) or free_split):
            # BatchNorm's backward then emits d(y) together with its TF32 split / bf16 copy
            y.want_split = "bf16" if self.wgrad_bf16_ok(x.c, co) else "tf32"
        splits = self.fwd_splits(M, co, K)
        if splits > 1 and (save_pre or res_mod or res_div):
            splits = 1
        if splits > 1:
            # forward split-K is deterministic: every split stores its partial product in a private slab and the slabs
            # are summed in a fixed order (red.global.add would make the logits - and near-tie argmax decisions -
            # depend on the arrival order of the atomics)
            simple = bias_t is None and act == ACT_NONE and res is None
            raw = y if (simple and y.ld == co and y.off == 0) else new_act(x.n, ho, wo, co, self.device)
            simple = raw is y
            nk = (K + 31) // 32
            splits = min(splits, nk)
            slabs = new_act(splits * x.n, ho, wo, co, self.device)
            self._igemm(x, wr.operand(), co, K, Act(slabs.buf[:M], x.n, ho, wo, co), geom=geom, splits=-splits)
            self.call("cavp_partials_sum", slabs.ptr, splits, M * co, M * co, 1, raw.ptr)
            if not simple:
                self.call("cavp_bn_apply", raw.ptr, raw.ld, self.const_vec(co, 1).data_ptr(),
                          (bias_t if bias_t is not None else self.const_vec(co, 0)).data_ptr(),
                          0 if res is None else res.ptr, 0 if res is None else res.ld, y.ptr, y.ld, M, co, act,
                          LEAKY_SLOPE)
            if want_stats:
                self.stats_from_tensor(y, stats)
        elif act == ACT_GELU and y.ld == co and y.off == 0:
            # GELU runs as its own HBM-bound pass (csrc/fusion.cu:gelu_fwd_kernel): the GEMM writes the pre-activation
            # (kept for the backward) with the plain epilogue
            tgt = pre if pre is not None else y
            self._igemm(x, wr.operand(), co, K, tgt, geom=geom, shift=bias_t, res=res, res_mod=res_mod,
                        res_div=res_div, stats_ptr=0, ldstat=0, act=ACT_NONE)
            self.work(nbytes=8.0 * M * co)
            self.call("cavp_gelu_fwd", tgt.ptr, y.ptr, M * co)
            assert stats is None
        elif (res is not None and res.ptr != y.ptr and not res_div and act == ACT_NONE and pre is None
              and stats is None and M >= 65536):
            # large GEMM + residual: the persistent / pair kernels with the plain epilogue, then one HBM-bound add
            # (a residual fetched inside the epilogue serialises four exposed global-latency waits per tile)
            self._igemm(x, wr.operand(), co, K, y, geom=geom, shift=bias_t)
            reps = 1 if res_mod == 0 else M // res_mod
            nrows = M // reps
            for i in range(reps):
                part = Act(y.buf[i * nrows:(i + 1) * nrows], 1, 1, nrows, y.c, y.off)
                rpart = Act(res.buf[:nrows], 1, 1, nrows, res.c, res.off)
                self.add_act(part, rpart)
        else:
            self._igemm(x, wr.operand(), co, K, y, geom=geom, y_pre=pre, shift=bias_t, res=res, res_mod=res_mod,
                        res_div=res_div, stats_ptr=0 if stats is None else stats[1],
                        ldstat=0 if stats is None else stats[3], act=act)

        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None:
                    return
                g = dy
                if act != ACT_NONE or bias is not None:
                    # g = dy * act'(.) and the bias gradient (column sums of g) from one pass
                    zin, gout = None, None
                    if act == ACT_GELU:
                        assert dy.ld == dy.c and pre is not None
                        g = new_act(y.n, y.h, y.w, co, self.device)
                        self.call("cavp_gelu_bwd", dy.ptr, pre.ptr, g.ptr, M * co)
                    elif act in (ACT_RELU, ACT_LEAKY):
                        zin = y
                        gout = new_act(y.n, y.h, y.w, co, self.device)
                    elif act != ACT_NONE:
                        raise _C.CavpError("unsupported fused activation in backward")
                    if zin is not None or bias is not None:
                        nblk = self.colreduce_blocks(M, co)
                        partials = self.empty(nblk, 2, co)
                        self.call("cavp_colreduce", g.ptr, g.ld, 0 if zin is None else zin.ptr,
                                  0 if zin is None else zin.ld, 0, 0, 0, 0, M, co, act if zin is not None else ACT_NONE,
                                  LEAKY_SLOPE, 0 if gout is None else gout.ptr, 0 if gout is None else gout.ld,
                                  partials.data_ptr(), co, nblk, 0, 0)
                        if gout is not None:
                            g = gout
                        if bias is not None and bias.requires_grad:
                            sums = self.empty(2, co)
                            self.call("cavp_partials_sum", partials.data_ptr(), nblk, co, co, 2, sums.data_ptr())
                            self.add_param_grad(bias, sums[0, :wr.cout])
                if res is not None and res.needs_grad:
                    self.accumulate_grad(res, g, res_mod, res_div)
                if wr.param.requires_grad:
                    dwk = self.weight_grad_buffer(wr, co, K)
                    wsplits = self.wgrad_splits(M, co, K)
                    if wsplits > 1 and not self._dw_prezeroed:  # (the flat gradient buffer is zeroed once per step)
                        self.call("cavp_zero", dwk.data_ptr(), dwk.numel() * 4)
                    if self.wgrad_bf16_ok(x.c, co):
                        g16 = g.split if (g is dy and g.split is not None and g.split.dtype == torch.bfloat16) else None
                        if g16 is None:
                            g16 = self.empty(M, co, dtype=torch.bfloat16)
                            self.call("cavp_cvt_bf16_2d", g.ptr, g.ld, M, co, g16.data_ptr())
                        self.work(flops=2.0 * M * co * K,
                                  tag=f"wgrad(bf16) P{M} Cout{co} K{K} k{r} s{stride} d{dil} splits{wsplits}")
                        self.call("cavp_igemm_wgrad_bf16", g16.data_ptr(), x.ptr, dwk.data_ptr(), x.n, x.h, x.w, x.c,
                                  x.ld, ho, wo, r, s, stride, pad, dil, co, wsplits)
                    elif self.wgrad_via_tma(M, co, K) or (g is dy and g.split is not None
                                                          and g.split.dtype == torch.float32):
                        # dY pre-split once (dense hi | lo) and fetched by TMA by every one of the K/128 column tiles
                        gsp = g.split if (g is dy and g.split is not None and g.split.dtype == torch.float32) else None
                        if gsp is None:
                            gsp = self.empty(2, M, co)
                            self.call("cavp_split_tf32_2d", g.ptr, g.ld, M, co, gsp[0].data_ptr(), gsp[1].data_ptr())
                        self.work(flops=2.0 * M * co * K,
                                  tag=f"wgrad(tma) P{M} Cout{co} K{K} k{r} s{stride} d{dil} splits{wsplits}")
                        self.call("cavp_igemm_wgrad_tma", gsp[0].data_ptr(), gsp[0].numel(), x.ptr, dwk.data_ptr(), x.n,
                                  x.h, x.w, x.c, x.ld, ho, wo, r, s, stride, pad, dil, co, wsplits, self.prec_wg)
                    else:
                        self.work(flops=2.0 * M * co * K,
                                  tag=f"wgrad P{M} Cout{co} K{K} k{r} s{stride} d{dil} splits{wsplits}")
                        self.call("cavp_igemm_wgrad", g.ptr, x.ptr, dwk.data_ptr(), x.n, x.h, x.w, x.c, x.ld, ho, wo, r,
                                  s, stride, pad, dil, co, g.ld, wsplits, self.prec_wg)
                    wr.deliver_grad(dwk)
                if x.needs_grad:
                    wt = wr.operand_t()
                    dx, accumulate = self.grad_target(x)
                    dsplits = self.fwd_splits(x.rows, wr.cin, r * s * co)
                    gs = Act(g.buf, y.n, ho, wo, co, g.off)
                    dgeom = (x.h, x.w, r, s, stride, pad, dil)
                    if dsplits > 1:
                        if not accumulate:
                            self.zero_act(dx)
                        self._igemm(gs, wt, wr.cin, r * s * co, dx, geom=dgeom, dgrad=1, splits=dsplits)
                    else:
                        self._igemm(gs, wt, wr.cin, r * s * co, dx, geom=dgeom, dgrad=1, res=dx if accumulate else None)
            self.tape.append(bwd)
        return y, stats

    # ------------------------------------------------------------------ BatchNorm (+ activation, + residual)
    def bn_act(self, y, stats, bn, *, act=ACT_RELU, res=None, out=None):
        """z = act(BN(y) + res) with batch statistics taken from the igemm epilogue partials (train mode)."""
        C = y.c
        z = out if out is not None else new_act(y.n, y.h, y.w, C, self.device)
        coeffs = self.empty(4, C)  # mean, invstd, scale, shift
        stats_t, stats_ptr, nparts, ldstat = stats
        count = float(y.rows)
        momentum = bn.momentum if bn.momentum is not None else 0.1
        sync = self.sync_bn_group is not None and isinstance(bn, torch.nn.SyncBatchNorm)
        head = (stats_ptr, nparts, ldstat, C)
        tail = (bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                momentum, bn.eps, coeffs[0].data_ptr(), coeffs[1].data_ptr(), coeffs[2].data_ptr(),
                coeffs[3].data_ptr())
        nbt = bn.num_batches_tracked.data_ptr() if bn.num_batches_tracked is not None else 0
        count_dev, sync_sums = 0, None
        if sync:
            # SyncBatchNorm (main_vpo_mono.py:130): local fp64 sums + local count -> ONE all-reduce -> global statistics.
            # The global count stays on the device (finalize mode 2 and the backward read it there): no host sync.
            import torch.distributed as dist
            sums = self.empty(2 * C + 1, dtype=torch.float64)
            self.call("cavp_bn_finalize", *head, count, *tail, sums.data_ptr(), 1, 0)
            self.all_reduce_stats(sums)
            self.call("cavp_bn_finalize", *head, count, *tail, sums.data_ptr(), 2, nbt)
            count_dev = sums.data_ptr() + 8 * 2 * C
            sync_sums = sums
        else:
            self.call("cavp_bn_finalize", *head, count, *tail, 0, 0, nbt)
        self.call("cavp_bn_apply", y.ptr, y.ld, coeffs[2].data_ptr(), coeffs[3].data_ptr(),
                  0 if res is None else res.ptr, 0 if res is None else res.ld, z.ptr, z.ld, y.rows, C, act, LEAKY_SLOPE)

        def bwd(sync_sums=sync_sums):  # (keeps the all-reduced statistics buffer alive: count_dev points into it)
            dz = self.grad_of(z)
            if dz is None:
                return
            M = y.rows
            nblk = self.colreduce_blocks(M, C)
            partials = self.empty(nblk, 2, C)
            # without a residual the activation input is fma(y, scale, shift): the kernels recompute the ReLU mask from y
            # (which they read anyway) instead of reading z back; with a residual z is the only record of the sign
            remask = act != ACT_NONE and res is None
            zin = z if (act != ACT_NONE and not remask) else None
            zs, zb = (coeffs[2].data_ptr(), coeffs[3].data_ptr()) if remask else (0, 0)
            self.call("cavp_colreduce", dz.ptr, dz.ld, 0 if zin is None else zin.ptr, 0 if zin is None else zin.ld,
                      y.ptr, y.ld, coeffs[0].data_ptr(), coeffs[1].data_ptr(), M, C, act, LEAKY_SLOPE, 0, 0,
                      partials.data_ptr(), C, nblk, zs, zb)
            sums = self.empty(2, C)
            self.call("cavp_partials_sum", partials.data_ptr(), nblk, C, C, 2, sums.data_ptr())
            local = sums
            if sync:
                import torch.distributed as dist
                local = sums.clone()
                self.all_reduce_stats(sums)
            self.add_param_grad(bn.bias, local[0])
            self.add_param_grad(bn.weight, local[1])
            dy, acc_y = self.grad_target(y)
            assert not acc_y, "a BatchNorm input has exactly one consumer"
            dres, tmp = None, None
            if res is not None and res.needs_grad:
                dres, acc_r = self.grad_target(res)
                if acc_r:
                    tmp = new_act(res.n, res.h, res.w, res.c, self.device)
            tgt = tmp if tmp is not None else dres
            hi_ptr = lo_ptr = split_bf16 = 0
            if y.want_split and dy.off == 0 and dy.ld == C:
                if y.want_split == "bf16":  # consumer = the bf16 weight-gradient kernel
                    dy.split = self.empty(M, C, dtype=torch.bfloat16)
                    hi_ptr, split_bf16 = dy.split.data_ptr(), 1
                else:
                    dy.split = self.empty(2, M, C)
                    hi_ptr = dy.split[0].data_ptr()
                    lo_ptr = dy.split[1].data_ptr() if self.prec_wg == 2 else 0  # plain TF32 reads the hi part only
            self.call("cavp_bn_bwd_apply", dz.ptr, dz.ld, 0 if zin is None else zin.ptr, 0 if zin is None else zin.ld,
                      y.ptr, y.ld, coeffs[0].data_ptr(), coeffs[1].data_ptr(), bn.weight.data_ptr(), sums.data_ptr(),
                      1.0 / count, M, C, act, LEAKY_SLOPE, dy.ptr, dy.ld, 0 if tgt is None else tgt.ptr,
                      0 if tgt is None else tgt.ld, zs, zb, count_dev, hi_ptr, lo_ptr, split_bf16)
            if tmp is not None:
                self.add_act(dres, tmp)
        self.tape.append(bwd)
        return z

    def conv_bn(self, x, conv_w, bn, *, stride=1, pad=0, dil=1, act=ACT_RELU, res=None, out=None, name=""):
        """conv -> BN -> (+res) -> act.  Eval mode folds BN, residual and activation into the igemm epilogue."""
        if not self.train:
            wr = WeightRef(self, conv_w)
            co, r, s = wr.cout_eff, wr.r, wr.s
            coeffs = self.empty(2, co)
            self.call("cavp_bn_eval_coeffs", bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                      bn.running_var.data_ptr(), bn.eps, co, coeffs[0].data_ptr(), coeffs[1].data_ptr())
            ho = (x.h + 2 * pad - dil * (r - 1) - 1) // stride + 1
            wo = (x.w + 2 * pad - dil * (s - 1) - 1) // stride + 1
            y = out if out is not None else new_act(x.n, ho, wo, co, self.device)
            splits = self.fwd_splits(y.rows, co, wr.K)
            if splits > 1:  # deterministic split-K (private slabs + fixed-order sum), as in conv()
                raw = new_act(x.n, ho, wo, co, self.device)
                splits = min(splits, (wr.K + 31) // 32)
                slabs = new_act(splits * x.n, ho, wo, co, self.device)
                self._igemm(x, wr.operand(), co, wr.K, Act(slabs.buf[:y.rows], x.n, ho, wo, co),
                            geom=(ho, wo, r, s, stride, pad, dil), splits=-splits)
                self.call("cavp_partials_sum", slabs.ptr, splits, y.rows * co, y.rows * co, 1, raw.ptr)
                self.call("cavp_bn_apply", raw.ptr, raw.ld, coeffs[0].data_ptr(), coeffs[1].data_ptr(),
                          0 if res is None else res.ptr, 0 if res is None else res.ld, y.ptr, y.ld, y.rows, co, act,
                          LEAKY_SLOPE)
            else:
                self._igemm(x, wr.operand(), co, wr.K, y, geom=(ho, wo, r, s, stride, pad, dil), scale=coeffs[0],
                            shift=coeffs[1], res=res, act=act)
            return y
        y, stats = self.conv(x, conv_w, stride=stride, pad=pad, dil=dil, want_stats=True, name=name)
        return self.bn_act(y, stats, bn, act=act, res=res, out=out)

    def bn_eval(self, y, bn, *, act=ACT_RELU, out=None):
        """eval-mode BN (+act) of an already materialised tensor (ASPP map_bn over the concat)."""
        C = y.c
        z = out if out is not None else new_act(y.n, y.h, y.w, C, self.device)
        coeffs = self.empty(2, C)
        self.call("cavp_bn_eval_coeffs", bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                  bn.running_var.data_ptr(), bn.eps, C, coeffs[0].data_ptr(), coeffs[1].data_ptr())
        self.call("cavp_bn_apply", y.ptr, y.ld, coeffs[0].data_ptr(), coeffs[1].data_ptr(), 0, 0, z.ptr, z.ld, y.rows, C,
                  act, LEAKY_SLOPE)
        return z

    # ------------------------------------------------------------------ pooling / resampling / views
    def maxpool(self, x, k, stride, pad):
        ho = (x.h + 2 * pad - k) // stride + 1
        wo = (x.w + 2 * pad - k) // stride + 1
        y = new_act(x.n, ho, wo, x.c, self.device)
        idx = self.empty(y.rows, x.c, dtype=torch.int32) if self.train else None
        self.call("cavp_maxpool_fwd", x.ptr, x.ld, y.ptr, y.ld, _C.ptr(idx), x.n, x.h, x.w, x.c, k, stride, pad, ho, wo)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None or not x.needs_grad:
                    return
                dx, accumulate = self.grad_target(x)
                if not accumulate:
                    self.zero_act(dx)
                self.call("cavp_maxpool_bwd", dy.ptr, dy.ld, idx.data_ptr(), dx.ptr, dx.ld, y.rows, x.c)
            self.tape.append(bwd)
        return y

    def global_avgpool(self, x):
        y = new_act(x.n, 1, 1, x.c, self.device)
        hw = x.h * x.w
        self.call("cavp_pixel_sum", x.ptr, x.ld, y.ptr, x.n, hw, x.c, 1.0 / hw)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None or not x.needs_grad:
                    return
                dx, accumulate = self.grad_target(x)
                self.call("cavp_pixel_bcast", dy.ptr, dx.ptr, dx.ld, x.n, hw, x.c, 1.0 / hw, 1 if accumulate else 0)
            self.tape.append(bwd)
        return y

    def global_maxpool(self, x):
        y = new_act(x.n, 1, 1, x.c, self.device)
        hw = x.h * x.w
        arg = self.empty(x.n, x.c, dtype=torch.int32)
        self.call("cavp_pixel_max", x.ptr, x.ld, y.ptr, arg.data_ptr(), x.n, hw, x.c)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None or not x.needs_grad:
                    return
                dx, accumulate = self.grad_target(x)
                assert not accumulate
                self.zero_act(dx)
                self.call("cavp_pixel_max_bwd", dy.ptr, arg.data_ptr(), dx.ptr, dx.ld, x.n, hw, x.c)
            self.tape.append(bwd)
        return y

    def bilinear(self, x, hout, wout, align_corners, out=None):
        y = out if out is not None else new_act(x.n, hout, wout, x.c, self.device)
        self.call("cavp_bilinear_fwd", x.ptr, x.ld, x.h, x.w, y.ptr, y.ld, hout, wout, x.n, x.c, int(align_corners), 0)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None or not x.needs_grad:
                    return
                dx, accumulate = self.grad_target(x)
                tgt = dx if not accumulate else new_act(x.n, x.h, x.w, x.c, self.device)
                self.call("cavp_bilinear_bwd", dy.ptr, dy.ld, hout, wout, tgt.ptr, tgt.ld, x.h, x.w, x.n, x.c,
                          int(align_corners), 0, x.n)
                if accumulate:
                    self.add_act(dx, tgt)
            self.tape.append(bwd)
        return y

    def flatten(self, x):
        """[n, h, w, c] -> [n, 1, 1, h*w*c] (NHWC flatten of models/audio/backbones/vgg.py:19-22); no copy."""
        assert x.off == 0 and x.ld == x.c and x.parent is None
        f = Act(x.buf.view(x.n, x.h * x.w * x.c), x.n, 1, 1, x.h * x.w * x.c, needs_grad=x.needs_grad)
        if self.train:
            def bwd():
                df = self.grad_of(f)
                if df is None:
                    return
                assert x.grad is None
                x.grad = Act(df.buf.view(x.rows, x.c), x.n, x.h, x.w, x.c)
            self.tape.append(bwd)
        return f

    def upsample_to_nchw(self, x, nc, hout, wout):
        """Final prediction: bilinear (align_corners=False) of the first `nc` channels -> torch NCHW tensor."""
        pred = self.empty(x.n, nc, hout, wout)
        self.call("cavp_bilinear_fwd", x.ptr, x.ld, x.h, x.w, pred.data_ptr(), 0, hout, wout, x.n, nc, 0, 1)
        return pred

    def upsample_to_nchw_backward(self, x, nc, dpred, n_valid=None):
        """Seed x.grad from the gradient of the full-resolution prediction (NCHW, contiguous)."""
        assert dpred.is_contiguous()
        dx, accumulate = self.grad_target(x)
        assert not accumulate
        if x.c != nc:
            self.zero_act(dx)
        hout, wout = dpred.shape[-2:]
        self.call("cavp_bilinear_bwd", dpred.data_ptr(), 0, hout, wout, dx.ptr, dx.ld, x.h, x.w, x.n, nc, 0, 1,
                  x.n if n_valid is None else n_valid)

    # ------------------------------------------------------------------ fusion ops
    def layernorm(self, x, ln):
        assert x.ld == x.c and x.off == 0
        y = new_act(x.n, x.h, x.w, x.c, self.device)
        T = x.rows
        mr = self.empty(2, T)
        self.call("cavp_layernorm_fwd", x.ptr, ln.weight.data_ptr(), ln.bias.data_ptr(), y.ptr, mr[0].data_ptr(),
                  mr[1].data_ptr(), T, x.c, ln.eps)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None:
                    return
                assert dy.ld == dy.c
                nparts = _C.query("cavp_layernorm_bwd_nparts", T)
                partials = self.empty(nparts, 2, x.c)
                dx, accumulate = self.grad_target(x)
                self.call("cavp_layernorm_bwd", dy.ptr, x.ptr, ln.weight.data_ptr(), mr[0].data_ptr(), mr[1].data_ptr(),
                          dx.ptr if accumulate else 0, dx.ptr, partials.data_ptr(), T, x.c)
                sums = self.empty(2, x.c)
                self.call("cavp_partials_sum", partials.data_ptr(), nparts, x.c, x.c, 2, sums.data_ptr())
                self.add_param_grad(ln.weight, sums[0])
                self.add_param_grad(ln.bias, sums[1])
            self.tape.append(bwd)
        return y

    def gate(self, q, k, v, heads=4):
        """models/attn.py:73-106 with one key/value token per row.  q [Bq, N, C] (Act n = Bq), k / v [rows, C]."""
        Bq, N, C = q.n, q.h * q.w, q.c
        rows = k.rows
        rep = rows // Bq
        assert rep * Bq == rows and q.ld == C and k.ld == C and v.ld == C
        x = new_act(rows, q.h, q.w, C, self.device)
        attn = self.empty(rows, heads, N)
        self.work(nbytes=4.0 * (Bq * N * C + rows * N * C + rows * heads * N + 2 * rows * C))
        self.call("cavp_gate_fwd", q.ptr, k.ptr, v.ptr, x.ptr, attn.data_ptr(), Bq, rep, N, C, heads)
        if self.train:
            def bwd():
                dx = self.grad_of(x)
                if dx is None:
                    return
                dq, acc_q = self.grad_target(q)
                dk, acc_k = self.grad_target(k)
                dv, acc_v = self.grad_target(v)
                assert not (acc_q or acc_k or acc_v) and dx.ld == C
                self.zero_act(dk)
                self.zero_act(dv)
                self.work(nbytes=4.0 * (rows * N * C + 2 * Bq * N * C + rows * heads * N + 4 * rows * C))
                self.call("cavp_gate_bwd", dx.ptr, q.ptr, k.ptr, v.ptr, attn.data_ptr(), dq.ptr, dk.ptr, dv.ptr, Bq, rep,
                          N, C, heads)
            self.tape.append(bwd)
        return x, attn

    def gather_rows(self, x, idx):
        """y = x[idx] over rows (feature-level audio shuffle, models/cavp_model.py:171)."""
        assert x.ld == x.c
        y = new_act(idx.numel(), 1, 1, x.c, self.device)
        self.call("cavp_gather_rows", x.ptr, idx.data_ptr(), y.ptr, idx.numel(), x.c, 0)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None:
                    return
                dx, accumulate = self.grad_target(x)
                if not accumulate:
                    self.zero_act(dx)
                self.call("cavp_gather_rows", dy.ptr, idx.data_ptr(), dx.ptr, idx.numel(), x.c, 1)
            self.tape.append(bwd)
        return y

    def concat_rows(self, a, b):
        """row-wise concatenation (torch.cat(dim=0)) of two [rows, C] Acts."""
        y = new_act(a.rows + b.rows, 1, 1, a.c, self.device)
        self.copy_act(Act(y.buf[:a.rows], a.rows, 1, 1, a.c), a)
        self.copy_act(Act(y.buf[a.rows:], b.rows, 1, 1, b.c), b)
        if self.train:
            def bwd():
                dy = self.grad_of(y)
                if dy is None:
                    return
                self.accumulate_grad(a, Act(dy.buf[:a.rows], a.n, a.h, a.w, a.c))
                self.accumulate_grad(b, Act(dy.buf[a.rows:], b.n, b.h, b.w, b.c))
            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ backward driver
    def backward(self):
        # each closure is dropped as soon as it has run, so the activations and gradients only it still references go
        # back to the allocator during the backward pass instead of at its end
        tape = self.tape
        while tape:
            tape.pop()()
