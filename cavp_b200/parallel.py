"""Data-parallel plumbing: one process per GPU, the batch sharded across ranks, gradients averaged over NCCL.

The reference wraps the model in DDP (main_vpo_mono.py:131-141: bucketed all-reduce fired from autograd hooks while the
backward is still running).  Here the gradients of a step live in ONE flat fp32 buffer made of a few contiguous
buckets; the weight-gradient kernels write straight into it (`p.grad` are views, the optimisers read the reduced
values in place), and each bucket's all-reduce (average, DDP semantics) is launched the moment the backward tape has
passed the last op that contributes to it - the audio backbone (63 % of the gradient bytes) is complete after the
first quarter of the backward pass, so most of the exchange overlaps the ResNet backward.  Parameters that never
receive gradients (cross_att.pos_embed_*, audio_backbone.cls_head.* - the reason the reference needs
find_unused_parameters=True) are left out of the buffer.
"""
import ctypes
import os
import socket
import sys

import torch
import torch.distributed as dist

from . import _C


class FlatGradBuffer:
    """`params`: a list of parameters (one bucket) or a list of lists (buckets, each contiguous in the flat buffer)."""

    def __init__(self, params, device=None):
        buckets = params if params and isinstance(params[0], (list, tuple)) else [list(params)]
        self.params, self.bucket_of, self.bucket_range, self.offsets = [], {}, [], []
        n = 0
        for b, plist in enumerate(buckets):
            start = n
            for p in plist:
                if not p.requires_grad or id(p) in self.bucket_of:
                    continue
                self.bucket_of[id(p)] = b
                self.params.append(p)
                self.offsets.append(n)
                n += (p.numel() + 3) // 4 * 4  # every view starts 16-byte aligned (vectorised kernels write into it)
            self.bucket_range.append((start, n))
        device = device if device is not None else self.params[0].device
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.views = [self._make_view(i) for i in range(len(self.params))]
        self.used = [False] * len(self.params)
        self._pending = []
        self.prezeroed = False  # True between begin_step() and the end of the step: the views hold zeros

    def _make_view(self, i):
        p = self.params[i]
        v = self.flat[self.offsets[i]:self.offsets[i] + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            n, c, h, w = p.shape
            return v.view(n, h, w, c).permute(0, 3, 1, 2)  # same memory format as the parameter
        return v.view(p.shape)

    def view_for(self, i):
        return self.views[i]

    def view_of(self, p):
        """Gradient view (the parameter's shape and memory format) inside the flat buffer, or None."""
        i = self.index.get(id(p))
        return None if i is None else self.views[i]

    def bucket_tensor(self, b):
        lo, hi = self.bucket_range[b]
        return self.flat[lo:hi]

    # ------------------------------------------------------------------------------------------------ filling
    def pack(self, grads, bucket=None):
        """grads: dict id(param) -> gradient tensor.  Gradients the kernels already wrote into their view are left
        alone; every other one is copied by ONE table-driven kernel launch (cavp_copy_multi) when the buffer lives on a
        CUDA device.  Points p.grad at the views; parameters without a gradient keep p.grad = None and contribute
        zeros to the collective.  `bucket`: only the parameters of that bucket."""
        todo = []
        for i, p in enumerate(self.params):
            if bucket is not None and self.bucket_of[id(p)] != bucket:
                continue
            g = grads.get(id(p))
            v = self.views[i]
            if g is None:
                if self.used[i] and not self.prezeroed:
                    v.zero_()
                self.used[i] = False
                continue
            if g.data_ptr() != v.data_ptr():
                todo.append((i, g))
            p.grad = v
            self.used[i] = True
        if not todo:
            return
        if not self.flat.is_cuda:
            for i, g in todo:
                self.views[i].copy_(g)
            return
        keep, rows = [], []
        for i, g in todo:
            v = self.views[i]
            if g.shape != v.shape or g.stride() != v.stride() or g.dtype != torch.float32:
                g2 = torch.empty_like(self.params[i], memory_format=torch.preserve_format)
                g2.copy_(g)  # plumbing copy into the parameter's layout (rare: re-packed stems, padded classifier)
                g = g2
            keep.append(g)
            rows.append((g.data_ptr(), v.data_ptr(), g.numel()))
        chunk = _C.query("cavp_opt_chunk_elems")
        work = [(r, c) for r, (_, _, n) in enumerate(rows) for c in range((n + chunk - 1) // chunk)]
        table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(self.flat.device, non_blocking=True)
        workt = torch.tensor(work, dtype=torch.int32).reshape(-1, 2).pin_memory().to(self.flat.device, non_blocking=True)
        _C.call("cavp_copy_multi", table.data_ptr(), workt.data_ptr(), len(work),
                torch.cuda.current_stream(self.flat.device).cuda_stream)
        for t in keep + [table, workt]:
            t.record_stream(torch.cuda.current_stream(self.flat.device))

    # ------------------------------------------------------------------------------------------------ collectives
    def _reduce(self, t, group, average, async_op):
        world = dist.get_world_size(group)
        if average and dist.get_backend(group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op), None
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return w, (1.0 / world if average else None)

    def all_reduce_bucket(self, b, group=None, average=True):
        """Launch bucket b's all-reduce now, asynchronously: the collective waits for the kernels already queued on
        the current stream (the gradients of this bucket) and then runs beside whatever is queued next."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        t = self.bucket_tensor(b)
        if t.numel() == 0:
            return
        work, scale = self._reduce(t, group, average, async_op=True)
        self._pending.append((b, work, scale))

    def begin_step(self):
        """Start of a step: ONE memset of the whole buffer, so that split-K weight gradients (red.global.add into their
        view) need no per-layer zeroing and parameters without a gradient contribute zeros to the collective."""
        self._flushed = set()
        if self.flat.is_cuda:
            _C.call("cavp_zero", self.flat.data_ptr(), self.flat.numel() * 4,
                    torch.cuda.current_stream(self.flat.device).cuda_stream)
        else:
            self.flat.zero_()
        self.prezeroed = True

    def flush_bucket(self, b, grads, group=None, average=True):
        """Bucket b is complete (the backward tape passed its marker): pack what was not produced in place and start
        its all-reduce.  Idempotent within a step."""
        flushed = self.__dict__.setdefault("_flushed", set())
        if b in flushed:
            return
        flushed.add(b)
        self.pack(grads, bucket=b)
        self.all_reduce_bucket(b, group, average)

    def finish(self, grads, group=None, average=True):
        """End of the backward pass: flush the remaining buckets and make the current stream wait for all of them."""
        for b in range(len(self.bucket_range)):
            self.flush_bucket(b, grads, group, average)
        self.wait()

    def wait(self):
        """Make the current stream wait for every launched bucket (call before the optimisers read p.grad)."""
        for b, work, scale in self._pending:
            work.wait()
            if scale is not None:
                self.bucket_tensor(b).mul_(scale)
        done = {b for b, _, _ in self._pending}
        self._pending = []
        return done

    def all_reduce(self, group=None, average=True):
        """Reduce every bucket that has not been launched yet, then wait for all of them."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        launched = {b for b, _, _ in self._pending}
        for b in range(len(self.bucket_range)):
            if b not in launched:
                self.all_reduce_bucket(b, group, average)
        self.wait()


def cavp_buckets(model):
    """Bucket order = the order in which the backward pass completes them (cavp_b200.models.cavp_model.build_graph
    records the matching tape markers): 0 audio backbone, 1 decoder + fusion + projector + DeepLab head, 2 ResNet layer4,
    3 layer3, 4 the rest of the ResNet (the only bucket whose all-reduce cannot overlap the backward pass: 1.5 M
    parameters = 6 MB)."""
    m = model.module if hasattr(model, "module") else model
    audio = list(m.audio_backbone.parameters())
    backbone = list(m.backbone.parameters())
    taken = {id(p) for p in audio + backbone}
    head = [p for p in m.parameters() if id(p) not in taken]
    layer4 = list(m.backbone.backbone.layer4.parameters())
    layer3 = list(m.backbone.backbone.layer3.parameters())
    late = {id(p) for p in layer4 + layer3}
    rest = [p for p in backbone if id(p) not in late]
    return [audio, head, layer4, layer3, rest]


# bucket i is complete when the backward tape passes marker i (the last bucket: at the end of the backward pass)
BUCKET_MARKERS = ("audio_grads_done", "head_grads_done", "layer4_grads_done", "layer3_grads_done")


def shard_batch(global_batch, rank, world):
    """Contiguous, equal shards of the global batch (DistributedSampler(drop_last=True) semantics per step)."""
    per = global_batch // world
    return rank * per, (rank + 1) * per


class PeerReduce:
    """Small all-reduce (sum) over NVLink peer memory for the SyncBatchNorm statistics (csrc/peer.cu).

    torch.nn.SyncBatchNorm (main_vpo_mono.py:130) issues two small NCCL collectives per layer and step; a train step
    has ~120 of them, each on the critical path.  One single-CTA kernel per collective pushes the local vector into
    every peer's exchange buffer (mapped through CUDA IPC) and sums the received slots in rank order.  Every rank must
    issue the same sequence of calls (they do: the ranks run the same kernel graph)."""

    SLOT_BYTES = 64 * 1024  # [2C + 1] doubles up to C = 4095

    def __init__(self, group, device, bases, own, rank, world):
        self.group, self.device, self.rank, self.world = group, device, rank, world
        self.own = own
        self.bases = (ctypes.c_void_p * world)(*bases)
        self.bases_addr = ctypes.addressof(self.bases)
        self.seq = 0

    def args(self, tensor):
        """-> argument tuple of cavp_peer_allreduce (without the stream) for an in-place sum of `tensor`."""
        assert tensor.is_contiguous() and tensor.dtype in (torch.float32, torch.float64)
        assert tensor.numel() * tensor.element_size() <= self.SLOT_BYTES, "vector larger than an exchange slot"
        self.seq += 1
        s = self.seq & 0xFFFFFFFF
        return (tensor.data_ptr(), tensor.numel(), int(tensor.dtype == torch.float64), self.bases_addr, self.rank,
                self.world, self.SLOT_BYTES, s - (1 << 32) if s >= (1 << 31) else s)

    def all_reduce_(self, tensor):
        _C.call("cavp_peer_allreduce", *self.args(tensor), _C.stream())
        return tensor


_PEER_REDUCERS = {}


def peer_reduce_for(group, device):
    """The PeerReduce of (group, device), set up on first use (collective: every rank of `group` must call it), or None
    when the ranks cannot map each other's memory (other backend than NCCL, more than 8 ranks, several nodes, two
    ranks on one GPU, no peer access) or the path is not switched on - the caller then uses
    torch.distributed.all_reduce.  Opt-in with CAVP_SYNCBN_PEER=1: measured on 2 B200s (bench.py --sync-bn) the step
    through pinned host buffers gains 0.6 % (920-923 vs 915-917 images/s), while the device-resident loop scatters
    more from run to run (837 / 899 / 925 vs 913 / 918 images/s), so NCCL stays the default."""
    if os.environ.get("CAVP_SYNCBN_PEER", "0") != "1":
        return None
    device = torch.device(device)
    key = (id(group), device.index)
    if key in _PEER_REDUCERS:
        return _PEER_REDUCERS[key]
    red = None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == "nccl" and 1 < world <= 8 and device.type == "cuda":
        _C.lib()
        buf, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            rc = _C.query("cavp_peer_alloc", PeerReduce.SLOT_BYTES, ctypes.addressof(buf), ctypes.addressof(handle))
            info = [None] * world
            dist.all_gather_object(info, (socket.gethostname(), device.index, bytes(handle) if rc == 0 else None),
                                   group=group)
            same_node = len({h for h, _, _ in info}) == 1 and len({d for _, d, _ in info}) == world
            bases, ok = [], rc == 0 and same_node and all(h is not None for _, _, h in info)
            if ok:
                for r, (_, _, h) in enumerate(info):
                    if r == rank:
                        bases.append(buf.value)
                        continue
                    peer = ctypes.c_void_p()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                    if _C.query("cavp_peer_open", ctypes.addressof(hb), ctypes.addressof(peer)) != 0:
                        ok = False
                        break
                    bases.append(peer.value)
            oks = [None] * world
            dist.all_gather_object(oks, bool(ok), group=group)  # (also: every buffer is zeroed before the first use)
            if all(oks):
                red = PeerReduce(group, device, bases, buf.value, rank, world)
            else:
                if rank == 0:
                    print("[cavp_b200] SyncBatchNorm statistics: peer memory unavailable, using NCCL all-reduce",
                          file=sys.stderr, flush=True)
                for b in bases:
                    if b != buf.value:
                        _C.query("cavp_peer_close", b, 0)
                if rc == 0:
                    _C.query("cavp_peer_close", buf.value, 1)
    _PEER_REDUCERS[key] = red
    return red
