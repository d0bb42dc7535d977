"""Data-parallel plumbing: one process per GPU, the batch sharded across ranks, ONE gradient all-reduce per step.

The reference wraps the model in DDP (main_vpo_mono.py:131-141: bucketed all-reduce from autograd hooks).  Here the
gradients of a step are packed into a single flat fp32 buffer (parameter order = model.parameters()) and reduced with
one NCCL all-reduce (average, DDP semantics) over NVLink/NVSwitch; `p.grad` are views into the flat buffer, so the
optimisers read the reduced values in place.  Parameters that never receive gradients (cross_att.pos_embed_*,
audio_backbone.cls_head.* - the reason the reference needs find_unused_parameters=True) are left out of the buffer.
"""
import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device if device is not None else self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.used = [False] * len(self.params)

    def view_for(self, i):
        p = self.params[i]
        v = self.flat[self.offsets[i]:self.offsets[i] + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            n, c, h, w = p.shape
            return v.view(n, h, w, c).permute(0, 3, 1, 2)  # same memory format as the parameter
        return v.view(p.shape)

    def pack(self, grads):
        """grads: dict id(param) -> gradient tensor (any strides).  Copies into the flat buffer and points p.grad at
        the views.  Parameters without a gradient keep p.grad = None (and contribute zeros to the collective)."""
        for i, p in enumerate(self.params):
            g = grads.get(id(p))
            if g is None:
                if self.used[i]:
                    self.view_for(i).zero_()
                continue
            v = self.view_for(i)
            v.copy_(g)
            p.grad = v
            self.used[i] = True

    def all_reduce(self, group=None, average=True):
        """ONE collective for the whole step.  NCCL: AVG in-switch when available; gloo (CPU tests): SUM then scale."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(world)


def shard_batch(global_batch, rank, world):
    """Contiguous, equal shards of the global batch (DistributedSampler(drop_last=True) semantics per step)."""
    per = global_batch // world
    return rank * per, (rank + 1) * per
