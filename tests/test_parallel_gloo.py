"""N>1 host logic on CPU: world_size-2 gloo process group, flat gradient buffer, one averaged all-reduce."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cavp_b200.parallel import FlatGradBuffer, shard_batch
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 4, 3)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    lin = torch.nn.Linear(5, 3)
    unused = torch.nn.Parameter(torch.zeros(7))
    params = list(conv.parameters()) + list(lin.parameters()) + [unused]
    buf = FlatGradBuffer(params)
    grads = {id(p): torch.full_like(p, float(rank + 1)) for p in params[:-1]}
    grads[id(conv.weight)] = torch.arange(conv.weight.numel(), dtype=torch.float32).view_as(conv.weight) * (rank + 1)
    buf.pack(grads)
    buf.all_reduce()
    ok = unused.grad is None
    ok &= bool(torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)))
    ok &= bool(torch.allclose(conv.weight.grad, torch.arange(conv.weight.numel(), dtype=torch.float32).view_as(conv.weight) * 1.5))
    ok &= conv.weight.grad.is_contiguous(memory_format=torch.channels_last)
    ok &= shard_batch(64, rank, world) == (rank * 32, rank * 32 + 32)
    # bucketed buffer: per-bucket flush (pack + asynchronous all-reduce) in completion order, then finish()
    bb = FlatGradBuffer([[lin.weight, lin.bias], [conv.weight, conv.bias, unused]])
    assert bb.bucket_range[0][0] == 0 and bb.bucket_range[1][0] == bb.bucket_range[0][1]
    assert all(o % 4 == 0 for o in bb.offsets)
    bb.begin_step()
    # a gradient produced in place (what the weight-gradient kernels do through Graph.weight_grad_buffer)
    bb.view_of(lin.weight).fill_(float(rank + 1))
    g2 = {id(lin.weight): bb.view_of(lin.weight), id(lin.bias): torch.full_like(lin.bias, float(2 * rank)),
          id(conv.weight): grads[id(conv.weight)], id(conv.bias): torch.full_like(conv.bias, float(rank + 3))}
    bb.flush_bucket(0, g2)
    bb.flush_bucket(0, g2)  # idempotent
    bb.finish(g2)
    ok &= bool(torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)))
    ok &= bool(torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 1.0)))
    ok &= bool(torch.allclose(conv.bias.grad, torch.full_like(conv.bias, 3.5)))
    ok &= bool(torch.allclose(conv.weight.grad, torch.arange(conv.weight.numel(), dtype=torch.float32).view_as(conv.weight) * 1.5))
    ok &= lin.weight.grad.data_ptr() == bb.view_of(lin.weight).data_ptr()
    # next step: a parameter that gets no gradient any more contributes zeros
    bb.begin_step()
    g3 = dict(g2); del g3[id(conv.bias)]
    conv.bias.grad = None
    bb.view_of(lin.weight).fill_(float(rank + 1))
    bb.finish(g3)
    ok &= conv.bias.grad is None and float(bb.view_of(conv.bias).abs().max()) == 0.0
    ret[rank] = ok
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
