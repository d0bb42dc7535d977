"""SyncBatchNorm semantics (reference DDP mode, main_vpo_mono.py:130): with the BNs converted to nn.SyncBatchNorm and
two ranks each holding half of the batch, the BN layer must reproduce single-process statistics of the full batch.
Needs 2 GPUs: run with  torchrun --nproc-per-node 2 -m pytest tests/test_syncbn_2gpu.py -m gpu2  (skipped otherwise)."""
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu2


@pytest.mark.parametrize("peer", ["0", "1"], ids=["nccl", "peer_memory"])
@pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) != 2 or not torch.cuda.is_available(),
                    reason="needs torchrun with 2 GPUs")
def test_syncbn_conv_bn_matches_full_batch(peer, monkeypatch):
    monkeypatch.setenv("CAVP_SYNCBN_PEER", peer)  # statistics over NCCL / over the peer-memory kernel (csrc/peer.cu)
    import torch.distributed as dist
    from cavp_b200.engine import Graph, new_act
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    torch.manual_seed(0)
    conv = nn.Conv2d(32, 64, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(64)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    x = torch.randn(4, 32, 10, 10)
    dz = torch.randn(4, 64, 10, 10)
    # full-batch reference (fp64, CPU)
    xr = x.double().requires_grad_(True)
    convd, bnd = nn.Conv2d(32, 64, 3, padding=1, bias=False).double(), nn.BatchNorm2d(64).double()
    convd.weight.data.copy_(conv.weight.data); bnd.weight.data.copy_(bn.weight.data); bnd.bias.data.copy_(bn.bias.data)
    z = F.relu(bnd(convd(xr)))
    z.backward(dz.double())
    # two ranks, half the batch each, SyncBatchNorm
    sbn = nn.SyncBatchNorm.convert_sync_batchnorm(bn).cuda()
    convc = conv.cuda()
    convc.weight.data = convc.weight.data.contiguous(memory_format=torch.channels_last)
    g = Graph(torch.device("cuda", rank), prec=2, train=True, sync_bn_group=dist.group.WORLD)
    xs = x[2 * rank:2 * rank + 2].cuda().contiguous()
    xa = new_act(2, 10, 10, 32, g.device)
    g.call("cavp_nchw_to_nhwc", xs.data_ptr(), xa.ptr, 2, 32, 100, 32)
    za = g.conv_bn(xa, convc.weight, sbn, pad=1)
    got = za.nchw().double().cpu()
    ref = z[2 * rank:2 * rank + 2]
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5
    assert float((sbn.running_var.double().cpu() - bnd.running_var).abs().max()) < 1e-5
    d, _ = g.grad_target(za)
    ds = dz[2 * rank:2 * rank + 2].cuda().contiguous()
    g.call("cavp_nchw_to_nhwc", ds.data_ptr(), d.ptr, 2, 64, 100, 64)
    g.backward()
    gx = g.grad_of(xa).nchw().double().cpu()
    assert float((gx - xr.grad[2 * rank:2 * rank + 2]).abs().max() / xr.grad.abs().max()) < 2e-5
    # local parameter gradients sum (over ranks) to the full-batch gradient (DDP would then average them)
    gw = g.param_grads[id(convc.weight)].contiguous().double()
    dist.all_reduce(gw)
    assert float((gw.cpu() - convd.weight.grad).abs().max() / convd.weight.grad.abs().max()) < 2e-5
    gg = g.param_grads[id(sbn.weight)].double().clone()
    dist.all_reduce(gg)
    assert float((gg.cpu() - bnd.weight.grad).abs().max() / bnd.weight.grad.abs().max()) < 2e-5


@pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) != 2 or not torch.cuda.is_available(),
                    reason="needs torchrun with 2 GPUs")
def test_peer_memory_allreduce_matches_nccl_bit_for_bit(monkeypatch):
    """csrc/peer.cu: the single-kernel all-reduce over CUDA-IPC peer memory that carries the SyncBatchNorm statistics.
    With two ranks a + b has one rounding, so it must equal NCCL's sum bit for bit - for fp64 [2C+1] and fp32 [2C]
    vectors of every BatchNorm width of the model, back to back (the double-buffered slots and the sequence flags)."""
    import torch.distributed as dist
    from cavp_b200.parallel import peer_reduce_for
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    monkeypatch.setenv("CAVP_SYNCBN_PEER", "1")  # opt-in path
    red = peer_reduce_for(dist.group.WORLD, dev)
    assert red is not None, "two GPUs of one node must be able to map each other's memory"
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    pairs = []
    for it in range(300):
        C = (1, 3, 64, 256, 304, 1024, 2048, 4095)[it % 8]
        dt, n = ((torch.float64, 2 * C + 1) if it % 3 else (torch.float32, 2 * C))
        x = torch.randn(n, device=dev, dtype=dt, generator=gen)
        want = x.clone()
        dist.all_reduce(want)
        pairs.append((red.all_reduce_(x), want))
        if it % 50 == 7:  # let one rank run ahead / fall behind
            torch.cuda.synchronize()
            if rank == it % 2:
                torch.cuda._sleep(20_000_000)
    torch.cuda.synchronize()
    for got, want in pairs:
        assert torch.equal(got, want)
