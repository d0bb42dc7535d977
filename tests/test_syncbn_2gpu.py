"""SyncBatchNorm semantics (reference DDP mode, main_vpo_mono.py:130): with the BNs converted to nn.SyncBatchNorm and
two ranks each holding half of the batch, the BN layer must reproduce single-process statistics of the full batch.
Needs 2 GPUs: run with  torchrun --nproc-per-node 2 -m pytest tests/test_syncbn_2gpu.py -m gpu2  (skipped otherwise)."""
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu2


@pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) != 2 or not torch.cuda.is_available(),
                    reason="needs torchrun with 2 GPUs")
def test_syncbn_conv_bn_matches_full_batch():
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
