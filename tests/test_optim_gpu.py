"""Fused optimiser steps (csrc/optim.cu, cavp_b200/optim.py) against torch.optim.SGD / torch.optim.Adam - the
optimisers the reference constructs (main_vpo_mono.py:118-125) - over several steps, mixed layouts and group settings."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def make_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 32, 3, 3), (40000,), (17,), (304, 1216), (5, 3, 3, 3), (1,), (70001,)]
    ps = []
    for i, s in enumerate(shapes):
        t = torch.randn(*s, generator=g).cuda()
        if len(s) == 4 and s[1] % 4 == 0:
            t = t.contiguous(memory_format=torch.channels_last)  # conv weights live channels_last in the product
        ps.append(torch.nn.Parameter(t))
    return ps


def set_grads(ps, seed, skip=()):
    g = torch.Generator().manual_seed(seed)
    for i, p in enumerate(ps):
        if i in skip:
            p.grad = None
            continue
        gr = torch.randn(*p.shape, generator=g).cuda()
        if i == 3:
            gr = gr.t().contiguous().t()  # a gradient whose strides differ from the parameter's
        elif p.dim() == 4 and not p.is_contiguous():
            gr = gr.contiguous(memory_format=torch.channels_last)
        p.grad = gr


def test_sgd_momentum_weight_decay_groups_match_torch():
    from cavp_b200.optim import SGD
    pa, pb = make_params(0), make_params(0)

    def groups(ps):
        return [dict(params=ps[:3], lr=1e-2), dict(params=ps[3:5], lr=1e-1, weight_decay=0.0), dict(params=ps[5:])]
    ref = torch.optim.SGD(groups(pa), lr=3e-3, momentum=0.9, weight_decay=5e-4)
    ours = SGD(groups(pb), lr=3e-3, momentum=0.9, weight_decay=5e-4)
    for step in range(4):
        skip = (2,) if step == 1 else ()
        set_grads(pa, 10 + step, skip)
        set_grads(pb, 10 + step, skip)
        if step == 2:  # trainers rewrite the group learning rates every iteration (trainer_cavp_vpo_mono.py:73-83)
            for o in (ref, ours):
                o.param_groups[0]["lr"] = 5e-3
                o.param_groups[1]["lr"] = 5e-2
        ref.step()
        ours.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert a.stride() == b.stride()
        assert rel_err(b, a) < 1e-6
    for a, b in zip(pa, pb):
        if "momentum_buffer" in ref.state[a]:
            assert rel_err(ours.state[b]["momentum_buffer"], ref.state[a]["momentum_buffer"]) < 1e-6
    assert set(ours.state_dict()["state"][0]) >= {"momentum_buffer"}


def test_adam_matches_torch():
    from cavp_b200.optim import Adam
    pa, pb = make_params(1), make_params(1)
    ref = torch.optim.Adam(pa, lr=1e-3)
    ours = Adam(pb, lr=1e-3)
    for step in range(5):
        set_grads(pa, 20 + step)
        set_grads(pb, 20 + step)
        ref.step()
        ours.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert rel_err(b, a) < 2e-6
        assert rel_err(ours.state[b]["exp_avg"], ref.state[a]["exp_avg"]) < 2e-6
        assert rel_err(ours.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"]) < 2e-6


def test_resume_from_torch_optim_checkpoint():
    """engine/engine.py:93-94 saves optimizer_v / optimizer_a state_dicts of torch.optim.SGD / Adam: per-parameter
    tensor `step`, NCHW-contiguous moment buffers (our conv parameters are channels_last).  Loading such a checkpoint
    into the fused optimisers and continuing must track torch step for step."""
    from cavp_b200.optim import SGD, Adam
    for make_ref, make_ours, keys, tol in (
            (lambda ps: torch.optim.SGD(ps, lr=3e-3, momentum=0.9, weight_decay=5e-4),
             lambda ps: SGD(ps, lr=3e-3, momentum=0.9, weight_decay=5e-4), ("momentum_buffer",), 1e-6),
            (lambda ps: torch.optim.Adam(ps, lr=1e-3), lambda ps: Adam(ps, lr=1e-3), ("exp_avg", "exp_avg_sq"), 2e-6)):
        # the checkpoint is written by torch.optim on NCHW-contiguous parameters (the reference's layout)
        pa = [torch.nn.Parameter(p.detach().contiguous()) for p in make_params(2)]
        pb = make_params(2)
        ref = make_ref(pa)
        for step in range(2):
            set_grads(pa, 30 + step)
            for p in pa:
                p.grad = p.grad.contiguous()
            ref.step()
        ckpt = torch.load(__import__("io").BytesIO(_dumps(ref.state_dict())), weights_only=False)
        with torch.no_grad():
            for a, b in zip(pa, pb):
                b.copy_(a)
        ours = make_ours(pb)
        ours.load_state_dict(ckpt)
        for step in range(2, 5):
            set_grads(pa, 30 + step)
            for p in pa:
                p.grad = p.grad.contiguous()
            set_grads(pb, 30 + step)
            ref.step()
            ours.step()
        torch.cuda.synchronize()
        for a, b in zip(pa, pb):
            assert rel_err(b, a) < tol
            for k in keys:
                assert rel_err(ours.state[b][k], ref.state[a][k]) < tol
            assert ours.state[b][k].stride() == b.stride()
            assert ours.state[b]["step"] == 5 if "step" in ref.state[a] else True


def _dumps(obj):
    import io
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()
