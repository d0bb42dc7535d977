"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads without a GPU and exports every symbol that
include/cavp_b200.h declares; the Python mirror keeps the reference's module tree; host-side loss logic matches the
oracle's restatement of the reference."""
import ctypes
import os
import re
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cavp_b200 import _C
    protos = _C.parse_header()
    assert len(protos) >= 37
    text = open(os.path.join(ROOT, "include", "cavp_b200.h")).read()
    declared = set(re.findall(r"\bint\s+(cavp_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(protos)
    lib = ctypes.CDLL(_C.LIB_PATH) if os.path.exists(_C.LIB_PATH) else _C.lib()
    for name in declared:
        assert hasattr(lib, name), name


def test_argument_validation_without_gpu():
    """Launchers validate pointers / alignment before touching the device (status codes of include/cavp_b200.h)."""
    from cavp_b200 import _C
    lib = _C.lib()
    assert lib.cavp_igemm(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 4, 4, 1, 1, 1, 1, 1, 0, 1, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0.0, 1,
                          2, 0, 0) == -1  # CAVP_ERR_NULL
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert lib.cavp_igemm(p, p, p, 0, 0, 0, 0, 0, 1, 1, 1, 3, 3, 1, 1, 1, 1, 1, 0, 1, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0.0, 1,
                          2, 0, 0) == -2  # channels not a multiple of 4 -> CAVP_ERR_ALIGN
    assert lib.cavp_igemm(p, p, p, 0, 0, 0, 0, 0, 1, 1, 1, 4, 4, 1, 1, 1, 1, 1, 0, 1, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0.0, 1,
                          7, 0, 0) == -3  # bad precision mode -> CAVP_ERR_ARG


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cavp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_model_tree_matches_reference_schema_on_cpu():
    from cavp_b200.models.cavp_model import CAVP
    from oracle import schema
    for audio, in_plane, nc in (("vgg", 1, 22), ("vgg", 1, 71)):
        args = SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=[False, True, True],
                               audio_backbone=audio, num_classes=nc, batch_size=2, local_rank="cpu")
        m = CAVP(50, None, num_classes=nc, args=args, in_plane=in_plane)
        ref = schema.cavp_schema(nc, audio, in_plane).shapes
        sd = m.state_dict()
        assert sorted(sd) == sorted(ref)
        for k, shape in ref.items():
            assert tuple(sd[k].shape) == tuple(shape), k
        # conv weights live in OHWI storage but keep the OIHW logical shape / values
        w = m.backbone.backbone.layer1[0].conv2.weight
        assert w.shape == (64, 64, 3, 3) and w.permute(0, 2, 3, 1).is_contiguous()
        # group_weight-style walk (engine/utils.py:642-688): every parameter sits in a Conv2d / Linear / norm leaf
        seen = set()
        for mod in m.modules():
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear, torch.nn.modules.batchnorm._BatchNorm,
                                torch.nn.LayerNorm)):
                seen.update(id(p) for p in mod.parameters(recurse=False))
        rest = [n for n, p in m.named_parameters() if id(p) not in seen]
        assert rest == ["cross_att.pos_embed_v", "cross_att.pos_embed_a"]  # same two as in the reference


def test_cpu_tensors_are_rejected_loudly():
    from cavp_b200.loss import ContrastLoss, CrossEntropyLoss
    with pytest.raises(RuntimeError):
        CrossEntropyLoss()(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        ContrastLoss(0.1, 255, 4)(torch.zeros(1, 304, 4, 4), torch.zeros(1, 16, 16, dtype=torch.long),
                                  torch.zeros(1, 304, 4, 4), torch.zeros(1, 16, 16, dtype=torch.long))


def test_contrast_select_matches_oracle_and_rng_order():
    from cavp_b200.loss import contrast_select
    from oracle import cavp_oracle as O
    g = torch.Generator().manual_seed(3)
    gt = torch.zeros(3, 64, 64, dtype=torch.int64)
    gt[0, 8:56, 8:56] = 3; gt[1, 4:60, 10:50] = 5; gt[2, 10:50, 4:60] = 3; gt[:, :2, :2] = 255
    gs = gt.clone(); gs[1] = 0
    torch.manual_seed(11); a = contrast_select(gt, gs, (16, 16), max_views=32)
    torch.manual_seed(11); b = O.contrast_select(gt, gs, (16, 16), max_views=32)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert contrast_select(torch.zeros(2, 32, 32, dtype=torch.int64), gs[:2, :32, :32], (8, 8), 512) is None


def test_shuffled_labels_match_oracle():
    from cavp_b200.trainer import shuffled_labels
    from oracle import seeded
    b = seeded.synthetic_batch(6, 32, 32, 22, seed=5)
    assert torch.equal(shuffled_labels(b["pix_label"], b["img_label"], b["shuffle_idx"]),
                       seeded.shuffled_labels(b["pix_label"], b["img_label"], b["shuffle_idx"]))


def test_optimizer_work_list_and_cpu_refusal():
    """host side of the fused optimisers: the (tensor, chunk) work list covers every element exactly once, and a CPU
    parameter is refused instead of silently stepping on the host."""
    from cavp_b200 import _C
    from cavp_b200.optim import SGD, Adam, build_work
    chunk = _C.query("cavp_opt_chunk_elems")
    sizes = [1, chunk - 1, chunk, chunk + 1, 5 * chunk + 7]
    work = build_work(sizes, chunk)
    covered = [0] * len(sizes)
    for t, c in work.tolist():
        covered[t] += min(chunk, sizes[t] - c * chunk)
    assert covered == sizes
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    for opt in (SGD([p], lr=0.1, momentum=0.9), Adam([p], lr=0.1)):
        with pytest.raises(RuntimeError):
            opt.step()
    assert [g["lr"] for g in SGD([dict(params=[p], lr=0.5)], lr=0.1, momentum=0.9).param_groups] == [0.5]


def test_wgrad_split_heuristic_avoids_wave_spill():
    """host logic: the weight-gradient split count never lands just past a multiple of 148 CTAs (a 2.03-wave launch
    costs three waves), and the TMA path is chosen only where enough column tiles re-read dY."""
    from cavp_b200.engine import Graph, NUM_SMS
    for P, cout, K in [(200704, 256, 2736), (200704, 1216, 304), (200704, 304, 1216), (25088, 256, 18432),
                       (25088, 2048, 512), (401408, 128, 576), (200704, 304, 304), (100352, 64, 576), (64, 4096, 12288)]:
        s = Graph.wgrad_splits(P, cout, K)
        bn = 128 if K > 64 else 64
        tiles = ((cout + 127) // 128) * ((K + bn - 1) // bn)
        ctas = tiles * s
        waves = (ctas + NUM_SMS - 1) // NUM_SMS
        assert 1 <= s <= max(1, (P + 31) // 32 // 4)
        assert ctas > (waves - 1) * NUM_SMS + NUM_SMS // 3 or waves == 1, (P, cout, K, s, ctas)
    assert Graph.wgrad_via_tma(200704, 256, 2736) and Graph.wgrad_via_tma(25088, 2048, 512)
    assert not Graph.wgrad_via_tma(200704, 1216, 304) and not Graph.wgrad_via_tma(401408, 16, 4096)


def test_integration_md_ctypes_stub_matches_the_header():
    """INTEGRATION.md section 4 shows a reference-side ctypes stub for cavp_igemm; its argtypes must be the header's."""
    import ctypes
    import re
    from cavp_b200 import _C
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"lib\.cavp_igemm\.argtypes = (.+?)\s+#", text)
    assert m, "the stub moved: update this test"
    ns = {"P": ctypes.c_void_p, "I": ctypes.c_int, "F": ctypes.c_float, "LL": ctypes.c_longlong}
    stub = eval(m.group(1), {"__builtins__": {}}, ns)
    assert stub == _C.parse_header()["cavp_igemm"]


def test_gradient_buckets_partition_the_model_in_backward_completion_order():
    """cavp_b200.parallel.cavp_buckets: every trainable parameter in exactly one bucket; bucket order = the order in
    which the backward tape completes them (audio, head + fusion, layer4, layer3, rest of the ResNet); one marker per
    bucket except the last."""
    from cavp_b200.models.cavp_model import CAVP
    from cavp_b200.parallel import BUCKET_MARKERS, FlatGradBuffer, cavp_buckets
    args = SimpleNamespace(seg_model="DeepLabV3Plus", last_three_dilation_stride=[False, True, True],
                           audio_backbone="vgg", num_classes=22, batch_size=2, local_rank="cpu")
    m = CAVP(50, None, num_classes=22, args=args, in_plane=1)
    buckets = cavp_buckets(m)
    assert len(buckets) == len(BUCKET_MARKERS) + 1 == 5
    ids = [id(p) for b in buckets for p in b]
    assert len(ids) == len(set(ids)) == len(list(m.parameters()))
    assert {id(p) for p in buckets[0]} == {id(p) for p in m.audio_backbone.parameters()}
    assert {id(p) for p in buckets[2]} == {id(p) for p in m.backbone.backbone.layer4.parameters()}
    names = {id(p): n for n, p in m.named_parameters()}
    assert all(names[i].startswith("backbone.") for i in map(id, buckets[4]))
    assert not any(names[i].startswith(("backbone.", "audio_backbone.")) for i in map(id, buckets[1]))
    flat = FlatGradBuffer(buckets, "cpu")
    assert [lo for lo, _ in flat.bucket_range] == sorted(lo for lo, _ in flat.bucket_range)
    sizes = [hi - lo for lo, hi in flat.bucket_range]
    assert sizes[0] > 0.6 * sum(sizes) and sizes[4] < 0.02 * sum(sizes)  # audio = 63 % of the bytes, the tail 1.3 %


def test_forward_split_k_rule_for_long_k_256_column_layers():
    from cavp_b200.engine import Graph
    assert Graph.fwd_splits(25088, 256, 18432) == 2      # ASPP 3x3: 98 pair tiles of 256 columns -> 196 work items
    assert Graph.fwd_splits(25088, 2048, 2304) == 1      # 784 items: fills the 74 CTA pairs without a split
    assert Graph.fwd_splits(200704, 256, 2304) == 1
    assert Graph.fwd_splits(25088, 256, 2304) == 1       # K too short for a split to pay
    assert Graph.fwd_splits(64, 4096, 12288) > 1         # VGG fc: weight-bandwidth bound, deterministic slabs
