"""Deterministic, construction-order-independent parameter and input fill.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under cavp_b200/ may import this.

The reference initialises weights from checkpoints we do not have (resnet50.pth, vgg.pth).
For parity work every implementation (the real reference imported in the build container,
the restated oracle, the CUDA product) is loaded with the SAME state, generated here from
a per-key seed so that it does not depend on module construction order or on any RNG the
constructors consume.  Values are non-trivial on purpose (BN gamma != 1, running stats != 0/1,
non-zero biases) so that a wrong epilogue shows up in the outputs.
"""
import zlib

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 1000003 * seed) % (2**31 - 1))
    return g


def seeded_tensor(key: str, shape, seed: int, kind: str) -> torch.Tensor:
    g = _gen(key, seed)
    shape = tuple(shape)
    if kind == "weight":  # conv / linear weight: He-style scale on fan_in keeps activations O(1)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5
    if kind == "gamma":  # BN / LN scale
        return torch.rand(shape, generator=g) + 0.5
    if kind == "beta":  # BN / LN shift, linear / conv bias
        return torch.randn(shape, generator=g) * 0.1
    if kind == "mean":
        return torch.randn(shape, generator=g) * 0.1
    if kind == "var":
        return torch.rand(shape, generator=g) + 0.5
    if kind == "embed":
        return torch.randn(shape, generator=g) * 0.02
    raise ValueError(kind)


def classify(key: str, tensor: torch.Tensor, norm_keys) -> str:
    """norm_keys: set of module prefixes that are BatchNorm/LayerNorm instances."""
    prefix, _, leaf = key.rpartition(".")
    if leaf == "num_batches_tracked":
        return "zero"
    if leaf == "running_mean":
        return "mean"
    if leaf == "running_var":
        return "var"
    if "pos_embed" in key:
        return "embed"
    if prefix in norm_keys:
        return "gamma" if leaf == "weight" else "beta"
    if leaf == "weight" and tensor.dim() >= 2:
        return "weight"
    return "beta"


def fill_module_(module: torch.nn.Module, seed: int = 0) -> None:
    """In-place seeded fill of every parameter and buffer of `module` (keys = state_dict keys)."""
    norm_types = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.LayerNorm, torch.nn.GroupNorm)
    norm_keys = {name for name, m in module.named_modules() if isinstance(m, norm_types)}
    sd = module.state_dict()
    with torch.no_grad():
        for key, t in sd.items():
            kind = classify(key, t, norm_keys)
            if kind == "zero":
                t.zero_()
            else:
                t.copy_(seeded_tensor(key, t.shape, seed, kind).to(t.dtype))


def seeded_state_dict(module: torch.nn.Module, seed: int = 0):
    fill_module_(module, seed)
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


# ----------------------------------------------------------------------------------------------
# synthetic inputs (seed 666 + rank, rectangles of one class per image, 8x8 corner of 255)
# ----------------------------------------------------------------------------------------------
def synthetic_batch(B: int, H: int, W: int, num_classes: int, seed: int = 666, audio_frames: int = 96,
                    in_plane: int = 1):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    image = torch.randn(B, 3, H, W, generator=g)
    audio_m = torch.randn(B, in_plane, audio_frames, 64, generator=g)
    pix_label = torch.zeros(B, H, W, dtype=torch.int64)
    img_label = torch.zeros(B, num_classes, dtype=torch.int64)
    cls = torch.randint(1, num_classes, (B,), generator=g)
    for b in range(B):
        # centred rectangle covering a bit more than half the image
        h0, h1 = H // 8, H - H // 8
        w0, w1 = W // 6, W - W // 6
        pix_label[b, h0:h1, w0:w1] = cls[b]
        pix_label[b, : max(H // 28, 1), : max(W // 28, 1)] = 255
        img_label[b, cls[b]] = 1
    img_label[:, 0] = 1
    shuffle_idx = torch.randperm(B, generator=g)
    # trainer_cavp_vpo_mono.py:152,166: the second half of the audio batch is the shuffled first half
    audio = torch.cat((audio_m, audio_m[shuffle_idx]), dim=0)
    return {
        "image": image,
        "audio": audio,
        "pix_label": pix_label,
        "img_label": img_label,
        "shuffle_idx": shuffle_idx,
    }


def shuffled_labels(pix_label, img_label, shuffle_idx):
    """trainer/trainer_cavp_vpo_mono.py:148-151,178-180 (epoch 0 branch, no overwrite)."""
    shuffle_img_label = img_label.clone()[shuffle_idx]
    shuffle_pix_label = pix_label.clone()[shuffle_idx]
    if_match = torch.all(torch.eq(img_label, shuffle_img_label), dim=1)
    shuffle_pix_label[~if_match] = 0
    shuffle_pix_label[if_match] = pix_label[if_match]
    return shuffle_pix_label
