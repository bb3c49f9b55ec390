"""State-dict schema of the reference CNN and a seeded random-init generator.

The key names/shapes are the ones `MVLMModel(...).state_dict()` produces
(reference: src/mvlm/prediction/paulsenpredictor.py:251-273 ResidualBlock,
:276-299 HourGlassModule, :364-402 MVLMModel), so that real checkpoints
(`models_urls`, paulsenpredictor.py:15-26) and these synthetic ones are
interchangeable.  There is no network in this environment, so benchmarks and
tests use `seeded_state_dict`.
"""
from __future__ import annotations

import math

import torch

IMAGE_CHANNELS = {"geometry": 1, "RGB": 3, "depth": 1, "RGB+depth": 4, "geometry+depth": 2}
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by the reference


def _bn_keys(prefix: str, c: int):
    return [
        (f"{prefix}.weight", (c,), "bn_w"),
        (f"{prefix}.bias", (c,), "bn_b"),
        (f"{prefix}.running_mean", (c,), "bn_m"),
        (f"{prefix}.running_var", (c,), "bn_v"),
        (f"{prefix}.num_batches_tracked", (), "bn_n"),
    ]


def _rb_keys(prefix: str, cin: int, cout: int):
    ks = []
    ks += _bn_keys(f"{prefix}.bn1", cin)
    ks.append((f"{prefix}.conv1.weight", (cout // 2, cin, 3, 3), "conv_w"))
    ks += _bn_keys(f"{prefix}.bn2", cout // 2)
    ks.append((f"{prefix}.conv2.weight", (cout // 4, cout // 2, 3, 3), "conv_w"))
    ks += _bn_keys(f"{prefix}.bn3", cout // 4)
    ks.append((f"{prefix}.conv3.weight", (cout // 4, cout // 4, 3, 3), "conv_w"))
    if cin != cout:
        ks += _bn_keys(f"{prefix}.resample.0", cin)
        ks.append((f"{prefix}.resample.2.weight", (cout, cin, 1, 1), "conv_w"))
    return ks


def schema(n_landmarks: int, image_channels: str = "RGB+depth", n_features: int = 256):
    """[(key, shape, kind)] in state_dict order."""
    cin = IMAGE_CHANNELS[image_channels]
    f, L = n_features, n_landmarks
    ks = []

    def conv(name, co, ci):
        ks.append((f"{name}.weight", (co, ci, 3, 3), "conv_w"))
        ks.append((f"{name}.bias", (co,), "conv_b"))

    conv("conv1", f // 4, cin)
    ks.extend(_bn_keys("bn1", f // 4))
    ks.extend(_rb_keys("conv2", f // 4, f // 2))
    ks.extend(_rb_keys("conv3", f // 2, f // 2))
    ks.extend(_rb_keys("conv4", f // 2, f))
    for hg in ("hg1", "hg2"):
        for i in range(1, 21):
            ks.extend(_rb_keys(f"{hg}.rb{i}", f, f))
    conv("conv5", f, f)
    ks.extend(_bn_keys("bn2", f))
    conv("conv6", L, f)
    conv("conv7", f, L)
    conv("conv8", L, L)
    conv("conv9", f, f)
    ks.extend(_bn_keys("bn3", f))
    conv("conv10", L, f)
    conv("conv11", L, L)
    return ks


def seeded_state_dict(n_landmarks: int = 73, image_channels: str = "RGB+depth", seed: int = 1234,
                      randomize_bn: bool = True) -> dict:
    """Deterministic random-init weights (CPU fp32).

    Convs follow torch's default init bound 1/sqrt(fan_in); BatchNorm affine and
    running statistics are randomised (unlike torch's 1/0/0/1 defaults) so that
    BN folding is actually exercised by parity tests.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    last_fan_in = 1
    for key, shape, kind in schema(n_landmarks, image_channels):
        if kind == "conv_w":
            last_fan_in = shape[1] * shape[2] * shape[3]
            b = 1.0 / math.sqrt(last_fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "conv_b":
            b = 1.0 / math.sqrt(last_fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "bn_w":
            sd[key] = 0.5 + torch.rand(shape, generator=g) if randomize_bn else torch.ones(shape)
        elif kind == "bn_b":
            sd[key] = 0.1 * torch.randn(shape, generator=g) if randomize_bn else torch.zeros(shape)
        elif kind == "bn_m":
            sd[key] = 0.1 * torch.randn(shape, generator=g) if randomize_bn else torch.zeros(shape)
        elif kind == "bn_v":
            sd[key] = 0.5 + torch.rand(shape, generator=g) if randomize_bn else torch.ones(shape)
        elif kind == "bn_n":
            sd[key] = torch.tensor(0, dtype=torch.long)
    return sd


def fold_bn(sd: dict, prefix: str):
    """Eval-mode BatchNorm as y = x*scale + shift (fp32)."""
    w, b = sd[f"{prefix}.weight"].float(), sd[f"{prefix}.bias"].float()
    m, v = sd[f"{prefix}.running_mean"].float(), sd[f"{prefix}.running_var"].float()
    scale = w / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale
