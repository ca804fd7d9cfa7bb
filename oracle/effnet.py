"""Oracle: EfficientNet-B0 ``extract_features`` in fp32 on the CPU (TEST INFRASTRUCTURE).

Restates the network pyspacer 0.14.0 vendors in ``spacer/models/effcientnet.py``
(an early lukemelas EfficientNet-PyTorch) and that the reference drives through
``net.extract_features(batch_t)`` (``/root/reference/scripts/build_feature_bucket.py:433-434``)
after ``EfficientNetExtractor.load_weights(stream)`` (``:405-408``).  pyspacer is a
third-party dependency absent from ``/root/reference`` -> PARITY UNPINNED (see
``oracle/__init__.py``); ``tests/test_oracle_effnet.py`` cross-checks this file against
an independent torchvision construction of the same topology.

Published algorithm restated here (functional form over a pyspacer-layout state_dict):

* every conv uses TF-"SAME" *dynamic, asymmetric* padding: total
  ``max((ceil(i/s)-1)*s + k - i, 0)``, ``before = total // 2``, ``after = total - before``;
* BatchNorm in inference form, ``eps = 1e-3``;
* swish ``x * sigmoid(x)``;
* MBConv: [expand 1x1 + BN + swish, absent when expand_ratio == 1] -> depthwise kxk + BN +
  swish -> SE (global mean -> 1x1+bias -> swish -> 1x1+bias -> sigmoid -> scale) ->
  project 1x1 + BN -> (+ input when stride == 1 and C_in == C_out);
* ``extract_features`` = stem -> 16 MBConv -> swish(bn1(conv_head)) -> global mean -> (N, 1280).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
FEATURE_DIM = 1280
NUM_CLASSES_FC = 1275  # pyspacer's `_fc` head (unused by extract_features)

# (repeats, kernel, stride, expand_ratio, in, out) -- the B0 stage table.
B0_STAGES = (
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
)
SE_RATIO = 0.25


@dataclass(frozen=True)
class BlockCfg:
    index: int
    kernel: int
    stride: int
    expand: int
    c_in: int
    c_out: int

    @property
    def c_mid(self) -> int:
        return self.c_in * self.expand

    @property
    def c_se(self) -> int:
        return max(1, int(self.c_in * SE_RATIO))

    @property
    def has_skip(self) -> bool:
        return self.stride == 1 and self.c_in == self.c_out


def b0_blocks() -> list[BlockCfg]:
    blocks: list[BlockCfg] = []
    for r, k, s, e, ci, co in B0_STAGES:
        for j in range(r):
            blocks.append(BlockCfg(len(blocks), k, s if j == 0 else 1, e, ci if j == 0 else co, co))
    return blocks


def same_pad(i: int, k: int, s: int) -> tuple[int, int]:
    total = max((math.ceil(i / s) - 1) * s + k - i, 0)
    return total // 2, total - total // 2


def conv_same(x: torch.Tensor, w: torch.Tensor, bias=None, stride: int = 1, groups: int = 1):
    k = w.shape[-1]
    pt, pb = same_pad(x.shape[-2], k, stride)
    pl, pr = same_pad(x.shape[-1], k, stride)
    if pt or pb or pl or pr:
        x = F.pad(x, [pl, pr, pt, pb])
    return F.conv2d(x, w, bias, stride=stride, padding=0, groups=groups)


def bn(x: torch.Tensor, sd: dict, prefix: str):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return F.batch_norm(x, m, v, w, b, training=False, eps=BN_EPS)


def swish(x: torch.Tensor):
    return x * torch.sigmoid(x)


def strip_module_prefix(sd: dict) -> dict:
    """pyspacer ``load_weights``: checkpoint is ``{'net': state_dict}`` saved from
    ``nn.DataParallel`` -> every key carries a 7-char ``module.`` prefix."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def mbconv(x: torch.Tensor, sd: dict, cfg: BlockCfg, taps: dict | None = None):
    p = f"_blocks.{cfg.index}."
    inp = x
    if cfg.expand != 1:
        x = swish(bn(conv_same(x, sd[p + "_expand_conv.weight"]), sd, p + "_bn0"))
        if taps is not None:
            taps[f"b{cfg.index}.expand"] = x
    x = swish(bn(conv_same(x, sd[p + "_depthwise_conv.weight"], stride=cfg.stride, groups=cfg.c_mid), sd, p + "_bn1"))
    if taps is not None:
        taps[f"b{cfg.index}.dw"] = x
    s = F.adaptive_avg_pool2d(x, 1)
    s = swish(conv_same(s, sd[p + "_se_reduce.weight"], sd[p + "_se_reduce.bias"]))
    s = conv_same(s, sd[p + "_se_expand.weight"], sd[p + "_se_expand.bias"])
    if taps is not None:
        taps[f"b{cfg.index}.gate"] = torch.sigmoid(s)
    x = torch.sigmoid(s) * x
    x = bn(conv_same(x, sd[p + "_project_conv.weight"]), sd, p + "_bn2")
    if cfg.has_skip:
        x = x + inp
    if taps is not None:
        taps[f"b{cfg.index}.out"] = x
    return x


@torch.no_grad()
def extract_features(sd: dict, x: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """``(N, 3, 224, 224) float32 -> (N, 1280) float32``.  ``taps`` (optional dict)
    receives per-layer NCHW activations for layer-by-layer kernel debugging."""
    sd = strip_module_prefix(sd)
    x = swish(bn(conv_same(x, sd["_conv_stem.weight"], stride=2), sd, "_bn0"))
    if taps is not None:
        taps["stem"] = x
    for cfg in b0_blocks():
        x = mbconv(x, sd, cfg, taps)
    x = swish(bn(conv_same(x, sd["_conv_head.weight"]), sd, "_bn1"))
    if taps is not None:
        taps["head"] = x
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


@torch.no_grad()
def extract_features_batched(sd: dict, x: torch.Tensor, batch_size: int = 10) -> torch.Tensor:
    """pyspacer ``TorchExtractor.patches_to_features`` batching (``BATCH_SIZE = 10``;
    reference override at ``scripts/build_feature_bucket.py:166-172,428-437``)."""
    outs = [extract_features(sd, x[i : i + batch_size]) for i in range(0, x.shape[0], batch_size)]
    return torch.cat(outs, 0) if outs else x.new_zeros((0, FEATURE_DIM))
