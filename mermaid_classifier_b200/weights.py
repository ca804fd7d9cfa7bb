"""Host-side weight handling: pyspacer checkpoint -> BN-folded packed blob for the C ABI.

Mirrors ``TorchExtractor.load_weights`` of pyspacer 0.14.0 (called at
``/root/reference/scripts/build_feature_bucket.py:405-408``): the checkpoint is
``torch.load(stream, map_location='cpu')['net']`` and every key carries the 7-character
DataParallel prefix ``module.``.

The packed layout is the one ``csrc/layers.h`` documents.  BatchNorm (eps 1e-3, inference
form) is folded to a per-channel ``scale``/``bias`` pair applied in each kernel's epilogue;
conv weights themselves are left untouched so bf16/tf32 rounding sees the original values.
"""

from __future__ import annotations

import hashlib
import io
from typing import Any

import numpy as np
import torch

BN_EPS = 1e-3

# (repeats, kernel, stride, expand, c_in, c_out)
B0_STAGES = (
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
)


def block_table() -> list[dict]:
    out = []
    for r, k, s, e, ci, co in B0_STAGES:
        for j in range(r):
            c_in = ci if j == 0 else co
            out.append(
                dict(index=len(out), k=k, stride=s if j == 0 else 1, expand=e, c_in=c_in, c_out=co,
                     c_mid=c_in * e, c_se=max(1, int(c_in * 0.25)))
            )
    return out


def load_checkpoint(stream_or_path: Any) -> dict:
    """``torch.load(...)['net']`` with the ``module.`` prefix stripped."""
    if isinstance(stream_or_path, (bytes, bytearray)):
        stream_or_path = io.BytesIO(stream_or_path)
    ckpt = torch.load(stream_or_path, map_location="cpu", weights_only=True)
    sd = ckpt["net"] if isinstance(ckpt, dict) and "net" in ckpt else ckpt
    return strip_prefix(sd)


def strip_prefix(sd: dict) -> dict:
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _fold(sd: dict, prefix: str) -> tuple[np.ndarray, np.ndarray]:
    g = sd[prefix + ".weight"].double()
    b = sd[prefix + ".bias"].double()
    m = sd[prefix + ".running_mean"].double()
    v = sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(v + BN_EPS)
    bias = b - m * scale
    return scale.float().numpy(), bias.float().numpy()


def pack_backbone(sd: dict) -> np.ndarray:
    """Flat fp32 array in the canonical order of ``csrc/layers.h``."""
    sd = strip_prefix(sd)
    parts: list[np.ndarray] = []

    def add(x):
        flat = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).reshape(-1)
        pad = (-flat.size) % 4  # every segment starts 16-byte aligned (csrc/layers.h)
        parts.append(np.concatenate([flat, np.zeros(pad, np.float32)]) if pad else flat)

    # stem: [co][ci][ky][kx] -> [ky][kx][ci][co]
    add(sd["_conv_stem.weight"].float().permute(2, 3, 1, 0).contiguous().numpy())
    s, b = _fold(sd, "_bn0")
    add(s), add(b)
    for cfg in block_table():
        p = f"_blocks.{cfg['index']}."
        if cfg["expand"] != 1:
            add(sd[p + "_expand_conv.weight"].float().reshape(cfg["c_mid"], cfg["c_in"]).numpy())
            s, b = _fold(sd, p + "_bn0")
            add(s), add(b)
        # depthwise: [c][1][ky][kx] -> [ky*k+kx][c]
        add(sd[p + "_depthwise_conv.weight"].float().reshape(cfg["c_mid"], cfg["k"] * cfg["k"]).t().contiguous().numpy())
        s, b = _fold(sd, p + "_bn1")
        add(s), add(b)
        add(sd[p + "_se_reduce.weight"].float().reshape(cfg["c_se"], cfg["c_mid"]).numpy())
        add(sd[p + "_se_reduce.bias"].float().numpy())
        # transposed to [c_se][c_mid] so the SE kernel reads it coalesced
        add(sd[p + "_se_expand.weight"].float().reshape(cfg["c_mid"], cfg["c_se"]).t().contiguous().numpy())
        add(sd[p + "_se_expand.bias"].float().numpy())
        add(sd[p + "_project_conv.weight"].float().reshape(cfg["c_out"], cfg["c_mid"]).numpy())
        s, b = _fold(sd, p + "_bn2")
        add(s), add(b)
    add(sd["_conv_head.weight"].float().reshape(1280, 320).numpy())
    s, b = _fold(sd, "_bn1")
    add(s), add(b)
    return np.concatenate(parts)


def sha256_of(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()
