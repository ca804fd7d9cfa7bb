"""Oracle: point-patch cropping and per-patch normalisation (TEST INFRASTRUCTURE).

Restates pyspacer 0.14.0 ``spacer/extract_features_utils.py::crop_patches`` /
``crop_simple`` and ``spacer/extractors/torch_extractors.py::transformation``
(third-party dependency of the reference, pinned at ``/root/reference/pyproject.toml:23-29``;
not present under ``/root/reference`` -> PARITY UNPINNED for this module, see
``oracle/__init__.py``).  Reference call sites that fix the contract:

* ``extractor(pil_image, rowcols)`` -- ``mermaid_classifier/pyspacer/annotation.py:241``
* patch = ``(224, 224, 3) uint8`` PIL image -- ``scripts/build_feature_bucket.py:470-473``
* ``torch.stack([transformer(p) for p in batch])`` -- ``scripts/build_feature_bucket.py:430``
* rowcols are the sorted unique (row, col) pairs -- ``scripts/build_feature_bucket.py:658-665``
"""

from __future__ import annotations

import numpy as np

CROP_SIZE = 224
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def reflect_index(t: np.ndarray | int, n: int):
    """NumPy ``mode='reflect'`` index map (mirror without repeating the edge
    sample).  One reflection is ``-t`` for ``t < 0`` and ``2(n-1) - t`` for
    ``t >= n``; images narrower than the 224 pad reflect repeatedly, i.e. the
    map is periodic with period ``2(n-1)`` (``n == 1`` maps everything to 0)."""
    t = np.asarray(t, dtype=np.int64)
    if n == 1:
        return np.zeros_like(t)
    period = 2 * (n - 1)
    t = np.mod(t, period)
    return np.where(t >= n, period - t, t)


def crop_patches_padded(im: np.ndarray, rowcols, crop_size: int = CROP_SIZE) -> list[np.ndarray]:
    """Literal restatement: pad the WHOLE image by ``crop_size`` with
    ``np.pad(..., mode='reflect')`` and slice ``[upper:upper+crop, left:left+crop]``
    with ``upper = int(row + pad - crop/2)``."""
    im = np.asarray(im)
    pad = crop_size
    padded = np.pad(im, ((pad, pad), (pad, pad), (0, 0)), mode="reflect")
    out = []
    for row, col in rowcols:
        upper = int((row + pad) - crop_size / 2)
        left = int((col + pad) - crop_size / 2)
        out.append(padded[upper : upper + crop_size, left : left + crop_size, :])
    return out


def patch_source_coords(row: int, col: int, H: int, W: int, crop_size: int = CROP_SIZE):
    """Source pixel coordinates of every patch pixel: patch[i, j] = im[ys[i], xs[j]]."""
    half = crop_size // 2
    ys = reflect_index(np.arange(crop_size) + (int(row) - half), H)
    xs = reflect_index(np.arange(crop_size) + (int(col) - half), W)
    return ys.astype(np.int64), xs.astype(np.int64)


def crop_patches(im: np.ndarray, rowcols, crop_size: int = CROP_SIZE) -> np.ndarray:
    """Gather form of :func:`crop_patches_padded` (no whole-image copy);
    returns ``(n, crop, crop, 3) uint8``.  ``tests/test_oracle_crop.py`` proves
    the two forms identical, including every border and corner."""
    im = np.asarray(im)
    H, W = im.shape[:2]
    out = np.empty((len(rowcols), crop_size, crop_size, im.shape[2]), dtype=im.dtype)
    for k, (row, col) in enumerate(rowcols):
        ys, xs = patch_source_coords(row, col, H, W, crop_size)
        out[k] = im[ys[:, None], xs[None, :], :]
    return out


def _fma32(a, b, c):
    """float32 fused multiply-add: the float64 product of two float32 values is exact, so one rounding remains."""
    return (np.asarray(a, np.float32).astype(np.float64) * np.asarray(b, np.float32).astype(np.float64)
            + np.asarray(c, np.float32).astype(np.float64)).astype(np.float32)


def bilinear_taps(in_size: int, out_size: int):
    """Taps of ``torch.nn.functional.interpolate(mode="bilinear", align_corners=False, antialias=False)`` along one axis:
    ``(i0, i1, l0, l1)`` with source index ``fma(scale, i + 0.5, -0.5)`` clamped at 0 (torch's CPU kernel contracts the
    multiply-add), ``l1`` its fraction, ``l0 = 1 - l1``."""
    scale = np.float32(in_size) / np.float32(out_size)
    i = np.arange(out_size, dtype=np.float32) + np.float32(0.5)
    real = np.maximum(_fma32(np.full_like(i, scale), i, np.full_like(i, -0.5)), np.float32(0))
    i0 = np.minimum(real.astype(np.int64), in_size - 1)
    l1 = np.clip((real - i0.astype(np.float32)).astype(np.float32), 0, 1).astype(np.float32)
    return i0, np.minimum(i0 + 1, in_size - 1), (np.float32(1) - l1).astype(np.float32), l1


def resize_patches_bilinear(patches_u8: np.ndarray, out_size: int = CROP_SIZE) -> np.ndarray:
    """``(n, P, P, C) uint8 -> (n, out, out, C) uint8``: torch's float bilinear resize, bit for bit (x pass then y pass, each
    ``fma(l0, a, l1 * b)``), rounded half-to-even.  Pinned to torch itself by ``tests/test_oracle_crop.py``."""
    x = np.asarray(patches_u8).astype(np.float32)
    P = x.shape[1]
    i0, i1, l0, l1 = bilinear_taps(P, out_size)
    def along_x(rows):
        a, b = rows[:, :, i0, :], rows[:, :, i1, :]
        w0 = np.broadcast_to(l0[None, None, :, None], a.shape)
        w1 = np.broadcast_to(l1[None, None, :, None], a.shape)
        return _fma32(w0, a, (w1 * b).astype(np.float32))
    t0, t1 = along_x(x[:, i0]), along_x(x[:, i1])
    w0 = np.broadcast_to(l0[None, :, None, None], t0.shape)
    w1 = np.broadcast_to(l1[None, :, None, None], t0.shape)
    v = _fma32(w0, t0, (w1 * t1).astype(np.float32))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def crop_resize_patches(im: np.ndarray, rowcols, crop_size: int, out_size: int = CROP_SIZE) -> np.ndarray:
    """The patch-size != 224 path: ``crop_size`` window (even sizes; centre offset ``crop_size // 2``) -> bilinear -> ``out_size``."""
    if crop_size % 2:
        raise ValueError("crop_size must be even (the window centre rule of crop_patches is only pinned for even sizes)")
    p = crop_patches(im, rowcols, crop_size)
    return p if crop_size == out_size else resize_patches_bilinear(p, out_size)


def normalize_patches(patches_u8: np.ndarray) -> np.ndarray:
    """torchvision ``ToTensor`` + ``Normalize(IMAGENET_MEAN, IMAGENET_STD)``:
    ``y[c,i,j] = (u8[i,j,c] / 255 - mean[c]) / std[c]`` in float32, HWC -> CHW.
    Returns ``(n, 3, crop, crop) float32``.  Mirrors torch's op order exactly
    (``.div(255)`` then ``.sub_(mean).div_(std)`` with fp32 mean/std)."""
    x = np.asarray(patches_u8).astype(np.float32) / np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)
    std = np.asarray(IMAGENET_STD, dtype=np.float32)
    x = (x - mean) / std
    return np.ascontiguousarray(np.moveaxis(x, -1, 1)).astype(np.float32)


class RowColumnInvalidError(ValueError):
    pass


class DataLimitError(ValueError):
    pass


def check_extract_inputs(H: int, W: int, rowcols, max_pixels: int = 10**8, max_points: int = 1000):
    """pyspacer ``task_utils.check_extract_inputs`` (called at
    ``mermaid_classifier/pyspacer/annotation.py:240``)."""
    if H * W > max_pixels:
        raise DataLimitError(f"image has {H * W} pixels, max {max_pixels}")
    if len(rowcols) > max_points:
        raise DataLimitError(f"{len(rowcols)} points, max {max_points}")
    for row, col in rowcols:
        if row < 0 or row > H - 1:
            raise RowColumnInvalidError(f"row {row} outside [0, {H - 1}]")
        if col < 0 or col > W - 1:
            raise RowColumnInvalidError(f"col {col} outside [0, {W - 1}]")
