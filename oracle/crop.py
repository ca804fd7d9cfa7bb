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
