"""``EfficientNetExtractor``: the pyspacer extractor surface, running on libmermaid_b200.

Drop-in for ``spacer.extractors.EfficientNetExtractor`` as the reference uses it:

* constructed with ``data_locations={"weights": DataLocation}`` (+ optional ``data_hashes``)
  -- ``/root/reference/mermaid_classifier/pyspacer/annotation.py:236-238``;
* ``extractor(pil_image, rowcols) -> (ImageFeatures, ExtractFeaturesReturnMsg)`` -- ``annotation.py:241``;
* ``patches_to_features(patch_list) -> (list[list[float]], loaded_remote)`` and the
  ``device=`` / ``batch_size=`` keywords of the reference's ``_DeviceCachingExtractor``
  -- ``/root/reference/scripts/build_feature_bucket.py:375-448``;
* ``load_datastream("weights")``, ``load_weights(stream)``, ``feature_dim``, ``CROP_SIZE``.

Unlike the reference, cropping happens on the GPU (the whole image is uploaded once and the
reflect-padded 224x224 gathers are fused into the stem convolution), the network is cached
for the life of the object, and nothing falls back to the CPU.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import time
from typing import Any, Sequence

import numpy as np

from . import _lib, weights as _weights
from .spacer_compat import (
    ExtractFeaturesReturnMsg,
    ImageFeatures,
    check_extract_inputs,
    image_features_from_array,
    storage_factory,
)


def _as_hwc_u8(im: Any) -> np.ndarray:
    arr = np.asarray(im)
    if arr.ndim == 2:
        arr = np.stack([arr] * 3, axis=-1)
    if arr.ndim != 3 or arr.shape[2] < 3:
        raise ValueError(f"image must be HxWx3; got {arr.shape}")
    if arr.shape[2] > 3:
        arr = arr[:, :, :3]
    if arr.dtype != np.uint8:
        raise ValueError(f"image must be uint8; got {arr.dtype}")
    return np.ascontiguousarray(arr)


class EfficientNetExtractor:
    """EfficientNet-B0 point-patch feature extractor on a B200."""

    CROP_SIZE = _lib.CROP_SIZE
    BATCH_SIZE = 10  # pyspacer's constant; kept for API parity, not used for scheduling
    feature_dim = _lib.FEATURE_DIM
    DATA_LOCATION_KEYS = ["weights"]

    def __init__(
        self,
        data_locations: dict[str, Any] | None = None,
        data_hashes: dict[str, str] | None = None,
        *,
        device: str | int = "cuda",
        batch_size: int | None = None,
        mode: str = "fp32",
        max_batch: int = 256,
        state_dict: dict | None = None,
        crop_size: int = _lib.CROP_SIZE,
    ):
        if state_dict is None and (not data_locations or "weights" not in data_locations):
            raise ValueError("data_locations must contain a 'weights' DataLocation")
        if mode not in _lib.MODES:
            raise ValueError(f"mode must be one of {sorted(_lib.MODES)}; got {mode!r}")
        self.data_locations = dict(data_locations or {})
        self.data_hashes = dict(data_hashes or {})
        self.mode = mode
        # window cropped around each point; != 224 (model.json config.patch_size, inference/export.py:77) adds the bilinear
        # resize to the network's 224 x 224 input (mc_crop_resize_patches)
        self.crop_size = int(crop_size)
        if self.crop_size < 2 or self.crop_size % 2:
            raise ValueError(f"crop_size must be a positive even number; got {crop_size!r}")
        self.max_batch = int(max_batch)
        self._batch_size = int(batch_size) if batch_size else None
        self._device_arg = device
        self._state_dict = state_dict
        self._handle: C.c_void_p | None = None
        self._loaded_remote = False
        self._device_index = 0

    # -- pyspacer plumbing ------------------------------------------------------------
    def load_datastream(self, key: str):
        """``(stream, loaded_remotely)`` for ``data_locations[key]`` with the optional sha256 check."""
        loc = self.data_locations[key]
        stream = storage_factory(loc.storage_type, getattr(loc, "bucket_name", None)).load(loc.key)
        want = self.data_hashes.get(key)
        if want:
            got = hashlib.sha256(stream.getbuffer()).hexdigest()
            if got != want:
                raise ValueError(f"hash mismatch for {key}: expected {want}, got {got}")
            stream.seek(0)
        return stream, loc.storage_type in ("s3", "url")

    @classmethod
    def load_weights(cls, stream) -> dict:
        """pyspacer builds an ``nn.Module`` here; this build returns the stripped state_dict
        that :func:`weights.pack_backbone` turns into the device blob."""
        return _weights.load_checkpoint(stream)

    # -- device handle ------------------------------------------------------------------
    def _resolve_device(self) -> int:
        torch = _lib.require_cuda()
        d = self._device_arg
        if isinstance(d, int):
            return d
        d = str(d)
        if d in ("cuda", "auto"):
            return torch.cuda.current_device()
        if d.startswith("cuda:"):
            return int(d.split(":", 1)[1])
        raise RuntimeError(f"--device {d!r} requested but this extractor only runs on CUDA (B200); no CPU fallback")

    def _ensure_handle(self):
        if self._handle is not None:
            return self._handle
        lib = _lib.load()
        self._device_index = self._resolve_device()
        if self._state_dict is not None:
            sd, self._loaded_remote = self._state_dict, False
        else:
            stream, self._loaded_remote = self.load_datastream("weights")
            sd = self.load_weights(stream)
        blob = _weights.pack_backbone(sd)
        if blob.size != lib.mc_backbone_param_count():
            raise ValueError(f"packed backbone has {blob.size} floats, library expects {lib.mc_backbone_param_count()}")
        h = C.c_void_p()
        _lib.check(lib.mc_extractor_create(blob.ctypes.data, blob.size, _lib.MODES[self.mode], self._device_index,
                                           self.max_batch, C.byref(h)))
        self._handle = h
        return h

    def close(self) -> None:
        if self._handle is not None:
            _lib.load().mc_extractor_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(_lib.load().mc_extractor_launches(self._handle)) if self._handle else 0

    # -- the reference-facing calls --------------------------------------------------------
    def __call__(self, im: Any, rowcols: Sequence[tuple[int, int]]):
        t0 = time.time()
        feats = self.extract_array(im, rowcols)
        loaded_remote, self._loaded_remote = self._loaded_remote, False
        return image_features_from_array(rowcols, feats), ExtractFeaturesReturnMsg(
            extractor_loaded_remotely=loaded_remote, runtime=time.time() - t0)

    def extract_array(self, im: Any, rowcols: Sequence[tuple[int, int]]) -> np.ndarray:
        """Host image + rowcols -> ``(n, 1280) float32`` through ``mc_extract_image_host``
        (H2D of the image and D2H of the features inside the call)."""
        h = self._ensure_handle()
        arr = _as_hwc_u8(im)
        rc = np.ascontiguousarray(np.asarray(rowcols, dtype=np.int32).reshape(-1, 2))
        out = np.empty((rc.shape[0], self.feature_dim), dtype=np.float32)
        torch = _lib.require_cuda()
        if self.crop_size != self.CROP_SIZE:
            check_extract_inputs(arr, [tuple(x) for x in rc.tolist()], "")
            with torch.cuda.device(self._device_index):
                dev = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
                pts = np.concatenate([np.zeros((rc.shape[0], 1), np.int32), rc], axis=1)
                return self.extract_device([dev], pts).cpu().numpy()
        with torch.cuda.device(self._device_index):
            _lib.check(_lib.load().mc_extract_image_host(
                h, arr.ctypes.data, arr.shape[0], arr.shape[1], arr.strides[0], rc.ctypes.data, rc.shape[0],
                out.ctypes.data, _lib.stream_ptr()))
        return out

    def extract_many(self, images: Sequence[Any], rowcols_list: Sequence[Sequence[tuple[int, int]]], head: Any = None,
                     out: np.ndarray | None = None, labels_out: np.ndarray | None = None, want_features: bool = True):
        """A list of host images + their rowcols through ``mc_extract_images_host``: the per-image loop of
        ``process_source`` (``scripts/build_feature_bucket.py:749-788``) as ONE call.  The library copies image
        ``i+1`` while it convolves image ``i`` and reads the features of image ``i-1`` back (pinned staging ring,
        copy streams); with ``head`` (a :class:`~mermaid_classifier_b200.inference.DeviceHead`) the labels of every
        point come back too (extract + classify, ``pyspacer/annotation.py:235-251``).

        ``images``: HxWx3 uint8 NumPy arrays or CPU torch tensors (pinned tensors are DMA'd without a staging copy).
        Returns ``(features, labels)``: ``(n, 1280) float32`` rows in image order then rowcol order (``None`` when
        ``want_features`` is false), ``(n,) int32`` class indices (``None`` without a head)."""
        h = self._ensure_handle()
        torch = _lib.require_cuda()
        if len(images) != len(rowcols_list):
            raise ValueError("images and rowcols_list differ in length")
        if self.crop_size != self.CROP_SIZE:
            # resize path: image by image through crop + bilinear resize + the pre-cropped-patch call
            parts = [self.extract_array(im.numpy() if hasattr(im, "data_ptr") else im, rc)
                     for im, rc in zip(images, rowcols_list) if len(rc)]
            feats = np.concatenate(parts) if parts else np.zeros((0, self.feature_dim), dtype=np.float32)
            labels = None
            if head is not None and feats.shape[0]:
                labels = head.scores_host(feats, want_proba=False, want_labels=True)[1].astype(np.int32)
            return (feats if want_features else None), labels
        keep = []   # keep the arrays alive for the duration of the call
        tab = (_lib.McImage * max(len(images), 1))()
        pts = []
        for i, (im, rcs) in enumerate(zip(images, rowcols_list)):
            if hasattr(im, "data_ptr"):   # torch CPU tensor
                if im.is_cuda or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or im.stride(2) != 1 or im.stride(1) != 3:
                    raise ValueError("torch images must be CPU uint8 HxWx3 tensors with packed pixels")
                tab[i] = _lib.McImage(im.data_ptr(), im.shape[0], im.shape[1], im.stride(0))
                keep.append(im)
            else:
                arr = _as_hwc_u8(im)
                tab[i] = _lib.McImage(arr.ctypes.data, arr.shape[0], arr.shape[1], arr.strides[0])
                keep.append(arr)
            rc = np.asarray(rcs, dtype=np.int32).reshape(-1, 2)
            pts.append(np.concatenate([np.full((rc.shape[0], 1), i, dtype=np.int32), rc], axis=1))
        pts = np.ascontiguousarray(np.concatenate(pts, axis=0)) if pts else np.zeros((0, 3), dtype=np.int32)
        n = pts.shape[0]
        feats = None
        if want_features:
            # default output: pinned host memory (torch's caching host allocator), so the read-back stays asynchronous
            feats = out if out is not None else (
                torch.empty((n, self.feature_dim), dtype=torch.float32, pin_memory=True).numpy() if n else
                np.empty((0, self.feature_dim), dtype=np.float32))
            if feats.shape != (n, self.feature_dim) or feats.dtype != np.float32 or not feats.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous (n, 1280) float32 array")
        labels = None
        if head is not None:
            labels = labels_out if labels_out is not None else (
                torch.empty((n,), dtype=torch.int32, pin_memory=True).numpy() if n else np.empty((0,), dtype=np.int32))
            if labels.shape != (n,) or labels.dtype != np.int32:
                raise ValueError("labels_out must be an (n,) int32 array")
        if not want_features and head is None:
            raise ValueError("nothing to compute: want_features is false and no head was given")
        if n == 0:
            return feats, labels
        if head is not None:
            head._set_exact(False)   # labels: the tensor-core Linear chain, as DeviceHead.scores_device
        with torch.cuda.device(self._device_index):
            _lib.check(_lib.load().mc_extract_images_host(
                h, head._h if head is not None else None, C.addressof(tab), len(images), pts.ctypes.data, n,
                feats.ctypes.data if feats is not None else None, labels.ctypes.data if labels is not None else None,
                _lib.stream_ptr()))
        del keep
        return feats, labels

    def pipe_stats(self) -> dict:
        """Bytes moved by the last :meth:`extract_many` call (``h2d``, ``d2h``) and its image ``groups``."""
        a, b, g = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.load().mc_extractor_pipe_stats(self._ensure_handle(), C.byref(a), C.byref(b), C.byref(g)))
        return {"h2d": a.value, "d2h": b.value, "groups": g.value}

    def patches_to_features(self, patch_list: Sequence[Any]):
        """Pre-cropped 224x224x3 patches -> ``(list[list[float]], loaded_remote)``."""
        feats = self.patches_to_array(patch_list)
        loaded_remote, self._loaded_remote = self._loaded_remote, False
        return feats.tolist(), loaded_remote

    def patches_to_array(self, patch_list: Sequence[Any]) -> np.ndarray:
        h = self._ensure_handle()
        torch = _lib.require_cuda()
        n = len(patch_list)
        if n == 0:
            return np.zeros((0, self.feature_dim), dtype=np.float32)
        stack = np.stack([_as_hwc_u8(p) for p in patch_list])
        if stack.shape[1:] != (224, 224, 3):
            raise ValueError(f"patches must be 224x224x3; got {stack.shape[1:]}")
        with torch.cuda.device(self._device_index):
            dev = torch.from_numpy(stack).cuda()
            out = torch.empty((n, self.feature_dim), dtype=torch.float32, device="cuda")
            _lib.check(_lib.load().mc_extract_patches(h, dev.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
            return out.cpu().numpy()

    # -- device-resident API (no host copies) -----------------------------------------------------
    def extract_device(self, images: Sequence[Any], points: np.ndarray, out=None):
        """``images``: list of CUDA uint8 HWC tensors; ``points``: ``(n, 3) int32`` rows of
        ``(image_index, row, col)``.  Returns a CUDA ``(n, 1280) float32`` tensor; asynchronous
        on the current stream."""
        h = self._ensure_handle()
        torch = _lib.require_cuda()
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, 3))
        tab = (_lib.McImage * len(images))()
        for i, t in enumerate(images):
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3 or t.stride(2) != 1 or t.stride(1) != 3:
                raise ValueError("images must be CUDA uint8 HxWx3 tensors with packed pixels")
            tab[i] = _lib.McImage(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0))
        with torch.cuda.device(self._device_index):
            if out is None:
                out = torch.empty((pts.shape[0], self.feature_dim), dtype=torch.float32, device="cuda")
            if self.crop_size != self.CROP_SIZE:   # crop + bilinear resize, then the pre-cropped-patch path, a sub-batch at a time
                for s in range(0, pts.shape[0], self.max_batch):
                    sub = pts[s:s + self.max_batch]
                    patches = crop_resize_patches_device(images, sub, self.crop_size)
                    _lib.check(_lib.load().mc_extract_patches(h, patches.data_ptr(), sub.shape[0], out[s:].data_ptr(),
                                                              _lib.stream_ptr()))
                return out
            _lib.check(_lib.load().mc_extract_points(h, C.addressof(tab), len(images), pts.ctypes.data, pts.shape[0],
                                                     out.data_ptr(), _lib.stream_ptr()))
        return out

    def set_tap(self, layer: int, out_tensor) -> None:
        h = self._ensure_handle()
        ptr = out_tensor.data_ptr() if out_tensor is not None else None
        cap = out_tensor.numel() if out_tensor is not None else 0
        _lib.check(_lib.load().mc_extractor_set_tap(h, layer, ptr, cap))


def crop_patches_device(images: Sequence[Any], points: np.ndarray):
    """Bit-exact reflect-padded crop on the device (``mc_crop_patches``): returns a CUDA
    ``(n, 224, 224, 3) uint8`` tensor."""
    torch = _lib.require_cuda()
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, 3))
    tab = (_lib.McImage * len(images))()
    for i, t in enumerate(images):
        tab[i] = _lib.McImage(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0))
    out = torch.empty((pts.shape[0], 224, 224, 3), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mc_crop_patches(C.addressof(tab), len(images), pts.ctypes.data, pts.shape[0], out.data_ptr(),
                                           _lib.stream_ptr()))
    return out


def crop_resize_patches_device(images: Sequence[Any], points: np.ndarray, crop_size: int):
    """``crop_size`` window around each point, bilinear-resized to 224 x 224 on the device (``mc_crop_resize_patches``):
    returns a CUDA ``(n, 224, 224, 3) uint8`` tensor; ``crop_size == 224`` is the plain crop."""
    torch = _lib.require_cuda()
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, 3))
    tab = (_lib.McImage * len(images))()
    for i, t in enumerate(images):
        tab[i] = _lib.McImage(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0))
    out = torch.empty((pts.shape[0], 224, 224, 3), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mc_crop_resize_patches(C.addressof(tab), len(images), pts.ctypes.data, pts.shape[0], int(crop_size),
                                                  out.data_ptr(), _lib.stream_ptr()))
    return out


def normalize_patches_device(patches):
    torch = _lib.require_cuda()
    n = patches.shape[0]
    out = torch.empty((n, 3, 224, 224), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mc_normalize_patches(patches.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    return out


def synth_image_device(seed: int, image_id: int, H: int, W: int):
    """Device-side synthetic image (same hash as ``synth.synth_image``)."""
    torch = _lib.require_cuda()
    out = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mc_synth_image(out.data_ptr(), H, W, W * 3, seed & 0xFFFFFFFF, image_id & 0xFFFFFFFF,
                                          _lib.stream_ptr()))
    return out
