"""Stand-ins for the pyspacer types that cross the drop-in boundary.

The reference passes pyspacer (``spacer``) data classes in and out of the hot path:
``DataLocation``, ``ImageFeatures`` / ``PointFeatures``, ``ExtractFeaturesMsg`` /
``ExtractFeaturesReturnMsg``, ``storage_factory`` / ``load_image`` and
``check_extract_inputs`` (every call site is listed in SURVEY.md §2.2).  When ``spacer``
is importable its own classes are re-exported, so objects produced here are *the*
pyspacer objects; otherwise the minimal mirrors below are used.  pyspacer 0.14.0 is not
available offline, so the mirrors follow its published behaviour as recalled in SURVEY.md
§2.2 (UPSTREAM-RECALLED) -- in particular the ``.featurevector`` payload:
``np.savez_compressed(meta=[valid_rowcol, feature_dim, npoints], rows, cols, feat)``
with a legacy-JSON reader tried first.
"""

from __future__ import annotations

import io
import json
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from ._lib import DataLimitError, RowColumnInvalidError

try:  # pragma: no cover - exercised only where pyspacer is installed
    from spacer.data_classes import DataLocation, ImageFeatures, PointFeatures  # type: ignore
    from spacer.messages import ExtractFeaturesMsg, ExtractFeaturesReturnMsg  # type: ignore
    from spacer.storage import load_image, storage_factory  # type: ignore

    HAVE_SPACER = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_SPACER = False

MAX_IMAGE_PIXELS = 10**8
MAX_POINTS_PER_IMAGE = 1000

if not HAVE_SPACER:

    @dataclass(frozen=True)
    class DataLocation:
        storage_type: str
        key: str
        bucket_name: str | None = None

        def __post_init__(self):
            if self.storage_type not in ("s3", "filesystem", "memory", "url"):
                raise ValueError(f"unknown storage_type {self.storage_type!r}")

        @property
        def filename(self) -> str:
            return Path(self.key).name

    class _FileSystemStorage:
        def load(self, key: str) -> io.BytesIO:
            return io.BytesIO(Path(key).read_bytes())

        def store(self, key: str, stream: io.BytesIO) -> None:
            path = Path(key)
            path.parent.mkdir(parents=True, exist_ok=True)
            tmp = path.with_name(path.name + ".part")
            tmp.write_bytes(stream.getvalue())
            tmp.replace(path)

        def exists(self, key: str) -> bool:
            return Path(key).exists()

        def delete(self, key: str) -> None:
            Path(key).unlink()

    class _MemoryStorage:
        _blobs: dict[str, bytes] = {}

        def load(self, key: str) -> io.BytesIO:
            return io.BytesIO(self._blobs[key])

        def store(self, key: str, stream: io.BytesIO) -> None:
            self._blobs[key] = stream.getvalue()

        def exists(self, key: str) -> bool:
            return key in self._blobs

        def delete(self, key: str) -> None:
            del self._blobs[key]

    class _S3Storage:
        """``DataLocation('s3', key, bucket_name)`` -- the bucket layout of ``scripts/build_feature_bucket.py:530-544`` (images
        ``s{sid}/images/{iid}.jpg`` in, ``s{sid}/features/i{iid}.featurevector`` out).  boto3 is imported on first use; a
        client can be injected (``_S3Storage.client_factory``) for tests and for callers that configure their own session."""

        client_factory = None   # callable() -> boto3-style client

        def __init__(self, bucket_name: str | None):
            if not bucket_name:
                raise ValueError("s3 storage needs a bucket_name")
            self.bucket = bucket_name
            self._client = None

        @property
        def client(self):
            if self._client is None:
                if type(self).client_factory is not None:
                    self._client = type(self).client_factory()
                else:
                    try:
                        import boto3
                    except ImportError as e:   # pragma: no cover - boto3 is not in the build image
                        raise RuntimeError("storage_type 's3' needs boto3 (or an injected _S3Storage.client_factory)") from e
                    self._client = boto3.client("s3")
            return self._client

        def load(self, key: str) -> io.BytesIO:
            return io.BytesIO(self.client.get_object(Bucket=self.bucket, Key=key)["Body"].read())

        def store(self, key: str, stream: io.BytesIO) -> None:
            self.client.put_object(Bucket=self.bucket, Key=key, Body=stream.getvalue())

        def exists(self, key: str) -> bool:
            try:
                self.client.head_object(Bucket=self.bucket, Key=key)
                return True
            except Exception as e:   # botocore ClientError 404 / NoSuchKey; anything else is a real failure
                code = str(getattr(e, "response", {}).get("Error", {}).get("Code", ""))
                if code in ("404", "NoSuchKey", "NotFound") or isinstance(e, KeyError):
                    return False
                raise

        def delete(self, key: str) -> None:
            self.client.delete_object(Bucket=self.bucket, Key=key)

    class _URLStorage:
        """Read-only ``DataLocation('url', key)``."""

        def load(self, key: str) -> io.BytesIO:
            from urllib.request import urlopen

            with urlopen(key) as r:   # noqa: S310 - the caller names the URL
                return io.BytesIO(r.read())

        def store(self, key: str, stream: io.BytesIO) -> None:
            raise TypeError("url storage is read-only")

        def exists(self, key: str) -> bool:
            try:
                self.load(key)
                return True
            except Exception:
                return False

        def delete(self, key: str) -> None:
            raise TypeError("url storage is read-only")

    def storage_factory(storage_type: str, bucket_name: str | None = None):
        if storage_type == "filesystem":
            return _FileSystemStorage()
        if storage_type == "memory":
            return _MemoryStorage()
        if storage_type == "s3":
            return _S3Storage(bucket_name)
        if storage_type == "url":
            return _URLStorage()
        raise ValueError(f"unknown storage_type {storage_type!r}")

    def load_image(loc: "DataLocation"):
        from PIL import Image

        stream = storage_factory(loc.storage_type, loc.bucket_name).load(loc.key)
        img = Image.open(stream)
        img.load()
        return img.convert("RGB")

    @dataclass
    class PointFeatures:
        row: int | None
        col: int | None
        data: Any  # list[float] / 1-D array

    class ImageFeatures:
        def __init__(self, point_features, valid_rowcol: bool, feature_dim: int, npoints: int):
            self.point_features = list(point_features)
            self.valid_rowcol = bool(valid_rowcol)
            self.feature_dim = int(feature_dim)
            self.npoints = int(npoints)
            self._rchash = (
                {(pf.row, pf.col): i for i, pf in enumerate(self.point_features)} if self.valid_rowcol else {}
            )

        def __getitem__(self, rowcol):
            if not self.valid_rowcol:
                raise ValueError("Method requires valid rows and columns")
            return self.point_features[self._rchash[tuple(rowcol)]].data

        def get_array(self, rowcol) -> np.ndarray:
            return np.asarray(self[rowcol], dtype=np.float32)

        # -- (de)serialisation ------------------------------------------------------
        def serialize(self) -> io.BytesIO:
            rows = np.asarray([pf.row for pf in self.point_features], dtype=np.int64)
            cols = np.asarray([pf.col for pf in self.point_features], dtype=np.int64)
            feat = np.asarray([pf.data for pf in self.point_features], dtype=np.float32).reshape(
                self.npoints, self.feature_dim
            )
            meta = np.asarray([int(self.valid_rowcol), self.feature_dim, self.npoints], dtype=np.int64)
            out = io.BytesIO()
            np.savez_compressed(out, meta=meta, rows=rows, cols=cols, feat=feat)
            out.seek(0)
            return out

        def store(self, loc: "DataLocation") -> None:
            storage_factory(loc.storage_type, loc.bucket_name).store(loc.key, self.serialize())

        @classmethod
        def deserialize(cls, data: Any) -> "ImageFeatures":
            if isinstance(data, list):  # oldest legacy layout: bare list of vectors
                return cls([PointFeatures(None, None, d) for d in data], False, len(data[0]), len(data))
            return cls(
                [PointFeatures(p["row"], p["col"], p["data"]) for p in data["point_features"]],
                data["valid_rowcol"], data["feature_dim"], data["npoints"],
            )

        @classmethod
        def load_from_stream(cls, stream: io.BytesIO) -> "ImageFeatures":
            stream.seek(0)
            try:
                return cls.deserialize(json.load(stream))
            except (UnicodeDecodeError, json.JSONDecodeError, ValueError):
                stream.seek(0)
            z = np.load(stream, allow_pickle=False)
            valid, dim, npts = (int(v) for v in z["meta"])
            feat = z["feat"]
            pfs = [PointFeatures(int(r), int(c), feat[i]) for i, (r, c) in enumerate(zip(z["rows"], z["cols"]))]
            return cls(pfs, bool(valid), dim, npts)

        @classmethod
        def load(cls, loc: "DataLocation") -> "ImageFeatures":
            return cls.load_from_stream(storage_factory(loc.storage_type, loc.bucket_name).load(loc.key))

    @dataclass
    class ExtractFeaturesMsg:
        job_token: str
        extractor: Any
        rowcols: list
        image_loc: Any
        feature_loc: Any

    @dataclass
    class ExtractFeaturesReturnMsg:
        extractor_loaded_remotely: bool
        runtime: float


def image_features_from_array(rowcols, feats: np.ndarray):
    """Build an ``ImageFeatures`` from an ``(n, D) float32`` matrix without per-element
    Python floats (the reference's ``.tolist()`` at build_feature_bucket.py:437)."""
    feats = np.asarray(feats, dtype=np.float32)
    pfs = [PointFeatures(int(r), int(c), feats[i]) for i, (r, c) in enumerate(rowcols)]
    return ImageFeatures(pfs, True, int(feats.shape[1]) if feats.ndim == 2 else 0, len(pfs))


def check_extract_inputs(image, rowcols, image_key: str = "") -> None:
    """pyspacer ``task_utils.check_extract_inputs`` (call site annotation.py:240)."""
    width, height = image.size if hasattr(image, "size") and not isinstance(image, np.ndarray) else (
        image.shape[1], image.shape[0])
    if width * height > MAX_IMAGE_PIXELS:
        raise DataLimitError(
            f"Image {image_key} has {width} x {height} = {width * height} total pixels, which is larger"
            f" than the max allowed of {MAX_IMAGE_PIXELS}.")
    if len(rowcols) > MAX_POINTS_PER_IMAGE:
        raise DataLimitError(
            f"{len(rowcols)} point locations were specified for image {image_key}, and that's larger"
            f" than the max allowed of {MAX_POINTS_PER_IMAGE}.")
    for row, col in rowcols:
        if row < 0 or row > height - 1:
            raise RowColumnInvalidError(f"{image_key}: Row value {row} falls outside this image's valid range of 0-{height - 1}.")
        if col < 0 or col > width - 1:
            raise RowColumnInvalidError(f"{image_key}: Column value {col} falls outside this image's valid range of 0-{width - 1}.")


__all__ = [
    "HAVE_SPACER", "DataLocation", "ImageFeatures", "PointFeatures", "ExtractFeaturesMsg",
    "ExtractFeaturesReturnMsg", "storage_factory", "load_image", "check_extract_inputs",
    "image_features_from_array", "RowColumnInvalidError", "DataLimitError",
]
