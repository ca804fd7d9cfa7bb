"""The callers either side of the hot path, restated over the B200 extractor / predictor.

* :func:`prepare_points`           -- ``prepare_source`` grouping, ``scripts/build_feature_bucket.py:658-665``
* :func:`extract_features`         -- ``spacer.tasks.extract_features(msg)`` as the reference calls it,
                                      ``scripts/build_feature_bucket.py:775`` (load image, validate, extract, store)
* :func:`classify_features`        -- pyspacer ``classify_features``-shaped scorer over a ``Predictor``
                                      (``pyspacer/annotation.py:243-261``: per-point top scores)
* :func:`build_feature_bucket`     -- the per-source loop of ``process_source``,
                                      ``scripts/build_feature_bucket.py:691-788``: key layout
                                      ``s{sid}/features/i{iid}.featurevector``, skip-existing, per-image error
                                      capture, counters; images are sharded round-robin over ranks
* :func:`stack_feature_files`      -- ``scripts/extract_reference_features.py:40-61`` (``(N, 1280) float32 .npy``)

Storage goes through the ``DataLocation`` / ``storage_factory`` seam of :mod:`spacer_compat`: a root is a directory, an
``s3://bucket/prefix`` URI (boto3, as the reference's buckets) or ``memory://prefix``; everything numerical runs in
libmermaid_b200.
"""

from __future__ import annotations

import csv
import json
import time
from dataclasses import dataclass, field
from datetime import datetime, timezone
from pathlib import Path
from typing import Any, Iterable, Mapping, Sequence

import numpy as np

from .sharding import images_for_rank
from .spacer_compat import (
    DataLocation,
    ExtractFeaturesMsg,
    ExtractFeaturesReturnMsg,
    ImageFeatures,
    check_extract_inputs,
    image_features_from_array,
    load_image,
    storage_factory,
)


class StorageRoot:
    """A bucket / directory root: ``loc(key)`` is the ``DataLocation`` of ``key`` under it.

    ``/some/dir`` -> filesystem, ``s3://bucket/prefix`` -> S3 (the reference's layout, ``build_feature_bucket.py:530-544``),
    ``memory://prefix`` -> the in-process store (tests)."""

    def __init__(self, root: "str | Path | StorageRoot"):
        if isinstance(root, StorageRoot):
            self.storage_type, self.bucket, self.prefix = root.storage_type, root.bucket, root.prefix
            return
        text = str(root)
        if text.startswith("s3://"):
            bucket, _, prefix = text[5:].partition("/")
            if not bucket:
                raise ValueError(f"no bucket in {text!r}")
            self.storage_type, self.bucket, self.prefix = "s3", bucket, prefix
        elif text.startswith("memory://"):
            self.storage_type, self.bucket, self.prefix = "memory", None, text[9:]
        else:
            self.storage_type, self.bucket, self.prefix = "filesystem", None, text
        if self.prefix and not self.prefix.endswith("/"):
            self.prefix += "/"

    def loc(self, key: str) -> DataLocation:
        return DataLocation(self.storage_type, self.prefix + key, self.bucket)

    def storage(self):
        return storage_factory(self.storage_type, self.bucket)


def prepare_points(rows: Iterable[int], cols: Iterable[int]) -> list[tuple[int, int]]:
    """Sorted set of unique ``(row, col)`` int pairs -- defines the output row order."""
    return sorted({(int(r), int(c)) for r, c in zip(rows, cols)})


def feature_key(source_id: str | int, image_id: str | int) -> str:
    return f"s{source_id}/features/i{image_id}.featurevector"


def image_key(source_prefix: str, source_id: str | int, image_id: str | int) -> str:
    return f"{source_prefix}s{source_id}/images/{image_id}.jpg"


def extract_features(msg: ExtractFeaturesMsg) -> ExtractFeaturesReturnMsg:
    """Load ``msg.image_loc``, validate the points, run ``msg.extractor`` and store ``msg.feature_loc``."""
    t0 = time.time()
    img = load_image(msg.image_loc)
    check_extract_inputs(img, msg.rowcols, getattr(msg.image_loc, "key", ""))
    features, ret = msg.extractor(img, msg.rowcols)
    features.store(msg.feature_loc)
    return ExtractFeaturesReturnMsg(extractor_loaded_remotely=ret.extractor_loaded_remotely, runtime=time.time() - t0)


def classify_features(features: Any, predictor: Any, top_k: int | None = None) -> dict[str, Any]:
    """Score an ``ImageFeatures`` (or an ``(N, D)`` array) with a ``Predictor``.

    Returns ``{"classes": [...], "scores": [(row, col, [p_0..p_K-1]), ...]}``; with ``top_k`` each entry
    carries instead the ``top_k`` ``(label, score)`` pairs in descending order, ties in class order
    (``annotation.py:252-261``), selected on the device."""
    if isinstance(features, ImageFeatures):
        rowcols = [(pf.row, pf.col) for pf in features.point_features]
        X = np.asarray([pf.data for pf in features.point_features], dtype=np.float32).reshape(len(rowcols), -1)
    else:
        X = np.asarray(features, dtype=np.float32)
        rowcols = [(None, None)] * X.shape[0]
    if top_k:
        labels, scores = predictor.predict_topk(X, top_k)
        out = [(r, c, list(zip(labels[i].tolist(), scores[i].tolist()))) for i, (r, c) in enumerate(rowcols)]
    else:
        proba = predictor.predict_proba(X)
        out = [(r, c, proba[i].tolist()) for i, (r, c) in enumerate(rowcols)]
    return {"classes": list(predictor.classes), "scores": out}


@dataclass
class RunCounters:
    sources_done: int = 0
    sources_skipped: int = 0
    images_ok: int = 0
    images_skipped: int = 0
    images_failed: int = 0
    patches: int = 0
    started: float = field(default_factory=time.monotonic)


def _capture(fn, arg):
    """``(result, None)`` or ``(None, exception)`` -- lets a thread pool map over items that may fail."""
    try:
        return fn(arg), None
    except KeyboardInterrupt:
        raise
    except Exception as exc:
        return None, exc


def _ts() -> str:
    return datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")


def build_feature_bucket(
    sources: Mapping[str, Mapping[str, Sequence[tuple[int, int]]]],
    extractor: Any,
    *,
    source_root: "str | Path | StorageRoot",
    target_root: "str | Path | StorageRoot",
    source_prefix: str = "",
    skip_existing: bool = True,
    dry_run: bool = False,
    rank: int = 0,
    world: int = 1,
    error_csv: str | Path | None = None,
    progress_jsonl: str | Path | None = None,
    batch_images: int = 16,
    io_threads: int = 8,
    decode: str = "host",
) -> RunCounters:
    """Extract every image of every source into ``target_root/s{sid}/features/i{iid}.featurevector``.

    ``sources[sid][iid]`` is the image's rowcols (see :func:`prepare_points`).  Images of a source are walked in
    sorted id order and dealt round-robin to ranks (``index % world == rank``); a failing image is recorded and
    the run continues (``build_feature_bucket.py:774-786``).  The two logs follow the reference's records:
    ``progress_jsonl`` one JSON object per image -- ``ts, source_id, image_id, outcome`` (``ok`` / ``skipped`` /
    ``failed``) plus ``reason`` (``no_rowcols`` / ``exists``), ``error_type`` or ``dry_run``
    (``record_progress``, ``:794-808``); ``error_csv`` rows ``ts, source_id, image_id, error_type, error_msg`` under
    that header (``record_failure``, ``:810-822``, header ``:883``).  Both append, so a resumed run extends them.

    Where the reference makes one synchronous ``extract_features`` call per image, images are taken ``batch_images`` at a
    time: a thread pool loads and decodes them (the reference uses thread pools for its S3 I/O too, ``:308-309``), ONE
    ``extractor.extract_many`` call runs the batch through the library's pinned-staging pipeline (copy of image i+1
    overlaps the convolution of image i), and the ``.featurevector`` files are written by the pool.  Records keep image
    order; an image that fails to load, validate or store is logged and the rest of its batch is unaffected.  Extractors
    without ``extract_many`` fall back to one ``extract_features`` call per image.

    ``decode="device"`` reads the image files as bytes and decodes JPEG streams ON the GPU
    (:mod:`mermaid_classifier_b200.decode`, one nvJPEG decoder per pool thread): the decoded image never crosses PCIe and
    ``extractor.extract_device`` reads it where it lands."""
    counters = RunCounters()
    source_root, target_root = StorageRoot(source_root), StorageRoot(target_root)
    new_err = error_csv is not None and (not Path(error_csv).exists() or Path(error_csv).stat().st_size == 0)
    err_file = open(error_csv, "a", newline="") if error_csv else None
    err = csv.writer(err_file) if err_file else None
    if err and new_err:
        err.writerow(["ts", "source_id", "image_id", "error_type", "error_msg"])
    prog = open(progress_jsonl, "a") if progress_jsonl else None

    def progress(sid, iid, outcome, **extra):
        if prog:
            prog.write(json.dumps({"ts": _ts(), "source_id": sid, "image_id": iid, "outcome": outcome, **extra}) + "\n")
            prog.flush()

    from concurrent.futures import ThreadPoolExecutor

    if decode not in ("host", "device"):
        raise ValueError("decode must be 'host' or 'device'")
    batched = (hasattr(extractor, "extract_many") and batch_images > 1) or decode == "device"
    pool = ThreadPoolExecutor(max_workers=max(1, io_threads)) if batched else None
    decode_pool = None
    if decode == "device":
        from .decode import DecodePool

        extractor._ensure_handle()
        decode_pool = DecodePool(max(1, io_threads), device=extractor._device_index)

    def fail_image(sid, iid, exc):
        counters.images_failed += 1
        if err:
            err.writerow([_ts(), sid, iid, type(exc).__name__, str(exc)])
        progress(sid, iid, "failed", error_type=type(exc).__name__)

    def load_one(item):
        sid, iid, rowcols, floc = item
        loc = source_root.loc(image_key(source_prefix, sid, iid))
        img = load_image(loc)
        check_extract_inputs(img, rowcols, loc.key)
        return np.asarray(img)

    def run_batch_device(items):
        """Device decode: file bytes -> nvJPEG -> extract_device -> features back -> threaded stores."""
        def read_bytes(it):
            return source_root.storage().load(source_root.loc(image_key(source_prefix, it[0], it[1])).key).getvalue()

        blobs = list(pool.map(lambda it: _capture(read_bytes, it), items))
        outcome = {id(it): exc for it, (_, exc) in zip(items, blobs) if exc is not None}
        todo = [(it, b) for it, (b, exc) in zip(items, blobs) if exc is None]
        decoded = decode_pool.decode_many([b for _, b in todo])
        good = []
        for (it, _), (img, exc) in zip(todo, decoded):
            if exc is None:
                try:
                    check_extract_inputs(np.empty((img.shape[0], img.shape[1], 0), np.uint8), it[2], str(it[1]))
                except Exception as e2:
                    exc = e2
            if exc is not None:
                outcome[id(it)] = exc
            else:
                good.append((it, img))
        if good:
            try:
                pts = np.array([(k, r, c) for k, (it, _) in enumerate(good) for r, c in it[2]], dtype=np.int32)
                feats = extractor.extract_device([img for _, img in good], pts).cpu().numpy()
                o = 0
                stores = []
                for it, _ in good:
                    stores.append((it, feats[o:o + len(it[2])]))
                    o += len(it[2])
                results = list(pool.map(lambda sf: _capture(
                    lambda x: image_features_from_array(x[0][2], x[1]).store(x[0][3]), sf), stores))
                for (it, _), (_, exc) in zip(stores, results):
                    if exc is not None:
                        outcome[id(it)] = exc
            except KeyboardInterrupt:
                raise
            except Exception as exc:
                for it, _ in good:
                    outcome[id(it)] = exc
        for it in items:
            exc = outcome.get(id(it))
            if exc is None:
                counters.images_ok += 1
                counters.patches += len(it[2])
                progress(it[0], it[1], "ok")
            else:
                fail_image(it[0], it[1], exc)

    def run_batch(items):
        """items: (sid, iid, rowcols, feature_loc) in image order."""
        if decode_pool is not None:
            return run_batch_device(items)
        loaded = list(pool.map(lambda it: _capture(load_one, it), items))
        good = [(it, arr) for it, (arr, exc) in zip(items, loaded) if exc is None]
        outcome = {id(it): exc for it, (arr, exc) in zip(items, loaded) if exc is not None}
        if good:
            try:
                feats, _ = extractor.extract_many([arr for _, arr in good], [it[2] for it, _ in good])
                o = 0
                stores = []
                for it, _ in good:
                    n = len(it[2])
                    stores.append((it, feats[o:o + n]))
                    o += n
                results = list(pool.map(lambda sf: _capture(
                    lambda x: image_features_from_array(x[0][2], x[1]).store(x[0][3]), sf), stores))
                for (it, _), (_, exc) in zip(stores, results):
                    if exc is not None:
                        outcome[id(it)] = exc
            except KeyboardInterrupt:
                raise
            except Exception as exc:   # the whole batch failed on the device: every image of it is recorded
                for it, _ in good:
                    outcome[id(it)] = exc
        for it in items:
            exc = outcome.get(id(it))
            if exc is None:
                counters.images_ok += 1
                counters.patches += len(it[2])
                progress(it[0], it[1], "ok")
            else:
                fail_image(it[0], it[1], exc)

    try:
        for sid in sorted(sources):
            grouped = sources[sid]
            if not grouped:
                counters.sources_skipped += 1
                continue
            ids = sorted(grouped)
            pending: list[tuple] = []
            for k in images_for_rank(len(ids), rank, world):
                iid = ids[k]
                rowcols = list(grouped[iid])
                floc = target_root.loc(feature_key(sid, iid))
                if not rowcols:
                    counters.images_skipped += 1
                    progress(sid, iid, "skipped", reason="no_rowcols")
                    continue
                if skip_existing and target_root.storage().exists(floc.key):
                    counters.images_skipped += 1
                    progress(sid, iid, "skipped", reason="exists")
                    continue
                if dry_run:
                    counters.images_ok += 1
                    progress(sid, iid, "ok", dry_run=True)
                    continue
                if batched:
                    pending.append((sid, iid, rowcols, floc))
                    if len(pending) >= batch_images:
                        run_batch(pending)
                        pending = []
                    continue
                msg = ExtractFeaturesMsg(
                    job_token=f"s{sid}_i{iid}", extractor=extractor, rowcols=rowcols,
                    image_loc=source_root.loc(image_key(source_prefix, sid, iid)),
                    feature_loc=floc)
                try:
                    extract_features(msg)
                    counters.images_ok += 1
                    counters.patches += len(rowcols)
                    progress(sid, iid, "ok")
                except KeyboardInterrupt:
                    raise
                except Exception as exc:  # per-image failure: log and carry on
                    fail_image(sid, iid, exc)
            if pending:
                run_batch(pending)
            counters.sources_done += 1
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
        if decode_pool is not None:
            decode_pool.close()
        if err_file:
            err_file.close()
        if prog:
            prog.close()
    return counters


def stack_feature_files(paths: Sequence[str | Path], out: str | Path) -> np.ndarray:
    """Concatenate the per-point vectors of ``.featurevector`` files into an ``(N, D) float32 .npy``."""
    vectors: list[Any] = []
    for p in paths:
        feats = ImageFeatures.load(DataLocation("filesystem", str(p)))
        vectors.extend(pf.data for pf in feats.point_features)
    X = np.asarray(vectors, dtype=np.float32)
    if X.ndim != 2:
        raise SystemExit(f"expected a 2-D feature matrix; got shape {X.shape}")
    np.save(out, X)
    return X
