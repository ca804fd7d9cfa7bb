"""Host logic either side of the hot path (no GPU): point grouping, bucket key layout, skip-existing,
per-image error capture, rank sharding of the bucket driver, .featurevector round trip and stacking."""
import csv
import json

import numpy as np
from PIL import Image

from mermaid_classifier_b200 import drivers
from mermaid_classifier_b200.spacer_compat import (
    DataLocation,
    ExtractFeaturesReturnMsg,
    ImageFeatures,
    image_features_from_array,
)


class FakeExtractor:
    """Stands in for the GPU extractor: feature = (row, col, mean pixel) so results are checkable."""

    def __init__(self):
        self.calls = 0

    def __call__(self, im, rowcols):
        self.calls += 1
        arr = np.asarray(im)
        feats = np.array([[r, c, arr.mean()] + [0.0] * 5 for r, c in rowcols], dtype=np.float32)
        return image_features_from_array(rowcols, feats), ExtractFeaturesReturnMsg(False, 0.0)


def test_prepare_points_sorted_unique():
    assert drivers.prepare_points([5, 1, 5, 1], [2, 9, 2, 3]) == [(1, 3), (1, 9), (5, 2)]


def test_image_features_round_trip(tmp_path):
    feats = np.random.RandomState(0).randn(4, 1280).astype(np.float32)
    rc = [(3, 4), (10, 2), (10, 7), (99, 0)]
    f = image_features_from_array(rc, feats)
    loc = DataLocation("filesystem", str(tmp_path / "s1" / "features" / "i7.featurevector"))
    f.store(loc)
    g = ImageFeatures.load(loc)
    assert g.valid_rowcol and g.npoints == 4 and g.feature_dim == 1280
    assert np.array_equal(g.get_array((10, 7)), feats[2])  # lossless fp32
    assert [(p.row, p.col) for p in g.point_features] == rc


def test_bucket_driver_layout_skip_errors_and_sharding(tmp_path):
    src, tgt = tmp_path / "src", tmp_path / "tgt"
    sources = {"12": {}, "7": {}}
    for sid, n in (("12", 5), ("7", 2)):
        (src / f"s{sid}" / "images").mkdir(parents=True)
        for i in range(n):
            Image.fromarray(np.full((40, 50, 3), 10 * i, np.uint8)).save(src / f"s{sid}" / "images" / f"{i}.jpg")
            sources[sid][str(i)] = [(1, 2), (3, 4)]
    sources["12"]["3"] = [(1, 2), (400, 4)]      # out-of-bounds point -> per-image failure, run continues
    sources["12"]["4"] = []                       # no rowcols -> skipped
    ex = FakeExtractor()
    dry = drivers.build_feature_bucket(sources, ex, source_root=src, target_root=tgt, dry_run=True,
                                       progress_jsonl=tmp_path / "dry.jsonl")
    assert dry.images_ok == 6 and dry.images_skipped == 1 and ex.calls == 0 and not tgt.exists()
    assert all(r["dry_run"] for r in map(json.loads, (tmp_path / "dry.jsonl").read_text().splitlines()) if r["outcome"] == "ok")
    c0 = drivers.build_feature_bucket(sources, ex, source_root=src, target_root=tgt, rank=0, world=2,
                                      error_csv=tmp_path / "err.csv", progress_jsonl=tmp_path / "p0.jsonl")
    c1 = drivers.build_feature_bucket(sources, ex, source_root=src, target_root=tgt, rank=1, world=2,
                                      error_csv=tmp_path / "err.csv", progress_jsonl=tmp_path / "p1.jsonl")
    assert c0.images_ok + c1.images_ok == 5 and c0.images_failed + c1.images_failed == 1
    assert c0.images_skipped + c1.images_skipped == 1 and c0.sources_done == 2
    assert sorted(p.name for p in (tgt / "s12" / "features").iterdir()) == ["i0.featurevector", "i1.featurevector", "i2.featurevector"]
    err_rows = list(csv.reader((tmp_path / "err.csv").open()))
    assert err_rows[0] == ["ts", "source_id", "image_id", "error_type", "error_msg"] and len(err_rows) == 2
    assert err_rows[1][1:4] == ["12", "3", "RowColumnInvalidError"]
    recs = [json.loads(l) for f in ("p0.jsonl", "p1.jsonl") for l in (tmp_path / f).read_text().splitlines()]
    outcome = {(r["source_id"], r["image_id"]): r for r in recs}
    assert len(recs) == 7 and outcome[("12", "4")]["reason"] == "no_rowcols" and outcome[("12", "3")]["outcome"] == "failed"
    assert outcome[("12", "3")]["error_type"] == "RowColumnInvalidError" and outcome[("7", "1")]["outcome"] == "ok"
    assert set(outcome[("7", "1")]) == {"ts", "source_id", "image_id", "outcome"}
    calls = ex.calls
    again = drivers.build_feature_bucket(sources, ex, source_root=src, target_root=tgt, skip_existing=True,
                                         error_csv=tmp_path / "err.csv", progress_jsonl=tmp_path / "again.jsonl")
    again_recs = [json.loads(l) for l in (tmp_path / "again.jsonl").read_text().splitlines()]
    assert sum(r.get("reason") == "exists" for r in again_recs) == 5
    assert len(list(csv.reader((tmp_path / "err.csv").open()))) == 3   # appended, header written once
    # only the failing image is retried, and it fails validation again before reaching the extractor
    assert again.images_ok == 0 and again.images_skipped == 6 and again.images_failed == 1 and ex.calls == calls
    X = drivers.stack_feature_files(sorted((tgt / "s7" / "features").iterdir()), tmp_path / "ref.npy")
    assert X.shape == (4, 8) and X.dtype == np.float32 and np.load(tmp_path / "ref.npy").shape == (4, 8)


def test_classify_features_shapes():
    class P:
        classes = ["a", "b", "c"]

        def predict_proba(self, X):
            return np.tile(np.array([[0.2, 0.5, 0.3]]), (X.shape[0], 1))

        def predict_topk(self, X, k):
            return np.array([["b", "c"]] * X.shape[0], dtype=object), np.array([[0.5, 0.3]] * X.shape[0])

    f = image_features_from_array([(1, 2), (3, 4)], np.zeros((2, 8), np.float32))
    out = drivers.classify_features(f, P())
    assert out["classes"] == ["a", "b", "c"] and out["scores"][1][:2] == (3, 4) and len(out["scores"][0][2]) == 3
    top = drivers.classify_features(f, P(), top_k=2)
    assert top["scores"][0][2] == [("b", 0.5), ("c", 0.3)]


class FakeBatchExtractor(FakeExtractor):
    """Adds the batch surface of the GPU extractor (extract_many): one call per group of images."""

    def __init__(self):
        super().__init__()
        self.batches = []

    def extract_many(self, images, rowcols_list, head=None, **_kw):
        self.batches.append(len(images))
        rows = [[r, c, np.asarray(im).mean()] + [0.0] * 5 for im, rcs in zip(images, rowcols_list) for r, c in rcs]
        return np.asarray(rows, dtype=np.float32).reshape(-1, 8), None


def test_bucket_driver_batched_path_matches_per_image_path(tmp_path):
    """The driver hands images to extract_many in batches (threaded loads and stores); files, counters and records must
    equal the one-image-per-call path, and a bad image must not take its batch down."""
    src = tmp_path / "src"
    (src / "s5" / "images").mkdir(parents=True)
    sources = {"5": {}}
    for i in range(7):
        Image.fromarray(np.full((40, 50, 3), 10 * i, np.uint8)).save(src / "s5" / "images" / f"{i}.png")
        (src / "s5" / "images" / f"{i}.png").rename(src / "s5" / "images" / f"{i}.jpg")
        sources["5"][str(i)] = [(1, 2), (3, 4), (5, 6)]
    sources["5"]["2"] = [(1, 2), (400, 4)]          # fails validation inside a batch
    sources["5"]["6"] = []                           # skipped
    (src / "s5" / "images" / "4.jpg").unlink()      # fails to load inside a batch
    one, many = FakeExtractor(), FakeBatchExtractor()
    ca = drivers.build_feature_bucket(sources, one, source_root=src, target_root=tmp_path / "a", progress_jsonl=tmp_path / "a.jsonl")
    cb = drivers.build_feature_bucket(sources, many, source_root=src, target_root=tmp_path / "b", progress_jsonl=tmp_path / "b.jsonl",
                                      batch_images=3, io_threads=2)
    assert many.batches == [2, 2] and many.calls == 0   # images 0..2 -> (0, 1) after the bad one drops out; 3..5 -> (3, 5)
    for c in (ca, cb):
        assert (c.images_ok, c.images_failed, c.images_skipped, c.patches) == (4, 2, 1, 12)
    ra = [{k: v for k, v in json.loads(l).items() if k != "ts"} for l in (tmp_path / "a.jsonl").read_text().splitlines()]
    rb = [{k: v for k, v in json.loads(l).items() if k != "ts"} for l in (tmp_path / "b.jsonl").read_text().splitlines()]
    assert sorted(ra, key=lambda r: r["image_id"]) == sorted(rb, key=lambda r: r["image_id"])
    fa = sorted((tmp_path / "a" / "s5" / "features").iterdir())
    fb = sorted((tmp_path / "b" / "s5" / "features").iterdir())
    assert [f.name for f in fa] == [f.name for f in fb] == ["i0.featurevector", "i1.featurevector", "i3.featurevector", "i5.featurevector"]
    for x, y in zip(fa, fb):
        gx, gy = ImageFeatures.load(DataLocation("filesystem", str(x))), ImageFeatures.load(DataLocation("filesystem", str(y)))
        assert [(p.row, p.col) for p in gx.point_features] == [(p.row, p.col) for p in gy.point_features]
        assert np.array_equal(np.stack([p.data for p in gx.point_features]), np.stack([p.data for p in gy.point_features]))


class _FakeS3:
    """boto3-client-shaped dictionary: get_object / put_object / head_object / delete_object."""

    def __init__(self):
        self.objects = {}

    def get_object(self, Bucket, Key):
        import io
        return {"Body": io.BytesIO(self.objects[(Bucket, Key)])}

    def put_object(self, Bucket, Key, Body):
        self.objects[(Bucket, Key)] = bytes(Body)

    def head_object(self, Bucket, Key):
        if (Bucket, Key) not in self.objects:
            err = Exception("not found")
            err.response = {"Error": {"Code": "404"}}
            raise err
        return {}

    def delete_object(self, Bucket, Key):
        del self.objects[(Bucket, Key)]


def test_bucket_driver_over_s3_locations(monkeypatch):
    """The reference's deployment reads images from and writes features to S3 buckets
    (``scripts/build_feature_bucket.py:530-544``): same key layout through ``DataLocation('s3', key, bucket)``,
    skip-existing by ``head_object``; the per-image and the batched path write identical objects."""
    import io

    from mermaid_classifier_b200 import spacer_compat

    s3 = _FakeS3()
    monkeypatch.setattr(spacer_compat._S3Storage, "client_factory", staticmethod(lambda: s3))
    sources = {"5": {}}
    for i in range(5):
        buf = io.BytesIO()
        Image.fromarray(np.full((40, 50, 3), 20 * i, np.uint8)).save(buf, format="JPEG")
        s3.put_object("coralnet-src", f"pub/s5/images/{i}.jpg", buf.getvalue())
        sources["5"][str(i)] = [(1, 2), (3, 4)]
    c = drivers.build_feature_bucket(sources, FakeExtractor(), source_root="s3://coralnet-src", source_prefix="pub/",
                                     target_root="s3://mermaid-features/run1")
    assert c.images_ok == 5 and c.images_failed == 0
    keys = sorted(k for b, k in s3.objects if b == "mermaid-features")
    assert keys == [f"run1/s5/features/i{i}.featurevector" for i in range(5)]
    f = ImageFeatures.load(DataLocation("s3", "run1/s5/features/i3.featurevector", "mermaid-features"))
    assert [(p.row, p.col) for p in f.point_features] == [(1, 2), (3, 4)]
    again = drivers.build_feature_bucket(sources, FakeExtractor(), source_root="s3://coralnet-src", source_prefix="pub/",
                                         target_root="s3://mermaid-features/run1", skip_existing=True)
    assert again.images_skipped == 5 and again.images_ok == 0
    many = drivers.build_feature_bucket(sources, FakeBatchExtractor(), source_root="s3://coralnet-src", source_prefix="pub/",
                                        target_root="s3://mermaid-features/run2", batch_images=3)
    assert many.images_ok == 5
    for i in range(5):
        assert s3.objects[("mermaid-features", f"run2/s5/features/i{i}.featurevector")] == \
            s3.objects[("mermaid-features", f"run1/s5/features/i{i}.featurevector")]
    root = drivers.StorageRoot("memory://bucket")
    assert root.loc("a/b").storage_type == "memory" and root.loc("a/b").key == "bucket/a/b"
