"""GPU end-to-end tests of the callers either side of the hot path: spacer-style extract_features task ->
.featurevector bucket (bf16 mode, sharded over two ranks) -> stacked .npy -> classify_features with a
model.pt/model.json artifact, compared with the CPU oracle."""
import numpy as np
import pytest
import torch
from PIL import Image

from mermaid_classifier_b200 import drivers, synth
from mermaid_classifier_b200.extractor import EfficientNetExtractor
from mermaid_classifier_b200.inference import load_predictor
from mermaid_classifier_b200.spacer_compat import DataLocation, ImageFeatures
from oracle import crop as ocrop
from oracle import effnet as oeff
from oracle import head as ohead

pytestmark = pytest.mark.gpu


def test_bucket_build_stack_and_classify(tmp_path, backbone_sd, golden_dir):
    src, tgt = tmp_path / "src", tmp_path / "tgt"
    (src / "s3" / "images").mkdir(parents=True)
    sources = {"3": {}}
    ims = {}
    for i in range(4):
        im = synth.synth_image(synth.DEFAULT_SEED, 40 + i, 300, 360)
        Image.fromarray(im).save(src / "s3" / "images" / f"{i}.png")   # lossless stand-in for the .jpg
        (src / "s3" / "images" / f"{i}.png").rename(src / "s3" / "images" / f"{i}.jpg")
        pts = synth.synth_points(synth.DEFAULT_SEED, 40 + i, 300, 360, 6, corners=(i == 0))
        sources["3"][str(i)] = drivers.prepare_points([r for r, _ in pts], [c for _, c in pts])
        ims[str(i)] = im
    ext = EfficientNetExtractor(state_dict=backbone_sd, mode="bf16", max_batch=8)   # < points of image 0 -> several sub-batches
    counters = [drivers.build_feature_bucket(sources, ext, source_root=src, target_root=tgt, rank=r, world=2)
                for r in range(2)]
    assert sum(c.images_ok for c in counters) == 4 and sum(c.images_failed for c in counters) == 0
    assert [c.images_ok for c in counters] == [2, 2]   # round-robin over ranks
    files = sorted((tgt / "s3" / "features").iterdir())
    assert [f.name for f in files] == [f"i{i}.featurevector" for i in range(4)]
    # per-image parity with the oracle in bf16 mode: the stated bf16 bound (tests/test_gpu_extract.py)
    for i in range(4):
        feats = ImageFeatures.load(DataLocation("filesystem", str(files[i])))
        rc = sources["3"][str(i)]
        assert [(p.row, p.col) for p in feats.point_features] == rc and feats.feature_dim == 1280
        got = np.stack([feats.get_array(x) for x in rc])
        want = oeff.extract_features(backbone_sd, torch.from_numpy(ocrop.normalize_patches(ocrop.crop_patches(ims[str(i)], rc)))).numpy()
        cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
        assert cos.min() >= 0.998, cos.min()
    X = drivers.stack_feature_files(files, tmp_path / "bucket.npy")
    assert X.shape == (sum(len(v) for v in sources["3"].values()), 1280) and X.dtype == np.float32
    # skip-existing: a second pass extracts nothing
    again = drivers.build_feature_bucket(sources, ext, source_root=src, target_root=tgt)
    assert again.images_ok == 0 and again.images_skipped == 4


def test_classify_features_with_artifact(golden_dir):
    pred = load_predictor(golden_dir / "head_small" / "model.pt", golden_dir / "head_small" / "model.json")
    io = np.load(golden_dir / "head_small_io.npz")
    out = drivers.classify_features(io["X"][:10], pred)
    assert out["classes"] == pred.classes
    got = np.array([s for _, _, s in out["scores"]])
    assert np.max(np.abs(got - io["proba"][:10])) <= 1e-6
    top = drivers.classify_features(io["X"][:10], pred, top_k=3)
    want = ohead.topk_labels(io["proba"][:10], 3)
    assert [[lab for lab, _ in s] for _, _, s in top["scores"]] == np.asarray(pred.classes, dtype=object)[want].tolist()
