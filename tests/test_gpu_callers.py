"""The two callers SURVEY section 8f-4 names, driven with the B200 objects in the reference's own call order:

* ``AnnotationRun.__init__`` -- ``mermaid_classifier/pyspacer/annotation.py:231-261``: ``load_predictor`` ->
  ``load_image`` -> ``EfficientNetExtractor(data_locations=...)`` -> ``check_extract_inputs`` -> ``extractor(img, rowcols)``
  -> ``np.vstack([features.get_array(rc) ...])`` -> ``predictor.predict_proba`` -> per-point ``sorted(zip(labels, proba),
  key=itemgetter(1), reverse=True)`` top predictions;
* ``MetricsCoordinator._precompute_probabilities`` -- ``mermaid_classifier/pyspacer/metrics/coordinator.py:59-82``: for each
  validation batch ``clf.predict_proba(batch_x)``, ``np.vstack`` of the results, ground truth extended batch by batch.

The reference modules themselves import mlflow / duckdb / matplotlib (absent from this image) and live in the build
container only, so the loops are restated here line by line and checked against the CPU oracle and the reference-run
golden artifact (``tests/golden/head_small``: produced by the reference's own ``CalibratedHead``)."""
import io
from operator import itemgetter

import numpy as np
import pytest
import torch
from PIL import Image

from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.extractor import EfficientNetExtractor
from mermaid_classifier_b200.inference import load_predictor
from mermaid_classifier_b200.spacer_compat import DataLocation, check_extract_inputs, load_image
from oracle import crop as ocrop
from oracle import effnet as oeff
from oracle import head as ohead

pytestmark = pytest.mark.gpu


def test_annotation_run_sequence(tmp_path, backbone_sd):
    """extract + classify of one image exactly as annotation.py:231-261 sequences it, with an artifact written by
    export_artifact-compatible code (input_dim 1280) and a pyspacer-layout weights checkpoint on disk."""
    from mermaid_classifier_b200.export import write_head_artifact

    # the artifact: MLP(200,100)/Platt head as model.pt + model.json; the weights: a {'net': {'module.'+k: v}} checkpoint
    w, bb, a, b, classes = synth.synth_head(1280, (200, 100), 37, seed=3)
    model_pt, model_json = write_head_artifact(tmp_path / "artifact", w, bb, a, b, classes)
    ckpt = tmp_path / "efficientnet_weights.pt"
    torch.save({"net": {"module." + k: v for k, v in backbone_sd.items()}}, ckpt)
    im = synth.synth_image(synth.DEFAULT_SEED, 77, 500, 640)
    Image.fromarray(im).save(tmp_path / "img.png")
    annotations = {rc: None for rc in synth.synth_points(synth.DEFAULT_SEED, 77, 500, 640, 11, corners=True)}
    scores = {}
    num_predictions_to_save = 3

    # ---- annotation.py:231-261, restated -------------------------------------------------------------------------
    predictor = load_predictor(model_pt, model_json)
    image_loc = DataLocation("filesystem", str(tmp_path / "img.png"))
    loaded_image = load_image(image_loc)
    extractor = EfficientNetExtractor(data_locations={"weights": DataLocation("filesystem", str(ckpt))})
    rowcols = list(annotations.keys())
    check_extract_inputs(loaded_image, rowcols, image_loc.key)
    features, _ = extractor(loaded_image, rowcols)
    predictions_per_point = max(num_predictions_to_save, 1)
    labels = predictor.classes
    feature_batch = np.vstack([features.get_array(rowcol) for rowcol in rowcols])
    proba_batch = predictor.predict_proba(feature_batch).tolist()
    for (row, column), proba in zip(rowcols, proba_batch):
        top_predictions = sorted(zip(labels, proba), key=itemgetter(1), reverse=True)
        annotations[(row, column)] = [label for label, _ in top_predictions[:predictions_per_point]]
        scores[(row, column)] = [score for _, score in top_predictions[:predictions_per_point]]
    extractor.close()
    # ---------------------------------------------------------------------------------------------------------------
    want_f = oeff.extract_features(backbone_sd, torch.from_numpy(ocrop.normalize_patches(ocrop.crop_patches(im, rowcols)))).numpy()
    want_p = ohead.calibrated_proba(want_f, w, bb, a, b)
    assert np.abs(feature_batch - want_f).max() <= 1e-3
    assert np.abs(np.asarray(proba_batch) - want_p).max() <= 1e-4
    want_top = ohead.topk_labels(want_p, 3)
    assert [annotations[rc] for rc in rowcols] == np.asarray(classes, dtype=object)[want_top].tolist()
    # the device top-k path picks the same labels without moving the (N, K) matrix
    dev_labels, dev_scores = predictor.predict_topk(feature_batch, 3)
    assert dev_labels.tolist() == [annotations[rc] for rc in rowcols]
    assert np.abs(dev_scores - np.asarray([scores[rc] for rc in rowcols])).max() <= 1e-5


def test_metrics_coordinator_precompute_probabilities(golden_dir):
    """coordinator.py:59-82 with the B200 Predictor as ctx.clf: the stacked matrix equals the reference-run golden output."""
    pred = load_predictor(golden_dir / "head_small" / "model.pt", golden_dir / "head_small" / "model.json")
    io_ = np.load(golden_dir / "head_small_io.npz")
    X, want = io_["X"], io_["proba"]
    gt = [pred.classes[i % len(pred.classes)] for i in range(X.shape[0])]

    class _Val:
        def load_data_in_batches(self, batch_size=7, random_seed=None):
            for s in range(0, X.shape[0], batch_size):   # ImageLabels yields Python lists of vectors and labels
                yield [row.tolist() for row in X[s:s + batch_size]], gt[s:s + batch_size]

    class _Ctx:
        clf = pred
        val_proba = None
        val_gt_labels = None

    ctx = _Ctx()
    # ---- coordinator.py:71-76, restated --------------------------------------------------------------------------
    all_proba, all_gt = [], []
    for batch_x, batch_y in _Val().load_data_in_batches():
        all_proba.append(ctx.clf.predict_proba(batch_x))
        all_gt.extend(batch_y)
    ctx.val_proba = np.vstack(all_proba)
    ctx.val_gt_labels = all_gt
    # ---------------------------------------------------------------------------------------------------------------
    assert ctx.val_proba.dtype == np.float64 and ctx.val_proba.shape == want.shape
    assert np.abs(ctx.val_proba - want).max() <= 1e-6 and ctx.val_gt_labels == gt
    assert np.abs(ctx.val_proba.sum(1) - 1).max() <= 1e-5
