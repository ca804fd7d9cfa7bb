"""Pin the head / training oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py imports /root/reference)."""
import json

import numpy as np
import torch

from mermaid_classifier_b200 import synth
from oracle import head as ohead


def test_calibrated_head_small(golden_dir):
    io = np.load(golden_dir / "head_small_io.npz")
    w, bb, a, b, _ = synth.synth_head(input_dim=32, hidden=(24, 16), n_classes=12, seed=7)
    got = ohead.calibrated_proba(io["X"], w, bb, a, b)
    assert got.dtype == np.float64
    assert np.max(np.abs(got - io["proba"])) <= 1e-7
    assert np.array_equal(got.argmax(1), io["proba"].argmax(1))
    manifest = json.loads((golden_dir / "head_small" / "model.json").read_text())
    assert manifest["schema_version"] == 1 and manifest["input_dim"] == 32 and len(manifest["classes"]) == 12


def test_calibrated_head_full(golden_dir):
    for tag, hidden in (("h200_100", (200, 100)), ("h500_300_100", (500, 300, 100))):
        g = np.load(golden_dir / f"head_full_{tag}.npz")
        w, bb, a, b, _ = synth.synth_head(1280, hidden, 500, seed=0)
        X = synth.synth_features(512, 1280, seed=5).numpy()
        got = ohead.calibrated_proba(X, w, bb, a, b)
        assert np.max(np.abs(got[g["rows"]] - g["proba_rows"])) <= 1e-6
        assert (got.argmax(1) == g["labels"]).mean() >= 0.999
        assert np.allclose(got.sum(1), 1.0, atol=1e-5)


def test_topk_is_stable_descending():
    p = np.array([[0.2, 0.5, 0.2, 0.1], [0.25, 0.25, 0.25, 0.25]])
    assert ohead.topk_labels(p, 3).tolist() == [[1, 0, 2], [0, 1, 2]]
    assert ohead.argmax_labels(p).tolist() == [1, 0]


def test_mlp_init_matches_reference(golden_dir):
    g = np.load(golden_dir / "mlp_train.npz")
    w, b = ohead.init_mlp(32, (16, 8), 5, 0)
    for i in range(3):
        assert np.array_equal(w[i].numpy(), g[f"init_W{i}"])
        assert not b[i].any()


def test_partial_fit_matches_reference(golden_dir):
    g = np.load(golden_dir / "mlp_train.npz")
    for tag in ("plain", "weighted"):
        w, b = ohead.init_mlp(32, (16, 8), 5, 0)
        adam = ohead.AdamState(w + b)
        cw = None if tag == "plain" else torch.tensor([0.5 + 0.5 * i for i in range(5)], dtype=torch.float32)
        curve = [ohead.partial_fit(w, b, adam, g["X"], g["y_idx"], lr=1e-3, random_state=0, class_weight=cw)
                 for _ in range(3)]
        assert adam.t == 3 * 4  # ceil(650 / 200) Adam steps per call
        np.testing.assert_allclose(curve, g[f"{tag}_loss_curve"], rtol=2e-5)
        for i in range(3):
            np.testing.assert_allclose(w[i].numpy(), g[f"{tag}_W{i}"], atol=2e-5, rtol=1e-4)
            np.testing.assert_allclose(b[i].numpy(), g[f"{tag}_b{i}"], atol=2e-5, rtol=1e-4)
        proba = ohead.softmax_proba(g["X"][:16], w, b)
        np.testing.assert_allclose(proba, g[f"{tag}_proba"], atol=1e-5)
        assert np.abs(proba.sum(1) - 1).max() < 1e-12
