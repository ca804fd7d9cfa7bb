"""Host-side logic of the trainer / export mirrors that needs no GPU: argument validation, the calibrated-model
attribute surface, the portable TorchScript head and the export parity gate (checked against the oracle)."""
import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.export import PortableHead, build_calibrated_head, export_artifact
from mermaid_classifier_b200.inference import ParityError, extract_head_params
from mermaid_classifier_b200.trainer import MermaidTrainer, SigmoidCalibrator
from oracle import head as ohead


def _stub_model(n_classes=12, input_dim=32, hidden=(24, 16), skew=0.0):
    """A calibrated model with the CalibratedClassifierCV(prefit) attribute surface; predict_proba is the oracle's."""
    w, bb, a, b, _ = synth.synth_head(input_dim=input_dim, hidden=hidden, n_classes=n_classes, seed=7)
    classes = np.asarray([f"ba{i:02d}::gf{i:02d}" for i in range(n_classes)])
    linears = [SimpleNamespace(weight=x, bias=y) for x, y in zip(w, bb)]
    est = SimpleNamespace(classes_=classes, n_features_in_=input_dim, _module=SimpleNamespace(linears=linears))
    inner = SimpleNamespace(estimator=est, calibrators=[SigmoidCalibrator(float(x), float(y)) for x, y in zip(a, b)])
    model = SimpleNamespace(calibrated_classifiers_=[inner], classes_=classes,
                            predict_proba=lambda X: ohead.calibrated_proba(X, w, bb, a, b) + skew)
    return model, (w, bb, a, b)


def test_trainer_rejects_bad_patience():
    for bad in (0, -1):
        with pytest.raises(ValueError, match="early_stopping_patience"):
            MermaidTrainer(batch_size=100, early_stopping_patience=bad)
    t = MermaidTrainer(batch_size=100, early_stopping_patience=None)
    assert t._early_stop_info is None and t.serialize() == {"batch_size": 100}


def test_sigmoid_calibrator_predict():
    c = SigmoidCalibrator(-7.5, 2.0)
    p = np.linspace(0, 1, 11)
    np.testing.assert_allclose(c.predict(p), 1.0 / (1.0 + np.exp(-7.5 * p + 2.0)), rtol=1e-15)
    assert c.predict(p).dtype == np.float64


def test_portable_head_is_the_oracle_head_and_is_readable():
    model, (w, bb, a, b) = _stub_model()
    head = build_calibrated_head(model).eval()
    x = np.random.default_rng(0).standard_normal((64, 32)).astype(np.float32)
    with torch.no_grad():
        got = head(torch.from_numpy(x)).numpy().astype(np.float64)
    assert np.array_equal(got, ohead.calibrated_proba(x, w, bb, a, b))
    frozen = torch.jit.freeze(torch.jit.script(head))
    ws, bs, a2, b2 = extract_head_params(frozen)
    assert all(torch.equal(p, q) for p, q in zip(ws, w)) and all(torch.equal(p, q) for p, q in zip(bs, bb))
    assert torch.equal(a2, a) and torch.equal(b2, b)


def test_build_calibrated_head_rejections():
    model, _ = _stub_model()
    model.calibrated_classifiers_ = model.calibrated_classifiers_ * 2
    with pytest.raises(ValueError, match="exactly one"):
        build_calibrated_head(model)
    model, _ = _stub_model()
    model.classes_ = model.classes_[::-1]
    with pytest.raises(ValueError, match="classes_"):
        build_calibrated_head(model)
    model, _ = _stub_model()
    model.calibrated_classifiers_[0].calibrators.pop()
    with pytest.raises(ValueError, match="per-class calibrators"):
        build_calibrated_head(model)
    model, _ = _stub_model(n_classes=2)
    with pytest.raises(ValueError, match="K > 2"):
        build_calibrated_head(model)
    with pytest.raises(ValueError):
        PortableHead([], [], torch.zeros(3), torch.zeros(3))
    with pytest.raises(ValueError):
        PortableHead([torch.zeros(3, 4)], [torch.zeros(3)], torch.zeros(3, 1), torch.zeros(3, 1))


def test_export_artifact_manifest_and_gate(tmp_path):
    model, (w, bb, a, b) = _stub_model()
    x = np.random.default_rng(1).standard_normal((32, 32)).astype(np.float32)
    path, manifest, diff = export_artifact(model, tmp_path / "art", x, config={"patch_size": 224, "note": "t"})
    assert path == tmp_path / "art" / "model.pt" and diff <= 1e-6
    on_disk = json.loads((tmp_path / "art" / "model.json").read_text())
    assert on_disk == manifest
    assert manifest["schema_version"] == 1 and manifest["task"] == "pyspacer_mlp_classifier"
    assert manifest["classes"] == model.classes_.tolist() and manifest["input_dim"] == 32
    assert manifest["config"] == {"patch_size": 224, "note": "t"}
    assert set(manifest["trained_with"]) >= {"torch", "sklearn", "pyspacer"}
    graph = torch.jit.load(str(path), map_location="cpu")
    with torch.no_grad():
        out = graph(torch.from_numpy(x)).numpy().astype(np.float64)
    assert np.array_equal(out, ohead.calibrated_proba(x, w, bb, a, b))
    # default config, and the gate refusing a model that disagrees with its own frozen graph
    _, manifest2, _ = export_artifact(model, tmp_path / "art2", x)
    assert manifest2["config"] == {"patch_size": 224}
    skewed, _ = _stub_model(skew=5e-6)
    with pytest.raises(ParityError, match="Refusing to ship"):
        export_artifact(skewed, tmp_path / "art3", x)
    assert not (tmp_path / "art3" / "model.pt").exists()


# ---- the reference's own early-stopping cases (tests/pyspacer/test_trainer.py:171-359), scripted the same way ----------
class _FakeClf:
    """Stands in for the GPU estimator: the loop bookkeeping under test never looks inside it."""

    def __init__(self, **_kw):
        self.loss_curve_ = []
        self.classes_ = np.asarray(["c0", "c1", "c2"])
        self.fits = 0

    def partial_fit(self, x, y, classes=None):
        self.fits += 1
        self.loss_curve_.append(1.0 / self.fits)
        return self

    def predict_proba(self, x):
        return np.full((len(x), 3), 1.0 / 3.0)


def _mock_labels(seed=0):
    rng = np.random.RandomState(seed)
    classes = ["c0", "c1", "c2"]

    def split(n):
        X = rng.randn(n, 16).astype(np.float32)
        y = list(rng.choice(classes, size=n))

        def gen(batch_size=10, random_seed=None):
            for i in range(0, n, batch_size):
                yield [X[j] for j in range(i, min(i + batch_size, n))], list(y[i:i + batch_size])

        return SimpleNamespace(classes_set=set(classes), label_count=n, load_data_in_batches=gen)

    return SimpleNamespace(train=split(30), ref=split(15), val=split(15))


def _scripted_run(schedule, patience, n_epochs):
    remaining = list(schedule)
    snapshots = []

    class Scripted(MermaidTrainer):
        def _calc_acc_batched(self, clf, labels):
            return 0.25

        def _calc_acc_and_log_loss_batched(self, clf, labels, classes_list):
            assert remaining, "val_loss schedule exhausted; trainer ran for more epochs than expected"
            return 0.5, float(remaining.pop(0))

        def _calibrate_in_batches(self, clf, ref_labels):
            snapshots.append(clf.fits)   # how many partial_fit calls the estimator handed to calibration had seen
            return clf

    captured = []
    trainer = Scripted(batch_size=10, on_epoch_end=lambda m: captured.append(dict(m)), early_stopping_patience=patience,
                       clf_factory=_FakeClf)
    model, val_results, msg = trainer(_mock_labels(), n_epochs, [])
    return trainer, captured, snapshots[0], msg


def test_reference_case_no_patience_runs_full_budget():
    trainer, captured, fits, msg = _scripted_run([1.0, 0.9, 0.8, 0.7, 0.6], None, 5)
    assert len(captured) == 5 and len(msg.ref_accs) == 5
    assert trainer._early_stop_info == {"enabled": False, "patience": None, "stop_reason": "budget_exhausted", "final_epoch": 5,
                                        "best_val_epoch": None, "best_val_loss": None}
    assert fits == 15   # 3 chunks x 5 epochs: the last-epoch estimator is calibrated


def test_reference_case_monotone_down_no_stop():
    trainer, captured, fits, _ = _scripted_run([1.0, 0.9, 0.8, 0.7, 0.6], 2, 5)
    info = trainer._early_stop_info
    assert len(captured) == 5 and info["enabled"] and info["stop_reason"] == "budget_exhausted"
    assert info["final_epoch"] == 5 and info["best_val_epoch"] == 5 and info["best_val_loss"] == pytest.approx(0.6)
    assert fits == 15


def test_reference_case_v_shape_triggers_after_patience():
    trainer, captured, fits, _ = _scripted_run([1.0, 0.9, 0.8, 0.85, 0.9, 1.0, 1.1], 2, 10)
    info = trainer._early_stop_info
    assert info["enabled"] and info["stop_reason"] == "early_stopping"
    assert info["final_epoch"] == 5 and info["best_val_epoch"] == 3 and info["best_val_loss"] == pytest.approx(0.8)
    assert len(captured) == 5
    assert captured[-1]["early_stopped"] is True and captured[-1]["final_epoch"] == 5 and captured[-1]["best_val_epoch"] == 3
    assert fits == 9    # the snapshot taken after epoch 3 (3 chunks x 3 epochs) is what gets calibrated, not the epoch-5 state


def test_reference_case_patience_one_immediate_stop():
    trainer, _, fits, _ = _scripted_run([1.0, 0.5, 0.6], 1, 10)
    info = trainer._early_stop_info
    assert (info["stop_reason"], info["final_epoch"], info["best_val_epoch"]) == ("early_stopping", 3, 2)
    assert fits == 6


def test_reference_case_summary_only_on_final_epoch():
    _, captured, _, _ = _scripted_run([1.0, 0.9, 0.95, 0.96], 2, 10)
    for cb in captured[:-1]:
        assert "early_stopped" not in cb and "final_epoch" not in cb
        assert set(cb) == {"epoch", "ref_accuracy", "val_accuracy", "val_loss", "training_loss", "cumulative_seconds"}
    for key in ("early_stopped", "final_epoch", "best_val_epoch", "best_val_loss"):
        assert key in captured[-1]


def test_reference_case_constructor():
    assert MermaidTrainer(batch_size=100).early_stopping_patience is None
    assert MermaidTrainer(batch_size=100, early_stopping_patience=1).early_stopping_patience == 1


# ---- host-side contracts of the estimator mirror (reference: tests/pyspacer/test_mlp_benchmark.py:409-570) -----------------
def test_estimator_batching_and_label_contracts_host_side():
    from mermaid_classifier_b200.torch_classifier import TorchMLPClassifier, split_steps

    auto = TorchMLPClassifier(hidden_layer_sizes=(16,), batch_size="auto", random_state=42)
    assert auto._resolve_batch_size(500) == 200 and auto._resolve_batch_size(50) == 50      # min(200, n_samples)
    assert TorchMLPClassifier(batch_size=128)._resolve_batch_size(50) == 50                  # explicit size is clipped
    # ceil(n / batch) Adam steps per partial_fit call, one step when the input is smaller than the batch
    for n, mb, steps in ((500, 100, 5), (500, 200, 3), (50, 200, 1), (650, 200, 4)):
        pos, offsets = split_steps(n, min(mb, n))
        assert len(offsets) - 1 == steps and offsets[0] == 0 and offsets[-1] == n and np.array_equal(pos, np.arange(n))
    # string labels handed over as an object array (what pyspacer's loaders yield) map onto sorted unique classes_
    y = np.array(["a", "b", "c"] * 166 + ["a", "b"], dtype=object)
    auto.classes_ = np.unique(np.asarray(["c", "a", "b"]))
    assert list(auto.classes_) == ["a", "b", "c"]
    idx = auto._labels_to_indices(y)
    assert idx.tolist() == ([0, 1, 2] * 166 + [0, 1])
    with pytest.raises(ValueError, match="not in classes_"):
        auto._labels_to_indices(np.array(["a", "z"], dtype=object))
    # class weights follow classes_ order; missing or negative weights are rejected (torch_classifier.py:188-214)
    auto.class_weight = {"b": 2.0, "a": 1.0, "c": 0.5}
    assert auto._build_class_weight().tolist() == [1.0, 2.0, 0.5]
    auto.class_weight = {"a": 1.0, "b": 2.0}
    with pytest.raises(ValueError, match="missing weights"):
        auto._build_class_weight()
    auto.class_weight = {"a": 1.0, "b": -2.0, "c": 1.0}
    with pytest.raises(ValueError, match="negative"):
        auto._build_class_weight()
    # the shuffle is re-seeded from random_state inside every call (same state -> same order, call after call)
    assert np.array_equal(auto._seed_rng().permutation(10), auto._seed_rng().permutation(10))
