"""GPU parity tests for the trainer level (A8 and the Platt / export follow-on): device evaluation, the epoch loop
with early stopping, device Platt calibration and the artifact round trip -- against the reference-run fixture
``tests/golden/trainer_eval.npz`` and the CPU oracle (``oracle/trainer.py``)."""
import json

import numpy as np
import pytest
import torch

from mermaid_classifier_b200.export import export_artifact
from mermaid_classifier_b200.inference import DeviceHead, load_predictor, platt_fit_device
from mermaid_classifier_b200.trainer import DeviceLabels, MermaidTrainer, TaskLabels
from oracle import head as ohead
from oracle import trainer as otr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "trainer_eval.npz")


def _params(g):
    n = len([k for k in g.files if k.startswith("W")])
    return [g[f"W{i}"] for i in range(n)], [g[f"b{i}"] for i in range(n)]


class HostLabels:
    """pyspacer-ImageLabels-shaped host split (lists out of ``load_data_in_batches``), as the reference run used."""

    def __init__(self, X, y):
        self.X, self.y = X, y
        self.label_count = len(y)
        self.classes_set = set(y.tolist())

    def __len__(self):
        return self.label_count

    def load_data_in_batches(self, batch_size, random_seed=None):
        order = np.arange(len(self.y)) if random_seed is None else np.random.default_rng(random_seed).permutation(len(self.y))
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            yield self.X[idx].tolist(), self.y[idx].tolist()


def _splits(g, kind):
    cls = g["classes"]
    mk = (lambda X, y: DeviceLabels(X, cls[y])) if kind == "device" else (lambda X, y: HostLabels(X, cls[y]))
    return TaskLabels(train=mk(g["Xt"], g["yt"]), ref=mk(g["Xr"], g["yr"]), val=mk(g["Xv"], g["yv"]))


def test_device_evaluation_matches_oracle_and_reference(g):
    w, b = _params(g)
    head = DeviceHead(w, b, None, None)
    xv, yv = torch.from_numpy(g["Xv"]).cuda(), torch.from_numpy(g["yv"]).cuda()
    hits, loss_sum = head.evaluate_device(xv, yv)
    tw, tb = [torch.from_numpy(x) for x in w], [torch.from_numpy(x) for x in b]
    pv = ohead.softmax_proba(g["Xv"], tw, tb)
    assert hits == int(round(otr.accuracy(g["yv"], pv) * len(g["yv"])))
    # fp32 Linear chain on both sides, different summation order: 1e-6 relative on the summed log-loss
    assert loss_sum == pytest.approx(float(otr.log_loss_terms(g["yv"], pv).sum()), rel=1e-6)
    assert loss_sum / len(g["yv"]) == pytest.approx(float(g["val_loss"][-1]), rel=1e-6)
    assert hits / len(g["yv"]) == pytest.approx(float(g["val_acc"][-1]), abs=1e-12)
    # bit-reproducible, ragged sizes, empty input, and the tensor-core chain agrees to fp32 class
    assert head.evaluate_device(xv, yv) == (hits, loss_sum)
    h1, l1 = head.evaluate_device(xv[:1].contiguous(), yv[:1].contiguous())
    assert l1 == pytest.approx(float(otr.log_loss_terms(g["yv"][:1], pv[:1])[0]), rel=1e-6) and h1 in (0, 1)
    assert head.evaluate_device(xv[:0].contiguous(), yv[:0].contiguous()) == (0, 0.0)
    ht, lt = head.evaluate_device(xv, yv, exact=False)
    assert abs(ht - hits) <= 1 and lt == pytest.approx(loss_sum, rel=1e-5)
    # a target outside [0, K) poisons the loss instead of reading out of bounds
    bad = yv.clone()
    bad[5] = 99
    assert np.isnan(head.evaluate_device(xv, bad)[1])
    with pytest.raises(ValueError):
        head.evaluate_device(xv, yv.long())


def test_log_loss_clip_on_saturated_rows():
    """A confidently wrong row has p_y == 0 in fp32 softmax; sklearn clips it to eps -> -log(eps) = 36.04."""
    w = [np.eye(4, dtype=np.float32) * 200.0]
    b = [np.zeros(4, dtype=np.float32)]
    head = DeviceHead(w, b, None, None)
    x = torch.eye(4, dtype=torch.float32).cuda()
    y = torch.tensor([0, 1, 3, 2], dtype=torch.int32).cuda()
    hits, loss_sum = head.evaluate_device(x, y)
    assert hits == 2
    p = ohead.softmax_proba(np.eye(4, dtype=np.float32), [torch.from_numpy(w[0])], [torch.from_numpy(b[0])])
    want = otr.log_loss_terms(np.array([0, 1, 3, 2]), p).sum()
    assert want > 72 and loss_sum == pytest.approx(float(want), rel=1e-12)


@pytest.mark.parametrize("kind", ["host", "device"])
def test_epoch_loop_reproduces_reference_run(g, kind):
    """Same data, chunking, seeds and hyper-parameters as the reference run behind the fixture: every per-epoch
    number the reference's helpers reported comes back from the GPU loop (host-batch and HBM-resident stores)."""
    seen = []
    trainer = MermaidTrainer(batch_size=int(g["chunk"]), on_epoch_end=seen.append, hidden_layer_sizes=(24, 16),
                             learning_rate_init=1e-3)
    clf_cal, val_results, msg = trainer(_splits(g, kind), int(g["epochs"]), [])
    assert [m["epoch"] for m in seen] == list(range(int(g["epochs"])))
    np.testing.assert_allclose([m["val_loss"] for m in seen], g["val_loss"], rtol=2e-5)
    np.testing.assert_allclose([m["training_loss"] for m in seen], g["train_loss"], rtol=2e-5)
    np.testing.assert_allclose([m["val_accuracy"] for m in seen], g["val_acc"], atol=1.01 / len(g["yv"]))
    np.testing.assert_allclose(msg.ref_accs, g["ref_acc"], atol=1.01 / len(g["yr"]))
    assert "final_epoch" not in seen[0] and seen[-1]["final_epoch"] == int(g["epochs"]) and seen[-1]["early_stopped"] is False
    assert "best_val_epoch" not in seen[-1]
    assert trainer._early_stop_info == {"enabled": False, "patience": None, "stop_reason": "budget_exhausted",
                                        "final_epoch": int(g["epochs"]), "best_val_epoch": None, "best_val_loss": None}
    # calibration of the trained head against the reference run's sklearn calibrators
    a, b = clf_cal.platt
    np.testing.assert_allclose(a, g["platt_a"], rtol=5e-3)
    np.testing.assert_allclose(b, g["platt_b"], rtol=5e-3)
    np.testing.assert_allclose(clf_cal.predict_proba(g["Xv"]), g["cal_proba_val"], atol=2e-4)
    assert list(clf_cal.classes_) == list(g["classes"]) and clf_cal.cv == "prefit"
    # val results of the calibrated model
    est = np.argmax(clf_cal.predict_proba(g["Xv"]), axis=1)
    assert val_results.est == est.tolist() and val_results.gt == g["yv"].tolist()
    assert val_results.classes == list(g["classes"])
    assert msg.acc == pytest.approx(float(np.mean(est == g["yv"])))
    assert len(val_results.scores) == len(est) and msg.pc_accs == []


def test_early_stopping_follows_the_oracle_walk(g):
    """Aggressive learning rate so the val loss turns: the loop must stop and restore exactly where the restated
    bookkeeping says, given the val losses it observed."""
    seen = []
    trainer = MermaidTrainer(batch_size=int(g["chunk"]), on_epoch_end=seen.append, early_stopping_patience=2,
                             hidden_layer_sizes=(24, 16), learning_rate_init=3e-2)
    labels = _splits(g, "device")
    clf_cal, _, msg = trainer(labels, 40, [])
    losses = [m["val_loss"] for m in seen]
    want = otr.early_stopping_walk(losses + [float("inf")] * 40, 40, 2)
    info = dict(trainer._early_stop_info)
    assert info == {k: want[k] for k in info}
    assert len(seen) == want["final_epoch"] and len(msg.ref_accs) == want["final_epoch"]
    assert seen[-1]["early_stopped"] == (want["stop_reason"] == "early_stopping")
    assert seen[-1]["best_val_epoch"] == want["best_val_epoch"] and seen[-1]["best_val_loss"] == want["best_val_loss"]
    # the estimator handed to calibration is the snapshot of the best epoch: its val loss is the best one seen
    _, restored_loss = trainer._calc_acc_and_log_loss_batched(clf_cal.estimator, labels.val, list(g["classes"]))
    assert restored_loss == pytest.approx(want["best_val_loss"], rel=1e-12)
    assert want["stop_reason"] == "early_stopping", losses


def test_platt_fit_matches_oracle_on_reference_probabilities(g):
    proba = torch.from_numpy(g["proba_ref"]).cuda()
    y = torch.from_numpy(g["yr"]).cuda()
    a, b, loss, passes = platt_fit_device(proba, y)
    # the device Newton iterate is at least as good a minimiser as sklearn's L-BFGS-B stopping point ...
    for k in range(proba.shape[1]):
        f, t = g["proba_ref"][:, k], otr.platt_targets((g["yr"] == k).astype(int))[0]
        l_ref = otr.platt_objective(g["platt_a"][k], g["platt_b"][k], f, t)[0]
        l_dev, grad = otr.platt_objective(a[k], b[k], f, t)
        assert l_dev <= l_ref + 1e-9 * abs(l_ref)
        assert loss[k] == pytest.approx(l_dev, rel=1e-12)
        assert np.abs(grad).max() < 1e-6
    # ... and the parameters agree within the tolerance the reference's own calibration-equivalence test uses
    # (tests/pyspacer/test_trainer.py:118-139: rtol 1e-3, atol 5e-3 on a_ / b_)
    np.testing.assert_allclose(a, g["platt_a"], rtol=1e-3, atol=5e-3)
    np.testing.assert_allclose(b, g["platt_b"], rtol=1e-3, atol=5e-3)
    cal = otr.calibrated_proba64(g["proba_ref"], a, b)
    assert np.abs(cal - otr.calibrated_proba64(g["proba_ref"], g["platt_a"], g["platt_b"])).max() < 1e-4
    assert 2 <= passes <= 100
    a2, b2, _, _ = platt_fit_device(proba, y)
    assert np.array_equal(a, a2) and np.array_equal(b, b2)  # fixed-order sums


def test_platt_fit_many_classes_and_empty_classes():
    """K = 300 over 20 000 rows (3 column blocks, ragged last block); some classes have no positive row at all."""
    rng = np.random.default_rng(3)
    n, K = 20000, 300
    logits = rng.standard_normal((n, K)) * 2.0
    y = rng.integers(0, K - 10, size=n)           # the last 10 classes never occur
    logits[np.arange(n), y] += rng.random(n) * 6.0
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    a, b, loss, passes = platt_fit_device(torch.from_numpy(p).cuda(), torch.from_numpy(y.astype(np.int32)).cuda())
    assert np.isfinite(a).all() and np.isfinite(b).all()
    for k in (0, 1, 127, 128, 255, 256, 289, 290, 299):
        y01 = (y == k).astype(int)
        t = otr.platt_targets(y01)[0]
        ra, rb = otr.sigmoid_calibration(p[:, k], y01)
        l_ref = otr.platt_objective(ra, rb, p[:, k], t)[0]
        l_dev, grad = otr.platt_objective(a[k], b[k], p[:, k], t)
        assert l_dev <= l_ref + 1e-9 * abs(l_ref), k
        assert loss[k] == pytest.approx(l_dev, rel=1e-11), k
        assert np.abs(grad).max() < 1e-5, (k, grad)
        q = np.linspace(0, 1, 101)
        assert np.abs(1 / (1 + np.exp(a[k] * q + b[k])) - 1 / (1 + np.exp(ra * q + rb))).max() < 2e-3, k
    with pytest.raises(ValueError):
        platt_fit_device(torch.from_numpy(p.astype(np.float32)).cuda(), torch.from_numpy(y.astype(np.int32)).cuda())


def test_export_round_trip_of_a_gpu_trained_model(g, tmp_path):
    trainer = MermaidTrainer(batch_size=int(g["chunk"]), hidden_layer_sizes=(24, 16), learning_rate_init=1e-3)
    clf_cal, _, _ = trainer(_splits(g, "device"), 3, [])
    path, manifest, diff = export_artifact(clf_cal, tmp_path, g["Xr"][:64])
    assert diff <= 1e-6 and manifest["classes"] == list(g["classes"]) and manifest["input_dim"] == 32
    assert json.loads((tmp_path / "model.json").read_text()) == manifest
    pred = load_predictor(path, tmp_path / "model.json")
    graph = torch.jit.load(str(path), map_location="cpu")
    with torch.no_grad():
        want = graph(torch.from_numpy(g["Xv"])).numpy().astype(np.float64)
    got = pred.predict_proba(g["Xv"])
    assert np.abs(got - want).max() <= 1e-6
    assert np.abs(got - clf_cal.predict_proba(g["Xv"])).max() <= 1e-6
    assert (pred.predict(g["Xv"]) == clf_cal.predict(g["Xv"])).all()


def test_full_size_evaluation_is_additive():
    """Production head (1280 -> 500 -> 300 -> 100 -> 500) over 200 000 rows: hits add up exactly and the log-loss sum
    to fp64 rounding when the rows are evaluated in two ragged pieces; per-row terms agree with the oracle on a
    subsample."""
    from mermaid_classifier_b200 import synth

    w, b, _, _, _ = synth.synth_head(input_dim=1280, hidden=(500, 300, 100), n_classes=500, seed=0)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in b], None, None)
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = 200_000
    X = torch.randn(n, 1280, device="cuda", generator=gen) * 0.25
    y = torch.randint(0, 500, (n,), device="cuda", generator=gen, dtype=torch.int32)
    hits, loss = head.evaluate_device(X, y)
    cut = 77_777
    h1, l1 = head.evaluate_device(X[:cut].contiguous(), y[:cut].contiguous())
    h2, l2 = head.evaluate_device(X[cut:].contiguous(), y[cut:].contiguous())
    assert h1 + h2 == hits
    assert l1 + l2 == pytest.approx(loss, rel=1e-12)
    sub = slice(1000, 1512)
    p = ohead.softmax_proba(X[sub].cpu().numpy(), w, b)
    ysub = y[sub].cpu().numpy()
    hs, ls = head.evaluate_device(X[sub].contiguous(), y[sub].contiguous())
    assert ls == pytest.approx(float(otr.log_loss_terms(ysub, p).sum()), rel=1e-5)
    assert abs(hs - int((p.argmax(1) == ysub).sum())) <= 1


def test_binary_calibration_single_calibrator():
    """K = 2 (reference ``trainer.py:365-374``): one sigmoid calibrator fitted on the positive-class column; calibrated
    probabilities are ``[1 - p, p]`` as scikit-learn's ``_CalibratedClassifier.predict_proba`` builds them."""
    from mermaid_classifier_b200.torch_classifier import TorchMLPClassifier

    rng = np.random.default_rng(5)
    n, d = 3000, 16
    y = rng.integers(0, 2, n)
    X = (rng.standard_normal((n, d)) + 1.2 * (2 * y[:, None] - 1) * (np.arange(d) < 4)).astype(np.float32)
    classes = np.asarray(["neg", "pos"])
    clf = TorchMLPClassifier(hidden_layer_sizes=(8,), learning_rate_init=1e-2, random_state=0)
    for _ in range(3):
        clf.partial_fit(X[:2000], classes[y[:2000]], classes=classes)
    trainer = MermaidTrainer(batch_size=700)
    cal = trainer._calibrate_in_batches(clf, DeviceLabels(X[2000:], classes[y[2000:]]))
    inner = cal.calibrated_classifiers_[0]
    assert cal.binary and len(inner.calibrators) == 1 and list(cal.classes_) == ["neg", "pos"]
    p1 = clf.predict_proba(X[2000:])[:, 1]
    ra, rb = otr.sigmoid_calibration(p1, y[2000:])
    t = otr.platt_targets(y[2000:])[0]
    l_ref = otr.platt_objective(ra, rb, p1, t)[0]
    l_dev, grad = otr.platt_objective(inner.calibrators[0].a_, inner.calibrators[0].b_, p1, t)
    assert l_dev <= l_ref + 1e-9 * abs(l_ref) and np.abs(grad).max() < 1e-5
    proba = cal.predict_proba(X[2000:])
    want1 = 1.0 / (1.0 + np.exp(ra * p1 + rb))
    assert proba.shape == (1000, 2) and np.abs(proba[:, 1] - want1).max() < 2e-3
    assert np.abs(proba.sum(1) - 1.0).max() < 1e-12
    assert (cal.predict(X[2000:]) == classes[np.argmax(proba, 1)]).all()
    gts, ests, scores = __import__("mermaid_classifier_b200.trainer", fromlist=["evaluate_classifier"]).evaluate_classifier(
        cal, DeviceLabels(X[2000:], classes[y[2000:]]), batch_size=400)
    assert len(ests) == 1000 and np.mean(np.asarray(gts) == np.asarray(ests)) > 0.8 and max(scores) <= 1.0
    with pytest.raises(ValueError):
        cal.head()
