"""Generate ``trainer_eval.npz`` by running the REFERENCE's trainer helpers (build container only).

    python tests/golden/make_golden_trainer.py

Reference code executed unmodified from ``/root/reference``:

* ``MermaidTrainer._calc_acc_batched`` / ``_calc_acc_and_log_loss_batched`` (trainer.py:295-342)
* ``MermaidTrainer._calibrate_in_batches`` (trainer.py:344-396) when the installed scikit-learn still accepts
  its private ``_fit_calibrator`` call; otherwise the same per-class ``_SigmoidCalibration().fit(p_k, y == k)``
  that call performs (recorded in the fixture as ``calib_source``)
* ``TorchMLPClassifier`` (torch_classifier.py) as the estimator driven through the epoch loop of
  trainer.py:138-165 (chunked ``partial_fit``, then ref accuracy, then val accuracy + log-loss)

``spacer`` (pyspacer) is not installed, so its four imports in trainer.py:20-23 are satisfied by empty stand-in
modules -- the helper methods above never touch them.  The label sets are a minimal stand-in for
``ImageLabels.load_data_in_batches`` (contiguous chunks in stored order; epoch-seeded permutation for training).
"""

import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

for name, attrs in {
    "spacer": [],
    "spacer.data_classes": ["ImageLabels", "ValResults"],
    "spacer.messages": ["TrainClassifierReturnMsg", "TrainingTaskLabels"],
    "spacer.train_classifier": ["ClassifierTrainer"],
    "spacer.train_utils": ["evaluate_classifier"],
}.items():
    mod = types.ModuleType(name)
    for a in attrs:
        setattr(mod, a, type(a, (), {}))
    sys.modules[name] = mod

from mermaid_classifier.pyspacer.torch_classifier import TorchMLPClassifier  # noqa: E402
from mermaid_classifier.pyspacer.trainer import MermaidTrainer  # noqa: E402


class Labels:
    """Stand-in for spacer ``ImageLabels``: only ``load_data_in_batches`` is used by the helpers."""

    def __init__(self, X, y):
        self.X, self.y = X, y

    def load_data_in_batches(self, batch_size, random_seed=None):
        order = np.arange(len(self.y))
        if random_seed is not None:
            order = np.random.default_rng(random_seed).permutation(len(self.y))
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            yield self.X[idx].tolist(), self.y[idx].tolist()


def cluster_data(n, n_features, n_classes, seed, centers=None):
    rng = np.random.RandomState(seed)
    if centers is None:
        centers = rng.randn(n_classes, n_features) * 1.1
    y = rng.randint(0, n_classes, size=n)
    X = (centers[y] + rng.randn(n, n_features) * 1.3).astype(np.float32)
    labels = np.array([f"class_{i:03d}" for i in range(n_classes)])
    return X, labels[y], labels, centers


def main():
    nf, K, chunk, epochs = 32, 6, 300, 5
    Xt, yt, classes, centers = cluster_data(1000, nf, K, 42)
    Xr, yr, _, _ = cluster_data(400, nf, K, 43, centers)
    Xv, yv, _, _ = cluster_data(350, nf, K, 44, centers)
    train, ref, val = Labels(Xt, yt), Labels(Xr, yr), Labels(Xv, yv)
    trainer = MermaidTrainer(batch_size=chunk)
    clf = TorchMLPClassifier(hidden_layer_sizes=(24, 16), learning_rate_init=1e-3, random_state=0)
    ref_acc, val_acc, val_loss, train_loss = [], [], [], []
    for epoch in range(epochs):
        for x, y in train.load_data_in_batches(batch_size=chunk, random_seed=epoch):
            clf.partial_fit(x, y, classes=list(classes))
        ref_acc.append(trainer._calc_acc_batched(clf, ref))
        a, l = trainer._calc_acc_and_log_loss_batched(clf, val, list(classes))
        val_acc.append(a)
        val_loss.append(l)
        train_loss.append(clf.loss_curve_[-1])
    proba_ref = clf.predict_proba(Xr)
    try:
        cal = trainer._calibrate_in_batches(clf, ref)
        calibs = cal.calibrated_classifiers_[0].calibrators
        source = "MermaidTrainer._calibrate_in_batches"
        cal_proba_val = cal.predict_proba(Xv)
    except Exception as exc:  # sklearn newer than the reference pin: private signature moved
        from sklearn.calibration import _SigmoidCalibration

        print("calibrate_in_batches unavailable under this sklearn:", type(exc).__name__, exc)
        calibs = [_SigmoidCalibration().fit(proba_ref[:, k], (yr == classes[k]).astype(int)) for k in range(K)]
        source = "sklearn._SigmoidCalibration per class"
        pv = clf.predict_proba(Xv)
        c = np.stack([calibs[k].predict(pv[:, k]) for k in range(K)], axis=1)
        cal_proba_val = c / c.sum(axis=1, keepdims=True)
    out = dict(
        Xt=Xt, yt=np.searchsorted(classes, yt).astype(np.int32), Xr=Xr, yr=np.searchsorted(classes, yr).astype(np.int32),
        Xv=Xv, yv=np.searchsorted(classes, yv).astype(np.int32), classes=classes, chunk=chunk, epochs=epochs,
        ref_acc=np.asarray(ref_acc), val_acc=np.asarray(val_acc), val_loss=np.asarray(val_loss),
        train_loss=np.asarray(train_loss), proba_ref=proba_ref,
        platt_a=np.asarray([float(c.a_) for c in calibs]), platt_b=np.asarray([float(c.b_) for c in calibs]),
        cal_proba_val=cal_proba_val, calib_source=np.asarray(source),
    )
    for i, lin in enumerate(clf._module.linears):
        out[f"W{i}"] = lin.weight.detach().numpy()
        out[f"b{i}"] = lin.bias.detach().numpy()
    np.savez_compressed(HERE / "trainer_eval.npz", **out)
    print("ref_acc", ref_acc, "\nval_acc", val_acc, "\nval_loss", val_loss, "\nsource", source)
    print("a", out["platt_a"], "\nb", out["platt_b"])


if __name__ == "__main__":
    main()
