"""Generate the golden fixtures in this directory by IMPORTING THE REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Reference code executed (unmodified, from ``/root/reference``):

* ``mermaid_classifier.pyspacer.inference.head.CalibratedHead``   (head.py:25-89)
* ``mermaid_classifier.pyspacer.inference.loader.load_predictor`` (loader.py:38-75)
* ``mermaid_classifier.pyspacer.torch_classifier.TorchMLPClassifier`` (torch_classifier.py:83-444)

``export_artifact`` itself cannot run here (it needs ``pyspacer`` metadata and sklearn 1.5.2,
export.py:41-49,86); the artifact is therefore written with the same three calls it makes
(``jit.script`` -> ``jit.freeze`` -> ``jit.save``, export.py:54-57,90-92) and a manifest with the
fields of export.py:71-88.
"""

import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

from mermaid_classifier.pyspacer.inference import SCHEMA_VERSION, TASK_NAME  # noqa: E402
from mermaid_classifier.pyspacer.inference.head import CalibratedHead  # noqa: E402
from mermaid_classifier.pyspacer.inference.loader import load_predictor  # noqa: E402
from mermaid_classifier.pyspacer.torch_classifier import TorchMLPClassifier  # noqa: E402

from mermaid_classifier_b200 import synth  # noqa: E402


def write_artifact(out_dir: Path, weights, biases, a, b, classes, input_dim):
    out_dir.mkdir(parents=True, exist_ok=True)
    head = CalibratedHead(weights, biases, a, b).eval()
    frozen = torch.jit.freeze(torch.jit.script(head))
    torch.jit.save(frozen, str(out_dir / "model.pt"))
    manifest = {
        "schema_version": SCHEMA_VERSION,
        "task": TASK_NAME,
        "classes": classes,
        "input_dim": int(input_dim),
        "config": {"patch_size": 224},
        "trained_with": {"torch": torch.__version__, "sklearn": "1.5.2", "pyspacer": "0.14.0"},
    }
    (out_dir / "model.json").write_text(json.dumps(manifest, indent=2))
    return load_predictor(out_dir / "model.pt", out_dir / "model.json")


def golden_head_small():
    w, bb, a, b, classes = synth.synth_head(input_dim=32, hidden=(24, 16), n_classes=12, seed=7)
    classes = [f"ba{i:02d}::gf{i:02d}" for i in range(12)]
    pred = write_artifact(HERE / "head_small", w, bb, a, b, classes, 32)
    X = synth.synth_features(96, 32, seed=11).numpy()
    X[0] = 0.0
    X[1] = 50.0  # saturating row
    proba = pred.predict_proba(X)
    np.savez_compressed(HERE / "head_small_io.npz", X=X, proba=proba)
    print("head_small", proba.shape, proba.sum(1)[:3])


def golden_head_full():
    """Full-size BASELINE head (1280 -> 200 -> 100 -> 500).  Weights are regenerated from the
    seed by ``synth.synth_head`` in the tests; only a few rows of reference output are stored."""
    for tag, hidden in (("h200_100", (200, 100)), ("h500_300_100", (500, 300, 100))):
        w, bb, a, b, classes = synth.synth_head(1280, hidden, 500, seed=0)
        head = CalibratedHead(w, bb, a, b).eval()
        frozen = torch.jit.freeze(torch.jit.script(head))
        X = synth.synth_features(512, 1280, seed=5)
        with torch.no_grad():
            proba = frozen(X).numpy().astype(np.float64)
        np.savez_compressed(
            HERE / f"head_full_{tag}.npz",
            rows=np.arange(0, 512, 32),
            proba_rows=proba[::32],
            labels=proba.argmax(1).astype(np.int32),
            top1=proba.max(1),
        )
        print("head_full", tag, proba.shape, np.bincount(proba.argmax(1)).max())


def cluster_data(n, n_features, n_classes, seed):
    rng = np.random.RandomState(seed)
    centers = rng.randn(n_classes, n_features) * 3.0
    y = rng.randint(0, n_classes, size=n)
    X = (centers[y] + rng.randn(n, n_features) * 1.3).astype(np.float32)
    labels = np.array([f"class_{i:03d}" for i in range(n_classes)])
    return X, labels[y], labels


def golden_mlp_train():
    X, y, labels = cluster_data(650, 32, 5, 42)  # 650 = 3 x 200 + ragged 50
    out = {"X": X, "y_idx": np.searchsorted(labels, y).astype(np.int32)}
    for tag, cw in (("plain", None), ("weighted", {c: 0.5 + 0.5 * i for i, c in enumerate(labels)})):
        clf = TorchMLPClassifier(hidden_layer_sizes=(16, 8), learning_rate_init=1e-3, random_state=0, class_weight=cw)
        w0 = None
        for _ in range(3):
            clf.partial_fit(X, y, classes=labels.tolist())
            if w0 is None:
                pass
        out[f"{tag}_loss_curve"] = np.asarray(clf.loss_curve_, dtype=np.float64)
        for i, lin in enumerate(clf._module.linears):
            out[f"{tag}_W{i}"] = lin.weight.detach().numpy().copy()
            out[f"{tag}_b{i}"] = lin.bias.detach().numpy().copy()
        out[f"{tag}_proba"] = clf.predict_proba(X[:16])
        out[f"{tag}_pred"] = np.searchsorted(labels, clf.predict(X)).astype(np.int32)
        print("mlp", tag, clf.loss_curve_, clf.n_iter_)
    # init-only snapshot (weights right after _init_module) for the init parity check
    clf = TorchMLPClassifier(hidden_layer_sizes=(16, 8), random_state=0)
    clf.classes_ = labels
    clf.n_features_in_ = 32
    clf._init_module()
    for i, lin in enumerate(clf._module.linears):
        out[f"init_W{i}"] = lin.weight.detach().numpy().copy()
    np.savez_compressed(HERE / "mlp_train.npz", **out)


if __name__ == "__main__":
    golden_head_small()
    golden_head_full()
    golden_mlp_train()
