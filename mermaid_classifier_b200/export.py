"""``export_artifact``: freeze a GPU-trained, calibrated head into the portable ``model.pt`` + ``model.json``.

Mirror of ``mermaid_classifier/pyspacer/inference/export.py:25-94`` and ``inference/head.py:25-123`` of the
reference: same signature and return triple, same parity gate (frozen TorchScript graph vs
``model.predict_proba`` on a representative batch, ``tol`` 1e-6, :class:`ParityError` beyond it), same manifest
fields.  The graph is the serving contract shared with the reference's ``load_predictor``
(``inference/loader.py:38-75``) -- module layout ``linears.{i}`` plus buffers ``a`` / ``b`` -- so an artifact
written here loads on either side.  ``model.predict_proba`` is the device path (:class:`trainer.CalibratedClassifier`).

No scikit-learn pin is enforced: nothing in this build's calibration depends on sklearn.
"""

# NB: no ``from __future__ import annotations`` -- TorchScript resolves class annotations at script time.
import json
from importlib import metadata
from pathlib import Path
from typing import Any, List

import numpy as np
import torch
from torch import nn

from .inference import SCHEMA_VERSION, TASK_NAME, ParityError


class PortableHead(nn.Module):
    """Linear/ReLU chain -> softmax -> per-class Platt sigmoid -> row normalise -> overshoot clip."""

    n_classes: int

    def __init__(self, weights: List[torch.Tensor], biases: List[torch.Tensor], a: torch.Tensor, b: torch.Tensor):
        super().__init__()
        if len(weights) == 0 or len(weights) != len(biases):
            raise ValueError(f"need one bias per weight and at least one layer; got {len(weights)} / {len(biases)}")
        if a.ndim != 1 or a.shape != b.shape or a.shape[0] != weights[-1].shape[0]:
            raise ValueError(f"a and b must be 1-D with one entry per class; got {tuple(a.shape)} / {tuple(b.shape)}")
        layers = []
        for w, bias in zip(weights, biases):
            lin = nn.Linear(int(w.shape[1]), int(w.shape[0]))
            with torch.no_grad():
                lin.weight.copy_(w)
                lin.bias.copy_(bias)
            layers.append(lin)
        self.linears = nn.ModuleList(layers)
        self.register_buffer("a", a.detach().clone().float())
        self.register_buffer("b", b.detach().clone().float())
        self.n_classes = int(a.shape[0])

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        x = features
        last = len(self.linears) - 1
        for i, lin in enumerate(self.linears):
            x = lin(x)
            if i < last:
                x = torch.relu(x)
        p = torch.softmax(x, dim=1)
        c = torch.sigmoid(-(self.a * p + self.b))
        total = c.sum(dim=1, keepdim=True)
        ok = total != 0
        proba = torch.where(ok, c / torch.where(ok, total, torch.ones_like(total)),
                            torch.full_like(c, 1.0 / float(self.n_classes)))
        return torch.where((proba > 1.0) & (proba <= 1.0 + 1e-5), torch.ones_like(proba), proba)


def build_calibrated_head(model: Any) -> PortableHead:
    """From a calibrated model with the ``CalibratedClassifierCV(cv="prefit")`` attribute surface."""
    calibrated = model.calibrated_classifiers_
    if len(calibrated) != 1:
        raise ValueError(f"Expected exactly one calibrated classifier (cv='prefit'), got {len(calibrated)}.")
    inner = calibrated[0]
    estimator, calibrators = inner.estimator, inner.calibrators
    if not np.array_equal(estimator.classes_, model.classes_):
        raise ValueError("estimator.classes_ does not match model.classes_")
    k = len(model.classes_)
    if k <= 2:
        raise ValueError(f"the portable head only supports the multiclass (K > 2) path; got K={k}.")
    if len(calibrators) != k:
        raise ValueError(f"Expected {k} per-class calibrators, got {len(calibrators)}.")
    module = estimator._module
    return PortableHead([lin.weight.detach().clone().float() for lin in module.linears],
                        [lin.bias.detach().clone().float() for lin in module.linears],
                        torch.tensor([float(c.a_) for c in calibrators], dtype=torch.float32),
                        torch.tensor([float(c.b_) for c in calibrators], dtype=torch.float32))


def _version(pkg: str):
    try:
        return metadata.version(pkg)
    except metadata.PackageNotFoundError:
        return "not-installed"   # explicit provenance: the artifact was produced without this package


def export_artifact(model: Any, output_dir, reference_features: Any, *, config=None, task: str = TASK_NAME,
                    tol: float = 1e-6, enforce_sklearn_pin: bool = True):
    """Returns ``(model_pt_path, manifest, max_abs_diff)``; raises :class:`ParityError` when the frozen graph and
    ``model.predict_proba`` disagree beyond ``tol`` on ``reference_features``.

    ``enforce_sklearn_pin`` is accepted for signature compatibility with the reference (``inference/export.py:32,41-49``,
    where it guards the private sklearn calibrator API).  The B200 trainer fits its Platt calibrators itself
    (``mc_platt_fit``) and reads none of scikit-learn's internals, so there is nothing to pin: the flag is a documented
    no-op here, and the manifest records the scikit-learn version as provenance when the package is installed."""
    del enforce_sklearn_pin
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    frozen = torch.jit.freeze(torch.jit.script(build_calibrated_head(model).eval()))
    ref = np.asarray(reference_features, dtype=np.float32)
    expected = model.predict_proba(ref)
    with torch.no_grad():
        got = frozen(torch.from_numpy(ref)).numpy().astype(np.float64)
    max_diff = float(np.max(np.abs(expected - got)))
    if not max_diff <= tol:
        raise ParityError(f"Frozen graph diverges from source model: max|d|={max_diff:.3e} exceeds tol={tol:.3e}."
                          " Refusing to ship.")
    estimator = model.calibrated_classifiers_[0].estimator
    manifest = {
        "schema_version": SCHEMA_VERSION,
        "task": task,
        "classes": [c.item() if hasattr(c, "item") else c for c in model.classes_],
        "input_dim": int(estimator.n_features_in_),
        "config": config if config is not None else {"patch_size": 224},
        "trained_with": {"torch": torch.__version__, "sklearn": _version("scikit-learn"), "pyspacer": _version("pyspacer"),
                         "trainer": "mermaid_classifier_b200"},
    }
    model_pt = output_dir / "model.pt"
    torch.jit.save(frozen, str(model_pt))
    (output_dir / "model.json").write_text(json.dumps(manifest, indent=2))
    return model_pt, manifest, max_diff


def write_head_artifact(output_dir, weights, biases, a, b, classes, *, config=None, task: str = TASK_NAME):
    """``model.pt`` + ``model.json`` from explicit head parameters (Linear weights / biases, per-class Platt ``a`` / ``b``):
    the same frozen TorchScript graph and manifest fields :func:`export_artifact` writes (``inference/export.py:54-57,71-92``
    in the reference), without a calibrated estimator object to read them from.  Returns ``(model_pt, model_json)``."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    ws = [torch.as_tensor(np.asarray(w), dtype=torch.float32) for w in weights]
    bs = [torch.as_tensor(np.asarray(x), dtype=torch.float32) for x in biases]
    head = PortableHead(ws, bs, torch.as_tensor(np.asarray(a), dtype=torch.float32), torch.as_tensor(np.asarray(b), dtype=torch.float32))
    frozen = torch.jit.freeze(torch.jit.script(head.eval()))
    manifest = {
        "schema_version": SCHEMA_VERSION,
        "task": task,
        "classes": [c.item() if hasattr(c, "item") else c for c in classes],
        "input_dim": int(ws[0].shape[1]),
        "config": config if config is not None else {"patch_size": 224},
        "trained_with": {"torch": torch.__version__, "sklearn": _version("scikit-learn"), "pyspacer": _version("pyspacer"),
                         "trainer": "mermaid_classifier_b200"},
    }
    model_pt, model_json = output_dir / "model.pt", output_dir / "model.json"
    torch.jit.save(frozen, str(model_pt))
    model_json.write_text(json.dumps(manifest, indent=2))
    return model_pt, model_json
