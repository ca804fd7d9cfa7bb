"""Cross-compatibility with the reference's own serving code, run only where the reference checkout exists (the build
container; ``/root/reference`` is absent on the GPU box, where these tests skip).  No GPU needed: the model handed to
``export_artifact`` is a stand-in whose ``predict_proba`` is the CPU oracle.

* an artifact written by ``mermaid_classifier_b200.export.export_artifact`` loads with the REFERENCE's ``load_predictor``
  (``mermaid_classifier/pyspacer/inference/loader.py:38-75``) and scores identically;
* the calibrated-model surface returned by the B200 trainer is what the reference's ``build_calibrated_head``
  (``inference/head.py:92-123``) reads, and the head it builds is bit-identical to ours.
"""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

REFERENCE = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REFERENCE / "mermaid_classifier").is_dir(), reason="reference checkout not present")

from mermaid_classifier_b200 import synth  # noqa: E402
from mermaid_classifier_b200.export import build_calibrated_head, export_artifact  # noqa: E402
from mermaid_classifier_b200.trainer import SigmoidCalibrator, _CalibratedInner  # noqa: E402
from oracle import head as ohead  # noqa: E402


@pytest.fixture(scope="module")
def reference():
    sys.path.insert(0, str(REFERENCE))
    try:
        from mermaid_classifier.pyspacer.inference import head as ref_head
        from mermaid_classifier.pyspacer.inference import loader as ref_loader
        yield SimpleNamespace(head=ref_head, loader=ref_loader)
    finally:
        sys.path.remove(str(REFERENCE))


def _model(n_classes=12, input_dim=32, hidden=(24, 16)):
    w, bb, a, b, _ = synth.synth_head(input_dim=input_dim, hidden=hidden, n_classes=n_classes, seed=7)
    classes = np.asarray([f"ba{i:02d}::gf{i:02d}" for i in range(n_classes)])
    linears = []
    for x, y in zip(w, bb):
        lin = torch.nn.Linear(x.shape[1], x.shape[0])
        with torch.no_grad():
            lin.weight.copy_(x)
            lin.bias.copy_(y)
        linears.append(lin)
    est = SimpleNamespace(classes_=classes, n_features_in_=input_dim, _module=SimpleNamespace(linears=linears))
    inner = _CalibratedInner(est, [SigmoidCalibrator(float(x), float(y)) for x, y in zip(a, b)], classes)
    model = SimpleNamespace(calibrated_classifiers_=[inner], classes_=classes,
                            predict_proba=lambda X: ohead.calibrated_proba(X, w, bb, a, b))
    return model, (w, bb, a, b)


def test_exported_artifact_loads_with_the_reference_loader(reference, tmp_path):
    model, (w, bb, a, b) = _model()
    x = np.random.default_rng(2).standard_normal((40, 32)).astype(np.float32)
    path, manifest, _ = export_artifact(model, tmp_path, x[:16])
    pred = reference.loader.load_predictor(path, tmp_path / "model.json")
    assert pred.classes == manifest["classes"] and pred.input_dim == 32 and pred.classes_ == pred.classes
    got = pred.predict_proba(x)
    assert got.dtype == np.float64 and got.shape == (40, 12)
    assert np.array_equal(got, ohead.calibrated_proba(x, w, bb, a, b))
    with pytest.raises(ValueError):
        pred.predict_proba(x[:, :5])


def test_reference_head_builder_accepts_the_trainer_output(reference):
    model, _ = _model()
    ref_head = reference.head.build_calibrated_head(model).eval()
    ours = build_calibrated_head(model).eval()
    x = torch.from_numpy(np.random.default_rng(3).standard_normal((25, 32)).astype(np.float32))
    with torch.no_grad():
        assert torch.equal(ref_head(x), ours(x))
    assert [tuple(l.weight.shape) for l in ref_head.linears] == [tuple(l.weight.shape) for l in ours.linears]
    assert torch.equal(ref_head.a, ours.a) and torch.equal(ref_head.b, ours.b)


def test_pickle_state_is_the_reference_layout(reference):
    """``TorchMLPClassifier.__getstate__`` entries (reference ``torch_classifier.py:411-420``): the state a REFERENCE
    classifier pickles converts to the flat device buffers and back without loss, and the converted state restores
    into a fresh reference instance that predicts identically."""
    sys.path.insert(0, str(REFERENCE))
    try:
        from mermaid_classifier.pyspacer.torch_classifier import TorchMLPClassifier as RefClf
    finally:
        sys.path.remove(str(REFERENCE))
    from mermaid_classifier_b200.torch_classifier import pack_reference_state, unpack_reference_state

    rng = np.random.default_rng(11)
    X = rng.standard_normal((300, 20)).astype(np.float32)
    y = np.asarray([f"c{i % 5}" for i in range(300)])
    ref = RefClf(hidden_layer_sizes=(12, 8), learning_rate_init=1e-3, random_state=0)
    ref.partial_fit(X, y, classes=np.unique(y))
    ref.partial_fit(X, y)
    state = ref.__getstate__()
    ws, bs, adam = unpack_reference_state(state["_module_state"], state["_optimizer_state"])
    assert [w.shape for w in ws] == [(12, 20), (8, 12), (5, 8)] and adam[4] == 4   # 2 x ceil(300 / 200) Adam steps
    mod2, opt2 = pack_reference_state(ws, bs, *adam, lr=ref.learning_rate_init, betas=(ref.beta_1, ref.beta_2), eps=ref.epsilon)
    assert list(mod2) == list(state["_module_state"])
    for k in mod2:
        assert torch.equal(mod2[k], state["_module_state"][k])
    ref_opt = state["_optimizer_state"]
    assert opt2["param_groups"] == ref_opt["param_groups"]
    assert sorted(opt2["state"]) == sorted(ref_opt["state"])
    for j in opt2["state"]:
        for key in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(torch.as_tensor(opt2["state"][j][key]).float(), torch.as_tensor(ref_opt["state"][j][key]).float())
    state2 = dict(state, _module_state=mod2, _optimizer_state=opt2)
    clone = RefClf.__new__(RefClf)
    clone.__setstate__(state2)
    assert np.array_equal(clone.predict_proba(X[:50]), ref.predict_proba(X[:50]))
    clone.partial_fit(X, y)
    ref.partial_fit(X, y)
    assert np.array_equal(clone.predict_proba(X[:50]), ref.predict_proba(X[:50]))   # resumed training continues identically
    # the round-1 flat layout still loads
    ws1, bs1, adam1 = unpack_reference_state({"weights": ws, "biases": bs},
                                             {"m_w": adam[0], "m_b": adam[1], "v_w": adam[2], "v_b": adam[3], "t": adam[4]})
    assert adam1[4] == 4 and np.array_equal(ws1[0], ws[0])
