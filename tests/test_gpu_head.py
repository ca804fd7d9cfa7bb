"""GPU parity tests for head scoring through the C ABI."""
import numpy as np
import pytest
import torch

from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.inference import DeviceHead, ManifestError, load_predictor
from oracle import head as ohead

pytestmark = pytest.mark.gpu


def test_load_predictor_golden_small(golden_dir):
    pred = load_predictor(golden_dir / "head_small" / "model.pt", golden_dir / "head_small" / "model.json")
    io = np.load(golden_dir / "head_small_io.npz")
    got = pred.predict_proba(io["X"])
    assert got.dtype == np.float64 and got.shape == io["proba"].shape
    assert np.max(np.abs(got - io["proba"])) <= 1e-6  # the reference's own export tolerance (export.py:32)
    assert np.array_equal(pred.predict_indices(io["X"]), io["proba"].argmax(1))
    assert pred.classes_ is pred.classes and pred.input_dim == 32
    labels, scores = pred.predict_topk(io["X"], 3)
    want = ohead.topk_labels(io["proba"], 3)
    assert np.array_equal(labels, np.asarray(pred.classes, dtype=object)[want])
    with pytest.raises(ValueError):
        pred.predict_proba(np.zeros((3, 31), dtype=np.float32))
    assert pred.predict_proba(np.zeros((0, 32), dtype=np.float32)).shape == (0, 12)


def test_manifest_errors(golden_dir, tmp_path):
    import json

    m = json.loads((golden_dir / "head_small" / "model.json").read_text())
    for patch in ({"schema_version": 2}, {"classes": m["classes"][:-1]}, {"input_dim": 31}):
        bad = dict(m, **patch)
        (tmp_path / "model.json").write_text(json.dumps(bad))
        with pytest.raises(ManifestError):
            load_predictor(golden_dir / "head_small" / "model.pt", tmp_path / "model.json")


@pytest.mark.parametrize("tag,hidden", [("h200_100", (200, 100)), ("h500_300_100", (500, 300, 100))])
def test_full_size_head_vs_golden_and_oracle(golden_dir, tag, hidden):
    g = np.load(golden_dir / f"head_full_{tag}.npz")
    w, bb, a, b, classes = synth.synth_head(1280, hidden, 500, seed=0)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
    X = synth.synth_features(512, 1280, seed=5).numpy()
    proba, labels = head.scores_host(X)
    assert np.max(np.abs(proba[g["rows"]] - g["proba_rows"])) <= 1e-6
    assert (labels == g["labels"]).mean() >= 0.999
    # larger sample against the oracle: >= 99.9 % top-1 agreement
    X2 = synth.synth_features(20000, 1280, seed=9).numpy()
    want = ohead.calibrated_proba(X2, w, bb, a, b)
    out = head.scores_device(torch.from_numpy(X2).cuda(), want_proba=True, topk=2)
    lab = out["labels"].cpu().numpy()
    assert (lab == want.argmax(1)).mean() >= 0.999
    assert np.max(np.abs(out["proba"].cpu().numpy() - want)) <= 1e-6
    assert np.array_equal(out["topk_idx"][:, 0].cpu().numpy(), lab)
    assert head.launches > 0


def test_softmax_path_matches_torch_classifier_semantics():
    w, bb, _, _, _ = synth.synth_head(64, (32,), 10, seed=3)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb])
    X = synth.synth_features(300, 64, seed=1).numpy()
    proba, labels = head.scores_host(X)
    want = ohead.softmax_proba(X, w, bb)
    assert np.max(np.abs(proba - want)) <= 1e-6
    assert np.abs(proba.sum(1) - 1).max() < 1e-12
    assert np.array_equal(labels, want.argmax(1))


def test_unpadded_dims():
    """Layer widths that are not multiples of 4 go through the padded path."""
    w, bb, a, b, _ = synth.synth_head(30, (13, 7), 5, seed=2)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
    X = synth.synth_features(100, 30, seed=4).numpy()
    proba, labels = head.scores_host(X)
    want = ohead.calibrated_proba(X, w, bb, a, b)
    assert np.max(np.abs(proba - want)) <= 1e-6
    assert np.array_equal(labels, want.argmax(1))


@pytest.mark.parametrize("hidden", [(200, 100), (500, 300, 100)])
def test_tensor_core_chain_matches_exact_chain(hidden):
    """Device scoring runs the Linear chain on tcgen05 (3xTF32); it must agree with the exact-fp32 chain and the oracle:
    proba within 1e-5, >= 99.9 % identical labels; ragged row counts exercise the TMA row bound."""
    w, bb, a, b, _ = synth.synth_head(1280, hidden, 500, seed=0)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
    for n in (1, 127, 129, 5000):
        X = synth.synth_features(n, 1280, seed=21 + n)
        xd = X.cuda()
        fast = head.scores_device(xd, want_proba=True, topk=3, exact=False)
        exact = head.scores_device(xd, want_proba=True, topk=3, exact=True)
        want = ohead.calibrated_proba(X.numpy(), w, bb, a, b)
        assert torch.max(torch.abs(fast["proba"] - exact["proba"])).item() <= 1e-5
        assert np.max(np.abs(fast["proba"].cpu().numpy() - want)) <= 1e-5
        assert (fast["labels"].cpu().numpy() == want.argmax(1)).mean() >= 0.999
        assert (fast["labels"] == exact["labels"]).float().mean().item() >= 0.999


def test_saturated_rows_and_steep_platt_slopes():
    """Confident rows (softmax p -> 1) with steep Platt slopes |a| ~ 10-20: the exact chain must hold the reference's 1e-6
    export gate there too (the library is built without --use_fast_math: expf / division in the head are IEEE)."""
    rng = np.random.default_rng(7)
    K, D = 40, 64
    w = [torch.from_numpy(rng.normal(0, 0.6, (K, D)).astype(np.float32))]
    bb = [torch.from_numpy(rng.normal(0, 0.1, (K,)).astype(np.float32))]
    a = torch.from_numpy(-rng.uniform(10, 20, K).astype(np.float32))
    b = torch.from_numpy(rng.uniform(0, 6, K).astype(np.float32))
    X = rng.normal(0, 1, (4000, D)).astype(np.float32)
    X[::3] *= 6.0   # a third of the rows saturate the softmax
    want = ohead.calibrated_proba(X, w, bb, a, b)
    assert (ohead.softmax_proba(X, w, bb).max(1) > 0.999).mean() > 0.2
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
    proba, labels = head.scores_host(X)
    assert np.max(np.abs(proba - want)) <= 1e-6
    assert np.abs(proba.sum(1) - 1).max() <= 1e-5
    assert (labels == want.argmax(1)).mean() >= 0.999
