"""GPU parity tests for the MLP-head training inner loop (mc_mlp_* through the C ABI) against
goldens produced by the reference's own TorchMLPClassifier and against the CPU oracle."""
import ctypes as C
import pickle

import numpy as np
import pytest
import torch

from mermaid_classifier_b200 import _lib
from mermaid_classifier_b200.torch_classifier import TorchMLPClassifier
from oracle import head as ohead

pytestmark = pytest.mark.gpu

LABELS5 = np.array([f"class_{i:03d}" for i in range(5)])


def cluster_data(n, n_features, n_classes, seed):
    rng = np.random.RandomState(seed)
    centers = rng.randn(n_classes, n_features) * 3.0
    y = rng.randint(0, n_classes, size=n)
    X = (centers[y] + rng.randn(n, n_features) * 1.3).astype(np.float32)
    return X, y


@pytest.mark.parametrize("tag", ["plain", "weighted"])
def test_partial_fit_matches_reference_golden(golden_dir, tag):
    """650 rows = 3 full mini-batches + ragged 50; K = 5 classes (padded to 8 on the device)."""
    g = np.load(golden_dir / "mlp_train.npz")
    cw = None if tag == "plain" else {c: 0.5 + 0.5 * i for i, c in enumerate(LABELS5)}
    clf = TorchMLPClassifier(hidden_layer_sizes=(16, 8), learning_rate_init=1e-3, random_state=0, class_weight=cw)
    y = LABELS5[g["y_idx"]]
    for _ in range(3):
        clf.partial_fit(g["X"], y, classes=LABELS5.tolist())
    assert clf.n_iter_ == 3 and clf.n_steps_ == 12  # ceil(650 / 200) Adam steps per pass
    assert np.array_equal(clf.classes_, LABELS5)
    # fp32 summation order differs from the CPU: tolerance 1e-4 relative on the loss, 1e-4 abs on weights
    np.testing.assert_allclose(clf.loss_curve_, g[f"{tag}_loss_curve"], rtol=1e-4)
    for i, lin in enumerate(clf._module.linears):
        np.testing.assert_allclose(lin.weight.detach().numpy(), g[f"{tag}_W{i}"], atol=1e-4, rtol=1e-3)
        np.testing.assert_allclose(lin.bias.detach().numpy(), g[f"{tag}_b{i}"], atol=1e-4, rtol=1e-3)
    proba = clf.predict_proba(g["X"][:16])
    assert proba.dtype == np.float64
    np.testing.assert_allclose(proba, g[f"{tag}_proba"], atol=1e-4)
    assert np.abs(proba.sum(1) - 1).max() < 1e-12
    pred = np.searchsorted(LABELS5, clf.predict(g["X"]))
    assert (pred == g[f"{tag}_pred"]).mean() >= 0.999
    assert clf.launches > 0


def test_init_is_bit_identical_to_reference(golden_dir):
    g = np.load(golden_dir / "mlp_train.npz")
    clf = TorchMLPClassifier(hidden_layer_sizes=(16, 8), random_state=0).init_for(32, LABELS5.tolist())
    for i, lin in enumerate(clf._module.linears):
        assert np.array_equal(lin.weight.detach().numpy(), g[f"init_W{i}"])
        assert not lin.bias.detach().numpy().any()


def test_first_step_gradient_buffer_matches_oracle():
    """Read the flat un-normalised gradient buffer through the grad_sync hook after the first backward."""
    X, y = cluster_data(200, 1280, 37, 3)   # K = 37 -> padded to 40
    hidden = (200, 100)
    clf = TorchMLPClassifier(hidden_layer_sizes=hidden, learning_rate_init=1e-4, random_state=0, shuffle=False)
    clf.init_for(1280, list(range(37)))
    n_grad = int(_lib.load().mc_mlp_grad_size(clf._h))
    got = torch.empty(n_grad, dtype=torch.float32, device="cuda")

    @_lib.GRAD_SYNC_FN
    def hook(ptr, n, stream, user):
        assert n == n_grad
        _copy(ptr, n)

    def _copy(ptr, n):
        # device-to-device copy on the current stream via torch (the hook runs on the launching thread)
        import cuda.bindings.runtime as rt

        (err,) = rt.cudaMemcpyAsync(got.data_ptr(), ptr, n * 4, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice,
                                    torch.cuda.current_stream().cuda_stream)
        assert int(err) == 0

    clf._grad_hook = hook
    clf.partial_fit(X, y, classes=list(range(37)))
    flat = got.cpu().numpy()
    w, b = ohead.init_mlp(1280, hidden, 37, 0)
    gW, gB, wsum, lsum = ohead.minibatch_sums(w, b, torch.from_numpy(X), torch.from_numpy(y.astype(np.int64)))
    dims = [1280, 200, 100, 37]
    dims_p = [(d + 3) // 4 * 4 for d in dims]
    off = 0
    for i in range(3):
        Wg = flat[off: off + dims_p[i + 1] * dims_p[i]].reshape(dims_p[i + 1], dims_p[i])
        off += Wg.size
        bg = flat[off: off + dims_p[i + 1]]
        off += dims_p[i + 1]
        np.testing.assert_allclose(Wg[: dims[i + 1], : dims[i]], gW[i].numpy(), atol=2e-4, rtol=1e-3)
        np.testing.assert_allclose(bg[: dims[i + 1]], gB[i].numpy(), atol=2e-4, rtol=1e-3)
        assert not Wg[dims[i + 1]:].any() and not Wg[:, dims[i]:].any() and not bg[dims[i + 1]:].any()
    assert off + 4 == n_grad
    np.testing.assert_allclose(flat[off: off + 3], [wsum, lsum, 200.0], rtol=1e-5)


@pytest.mark.parametrize("hidden", [(200, 100), (500, 300, 100)])
def test_full_size_pass_vs_oracle(hidden):
    """BASELINE head sizes, 500 classes, 10 mini-batches + ragged tail; compare with the pinned oracle."""
    X, y = cluster_data(2050, 1280, 500, 42)
    clf = TorchMLPClassifier(hidden_layer_sizes=hidden, learning_rate_init=1e-4, random_state=0, alpha=1e-4)
    classes = list(range(500))
    clf.partial_fit(X, y, classes=classes)
    clf.partial_fit(X, y)
    w, b = ohead.init_mlp(1280, hidden, 500, 0)
    adam = ohead.AdamState(w + b)
    curve = [ohead.partial_fit(w, b, adam, X, y, lr=1e-4, random_state=0) for _ in range(2)]
    np.testing.assert_allclose(clf.loss_curve_, curve, rtol=1e-4)
    # Adam moves a parameter by up to ~lr per step whatever the gradient's size, so entries whose
    # gradient is at rounding level may differ by a few lr (1e-4); everything else agrees to 2e-5.
    for lin, wi, bi in zip(clf._module.linears, w, b):
        for got, ref in ((lin.weight.detach().numpy(), wi.numpy()), (lin.bias.detach().numpy(), bi.numpy())):
            d = np.abs(got - ref)
            assert d.max() <= 3e-4 and (d <= 2e-5 + 1e-3 * np.abs(ref)).mean() >= 0.999
    assert clf.n_steps_ == adam.t == 22


def test_pickle_round_trip_resumes_identically():
    X, y = cluster_data(900, 64, 6, 1)
    a = TorchMLPClassifier(hidden_layer_sizes=(32,), random_state=3)
    a.partial_fit(X, y, classes=list(range(6)))
    b = pickle.loads(pickle.dumps(a))
    assert b.n_iter_ == 1 and b.loss_curve_ == a.loss_curve_
    # the pickle entries are the reference's: _MLPModule.state_dict() and torch.optim.Adam.state_dict()
    state = a.__getstate__()
    assert list(state["_module_state"]) == ["linears.0.weight", "linears.0.bias", "linears.1.weight", "linears.1.bias"]
    params = [torch.nn.Parameter(torch.zeros_like(v)) for v in state["_module_state"].values()]
    opt = torch.optim.Adam(params, lr=a.learning_rate_init)
    opt.load_state_dict(state["_optimizer_state"])
    assert int(opt.state[params[0]]["step"]) == a.n_steps_ and opt.state[params[0]]["exp_avg"].shape == (32, 64)
    a.partial_fit(X, y)
    b.partial_fit(X, y)
    assert a.loss_curve_[-1] == b.loss_curve_[-1]  # same kernels, same order: bit-identical
    for la, lb in zip(a._module.linears, b._module.linears):
        assert torch.equal(la.weight, lb.weight) and torch.equal(la.bias, lb.bias)


def test_errors_and_sklearn_protocol():
    with pytest.raises(ValueError):
        TorchMLPClassifier(activation="tanh")
    with pytest.raises(ValueError):
        TorchMLPClassifier(solver="sgd")
    clf = TorchMLPClassifier(hidden_layer_sizes=(8,), random_state=0)
    with pytest.raises(RuntimeError):
        clf.predict_proba(np.zeros((2, 4), np.float32))  # torch_classifier.py:333-337
    X, y = cluster_data(50, 12, 3, 0)
    clf.partial_fit(X, y, classes=[0, 1, 2])
    with pytest.raises(ValueError):
        clf.partial_fit(X[:, :11], y)
    with pytest.raises(ValueError):
        clf.partial_fit(X, np.full(50, 7))
    with pytest.raises(ValueError):
        clf.predict_proba(np.zeros((2, 11), np.float32))
    assert clf.get_params()["hidden_layer_sizes"] == (8,)
    assert clf.set_params(alpha=0.5).alpha == 0.5
    with pytest.raises(ValueError):
        clf.set_params(nope=1)
    with pytest.raises(ValueError):
        TorchMLPClassifier(class_weight={0: 1.0}).partial_fit(X, y, classes=[0, 1, 2])
    f = TorchMLPClassifier(hidden_layer_sizes=(16,), random_state=0, max_iter=5).fit(X, y)
    assert 1 <= f.n_iter_ <= 5 and len(f.loss_curve_) == f.n_iter_


def test_rowlocal_experiment_path_matches_default():
    """The opt-in row-local middle of the Adam step (``MC_MLP_ROWLOCAL=1``, read once per process -> a subprocess) walks
    the same trajectory as the default GEMM chain up to fp32 summation order, class weights and ragged tail included."""
    import json
    import os
    import subprocess
    import sys

    code = (
        "import json, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from mermaid_classifier_b200.torch_classifier import TorchMLPClassifier\n"
        "rng = np.random.RandomState(5); c = rng.randn(7, 48) * 3.0; y = rng.randint(0, 7, size=930)\n"
        "X = (c[y] + rng.randn(930, 48) * 1.3).astype(np.float32)\n"
        "cw = {k: 0.5 + 0.25 * k for k in range(7)}\n"
        "clf = TorchMLPClassifier(hidden_layer_sizes=(24, 16, 12), learning_rate_init=1e-3, random_state=0, class_weight=cw)\n"
        "for _ in range(3): clf.partial_fit(X, y, classes=list(range(7)))\n"
        "print(json.dumps({'loss': clf.loss_curve_, 'p': clf.predict_proba(X[:20]).tolist()}))\n"
    ) % str(__import__("pathlib").Path(__file__).resolve().parents[1])
    outs = []
    for flag in (None, "1"):
        env = {k: v for k, v in os.environ.items() if k != "MC_MLP_ROWLOCAL"}
        if flag:
            env["MC_MLP_ROWLOCAL"] = flag
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    np.testing.assert_allclose(outs[1]["loss"], outs[0]["loss"], rtol=2e-5)
    np.testing.assert_allclose(np.asarray(outs[1]["p"]), np.asarray(outs[0]["p"]), atol=2e-5)


def test_graph_replay_is_bit_identical_to_launched_steps():
    """Runs of >= 8 equal-sized Adam steps are replayed from one captured pair of steps (csrc/mlp_api.inl): the same kernels in
    the same order on a staged copy of each mini-batch, so parameters, Adam moments and the loss curve equal the
    kernel-by-kernel loop (``MC_MLP_GRAPH=0``, read once per process -> subprocesses) BIT FOR BIT -- ragged tail, class
    weights, several calls (the graph is reused), a call too short to replay, and an odd run length included."""
    import json
    import os
    import subprocess
    import sys

    code = (
        "import json, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from mermaid_classifier_b200.torch_classifier import TorchMLPClassifier\n"
        "rng = np.random.RandomState(11); c = rng.randn(9, 64) * 3.0; y = rng.randint(0, 9, size=5330)\n"
        "X = (c[y] + rng.randn(5330, 64) * 1.3).astype(np.float32)\n"
        "cw = {k: 0.5 + 0.125 * k for k in range(9)}\n"
        "clf = TorchMLPClassifier(hidden_layer_sizes=(40, 24), learning_rate_init=1e-3, random_state=0, class_weight=cw)\n"
        "clf.partial_fit(X, y, classes=list(range(9)))\n"          # 26 full steps + a tail of 130
        "clf.partial_fit(X[:4600], y[:4600])\n"                    # 23 full steps: odd run
        "clf.partial_fit(X[:900], y[:900])\n"                      # 4 full steps + tail: too short to replay
        "clf.partial_fit(X, y)\n"
        "ws, bs = clf._pull_params()\n"
        "print(json.dumps({'loss': clf.loss_curve_, 'w': [w.tolist() for w in ws], 'b': [b.tolist() for b in bs],"
        " 'steps': clf.n_steps_, 'graph_steps': clf.graph_steps_}))\n"
    ) % str(__import__("pathlib").Path(__file__).resolve().parents[1])
    outs = []
    for flag in ("0", None):
        env = {k: v for k, v in os.environ.items() if k not in ("MC_MLP_GRAPH", "MC_MLP_ROWLOCAL")}
        if flag:
            env["MC_MLP_GRAPH"] = flag
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    launched, replayed = outs
    assert launched["graph_steps"] == 0 and replayed["graph_steps"] >= 60, (launched["graph_steps"], replayed["graph_steps"])
    assert launched["steps"] == replayed["steps"] == 27 + 23 + 5 + 27
    assert launched["loss"] == replayed["loss"]
    for a, b in zip(launched["w"] + launched["b"], replayed["w"] + replayed["b"]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
