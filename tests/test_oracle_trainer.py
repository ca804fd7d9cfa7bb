"""Pin ``oracle/trainer.py`` (evaluation metrics, early stopping, Platt calibration) against scikit-learn and
against ``tests/golden/trainer_eval.npz`` -- produced by the reference's own ``MermaidTrainer`` helpers driving
the reference's ``TorchMLPClassifier`` (``tests/golden/make_golden_trainer.py``).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import head as ohead
from oracle import trainer as otr


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "trainer_eval.npz")


def _params(g):
    n = len([k for k in g.files if k.startswith("W")])
    return [torch.from_numpy(g[f"W{i}"]) for i in range(n)], [torch.from_numpy(g[f"b{i}"]) for i in range(n)]


def test_metrics_match_sklearn():
    from sklearn.metrics import accuracy_score
    from sklearn.metrics import log_loss as sk_log_loss

    rng = np.random.default_rng(0)
    p = rng.dirichlet(np.full(7, 0.3), size=500)
    p[3] = 0.0
    p[3, 2] = 1.0  # an exact 0 / 1 row exercises the eps clip
    y = rng.integers(0, 7, size=500)
    y[3] = 5
    assert otr.log_loss(y, p) == pytest.approx(sk_log_loss(y, p, labels=list(range(7))), rel=1e-13)
    assert otr.accuracy(y, p) == accuracy_score(y, p.argmax(1))
    assert otr.log_loss_terms(y, p)[3] == pytest.approx(-np.log(otr.EPS))


def test_final_epoch_metrics_match_reference_helpers(g):
    w, b = _params(g)
    proba_val = ohead.softmax_proba(g["Xv"], w, b)
    proba_ref = ohead.softmax_proba(g["Xr"], w, b)
    np.testing.assert_allclose(proba_ref, g["proba_ref"], atol=1e-7)
    assert otr.accuracy(g["yv"], proba_val) == pytest.approx(float(g["val_acc"][-1]), abs=1e-12)
    assert otr.accuracy(g["yr"], proba_ref) == pytest.approx(float(g["ref_acc"][-1]), abs=1e-12)
    assert otr.log_loss(g["yv"], proba_val) == pytest.approx(float(g["val_loss"][-1]), rel=1e-6)


def test_epoch_loop_replays_reference_run(g):
    """The oracle's partial_fit + metrics, driven by the restated epoch loop, reproduce every per-epoch number the
    reference's helpers reported (chunked training with epoch-seeded permutations, then ref / val evaluation)."""
    K, chunk, epochs = len(g["classes"]), int(g["chunk"]), int(g["epochs"])
    w, b = ohead.init_mlp(32, (24, 16), K, 0)
    adam = ohead.AdamState(w + b)

    class Clf:
        loss_curve_: list = []

        def partial_fit(self, x, y):
            self.loss_curve_.append(ohead.partial_fit(w, b, adam, x, y, lr=1e-3, random_state=0))

    def batches(epoch):
        order = np.random.default_rng(epoch).permutation(len(g["yt"]))
        for s in range(0, len(order), chunk):
            yield g["Xt"][order[s:s + chunk]], g["yt"][order[s:s + chunk]]

    def evaluate(_):
        pr, pv = ohead.softmax_proba(g["Xr"], w, b), ohead.softmax_proba(g["Xv"], w, b)
        return otr.accuracy(g["yr"], pr), otr.accuracy(g["yv"], pv), otr.log_loss(g["yv"], pv)

    _, hist, info = otr.run_epochs(Clf(), batches, evaluate, epochs, None, lambda c: c)
    assert info == {"enabled": False, "patience": None, "stop_reason": "budget_exhausted", "final_epoch": epochs,
                    "best_val_epoch": None, "best_val_loss": None}
    np.testing.assert_allclose([h["val_loss"] for h in hist], g["val_loss"], rtol=1e-5)
    np.testing.assert_allclose([h["val_accuracy"] for h in hist], g["val_acc"], atol=1.01 / len(g["yv"]))
    np.testing.assert_allclose([h["ref_accuracy"] for h in hist], g["ref_acc"], atol=1.01 / len(g["yr"]))
    # loss_curve_ gains one entry per chunk; the callback reports the last chunk of the epoch
    np.testing.assert_allclose([h["training_loss"] for h in hist], g["train_loss"], rtol=1e-5)


def test_early_stopping_walk():
    walk = otr.early_stopping_walk
    # improves, then two non-improving epochs with patience 2 -> stop after epoch 4, restore epoch 2
    r = walk([1.0, 0.8, 0.9, 0.85, 0.7], 5, 2)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"], r["restored_epoch"]) == ("early_stopping", 4, 2, 2)
    assert r["best_val_loss"] == 0.8
    # budget exhausted but the best epoch is not the last: still restored (trainer.py:238-252)
    r = walk([1.0, 0.7, 0.9], 3, 5)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"], r["restored_epoch"]) == ("budget_exhausted", 3, 2, 2)
    # disabled: nothing tracked
    r = walk([1.0, 2.0, 3.0], 3, None)
    assert (r["enabled"], r["final_epoch"], r["best_val_epoch"], r["best_val_loss"], r["restored_epoch"]) == (False, 3, None, None, 3)
    # the reference's own scripted cases (tests/pyspacer/test_trainer.py:289-359)
    r = walk([1.0, 0.9, 0.8, 0.7, 0.6], 5, 2)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"], r["best_val_loss"]) == ("budget_exhausted", 5, 5, 0.6)
    r = walk([1.0, 0.9, 0.8, 0.85, 0.9, 1.0, 1.1], 10, 2)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"], r["best_val_loss"]) == ("early_stopping", 5, 3, 0.8)
    r = walk([1.0, 0.5, 0.6], 10, 1)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"]) == ("early_stopping", 3, 2)
    r = walk([1.0, 0.9, 0.95, 0.96], 10, 2)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"]) == ("early_stopping", 4, 2)
    # a tie is not an improvement (strict <)
    r = walk([1.0, 1.0, 1.0], 3, 2)
    assert (r["stop_reason"], r["final_epoch"], r["best_val_epoch"]) == ("early_stopping", 3, 1)


def test_sigmoid_calibration_matches_sklearn(g):
    from sklearn.calibration import _SigmoidCalibration

    a, b = otr.calibrate(g["proba_ref"], g["yr"])
    np.testing.assert_allclose(a, g["platt_a"], rtol=1e-9)
    np.testing.assert_allclose(b, g["platt_b"], rtol=1e-9)
    rng = np.random.default_rng(1)
    f = rng.random(2000) ** 3
    y = (rng.random(2000) < 0.1 + 0.8 * f).astype(int)
    sk = _SigmoidCalibration().fit(f, y)
    assert otr.sigmoid_calibration(f, y) == pytest.approx((sk.a_, sk.b_), rel=1e-9)
    # no positive sample at all: still a well-defined minimiser
    sk0 = _SigmoidCalibration().fit(f, np.zeros_like(y))
    assert otr.sigmoid_calibration(f, np.zeros_like(y)) == pytest.approx((sk0.a_, sk0.b_), rel=1e-6, abs=1e-6)


def test_calibrated_proba_matches_reference_run(g):
    w, b = _params(g)
    pv = ohead.softmax_proba(g["Xv"], w, b)
    got = otr.calibrated_proba64(pv, g["platt_a"], g["platt_b"])
    np.testing.assert_allclose(got, g["cal_proba_val"], atol=1e-7)
    # and the fp32 head form (what the artifact computes) sits within the reference's 1e-6 export gate of it
    head = ohead.calibrated_proba(g["Xv"], w, b, torch.from_numpy(g["platt_a"]).float(), torch.from_numpy(g["platt_b"]).float())
    assert np.abs(head - got).max() < 1e-6


@pytest.mark.parametrize("case", ["skewed", "rare", "no_positive", "all_positive", "near_separable", "uninformative"])
def test_newton_iteration_reaches_the_lbfgsb_optimum(case):
    """The device algorithm (restated as ``platt_newton``) and sklearn's L-BFGS-B minimise the same convex objective:
    Newton's end point is never worse, its gradient is at rounding level, and the calibration curves coincide."""
    rng = np.random.default_rng(7)
    n = 20000
    f = rng.random(n) ** 6
    y = {"skewed": rng.random(n) < 0.05 + 0.9 * f,
         "rare": rng.random(n) < 0.002 + 0.3 * f,
         "no_positive": np.zeros(n, bool),
         "all_positive": np.ones(n, bool),
         "near_separable": (f > 0.3) ^ (rng.random(n) < 0.01),
         "uninformative": rng.random(n) < 0.1}[case].astype(int)
    a, b, loss, passes = otr.platt_newton(f, y)
    ra, rb = otr.sigmoid_calibration(f, y)
    t = otr.platt_targets(y)[0]
    l_newton, grad = otr.platt_objective(a, b, f, t)
    l_ref = otr.platt_objective(ra, rb, f, t)[0]
    assert passes <= 60
    assert l_newton <= l_ref + 1e-9 * max(1.0, abs(l_ref))
    assert np.abs(grad).max() < 1e-6
    assert loss == pytest.approx(l_newton, rel=1e-9, abs=1e-9)
    q = np.linspace(0.0, 1.0, 201)
    curve = lambda A, B: 1.0 / (1.0 + np.exp(A * q + B))
    assert np.abs(curve(a, b) - curve(ra, rb)).max() < 2e-3
