"""Oracle: per-epoch evaluation, early stopping and Platt calibration of the trainer (TEST INFRASTRUCTURE).

Restates on the CPU (numpy fp64; scipy for the optimiser sklearn itself calls):

* ``MermaidTrainer._calc_acc_batched`` / ``_calc_acc_and_log_loss_batched`` --
  ``/root/reference/mermaid_classifier/pyspacer/trainer.py:295-342`` (accuracy of the argmax labels,
  ``sklearn.metrics.log_loss(gt, proba, labels=classes)``)
* the early-stopping bookkeeping of ``MermaidTrainer.__call__`` -- ``trainer.py:125-265``
* ``MermaidTrainer._calibrate_in_batches`` -- ``trainer.py:344-396`` -> ``sklearn.calibration._fit_calibrator``
  with ``method="sigmoid"`` -> ``_sigmoid_calibration`` per class, and the calibrated ``predict_proba``
  of ``_CalibratedClassifier`` (multiclass form)

The metric / calibration arithmetic lives in a third-party dependency, scikit-learn (reference pin
``scikit-learn==1.5.2``, ``pyproject.toml:28,53``), which is not under ``/root/reference``.  Its published
algorithm is restated here; PINNED by ``tests/test_oracle_trainer.py`` against the scikit-learn installed in
this image (1.9.0 -- ``log_loss``, ``accuracy_score``, ``_SigmoidCalibration``; the calibration routine is
unchanged between 1.5.2 and 1.9.0) and against the reference's own ``MermaidTrainer`` helper methods, run
on the reference's ``TorchMLPClassifier`` (``tests/golden/make_golden_trainer.py``, fixture ``trainer_eval.npz``).
"""

from __future__ import annotations

from math import log
from typing import Any, Callable, Sequence

import numpy as np

EPS = float(np.finfo(np.float64).eps)


def accuracy(gt_idx: np.ndarray, proba: np.ndarray) -> float:
    """``accuracy_score(gt, classes[argmax(proba)])`` (trainer.py:305-308, 335-338)."""
    return float(np.mean(np.argmax(proba, axis=1) == np.asarray(gt_idx)))


def log_loss_terms(gt_idx: np.ndarray, proba: np.ndarray) -> np.ndarray:
    """Per-row terms of ``log_loss(labels=classes)``: clip to ``[eps, 1 - eps]``, ``-log p[i, y_i]``
    (scikit-learn ``_classification.py``; no renormalisation since 1.3)."""
    p = np.clip(np.asarray(proba, dtype=np.float64), EPS, 1.0 - EPS)
    return -np.log(p[np.arange(p.shape[0]), np.asarray(gt_idx)])


def log_loss(gt_idx: np.ndarray, proba: np.ndarray) -> float:
    return float(np.mean(log_loss_terms(gt_idx, proba)))


def early_stopping_walk(val_losses: Sequence[float], nbr_epochs: int, patience: int | None) -> dict[str, Any]:
    """The bookkeeping of trainer.py:125-265 driven by a given per-epoch val-loss sequence: which epoch is
    the last one run, which one is restored, and the summary dict the runner logs."""
    best, best_idx, since = float("inf"), None, 0
    stop_reason, epoch = "budget_exhausted", 0
    for epoch in range(nbr_epochs):
        v = val_losses[epoch]
        if patience is not None:
            if v < best:
                best, best_idx, since = v, epoch, 0
            else:
                since += 1
            if since >= patience:
                stop_reason = "early_stopping"
                break
    restored = best_idx if (patience is not None and best_idx is not None and best_idx != epoch) else epoch
    return {
        "enabled": patience is not None,
        "patience": patience,
        "stop_reason": stop_reason,
        "final_epoch": epoch + 1,
        "best_val_epoch": best_idx + 1 if best_idx is not None else None,
        "best_val_loss": best if best != float("inf") else None,
        "restored_epoch": restored + 1,
    }


def platt_targets(y01: np.ndarray) -> tuple[np.ndarray, float, float]:
    neg = y01 <= 0
    prior0 = float(np.sum(neg))
    prior1 = y01.shape[0] - prior0
    t = np.where(neg, 1.0 / (prior0 + 2.0), (prior1 + 1.0) / (prior1 + 2.0)).astype(np.float64)
    return t, prior0, prior1


def platt_objective(a: float, b: float, f: np.ndarray, t: np.ndarray) -> tuple[float, np.ndarray]:
    """``sum log(1 + exp(r)) - t r`` with ``r = -(a f + b)`` and its gradient wrt ``(a, b)``
    (HalfBinomialLoss.loss_gradient as ``_sigmoid_calibration.loss_grad`` uses it)."""
    r = -(a * f + b)
    loss = np.logaddexp(0.0, r) - t * r
    g = 1.0 / (1.0 + np.exp(-r)) - t
    return float(loss.sum()), np.asarray([-(g @ f), -g.sum()], dtype=np.float64)


def sigmoid_calibration(f: np.ndarray, y01: np.ndarray) -> tuple[float, float]:
    """``sklearn.calibration._sigmoid_calibration`` (1.5.2) for one class: returns ``(a, b)`` of
    ``expit(-(a f + b))``."""
    from scipy.optimize import minimize

    f = np.asarray(f, dtype=np.float64)
    scale = 1.0
    mx = float(np.max(np.abs(f)))
    if mx >= 30:
        scale = mx
        f = f / scale
    t, prior0, prior1 = platt_targets(np.asarray(y01))
    ab0 = np.array([0.0, log((prior0 + 1.0) / (prior1 + 1.0))])
    res = minimize(lambda ab: platt_objective(ab[0], ab[1], f, t), ab0, method="L-BFGS-B", jac=True,
                   options={"gtol": 1e-6, "ftol": 64 * np.finfo(float).eps})
    return float(res.x[0] / scale), float(res.x[1])


def calibrate(proba: np.ndarray, gt_idx: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """One sigmoid calibrator per class on the uncalibrated ``predict_proba`` of the reference split
    (trainer.py:360-384, multiclass: the full (N, K) matrix, class k against ``y == k``)."""
    k = proba.shape[1]
    ab = [sigmoid_calibration(proba[:, j], (np.asarray(gt_idx) == j).astype(np.int64)) for j in range(k)]
    return np.asarray([x[0] for x in ab]), np.asarray([x[1] for x in ab])


def calibrated_proba64(proba: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``_CalibratedClassifier.predict_proba`` (multiclass, fp64): per-class ``expit(-(a p + b))``, row
    normalise (uniform when the row sums to 0), clip (1, 1 + 1e-5] to 1."""
    c = 1.0 / (1.0 + np.exp(a[None, :] * proba + b[None, :]))
    denom = c.sum(axis=1)[:, None]
    k = proba.shape[1]
    out = np.divide(c, denom, out=np.full_like(c, 1.0 / k), where=denom != 0)
    out[(out > 1.0) & (out <= 1.0 + 1e-5)] = 1.0
    return out


def run_epochs(clf: Any, train_batches: Callable[[int], Any], evaluate: Callable[[Any], tuple[float, float, float]],
               nbr_epochs: int, patience: int | None, snapshot: Callable[[Any], Any]) -> tuple[Any, list[dict[str, Any]], dict[str, Any]]:
    """The epoch loop of trainer.py:138-265 over injected pieces: ``train_batches(epoch)`` yields ``(x, y)``
    chunks for ``clf.partial_fit``; ``evaluate(clf) -> (ref_acc, val_acc, val_loss)``."""
    history: list[dict[str, Any]] = []
    best, best_idx, since, best_snap = float("inf"), None, 0, None
    stop_reason, epoch = "budget_exhausted", 0
    for epoch in range(nbr_epochs):
        for x, y in train_batches(epoch):
            clf.partial_fit(x, y)
        ref_acc, val_acc, val_loss = evaluate(clf)
        history.append({"epoch": epoch, "ref_accuracy": ref_acc, "val_accuracy": val_acc, "val_loss": val_loss,
                        "training_loss": clf.loss_curve_[-1]})
        if patience is not None:
            if val_loss < best:
                best, best_idx, since, best_snap = val_loss, epoch, 0, snapshot(clf)
            else:
                since += 1
            if since >= patience:
                stop_reason = "early_stopping"
                break
    if patience is not None and best_snap is not None and best_idx != epoch:
        clf = best_snap
    info = {"enabled": patience is not None, "patience": patience, "stop_reason": stop_reason, "final_epoch": epoch + 1,
            "best_val_epoch": best_idx + 1 if best_idx is not None else None,
            "best_val_loss": best if best != float("inf") else None}
    return clf, history, info


def platt_newton(f: np.ndarray, y01: np.ndarray, gtol: float = 1e-9, max_passes: int = 400) -> tuple[float, float, float, int]:
    """The iteration ``mermaid_classifier_b200/csrc/platt.cuh`` runs per class, restated in numpy: Newton direction with
    a 1e-12 ridge, backtracking (sufficient decrease 1e-4, halving down to 1e-10), stop on ``|grad|_inf < gtol`` or, once
    the predicted decrease is below 1e-11 of the objective, one final un-searched Newton step.  Same objective and start as
    :func:`sigmoid_calibration`; returns ``(a, b, objective at the last evaluated point, passes)``.  Used to check on
    the CPU that this iteration lands on sklearn's optimum, and on the GPU that the kernel follows it."""
    f = np.asarray(f, dtype=np.float64)
    t, prior0, prior1 = platt_targets(np.asarray(y01))
    A, B = 0.0, log((prior0 + 1.0) / (prior1 + 1.0))
    tA, tB, first, step, fval, gd, dA, dB, passes, accepted = A, B, True, 1.0, 0.0, 0.0, 0.0, 0.0, 0, 0
    while passes < max_passes:
        r = -(tA * f + tB)
        e = np.exp(-np.abs(r))
        p = np.where(r >= 0, 1.0, e) / (1.0 + e)
        L = float((np.log1p(e) + np.maximum(r, 0.0) - t * r).sum())
        g, h = p - t, p * (1.0 - p)
        gA, gB = float(-(g * f).sum()), float(-g.sum())
        hAA, hAB, hBB = float((h * f * f).sum()) + 1e-12, float((h * f).sum()), float(h.sum()) + 1e-12
        passes += 1
        if first or L < fval + 1e-4 * step * gd:
            first, A, B, fval = False, tA, tB, L
            accepted += 1
            if (abs(gA) < gtol and abs(gB) < gtol) or accepted > 100:
                break
            det = hAA * hBB - hAB * hAB
            dA, dB = -(hBB * gA - hAB * gB) / det, -(-hAB * gA + hAA * gB) / det
            gd, step = gA * dA + gB * dB, 1.0
            if not gd < 0.0 or not np.isfinite(dA) or not np.isfinite(dB):
                break
            if -gd < 1e-11 * max(1.0, abs(fval)):
                A, B = A + dA, B + dB
                break
        else:
            step *= 0.5
            if step < 1e-10:
                break
        tA, tB = A + step * dA, B + step * dB
    return A, B, fval, passes
