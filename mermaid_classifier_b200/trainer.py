"""``MermaidTrainer`` over the B200 head kernels: epoch loop, batched evaluation, early stopping, Platt calibration.

Mirror of ``mermaid_classifier/pyspacer/trainer.py`` of the reference:

* ``MermaidTrainer.__init__`` / ``__call__``               -- ``trainer.py:56-293`` (same keywords, same callback
  dictionary, same ``_early_stop_info`` summary, same ``(clf_calibrated, val_results, return_message)`` triple)
* ``_calc_acc_batched``                                    -- ``trainer.py:295-308``
* ``_calc_acc_and_log_loss_batched``                       -- ``trainer.py:310-342``
* ``_calibrate_in_batches``                                -- ``trainer.py:344-396``

What changes is where the arithmetic runs.  Training chunks go to ``TorchMLPClassifier.partial_fit`` (fused
CUDA forward / CE / backward / Adam).  The two evaluation passes never build the ``(N, K)`` probability matrix:
``mc_head_evaluate`` reduces hits and log-loss terms on the device.  Calibration scores the reference split
once into a device float64 matrix and fits all K sigmoid calibrators together (``mc_platt_fit``).

Label sets are anything with pyspacer's ``ImageLabels.load_data_in_batches(batch_size, random_seed=None)``
(host lists, uploaded per chunk) or the device-resident :class:`DeviceLabels` below, which keeps a split's
features in HBM across epochs (the reference re-reads every feature file from disk each epoch,
``trainer.py:141-144``).
"""

from __future__ import annotations

import copy
import time
from dataclasses import dataclass
from logging import getLogger
from typing import Any, Callable, Iterator, Sequence

import numpy as np
import torch

from . import _lib
from .inference import DeviceHead, platt_fit_device
from .torch_classifier import TorchMLPClassifier

logger = getLogger(__name__)


# -- data side ---------------------------------------------------------------------------------------------------
class DeviceLabels:
    """One split (train / ref / val) resident in HBM: ``X`` ``(n, D)`` float32, ``y`` labels.

    Offers the two things the trainer reads from pyspacer's ``ImageLabels`` -- ``load_data_in_batches``,
    ``label_count`` / ``classes_set`` / ``len`` -- plus :meth:`device_batches`, which yields CUDA chunks."""

    def __init__(self, X: Any, y: Sequence[Any], *, device: int | None = None, n_images: int | None = None):
        t = _lib.require_cuda()
        self.device = t.cuda.current_device() if device is None else int(device)
        self.y = np.asarray(y)
        with t.cuda.device(self.device):
            if isinstance(X, torch.Tensor):
                self.X = X.to(device=f"cuda:{self.device}", dtype=torch.float32).contiguous()
            else:
                self.X = t.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float32))).cuda()
        if self.X.dim() != 2 or self.X.shape[0] != self.y.shape[0]:
            raise ValueError(f"X must be (n, D) with one label per row; got {tuple(self.X.shape)} and {self.y.shape}")
        self._n_images = n_images
        self._idx_cache: tuple[tuple[Any, ...], torch.Tensor] | None = None

    def __len__(self) -> int:  # ImageLabels: number of images
        return self._n_images if self._n_images is not None else self.label_count

    @property
    def label_count(self) -> int:
        return int(self.y.shape[0])

    @property
    def classes_set(self) -> set[Any]:
        return set(self.y.tolist())

    def _order(self, random_seed: int | None) -> np.ndarray:
        n = self.label_count
        return np.arange(n) if random_seed is None else np.random.default_rng(random_seed).permutation(n)

    def load_data_in_batches(self, batch_size: int, random_seed: int | None = None) -> Iterator[tuple[list[Any], list[Any]]]:
        order = self._order(random_seed)
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            yield self.X[torch.from_numpy(idx).to(self.X.device)].cpu().numpy().tolist(), self.y[idx].tolist()

    def indices_for(self, classes: Sequence[Any]) -> torch.Tensor:
        """int32 positions of every label in ``classes`` (sorted unique order of the estimator), on the device."""
        key = tuple(classes)
        if self._idx_cache is None or self._idx_cache[0] != key:
            arr = np.asarray(classes)
            pos = np.searchsorted(arr, self.y)
            bad = (pos >= len(arr)) | (arr[np.minimum(pos, len(arr) - 1)] != self.y)
            if bad.any():
                raise ValueError(f"y contains labels not in classes: {sorted(set(self.y[bad].tolist()))[:5]}")
            with torch.cuda.device(self.device):
                self._idx_cache = (key, torch.from_numpy(pos.astype(np.int32)).cuda())
        return self._idx_cache[1]

    def device_batches(self, batch_size: int, classes: Sequence[Any], random_seed: int | None = None):
        yi = self.indices_for(classes)
        if random_seed is None:
            for s in range(0, self.label_count, batch_size):
                yield self.X[s:s + batch_size], yi[s:s + batch_size]
            return
        order = torch.from_numpy(self._order(random_seed)).to(self.X.device)
        for s in range(0, self.label_count, batch_size):
            idx = order[s:s + batch_size]
            yield self.X.index_select(0, idx), yi.index_select(0, idx).contiguous()


@dataclass
class TaskLabels:
    """``TrainingTaskLabels``-shaped bundle of the three splits."""

    train: Any
    ref: Any
    val: Any

    @property
    def label_count(self) -> int:
        return self.train.label_count + self.ref.label_count + self.val.label_count


@dataclass
class ValResults:
    scores: list[float]
    gt: list[int]
    est: list[int]
    classes: list[Any]


@dataclass
class TrainClassifierReturnMsg:
    acc: float
    pc_accs: list[float]
    ref_accs: list[float]
    runtime: float


# -- calibrated estimator ------------------------------------------------------------------------------------------
@dataclass
class SigmoidCalibrator:
    """``sklearn.calibration._SigmoidCalibration`` surface: ``a_``, ``b_``, ``predict``."""

    a_: float
    b_: float

    def predict(self, T: Any) -> np.ndarray:
        return 1.0 / (1.0 + np.exp(self.a_ * np.asarray(T, dtype=np.float64) + self.b_))


@dataclass
class _CalibratedInner:
    estimator: TorchMLPClassifier
    calibrators: list[SigmoidCalibrator]
    classes: np.ndarray
    method: str = "sigmoid"


class CalibratedClassifier:
    """What ``_calibrate_in_batches`` returns, with the attribute surface of ``CalibratedClassifierCV(cv="prefit")``
    that the reference reads afterwards: ``calibrated_classifiers_[0].{estimator, calibrators[k].a_/b_}``,
    ``classes_``, ``estimator``, ``cv`` (``inference/head.py:92-123``, ``export.py:71-77``).  ``predict_proba`` runs the
    calibrated head on the GPU (``CalibratedHead.forward`` arithmetic, fp32 -> fp64)."""

    cv = "prefit"
    method = "sigmoid"
    ensemble = True
    n_jobs = None

    def __init__(self, estimator: TorchMLPClassifier, a: np.ndarray, b: np.ndarray):
        self.estimator = estimator
        self.classes_ = estimator.classes_
        # K == 2: scikit-learn keeps ONE calibrator, fitted on the positive-class column (reference ``trainer.py:365-374``)
        self.binary = len(self.classes_) == 2
        if len(a) != (1 if self.binary else len(self.classes_)) or len(a) != len(b):
            raise ValueError(f"{len(a)} calibrators for {len(self.classes_)} classes")
        self.calibrated_classifiers_ = [_CalibratedInner(
            estimator, [SigmoidCalibrator(float(x), float(z)) for x, z in zip(a, b)], estimator.classes_)]
        self._head: DeviceHead | None = None

    @property
    def platt(self) -> tuple[np.ndarray, np.ndarray]:
        cal = self.calibrated_classifiers_[0].calibrators
        return np.asarray([c.a_ for c in cal]), np.asarray([c.b_ for c in cal])

    def head(self) -> DeviceHead:
        if self.binary:   # same refusal as the reference's build_calibrated_head (inference/head.py:110-114)
            raise ValueError("the calibrated head only supports the multiclass (K > 2) path; got K=2")
        if self._head is None:
            ws, bs = self.estimator._pull_params()
            a, b = self.platt
            self._head = DeviceHead(ws, bs, a.astype(np.float32), b.astype(np.float32), device=self.estimator._dev_index)
        return self._head

    def _check(self, X: Any) -> np.ndarray:
        arr = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != self.estimator.n_features_in_:
            raise ValueError(f"X must be (N, {self.estimator.n_features_in_}); got {arr.shape}")
        return arr

    def predict_proba(self, X: Any) -> np.ndarray:
        if self.binary:
            # sklearn _CalibratedClassifier.predict_proba, n_classes == 2: the calibrator maps the positive column, the
            # negative one is its complement (no renormalisation), then the (1, 1 + 1e-5] clip
            p1 = self.calibrated_classifiers_[0].calibrators[0].predict(self.estimator.predict_proba(self._check(X))[:, 1])
            proba = np.stack([1.0 - p1, p1], axis=1)
            proba[(1.0 < proba) & (proba <= 1.0 + 1e-5)] = 1.0
            return proba
        return self.head().scores_host(self._check(X), want_proba=True, want_labels=False)[0]

    def predict(self, X: Any) -> np.ndarray:
        if self.binary:
            return self.classes_[np.argmax(self.predict_proba(X), axis=1)]
        return self.classes_[self.head().scores_host(self._check(X), want_proba=False, want_labels=True)[1]]

    def __getstate__(self) -> dict[str, Any]:
        return {k: v for k, v in self.__dict__.items() if k != "_head"}

    def __setstate__(self, state: dict[str, Any]) -> None:
        self.__dict__.update(state)
        self._head = None


# -- the trainer ---------------------------------------------------------------------------------------------------
def _chunks(labels: Any, batch_size: int, clf: TorchMLPClassifier, random_seed: int | None = None):
    """CUDA ``(X, y_idx)`` chunks of a split, from the device store or from host batches."""
    if hasattr(labels, "device_batches"):
        yield from labels.device_batches(batch_size, clf.classes_, random_seed)
        return
    kwargs = {} if random_seed is None else {"random_seed": random_seed}
    for x, y in labels.load_data_in_batches(batch_size=batch_size, **kwargs):
        arr = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        with torch.cuda.device(clf._dev_index):
            yield torch.from_numpy(arr).cuda(), torch.from_numpy(clf._labels_to_indices(np.asarray(y)).astype(np.int32)).cuda()


class MermaidTrainer:
    def __init__(
        self,
        batch_size: int,
        on_epoch_end: Callable[[dict[str, Any]], None] | None = None,
        class_weight: dict[str, float] | None = None,
        early_stopping_patience: int | None = None,
        *,
        hidden_layer_sizes: Sequence[int] = (500, 300, 100),
        learning_rate_init: float = 1e-4,
        device: int | None = None,
        data_parallel: Any = None,
        dp_mode: str = "parity",
        clf_factory: Callable[..., Any] | None = None,
    ):
        if early_stopping_patience is not None and early_stopping_patience < 1:
            raise ValueError(f"early_stopping_patience must be >= 1 or None, got {early_stopping_patience!r}")
        self.batch_size = batch_size
        self.on_epoch_end = on_epoch_end
        self.class_weight = class_weight
        self.early_stopping_patience = early_stopping_patience
        self.hidden_layer_sizes = tuple(hidden_layer_sizes)
        self.learning_rate_init = learning_rate_init
        self.device = device
        # One process per GPU (torch_classifier.DataParallel).  "parity": every rank holds the SAME splits, each
        # 200-row mini-batch is divided over the ranks and the gradient all-reduced -> the single-GPU trajectory;
        # evaluation and calibration run replicated.  "throughput": every rank holds ITS OWN shard of each split;
        # evaluation counts are summed and the reference-split probabilities gathered over the ranks, so all ranks
        # report the same metrics, take the same early-stopping decisions and fit the same calibrators.
        if dp_mode not in ("parity", "throughput"):
            raise ValueError("dp_mode must be 'parity' or 'throughput'")
        self.data_parallel = data_parallel
        self.dp_mode = dp_mode
        self.clf_factory = clf_factory   # estimator class; default: the GPU TorchMLPClassifier
        self._early_stop_info: dict[str, Any] | None = None

    @property
    def _sharded(self) -> bool:
        dp = self.data_parallel
        return dp is not None and dp.world > 1 and self.dp_mode == "throughput"

    def _collective_device(self) -> torch.device:
        """Where small control tensors of the collectives live: the GPU for NCCL groups, the host for gloo (CPU tests)."""
        import torch.distributed as dist

        backend = str(dist.get_backend(self.data_parallel.group)).lower()
        return torch.device("cuda", self.data_parallel.device) if "nccl" in backend else torch.device("cpu")

    def _gather_lists(self, *lists: list[Any]) -> tuple[list[Any], ...]:
        """Concatenate per-rank Python lists in rank order on every rank (final validation results of throughput mode)."""
        if not self._sharded:
            return lists
        import torch.distributed as dist

        parts: list[Any] = [None] * self.data_parallel.world
        dist.all_gather_object(parts, [list(x) for x in lists], group=self.data_parallel.group)
        return tuple([v for part in parts for v in part[i]] for i in range(len(lists)))

    def _sum_over_ranks(self, hits: int, loss: float, n: int, device: Any) -> tuple[int, float, int]:
        """``device``: CUDA index of the estimator (NCCL groups), or a ``torch.device`` (the gloo tests pass "cpu")."""
        if not self._sharded:
            return hits, loss, n
        import torch.distributed as dist

        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        t = torch.tensor([float(hits), loss, float(n)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.data_parallel.group)
        h, l, m = t.tolist()
        return int(round(h)), l, int(round(m))

    def _gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenate every rank's rows (ragged first dimension) in rank order, on every rank."""
        if not self._sharded:
            return t
        import torch.distributed as dist

        group, world = self.data_parallel.group, self.data_parallel.world
        sizes = torch.zeros(world, dtype=torch.int64, device=t.device)
        sizes[self.data_parallel.rank] = t.shape[0]
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
        cap = int(sizes.max().item())
        padded = torch.zeros((cap, *t.shape[1:]), dtype=t.dtype, device=t.device)
        padded[: t.shape[0]] = t
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        return torch.cat([p[: int(sizes[r].item())] for r, p in enumerate(parts)]).contiguous()

    def __call__(self, labels: Any, nbr_epochs: int, pc_models: Sequence[Any] = (), **_kwargs: Any):
        classes_list = list(labels.ref.classes_set)
        if self._sharded:  # a shard may miss classes: every rank trains on the union
            import torch.distributed as dist

            sets: list[Any] = [None] * self.data_parallel.world
            dist.all_gather_object(sets, sorted(classes_list), group=self.data_parallel.group)
            classes_list = sorted(set().union(*map(set, sets)))
        clf = (self.clf_factory or TorchMLPClassifier)(
            hidden_layer_sizes=self.hidden_layer_sizes, learning_rate_init=self.learning_rate_init,
            class_weight=self.class_weight, random_state=0)
        if self.device is not None:
            clf.set_device(self.device)
        if self.data_parallel is not None:
            clf.enable_data_parallel(self.data_parallel, self.dp_mode)
        ref_accs: list[float] = []
        t0 = time.time()
        best_val_loss = float("inf")
        best_snapshot = None
        best_epoch_idx: int | None = None
        epochs_since_best = 0
        stop_reason = "budget_exhausted"
        epoch = 0
        for epoch in range(nbr_epochs):
            self._train_epoch(clf, labels.train, classes_list, epoch)
            ref_accs.append(self._calc_acc_batched(clf, labels.ref))
            val_acc, val_loss = self._calc_acc_and_log_loss_batched(clf, labels.val, classes_list)
            logger.debug(f"Epoch {epoch}, acc: {ref_accs[-1]}, val_acc: {val_acc}, val_loss: {val_loss}")
            patience = self.early_stopping_patience
            if patience is not None:
                if val_loss < best_val_loss:
                    best_val_loss, best_epoch_idx, epochs_since_best = val_loss, epoch, 0
                    best_snapshot = copy.deepcopy(clf)
                else:
                    epochs_since_best += 1
            stopping = patience is not None and epochs_since_best >= patience
            if self.on_epoch_end is not None:
                cb: dict[str, Any] = {
                    "epoch": epoch, "ref_accuracy": ref_accs[-1], "val_accuracy": val_acc, "val_loss": val_loss,
                    "training_loss": clf.loss_curve_[-1] if clf.loss_curve_ else None,
                    "cumulative_seconds": time.time() - t0,
                }
                if epoch == nbr_epochs - 1 or stopping:
                    cb["final_epoch"] = epoch + 1
                    cb["early_stopped"] = stopping
                    if best_epoch_idx is not None:
                        cb["best_val_epoch"] = best_epoch_idx + 1
                        cb["best_val_loss"] = best_val_loss
                self.on_epoch_end(cb)
            if stopping:
                stop_reason = "early_stopping"
                logger.info(f"Early stopping at epoch {epoch + 1}: best was epoch {(best_epoch_idx or 0) + 1}"
                            f" (val_loss={best_val_loss:.4f}).")
                break
        if self.early_stopping_patience is not None and best_snapshot is not None and best_epoch_idx != epoch:
            clf = best_snapshot
        self._early_stop_info = {
            "enabled": self.early_stopping_patience is not None,
            "patience": self.early_stopping_patience,
            "stop_reason": stop_reason,
            "final_epoch": epoch + 1,
            "best_val_epoch": best_epoch_idx + 1 if best_epoch_idx is not None else None,
            "best_val_loss": best_val_loss if best_val_loss != float("inf") else None,
        }

        clf_calibrated = self._calibrate_in_batches(clf, labels.ref)
        classes = clf_calibrated.classes_.tolist()
        # throughput mode: every rank scored its own validation shard; gather so that all ranks return the SAME
        # val_results / accuracies over the whole validation split (rank order = shard order)
        val_gts, val_ests, val_scores = self._gather_lists(*evaluate_classifier(clf_calibrated, labels.val, self.batch_size))
        pc_accs = []
        for pc_model in pc_models:
            pc_gts, pc_ests = self._gather_lists(*evaluate_classifier(pc_model, labels.val, self.batch_size)[:2])
            pc_accs.append(float(np.mean(np.asarray(pc_gts) == np.asarray(pc_ests))))
        val_results = ValResults(scores=val_scores, gt=[classes.index(m) for m in val_gts],
                                 est=[classes.index(m) for m in val_ests], classes=classes)
        msg = TrainClassifierReturnMsg(acc=float(np.mean(np.asarray(val_gts) == np.asarray(val_ests))),
                                       pc_accs=pc_accs, ref_accs=ref_accs, runtime=time.time() - t0)
        return clf_calibrated, val_results, msg

    def _train_epoch(self, clf: TorchMLPClassifier, train: Any, classes_list: list[Any], epoch: int) -> None:
        """``trainer.py:138-145``: one ``partial_fit`` per chunk, chunks drawn with ``random_seed=epoch``."""
        if hasattr(train, "device_batches"):
            clf.init_for(int(train.X.shape[1]), classes_list)
            chunks = ((xd.contiguous(), yd) for xd, yd in train.device_batches(self.batch_size, clf.classes_, random_seed=epoch))
            fit = lambda c: clf.partial_fit_device(*c)   # noqa: E731
        else:
            chunks = iter(train.load_data_in_batches(batch_size=self.batch_size, random_seed=epoch))
            fit = lambda c: clf.partial_fit(c[0], c[1], classes=classes_list)   # noqa: E731
        if not self._sharded:
            for c in chunks:
                fit(c)
            return
        # Throughput mode: every partial_fit pass all-reduces once per Adam step, so the ranks must make the same number
        # of passes.  Host ImageLabels batch by whole images, so the count is only known by drawing the chunks: agree on
        # "one more chunk?" before every pass and fail on ALL ranks together instead of hanging in the collective.
        import torch.distributed as dist

        n_done = 0
        while True:
            c = next(chunks, None)
            has = 0 if c is None else 1
            t = torch.tensor([has, -has], dtype=torch.int64, device=self._collective_device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.data_parallel.group)
            if int(t[0]) != -int(t[1]):
                raise ValueError(f"throughput mode needs the same number of training chunks on every rank: after {n_done} "
                                 f"chunks this rank has {'another' if has else 'no more'} while another rank differs; "
                                 f"use equal-size shards")
            if c is None:
                return
            fit(c)
            n_done += 1

    def _calc_acc_batched(self, clf: TorchMLPClassifier, labels: Any) -> float:
        hits = n = 0
        for xd, yd in _chunks(labels, self.batch_size, clf):
            h, _ = clf.evaluate_device(xd.contiguous(), yd)
            hits += h
            n += int(yd.shape[0])
        hits, _, n = self._sum_over_ranks(hits, 0.0, n, clf._dev_index)
        return hits / n

    def _calc_acc_and_log_loss_batched(self, clf: TorchMLPClassifier, labels: Any, classes_list: list[Any]) -> tuple[float, float]:
        if sorted(classes_list) != list(clf.classes_):
            raise ValueError("classes_list does not match the estimator's classes_")
        hits = n = 0
        loss = 0.0
        for xd, yd in _chunks(labels, self.batch_size, clf):
            h, l = clf.evaluate_device(xd.contiguous(), yd)
            hits += h
            loss += l
            n += int(yd.shape[0])
        hits, loss, n = self._sum_over_ranks(hits, loss, n, clf._dev_index)
        return hits / n, loss / n

    def _calibrate_in_batches(self, clf: TorchMLPClassifier, ref_labels: Any) -> CalibratedClassifier:
        k = len(clf.classes_)
        if k < 2:
            raise ValueError(f"calibration needs at least two classes; got K={k}")
        probs, ys = [], []
        for xd, yd in _chunks(ref_labels, self.batch_size, clf):
            probs.append(clf.predict_proba_device(xd.contiguous()))
            ys.append(yd)
        proba = torch.cat(probs) if len(probs) > 1 else probs[0]
        y = (torch.cat(ys) if len(ys) > 1 else ys[0]).contiguous()
        proba, y = self._gather_rows(proba.contiguous()), self._gather_rows(y)
        if k == 2:
            # binary: one calibrator on the positive-class column (``preds[:, 1:]``, reference ``trainer.py:371-373``); as a
            # one-column problem the positive rows are "class 0" and the negatives match no column
            proba, y = proba[:, 1:].contiguous(), (1 - y).to(torch.int32).contiguous()
        a, b, _, passes = platt_fit_device(proba.contiguous(), y)
        logger.debug(f"Platt calibration: {k} classes, {proba.shape[0]} rows, {passes} matrix passes")
        return CalibratedClassifier(clf, a, b)

    def serialize(self) -> dict[str, Any]:
        return {"batch_size": self.batch_size}


def evaluate_classifier(clf: Any, labels: Any, batch_size: int = 5000) -> tuple[list[Any], list[Any], list[float]]:
    """pyspacer ``train_utils.evaluate_classifier`` as ``trainer.py:269,274`` uses it: ground truth, estimated
    label and top score per point, streamed in batches (UPSTREAM-RECALLED shape of the return triple)."""
    gts: list[Any] = []
    ests: list[Any] = []
    scores: list[float] = []
    classes = np.asarray(clf.classes_)
    if hasattr(labels, "device_batches") and hasattr(clf, "head") and not getattr(clf, "binary", False):  # labels and top scores picked on the device
        for xd, _ in labels.device_batches(batch_size, clf.classes_):
            out = clf.head().scores_device(xd.contiguous(), topk=1)
            ests.extend(classes[out["topk_idx"][:, 0].cpu().numpy()].tolist())
            scores.extend(out["topk_val"][:, 0].double().cpu().numpy().tolist())
        return labels.y.tolist(), ests, scores
    for x, y in labels.load_data_in_batches(batch_size=batch_size):
        proba = clf.predict_proba(x)
        top = np.argmax(proba, axis=1)
        ests.extend(classes[top].tolist())
        scores.extend(proba[np.arange(len(top)), top].tolist())
        gts.extend(list(y))
    return gts, ests, scores
