"""``TorchMLPClassifier`` with its training inner loop on the GPU (``mc_mlp_*`` in libmermaid_b200).

Drop-in for ``mermaid_classifier.pyspacer.torch_classifier.TorchMLPClassifier`` of the reference
(``/root/reference/mermaid_classifier/pyspacer/torch_classifier.py:83-444``): same constructor
keywords, ``partial_fit`` / ``fit`` / ``predict`` / ``predict_proba``, ``classes_`` (sorted
``np.unique``), ``loss_curve_``, ``n_iter_``, ``get_params`` / ``set_params``, picklable, and a
``_module`` whose ``linears[i].weight/.bias`` are what ``build_calibrated_head``
(``inference/head.py:118-123``) and ``export_artifact`` (``inference/export.py:71-77``) read.

What moved to the device: the whole mini-batch loop of ``partial_fit`` (``:270-297``) -- forward,
weighted cross-entropy + L2, backward, Adam -- and ``_forward_probs`` (``:332-370``).  What stays on
the host, bit-identical to the reference: class bookkeeping, the ``xavier_uniform_`` initialisation
drawn from torch's CPU generator after ``torch.manual_seed(random_state)`` (``:62-73,175-182``) and
the shuffle order from ``np.random.default_rng(random_state)`` (``:143-160,257-261``).

Data parallelism (one process per GPU, ``torch.distributed`` for the rendezvous, NCCL all-reduce of
the flat gradient inside the C library):

* ``dp_mode="parity"``  -- every rank is given the same ``X``/``y``; each global mini-batch of
  ``batch_size`` rows is split into contiguous per-rank slices, so the Adam trajectory is the
  reference's (up to fp32 summation order).
* ``dp_mode="throughput"`` -- every rank is given its own shard; each rank contributes a local
  mini-batch of ``batch_size`` rows per step (global mini-batch = ``batch_size * world``).
"""

from __future__ import annotations

import ctypes as C
import warnings
from collections.abc import Sequence
from typing import Any

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .inference import DeviceHead

_EXPECTED_FP_DRIFT_TOL = 1e-4


class _MLPModule(nn.Module):
    """Host-side parameter container with the reference's module layout (``linears.{i}``)."""

    def __init__(self, n_features_in: int, hidden_layer_sizes: Sequence[int], n_outputs: int):
        super().__init__()
        sizes = [n_features_in, *hidden_layer_sizes, n_outputs]
        self.linears = nn.ModuleList([nn.Linear(i, o) for i, o in zip(sizes[:-1], sizes[1:])])
        for lin in self.linears:
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)


def pack_reference_state(ws: list[np.ndarray], bs: list[np.ndarray], m_w, m_b, v_w, v_b, t: int, *, lr: float,
                         betas: tuple[float, float], eps: float) -> tuple[dict[str, torch.Tensor], dict[str, Any]]:
    """Device parameters and Adam moments -> the two pickle entries of the reference classifier.

    ``_module_state`` is ``_MLPModule.state_dict()`` (keys ``linears.{i}.weight`` / ``.bias``) and
    ``_optimizer_state`` is ``torch.optim.Adam.state_dict()`` over the parameters in module order
    (reference ``torch_classifier.py:411-420``), so a pickle written here restores into the reference class
    and the other way round."""
    module_state: dict[str, torch.Tensor] = {}
    params = []
    for i, (w, b) in enumerate(zip(ws, bs)):
        module_state[f"linears.{i}.weight"] = torch.from_numpy(np.array(w, dtype=np.float32, copy=True))
        module_state[f"linears.{i}.bias"] = torch.from_numpy(np.array(b, dtype=np.float32, copy=True))
        params += [nn.Parameter(torch.empty(w.shape)), nn.Parameter(torch.empty(b.shape))]
    # param_groups (hyper-parameters and torch-version-specific flags) come from a real Adam over same-shaped parameters
    opt_state = torch.optim.Adam(params, lr=lr, betas=tuple(betas), eps=eps).state_dict()
    if t > 0:
        moments = [x for pair in zip(zip(m_w, v_w), zip(m_b, v_b)) for x in pair]   # (m, v) per parameter, module order
        opt_state["state"] = {
            j: {"step": torch.tensor(float(t)),
                "exp_avg": torch.from_numpy(np.array(m, dtype=np.float32, copy=True)),
                "exp_avg_sq": torch.from_numpy(np.array(v, dtype=np.float32, copy=True))}
            for j, (m, v) in enumerate(moments)}
    return module_state, opt_state


def unpack_reference_state(module_state: dict[str, Any], optimizer_state: dict[str, Any] | None):
    """Inverse of :func:`pack_reference_state`; also accepts the flat layout this class wrote in round 1."""
    if "weights" in module_state:   # round-1 layout
        ws, bs = list(module_state["weights"]), list(module_state["biases"])
        if optimizer_state is None:
            return ws, bs, None
        return ws, bs, (optimizer_state["m_w"], optimizer_state["m_b"], optimizer_state["v_w"], optimizer_state["v_b"],
                        int(optimizer_state["t"]))
    n = len(module_state) // 2
    to_np = lambda x: np.ascontiguousarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=np.float32)
    ws = [to_np(module_state[f"linears.{i}.weight"]) for i in range(n)]
    bs = [to_np(module_state[f"linears.{i}.bias"]) for i in range(n)]
    if optimizer_state is None:
        return ws, bs, None
    st = optimizer_state.get("state", {})
    if not st:   # optimizer created, no step taken yet
        zeros = lambda xs: [np.zeros_like(x) for x in xs]
        return ws, bs, (zeros(ws), zeros(bs), zeros(ws), zeros(bs), 0)
    order = optimizer_state["param_groups"][0]["params"]
    ent = [st[j] for j in order]
    steps = {int(float(e["step"])) for e in ent}
    if len(steps) != 1:
        raise ValueError("Adam state with per-parameter step counts cannot be restored into the flat device optimizer")
    m = [to_np(e["exp_avg"]) for e in ent]
    v = [to_np(e["exp_avg_sq"]) for e in ent]
    return ws, bs, (m[0::2], m[1::2], v[0::2], v[1::2], steps.pop())


def split_steps(n_samples: int, batch_size: int, rank: int = 0, world: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """Rows of the shuffled order this rank trains on, and the per-step offsets into them.

    Global mini-batch ``s`` covers shuffled positions ``[s*B, min((s+1)*B, n))``; rank ``r`` takes
    the ``r``-th of ``world`` contiguous, near-equal slices of it (possibly empty in a ragged tail).
    Returns ``(positions, offsets)`` with ``positions`` indices into the shuffled order."""
    pos: list[np.ndarray] = []
    offsets = [0]
    for start in range(0, n_samples, batch_size):
        end = min(start + batch_size, n_samples)
        m = end - start
        lo = start + (m * rank) // world
        hi = start + (m * (rank + 1)) // world
        pos.append(np.arange(lo, hi, dtype=np.int64))
        offsets.append(offsets[-1] + (hi - lo))
    return (np.concatenate(pos) if pos else np.zeros(0, np.int64)), np.asarray(offsets, dtype=np.int64)


class DataParallel:
    """NCCL communicator of the C library, bootstrapped through ``torch.distributed``."""

    def __init__(self, group=None, device: int | None = None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before enabling data parallelism")
        lib = _lib.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else int(device)
        buf = (C.c_char * 128)()
        if self.rank == 0:
            _lib.check(lib.mc_dp_unique_id(buf))
        ids = [bytes(buf) if self.rank == 0 else None]
        dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        h = C.c_void_p()
        _lib.check(lib.mc_dp_create(ids[0], self.rank, self.world, self.device, C.byref(h)))
        self._h = h
        self.group = group

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().mc_dp_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class TorchMLPClassifier:
    """GPU-trained MLP classifier with the reference's sklearn-style surface."""

    _estimator_type = "classifier"

    def __init__(
        self,
        hidden_layer_sizes: Sequence[int] = (100,),
        activation: str = "relu",
        solver: str = "adam",
        alpha: float = 0.0001,
        batch_size: int | str = "auto",
        learning_rate_init: float = 0.001,
        max_iter: int = 200,
        shuffle: bool = True,
        random_state: int | None = None,
        tol: float = 1e-4,
        beta_1: float = 0.9,
        beta_2: float = 0.999,
        epsilon: float = 1e-8,
        class_weight: dict[str, float] | None = None,
    ):
        if activation != "relu":
            raise ValueError(f"TorchMLPClassifier only supports activation='relu', got {activation!r}.")
        if solver != "adam":
            raise ValueError(f"TorchMLPClassifier only supports solver='adam', got {solver!r}.")
        self.hidden_layer_sizes = tuple(hidden_layer_sizes)
        self.activation = activation
        self.solver = solver
        self.alpha = alpha
        self.batch_size = batch_size
        self.learning_rate_init = learning_rate_init
        self.max_iter = max_iter
        self.shuffle = shuffle
        self.random_state = random_state
        self.tol = tol
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.epsilon = epsilon
        self.class_weight = class_weight

    # -- non-sklearn knobs (not part of get_params) ---------------------------------------
    _device: int | None = None
    _dp: DataParallel | None = None
    _dp_mode: str = "parity"

    def set_device(self, device: int) -> "TorchMLPClassifier":
        self._device = int(device)
        return self

    def enable_data_parallel(self, dp: DataParallel, mode: str = "parity") -> "TorchMLPClassifier":
        if mode not in ("parity", "throughput"):
            raise ValueError("dp mode must be 'parity' or 'throughput'")
        self._dp = dp
        self._dp_mode = mode
        return self

    # -- helpers shared with the reference ----------------------------------------------------
    def _resolve_batch_size(self, n_samples: int) -> int:
        if self.batch_size == "auto":
            return min(200, n_samples)
        return min(int(self.batch_size), n_samples)

    def _seed_rng(self) -> np.random.Generator:
        if self.random_state is not None:
            return np.random.default_rng(int(self.random_state))
        if not hasattr(self, "_none_rng"):
            self._none_rng = np.random.default_rng(np.random.randint(0, np.iinfo(np.int32).max))
        return self._none_rng

    def _labels_to_indices(self, y: np.ndarray) -> np.ndarray:
        y = np.asarray(y)
        idx = np.searchsorted(self.classes_, y)
        missing = idx >= len(self.classes_)
        if missing.any() or not np.array_equal(self.classes_[idx], y):
            bad = set(np.asarray(y).tolist()) - set(self.classes_.tolist())
            raise ValueError(
                f"Labels {sorted(bad)} are not in classes_ {self.classes_.tolist()}."
                f" Pass all classes to the first partial_fit call.")
        return idx

    def _build_class_weight(self) -> np.ndarray | None:
        if self.class_weight is None:
            return None
        weights: list[float] = []
        for cls in self.classes_:
            if cls not in self.class_weight:
                bad = sorted(set(self.classes_.tolist()) - set(self.class_weight))
                raise ValueError(f"class_weight is missing weights for {bad!r}. Pass weights for every class in classes_.")
            w = float(self.class_weight[cls])
            if w < 0:
                raise ValueError(f"class_weight for {cls!r} is negative ({w!r}); weights must be >= 0.")
            weights.append(w)
        return np.asarray(weights, dtype=np.float32)

    # -- device handle ------------------------------------------------------------------------
    def _dims(self) -> list[int]:
        return [self.n_features_in_, *self.hidden_layer_sizes, len(self.classes_)]

    def _create_handle(self, weights: list[np.ndarray], biases: list[np.ndarray]) -> None:
        t = _lib.require_cuda()
        lib = _lib.load()
        dev = t.cuda.current_device() if self._device is None else self._device
        dims = self._dims()
        n = len(weights)
        self._keep = ([np.ascontiguousarray(w, dtype=np.float32) for w in weights],
                      [np.ascontiguousarray(b, dtype=np.float32) for b in biases])
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in self._keep[0]])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in self._keep[1]])
        cw = self._build_class_weight()
        self._class_weight_array = cw
        h = C.c_void_p()
        _lib.check(lib.mc_mlp_create(n, (C.c_int32 * (n + 1))(*dims), wp, bp, cw.ctypes.data if cw is not None else None,
                                     float(self.learning_rate_init), float(self.alpha), float(self.beta_1),
                                     float(self.beta_2), float(self.epsilon), dev, C.byref(h)))
        self._h = h
        self._dev_index = dev
        self._host_module = None
        self._head = None

    def _init_module(self) -> None:
        """Reference ``_init_module`` (``:175-182``): seed torch, build the layers (their default
        init consumes RNG), then xavier_uniform_ weights / zero biases -- on the host, then upload."""
        if self.random_state is not None:
            torch.manual_seed(int(self.random_state))
        mod = _MLPModule(self.n_features_in_, self.hidden_layer_sizes, len(self.classes_))
        self._create_handle([lin.weight.detach().numpy() for lin in mod.linears],
                            [lin.bias.detach().numpy() for lin in mod.linears])

    def _pull_params(self) -> tuple[list[np.ndarray], list[np.ndarray]]:
        dims = self._dims()
        ws = [np.empty((dims[i + 1], dims[i]), dtype=np.float32) for i in range(len(dims) - 1)]
        bs = [np.empty((dims[i + 1],), dtype=np.float32) for i in range(len(dims) - 1)]
        n = len(ws)
        _lib.check(_lib.load().mc_mlp_get_params(self._h, (C.c_void_p * n)(*[w.ctypes.data for w in ws]),
                                                 (C.c_void_p * n)(*[b.ctypes.data for b in bs])))
        return ws, bs

    @property
    def _module(self) -> _MLPModule:
        """Host copy of the trained network in the reference's layout (read by export/head code)."""
        if not hasattr(self, "_h"):
            raise AttributeError("_module")
        if self._host_module is None:
            ws, bs = self._pull_params()
            mod = _MLPModule(self.n_features_in_, self.hidden_layer_sizes, len(self.classes_))
            with torch.no_grad():
                for lin, w, b in zip(mod.linears, ws, bs):
                    lin.weight.copy_(torch.from_numpy(w))
                    lin.bias.copy_(torch.from_numpy(b))
            self._host_module = mod.eval()
        return self._host_module

    @property
    def launches(self) -> int:
        n = int(_lib.load().mc_mlp_launches(self._h)) if hasattr(self, "_h") else 0
        return n + (self._head.launches if getattr(self, "_head", None) is not None else 0)

    @property
    def graph_steps_(self) -> int:
        """Adam steps taken by CUDA-graph replay (runs of >= 8 equal-sized mini-batches of one process; ``MC_MLP_GRAPH=0``
        launches every step kernel by kernel)."""
        return int(_lib.load().mc_mlp_graph_steps(self._h)) if hasattr(self, "_h") else 0

    @property
    def n_steps_(self) -> int:
        return int(_lib.load().mc_mlp_steps(self._h)) if hasattr(self, "_h") else 0

    # -- training ----------------------------------------------------------------------------------
    def _first_call(self, n_features: int, y, classes) -> None:
        self.classes_ = np.unique(np.asarray(y)) if classes is None else np.unique(np.asarray(classes))
        self.n_features_in_ = int(n_features)
        self.n_iter_ = 0
        self.loss_curve_ = []
        self._init_module()

    def partial_fit(self, X, y, classes: Sequence[Any] | None = None) -> "TorchMLPClassifier":
        X_arr = np.asarray(X, dtype=np.float32)
        if X_arr.ndim != 2:
            raise ValueError(f"X must be 2D, got shape {X_arr.shape}")
        if not hasattr(self, "_h"):
            self._first_call(X_arr.shape[1], y, classes)
        elif X_arr.shape[1] != self.n_features_in_:
            raise ValueError(f"X has {X_arr.shape[1]} features, expected {self.n_features_in_}")
        y_idx = self._labels_to_indices(np.asarray(y)).astype(np.int32)
        t = _lib.require_cuda()
        with t.cuda.device(self._dev_index):
            xd = t.from_numpy(np.ascontiguousarray(X_arr)).cuda()
            yd = t.from_numpy(y_idx).cuda()
            return self.partial_fit_device(xd, yd)

    def partial_fit_device(self, X_dev: torch.Tensor, y_idx_dev: torch.Tensor) -> "TorchMLPClassifier":
        """One pass over device-resident data: ``X_dev`` CUDA fp32 ``(n, n_features)``, ``y_idx_dev``
        CUDA int32 class indices (positions in ``classes_``).  Requires a prior ``partial_fit`` call or
        :meth:`init_for` so the classes are known."""
        if not hasattr(self, "_h"):
            raise RuntimeError("call partial_fit (host data) or init_for(...) first so classes_ is known")
        if X_dev.dtype != torch.float32 or X_dev.dim() != 2 or not X_dev.is_contiguous() or X_dev.shape[1] != self.n_features_in_:
            raise ValueError(f"X must be contiguous CUDA float32 (n, {self.n_features_in_})")
        if y_idx_dev.dtype != torch.int32 or y_idx_dev.shape != (X_dev.shape[0],):
            raise ValueError("y must be CUDA int32 class indices, one per row")
        n = int(X_dev.shape[0])
        mb = self._resolve_batch_size(n)
        order = np.arange(n)
        if self.shuffle:
            self._seed_rng().shuffle(order)
        dp = self._dp
        if dp is not None and dp.world > 1 and self._dp_mode == "parity":
            pos, offsets = split_steps(n, mb, dp.rank, dp.world)
            local_order = order[pos]
        else:
            local_order = order
            offsets = np.asarray(list(range(0, n, mb)) + [n], dtype=np.int64) if n else np.zeros(1, np.int64)
            if dp is not None and dp.world > 1:
                import torch.distributed as dist

                steps = torch.tensor([len(offsets) - 1], dtype=torch.int64, device=X_dev.device)
                dist.all_reduce(steps, op=dist.ReduceOp.MAX, group=dp.group)
                pad = int(steps.item()) - (len(offsets) - 1)
                offsets = np.concatenate([offsets, np.full(pad, offsets[-1], dtype=np.int64)])
        loss = C.c_double(0.0)
        with torch.cuda.device(self._dev_index):
            need_order = self.shuffle or local_order.shape[0] != n
            od = torch.from_numpy(np.ascontiguousarray(local_order, dtype=np.int64)).cuda() if need_order else None
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            _lib.check(_lib.load().mc_mlp_partial_fit(
                self._h, X_dev.data_ptr(), y_idx_dev.data_ptr(), od.data_ptr() if od is not None else None,
                offsets.ctypes.data, len(offsets) - 1, dp._h if dp is not None else None,
                self._grad_hook_ptr(), None, C.byref(loss), _lib.stream_ptr()))
        self._host_module = None
        self._head = None
        self.loss_curve_.append(float(loss.value))
        self.n_iter_ += 1
        return self

    _grad_hook = None  # optional GRAD_SYNC_FN instance (tests use it to read the flat gradient buffer)

    def _grad_hook_ptr(self):
        return C.cast(self._grad_hook, C.c_void_p) if self._grad_hook is not None else None

    def init_for(self, n_features: int, classes: Sequence[Any]) -> "TorchMLPClassifier":
        """Prepare for device-resident training without a host batch (what the first
        ``partial_fit`` call does at ``torch_classifier.py:236-248``)."""
        if not hasattr(self, "_h"):
            self._first_call(n_features, classes, classes)
        return self

    def fit(self, X, y) -> "TorchMLPClassifier":
        y_arr = np.asarray(y)
        classes = np.unique(y_arr).tolist()
        self._release()
        for attr in ("classes_", "n_features_in_", "n_iter_", "loss_curve_"):
            if hasattr(self, attr):
                delattr(self, attr)
        prev = np.inf
        for _ in range(self.max_iter):
            self.partial_fit(X, y_arr, classes=classes)
            cur = self.loss_curve_[-1]
            if abs(prev - cur) < self.tol:
                break
            prev = cur
        return self

    # -- inference -----------------------------------------------------------------------------------
    def _ensure_head(self) -> DeviceHead:
        """Uncalibrated device head over the current parameters (rebuilt after every training pass)."""
        if not hasattr(self, "_h"):
            raise RuntimeError("TorchMLPClassifier is not fitted. Call partial_fit or fit before predict/predict_proba.")
        if self._head is None:
            ws, bs = self._pull_params()
            self._head = DeviceHead(ws, bs, None, None, device=self._dev_index)
        return self._head

    def evaluate_device(self, X_dev: torch.Tensor, y_idx_dev: torch.Tensor) -> tuple[int, float]:
        """``(n_correct, sum of log-loss terms)`` on device-resident data: the two accumulators behind
        ``_calc_acc_batched`` / ``_calc_acc_and_log_loss_batched`` (``trainer.py:295-342``), no ``(N, K)`` matrix."""
        return self._ensure_head().evaluate_device(X_dev, y_idx_dev)

    def predict_proba_device(self, X_dev: torch.Tensor) -> torch.Tensor:
        """``predict_proba`` that stays on the device: CUDA float64 ``(n, K)``."""
        return self._ensure_head().scores_device(X_dev, want_proba=True)["proba"]

    def _forward_probs(self, X) -> np.ndarray:
        if not hasattr(self, "_h"):
            raise RuntimeError("TorchMLPClassifier is not fitted. Call partial_fit or fit before predict/predict_proba.")
        X_arr = np.asarray(X, dtype=np.float32)
        if X_arr.ndim != 2:
            raise ValueError(f"X must be 2D, got shape {X_arr.shape}")
        if X_arr.shape[1] != self.n_features_in_:
            raise ValueError(f"X has {X_arr.shape[1]} features, expected {self.n_features_in_}")
        proba, _ = self._ensure_head().scores_host(np.ascontiguousarray(X_arr), want_proba=True, want_labels=False)
        row_sums = proba.sum(axis=1)
        max_drift = float(np.max(np.abs(row_sums - 1.0))) if proba.size else 0.0
        if max_drift > _EXPECTED_FP_DRIFT_TOL:
            warnings.warn(
                f"predict_proba row sums deviate from 1.0 by up to {max_drift:.2e}, exceeding the expected float32 "
                f"softmax drift bound ({_EXPECTED_FP_DRIFT_TOL:.0e}).", RuntimeWarning, stacklevel=2)
        return proba

    def predict_proba(self, X) -> np.ndarray:
        return self._forward_probs(X)

    def predict(self, X) -> np.ndarray:
        return self.classes_[np.argmax(self._forward_probs(X), axis=1)]

    # -- sklearn parameter protocol -------------------------------------------------------------------
    def get_params(self, deep: bool = True) -> dict[str, Any]:
        return {
            "hidden_layer_sizes": self.hidden_layer_sizes, "activation": self.activation, "solver": self.solver,
            "alpha": self.alpha, "batch_size": self.batch_size, "learning_rate_init": self.learning_rate_init,
            "max_iter": self.max_iter, "shuffle": self.shuffle, "random_state": self.random_state, "tol": self.tol,
            "beta_1": self.beta_1, "beta_2": self.beta_2, "epsilon": self.epsilon,
            "class_weight": getattr(self, "class_weight", None),
        }

    def set_params(self, **params: Any) -> "TorchMLPClassifier":
        for key, value in params.items():
            if not hasattr(self, key):
                raise ValueError(f"Invalid parameter {key!r} for TorchMLPClassifier")
            setattr(self, key, value)
        return self

    # -- pickling: parameters + Adam state travel as host arrays --------------------------------------------
    def _release(self) -> None:
        if hasattr(self, "_h"):
            _lib.load().mc_mlp_destroy(self._h)
            del self._h
        self._head = None
        self._host_module = None

    def __del__(self):  # pragma: no cover
        try:
            self._release()
        except Exception:
            pass

    def __getstate__(self) -> dict[str, Any]:
        state = {k: v for k, v in self.__dict__.items()
                 if k not in ("_h", "_head", "_host_module", "_keep", "_dp", "_class_weight_array", "_grad_hook")}
        if hasattr(self, "_h"):
            ws, bs = self._pull_params()
            dims = self._dims()
            n = len(ws)
            mk = lambda: ([np.empty((dims[i + 1], dims[i]), np.float32) for i in range(n)],
                          [np.empty((dims[i + 1],), np.float32) for i in range(n)])
            (mw, mb_), (vw, vb) = mk(), mk()
            tcount = C.c_int64(0)
            arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
            _lib.check(_lib.load().mc_mlp_get_adam(self._h, arr(mw), arr(mb_), arr(vw), arr(vb), C.byref(tcount)))
            state["_module_state"], state["_optimizer_state"] = pack_reference_state(
                ws, bs, mw, mb_, vw, vb, int(tcount.value), lr=self.learning_rate_init,
                betas=(self.beta_1, self.beta_2), eps=self.epsilon)
        return state

    def __setstate__(self, state: dict[str, Any]) -> None:
        module_state = state.pop("_module_state", None)
        opt = state.pop("_optimizer_state", None)
        self.__dict__.update(state)
        self.__dict__.setdefault("class_weight", None)
        if module_state is not None:
            ws, bs, adam = unpack_reference_state(module_state, opt)
            self._create_handle(ws, bs)
            if adam is not None:
                n = len(ws)
                keep = [[np.ascontiguousarray(x, dtype=np.float32) for x in xs] for xs in adam[:4]]
                ptrs = [(C.c_void_p * n)(*[x.ctypes.data for x in ks]) for ks in keep]
                _lib.check(_lib.load().mc_mlp_set_adam(self._h, *ptrs, int(adam[4])))
