"""Serve-time loader and scorer for the portable classifier artifact (``model.pt`` + ``model.json``).

Drop-in for ``mermaid_classifier.pyspacer.inference.loader`` of the reference
(``/root/reference/mermaid_classifier/pyspacer/inference/loader.py:16-75``): same
``load_predictor(model_pt_path, model_json_path) -> Predictor``, same ``Predictor.predict_proba``
contract (fp32 ``(N, input_dim)`` in, fp64 ``(N, K)`` out, ``ValueError`` on a bad shape), same
``ManifestError`` conditions (schema version, class count, input_dim).  The Linear/ReLU chain,
softmax, per-class Platt sigmoid, row normalisation and overshoot clip of ``CalibratedHead.forward``
(``inference/head.py:66-89``) run in ``libmermaid_b200`` on the GPU; the TorchScript graph is only
read for its constants and for a 4-row load-time self-check.
"""

from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Any

import numpy as np
import torch

from . import _lib

SCHEMA_VERSION = 1
TASK_NAME = "pyspacer_mlp_classifier"


class ManifestError(Exception):
    """model.json is incompatible with the graph (schema version, class count, input_dim)."""


class ParityError(Exception):
    """The device head diverges from the TorchScript graph beyond tolerance."""


def _const(value):
    v = value.toIValue()
    if not isinstance(v, torch.Tensor):
        raise ManifestError("expected a tensor constant in the frozen graph")
    return v.detach().float().contiguous()


def extract_head_params(graph: Any):
    """``(weights, biases, a, b)`` from a TorchScript ``CalibratedHead``.

    Frozen artifacts (``export.py:54-57``) carry the parameters as graph constants: each
    ``aten::linear(x, W, b)`` in order, then ``aten::mul(a, softmax)`` and ``aten::add(., b)``
    (``head.py:69-76``).  Unfrozen scripted modules expose ``linears.{i}.weight/bias`` and the
    ``a``/``b`` buffers (``head.py:55-64``)."""
    names = dict(graph.named_parameters()) if hasattr(graph, "named_parameters") else {}
    if names:
        bufs = dict(graph.named_buffers())
        n = len([k for k in names if k.endswith(".weight")])
        ws = [names[f"linears.{i}.weight"].detach().float().contiguous() for i in range(n)]
        bs = [names[f"linears.{i}.bias"].detach().float().contiguous() for i in range(n)]
        return ws, bs, bufs["a"].detach().float().contiguous(), bufs["b"].detach().float().contiguous()
    g = graph.graph

    def tensor_consts(ins):
        return [i for i in ins if i.node().kind() == "prim::Constant" and isinstance(i.toIValue(), torch.Tensor)]

    ws, bs, a, b = [], [], None, None
    softmax_name = None
    for node in g.nodes():
        kind = node.kind()
        ins = list(node.inputs())
        if kind == "aten::linear":
            ws.append(_const(ins[1]))
            bs.append(_const(ins[2]))
        elif kind == "aten::softmax":
            softmax_name = node.output().debugName()
        elif kind == "aten::mul" and softmax_name is not None and a is None:
            if any(i.debugName() == softmax_name for i in ins) and tensor_consts(ins):
                a = _const(tensor_consts(ins)[0])
        elif kind == "aten::add" and a is not None and b is None and tensor_consts(ins):
            b = _const(tensor_consts(ins)[0])
    if not ws or a is None or b is None:
        raise ManifestError("model.pt is not a CalibratedHead graph this loader understands")
    return ws, bs, a.reshape(-1), b.reshape(-1)


class DeviceHead:
    """Owns an ``mc_head`` handle.  ``a``/``b`` None -> uncalibrated softmax path."""

    def __init__(self, weights, biases, a=None, b=None, device: int | None = None):
        t = _lib.require_cuda()
        lib = _lib.load()
        self.device = t.cuda.current_device() if device is None else int(device)
        self._w = [np.ascontiguousarray(np.asarray(w, dtype=np.float32)) for w in weights]
        self._b = [np.ascontiguousarray(np.asarray(x, dtype=np.float32)) for x in biases]
        dims = [self._w[0].shape[1]] + [w.shape[0] for w in self._w]
        for i, w in enumerate(self._w):
            if w.shape != (dims[i + 1], dims[i]) or self._b[i].shape != (dims[i + 1],):
                raise ValueError("inconsistent layer shapes")
        self.dims = dims
        n = len(self._w)
        dims_c = (C.c_int32 * (n + 1))(*dims)
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in self._w])
        bp = (C.c_void_p * n)(*[x.ctypes.data for x in self._b])
        pa = pb = None
        if a is not None:
            self._a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
            self._pb = np.ascontiguousarray(np.asarray(b, dtype=np.float32).reshape(-1))
            if self._a.shape != (dims[-1],) or self._pb.shape != (dims[-1],):
                raise ValueError("Platt parameters must have one entry per class")
            pa, pb = self._a.ctypes.data, self._pb.ctypes.data
        h = C.c_void_p()
        _lib.check(lib.mc_head_create(n, dims_c, wp, bp, pa, pb, self.device, C.byref(h)))
        self._h = h

    @property
    def n_classes(self) -> int:
        return self.dims[-1]

    @property
    def launches(self) -> int:
        return int(_lib.load().mc_head_launches(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().mc_head_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _set_exact(self, exact: bool) -> None:
        _lib.check(_lib.load().mc_head_set_exact(self._h, 1 if exact else 0))

    def scores_host(self, arr: np.ndarray, want_proba=True, want_labels=True, exact: bool | None = None):
        """``exact`` (default: whenever probabilities are returned) runs the Linear chain on the exact-fp32
        CUDA-core GEMM -- ``predict_proba`` keeps the reference's 1e-6 export tolerance -- otherwise on the
        tensor cores (tcgen05, 3xTF32 split, error ~2^-21)."""
        self._set_exact(want_proba if exact is None else exact)
        n = arr.shape[0]
        proba = np.empty((n, self.n_classes), dtype=np.float64) if want_proba else None
        labels = np.empty((n,), dtype=np.int32) if want_labels else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().mc_head_scores_host(
                self._h, arr.ctypes.data, n, proba.ctypes.data if want_proba else None,
                labels.ctypes.data if want_labels else None, _lib.stream_ptr()))
        return proba, labels

    def scores_device(self, feats: torch.Tensor, want_proba=False, topk: int = 0, exact: bool | None = None):
        """CUDA fp32 ``(n, input_dim)`` -> dict of CUDA tensors (labels / proba / topk).  Same rule as
        :meth:`scores_host`: exact chain when probabilities are returned (unless overridden), tensor cores otherwise."""
        self._set_exact(want_proba if exact is None else exact)
        if feats.dtype != torch.float32 or feats.dim() != 2 or feats.shape[1] != self.dims[0] or not feats.is_contiguous():
            raise ValueError(f"features must be contiguous CUDA float32 (N, {self.dims[0]}); got {tuple(feats.shape)}")
        n = feats.shape[0]
        out = {"labels": torch.empty((n,), dtype=torch.int32, device=feats.device)}
        if want_proba:
            out["proba"] = torch.empty((n, self.n_classes), dtype=torch.float64, device=feats.device)
        if topk:
            out["topk_idx"] = torch.empty((n, topk), dtype=torch.int32, device=feats.device)
            out["topk_val"] = torch.empty((n, topk), dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.check(_lib.load().mc_head_scores(
                self._h, feats.data_ptr(), n, out["proba"].data_ptr() if want_proba else None, out["labels"].data_ptr(),
                topk, out["topk_idx"].data_ptr() if topk else None, out["topk_val"].data_ptr() if topk else None,
                _lib.stream_ptr()))
        return out

    def evaluate_device(self, feats: torch.Tensor, y_idx: torch.Tensor, exact: bool = True) -> tuple[int, float]:
        """``(n_correct, loss_sum)`` of this head against int32 targets, reduced on the device
        (``trainer.py:295-342``: accuracy of the argmax labels and the un-averaged ``log_loss`` terms)."""
        self._set_exact(exact)
        if feats.dtype != torch.float32 or feats.dim() != 2 or feats.shape[1] != self.dims[0] or not feats.is_contiguous():
            raise ValueError(f"features must be contiguous CUDA float32 (N, {self.dims[0]}); got {tuple(feats.shape)}")
        if y_idx.dtype != torch.int32 or y_idx.shape != (feats.shape[0],) or not y_idx.is_contiguous():
            raise ValueError("targets must be contiguous CUDA int32 class indices, one per row")
        hits, loss = C.c_int64(0), C.c_double(0.0)
        with torch.cuda.device(feats.device):
            _lib.check(_lib.load().mc_head_evaluate(self._h, feats.data_ptr(), y_idx.data_ptr(), feats.shape[0],
                                                    C.byref(hits), C.byref(loss), _lib.stream_ptr()))
        return int(hits.value), float(loss.value)


def platt_fit_device(proba: torch.Tensor, y_idx: torch.Tensor, gtol: float = 1e-9, max_passes: int = 400):
    """Per-class Platt parameters ``(a, b, loss, passes)`` from a CUDA float64 ``(n, K)`` probability matrix and
    int32 targets -- every class's ``_sigmoid_calibration(proba[:, k], y == k)`` at once (``trainer.py:344-396``)."""
    if proba.dtype != torch.float64 or proba.dim() != 2 or not proba.is_contiguous() or not proba.is_cuda:
        raise ValueError("proba must be a contiguous CUDA float64 (n, K) matrix")
    if y_idx.dtype != torch.int32 or y_idx.shape != (proba.shape[0],) or not y_idx.is_contiguous():
        raise ValueError("targets must be contiguous CUDA int32 class indices, one per row")
    n, k = proba.shape
    a, b, loss = (np.empty(k, dtype=np.float64) for _ in range(3))
    passes = C.c_int32(0)
    with torch.cuda.device(proba.device):
        _lib.check(_lib.load().mc_platt_fit(proba.data_ptr(), y_idx.data_ptr(), n, k, proba.device.index, float(gtol),
                                            int(max_passes), a.ctypes.data, b.ctypes.data, loss.ctypes.data,
                                            C.byref(passes), _lib.stream_ptr()))
    return a, b, loss, int(passes.value)


class Predictor:
    """A loaded classifier head: feature batch -> calibrated probabilities (on the GPU)."""

    def __init__(self, head: DeviceHead, classes: list[str], input_dim: int) -> None:
        self._head = head
        self.classes = classes
        self.input_dim = input_dim

    @property
    def classes_(self) -> list[str]:
        return self.classes

    def _check(self, features: Any) -> np.ndarray:
        arr = np.ascontiguousarray(np.asarray(features, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != self.input_dim:
            raise ValueError(f"features must be (N, {self.input_dim}); got {arr.shape}.")
        return arr

    def predict_proba(self, features: Any) -> np.ndarray:
        proba, _ = self._head.scores_host(self._check(features), want_proba=True, want_labels=False)
        return proba

    def predict_indices(self, features: Any) -> np.ndarray:
        """``argmax(axis=1)`` of predict_proba, computed on the device (no N x K transfer)."""
        _, labels = self._head.scores_host(self._check(features), want_proba=False, want_labels=True)
        return labels

    def predict(self, features: Any) -> np.ndarray:
        return np.asarray(self.classes, dtype=object)[self.predict_indices(features)]

    def predict_topk(self, features: Any, k: int):
        """Top-k ``(labels, scores)`` per row, descending, ties in class order (annotation.py:252-261)."""
        arr = self._check(features)
        with torch.cuda.device(self._head.device):
            out = self._head.scores_device(torch.from_numpy(arr).cuda(), topk=int(k))
            idx = out["topk_idx"].cpu().numpy()
            val = out["topk_val"].cpu().numpy().astype(np.float64)
        return np.asarray(self.classes, dtype=object)[idx], val

    def predict_device(self, feats: torch.Tensor, want_proba: bool = False, topk: int = 0):
        return self._head.scores_device(feats, want_proba=want_proba, topk=topk)


def load_predictor(model_pt_path: str | Path, model_json_path: str | Path, *, device: int | None = None,
                   self_check: bool = True) -> Predictor:
    """Load model.pt + model.json onto the GPU, validating compatibility loudly
    (``ManifestError`` on schema-version, class-count or input_dim mismatch)."""
    manifest = json.loads(Path(model_json_path).read_text())
    schema_version = manifest.get("schema_version")
    if schema_version != SCHEMA_VERSION:
        raise ManifestError(
            f"model.json schema_version={schema_version!r} is incompatible with this loader (expects {SCHEMA_VERSION}).")
    classes = manifest["classes"]
    input_dim = int(manifest["input_dim"])
    graph = torch.jit.load(str(model_pt_path), map_location="cpu")
    graph.eval()
    weights, biases, a, b = extract_head_params(graph)
    if weights[0].shape[1] != input_dim:
        raise ManifestError(
            f"graph rejects input_dim={input_dim} declared in model.json: first layer expects {weights[0].shape[1]}")
    if weights[-1].shape[0] != len(classes):
        raise ManifestError(
            f"class-count mismatch: graph outputs {weights[-1].shape[0]} classes but model.json declares {len(classes)}.")
    head = DeviceHead([w.numpy() for w in weights], [x.numpy() for x in biases], a.numpy(), b.numpy(), device=device)
    pred = Predictor(head, list(classes), input_dim)
    if self_check:
        gen = torch.Generator().manual_seed(0)
        probe = torch.cat([torch.zeros(1, input_dim), torch.rand(3, input_dim, generator=gen)])
        with torch.no_grad():
            want = graph(probe).numpy().astype(np.float64)
        got = pred.predict_proba(probe.numpy())
        diff = float(np.max(np.abs(want - got)))
        if not diff <= 1e-5:
            raise ParityError(f"device head diverges from model.pt on the load-time probe: max|d|={diff:.3e}")
    return pred
