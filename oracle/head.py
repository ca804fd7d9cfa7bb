"""Oracle: MLP / Platt head scoring and the MLP training step (TEST INFRASTRUCTURE).

Restates, in plain torch fp32 on the CPU:

* ``CalibratedHead.forward`` -- ``/root/reference/mermaid_classifier/pyspacer/inference/head.py:66-89``
* ``Predictor.predict_proba`` -- ``.../inference/loader.py:30-35`` (fp32 in, fp64 out)
* ``TorchMLPClassifier._forward_probs`` -- ``.../torch_classifier.py:332-370`` (softmax, fp64 renorm)
* ``TorchMLPClassifier.partial_fit`` -- ``.../torch_classifier.py:226-303`` (shuffle, mini-batch
  weighted CE + 0.5*alpha/mb*sum(W^2), Adam)

PINNED: ``tests/test_oracle_head.py`` checks every function here against golden vectors
produced by importing the reference itself (``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def mlp_logits(x: torch.Tensor, weights, biases) -> torch.Tensor:
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < n - 1:
            x = F.relu(x)
    return x


@torch.no_grad()
def calibrated_proba(features, weights, biases, a, b) -> np.ndarray:
    """head.py:66-89 then loader.py:35 -> ``(N, K) float64``."""
    x = torch.from_numpy(np.asarray(features, dtype=np.float32))
    p = F.softmax(mlp_logits(x, weights, biases), dim=1)
    c = torch.sigmoid(-(a * p + b))
    denom = c.sum(dim=1, keepdim=True)
    nonzero = denom != 0
    safe = torch.where(nonzero, denom, torch.ones_like(denom))
    proba = torch.where(nonzero, c / safe, torch.full_like(c, 1.0 / float(a.shape[0])))
    proba = torch.where((proba > 1.0) & (proba <= 1.0 + 1e-5), torch.ones_like(proba), proba)
    return proba.numpy().astype(np.float64)


@torch.no_grad()
def softmax_proba(features, weights, biases) -> np.ndarray:
    """torch_classifier.py:332-370 -> ``(N, K) float64`` rows summing to 1 in fp64."""
    x = torch.from_numpy(np.asarray(features, dtype=np.float32))
    p = F.softmax(mlp_logits(x, weights, biases), dim=1).numpy().astype(np.float64)
    return p / p.sum(axis=1)[:, None]


def argmax_labels(proba: np.ndarray) -> np.ndarray:
    """``np.argmax(axis=1)`` -- lowest index wins ties (torch_classifier.py:375-376)."""
    return np.argmax(proba, axis=1)


def topk_labels(proba: np.ndarray, k: int) -> np.ndarray:
    """``sorted(zip(labels, proba), key=itemgetter(1), reverse=True)[:k]``
    (annotation.py:253-259): stable sort, descending -> ties keep ascending class order."""
    order = np.argsort(-proba, axis=1, kind="stable")
    return order[:, :k]


# ---------------------------------------------------------------------------------
# Training step
# ---------------------------------------------------------------------------------


def init_mlp(n_features: int, hidden, n_classes: int, random_state: int | None):
    """``_init_module`` (torch_classifier.py:175-182): ``torch.manual_seed`` then, layer by
    layer, ``nn.Linear`` construction (whose default init consumes RNG) followed -- after ALL
    layers exist -- by ``xavier_uniform_`` on each weight and zero biases (``:62-73``)."""
    if random_state is not None:
        torch.manual_seed(int(random_state))
    dims = [n_features, *hidden, n_classes]
    layers = [torch.nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:])]
    for lin in layers:
        torch.nn.init.xavier_uniform_(lin.weight)
        torch.nn.init.zeros_(lin.bias)
    return [lin.weight.detach().clone() for lin in layers], [lin.bias.detach().clone() for lin in layers]


class AdamState:
    def __init__(self, params):
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0


def minibatch_sums(weights, biases, xb: torch.Tensor, yb: torch.Tensor, class_weight: torch.Tensor | None = None):
    """UN-NORMALISED sums over the rows of one (share of a) mini-batch -- what each data-parallel
    rank contributes to the all-reduce:  ``gW[i] = sum_r w_r * dNLL_r/dW_i``, ``gB[i]`` likewise,
    ``wsum = sum_r w_r``, ``lsum = sum_r w_r * NLL_r`` with ``w_r = class_weight[y_r]`` (1 when
    unweighted).  Manual backward so the oracle does not lean on autograd."""
    L = len(weights)
    m = xb.shape[0]
    if m == 0:
        return [torch.zeros_like(w) for w in weights], [torch.zeros_like(b) for b in biases], 0.0, 0.0
    acts = [xb]
    for i in range(L):
        z = acts[-1] @ weights[i].T + biases[i]
        acts.append(torch.relu(z) if i < L - 1 else z)
    logp = F.log_softmax(acts[-1], dim=1)
    w_r = torch.ones(m) if class_weight is None else class_weight[yb]
    lsum = float(-(w_r * logp[torch.arange(m), yb]).sum())
    delta = torch.exp(logp)
    delta[torch.arange(m), yb] -= 1.0
    delta = delta * w_r[:, None]
    gW, gB = [None] * L, [None] * L
    for i in reversed(range(L)):
        gW[i] = delta.T @ acts[i]
        gB[i] = delta.sum(0)
        if i > 0:
            delta = (delta @ weights[i]) * (acts[i] > 0).to(delta.dtype)
    return gW, gB, float(w_r.sum()), lsum


def adam_step(weights, biases, adam: AdamState, gW, gB, wsum: float, lsum: float, nrows: int, *, lr, alpha, beta_1,
              beta_2, epsilon) -> float:
    """Normalise the (all-reduced) sums like ``F.cross_entropy(weight=...)``'s weighted mean
    (torch_classifier.py:283), add the per-mini-batch L2 term (``:288``) and apply one
    ``torch.optim.Adam`` step (eps added to ``sqrt(v_hat)``).  Returns the regularised loss."""
    loss = lsum / wsum + 0.5 * alpha / nrows * float(sum((w**2).sum() for w in weights))
    grads = [g / wsum + (alpha / nrows) * w for g, w in zip(gW, weights)] + [g / wsum for g in gB]
    adam.t += 1
    bc1 = 1.0 - beta_1**adam.t
    bc2 = 1.0 - beta_2**adam.t
    for j, (p, g) in enumerate(zip(list(weights) + list(biases), grads)):
        adam.m[j].mul_(beta_1).add_(g, alpha=1 - beta_1)
        adam.v[j].mul_(beta_2).addcmul_(g, g, value=1 - beta_2)
        denom = (adam.v[j].sqrt() / (bc2**0.5)).add_(epsilon)
        p.addcdiv_(adam.m[j], denom, value=-lr / bc1)
    return loss


def rank_slice(start: int, end: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous near-equal share of shuffled positions [start, end) owned by ``rank``."""
    m = end - start
    return start + (m * rank) // world, start + (m * (rank + 1)) // world


def partial_fit(
    weights,
    biases,
    adam: AdamState,
    X: np.ndarray,
    y_idx: np.ndarray,
    *,
    lr: float = 1e-3,
    alpha: float = 1e-4,
    beta_1: float = 0.9,
    beta_2: float = 0.999,
    epsilon: float = 1e-8,
    batch_size: int | str = "auto",
    shuffle: bool = True,
    random_state: int | None = 0,
    class_weight: torch.Tensor | None = None,
    rank: int = 0,
    world: int = 1,
    all_reduce=None,
) -> float:
    """One ``partial_fit`` call; updates ``weights``/``biases``/``adam`` in place and returns the
    ``loss_curve_`` entry.  With ``world > 1`` this is ONE RANK of the data-parallel form: the rank
    trains on its slice of every global mini-batch and ``all_reduce(flat_tensor)`` must sum the
    flat buffer over ranks in place (gloo in the CPU tests, NCCL on the GPUs)."""
    X = np.asarray(X, dtype=np.float32)
    n = X.shape[0]
    mb = min(200, n) if batch_size == "auto" else min(int(batch_size), n)
    order = np.arange(n)
    if shuffle:
        np.random.default_rng(int(random_state)).shuffle(order)
    Xt = torch.from_numpy(X[order])
    yt = torch.from_numpy(np.asarray(y_idx)[order].astype(np.int64))
    total, seen = 0.0, 0
    for start in range(0, n, mb):
        end = min(start + mb, n)
        lo, hi = rank_slice(start, end, rank, world)
        gW, gB, wsum, lsum = minibatch_sums(weights, biases, Xt[lo:hi], yt[lo:hi], class_weight)
        nrows = hi - lo
        if world > 1:
            flat = torch.cat([g.reshape(-1) for g in gW + gB] + [torch.tensor([wsum, lsum, float(nrows)])])
            all_reduce(flat)
            off = 0
            for lst in (gW, gB):
                for j, g in enumerate(lst):
                    lst[j] = flat[off : off + g.numel()].reshape(g.shape)
                    off += g.numel()
            wsum, lsum, nrows = float(flat[off]), float(flat[off + 1]), int(round(float(flat[off + 2])))
        loss = adam_step(weights, biases, adam, gW, gB, wsum, lsum, nrows, lr=lr, alpha=alpha, beta_1=beta_1,
                         beta_2=beta_2, epsilon=epsilon)
        total += loss * nrows
        seen += nrows
    return total / max(seen, 1)
