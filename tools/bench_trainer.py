"""Trainer-level timings on one B200 (not the bench.py metric): device evaluation (accuracy + log-loss) and the
device Platt fit at production head size.  Usage: python tools/bench_trainer.py [--rows 1000000] [--ref 200000]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import synth  # noqa: E402
from mermaid_classifier_b200.inference import DeviceHead, platt_fit_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--ref", type=int, default=200_000)
    ap.add_argument("--classes", type=int, default=500)
    ap.add_argument("--full", action="store_true", help="also run a whole MermaidTrainer call on device-resident splits")
    ap.add_argument("--train-rows", type=int, default=1_000_000)
    ap.add_argument("--epochs", type=int, default=3)
    args = ap.parse_args()
    K = args.classes
    w, b, _, _, _ = synth.synth_head(input_dim=1280, hidden=(500, 300, 100), n_classes=K, seed=0)
    head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in b], None, None)
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(args.rows, 1280, device="cuda", generator=g)
    # targets: the head's own argmax with 30 % of the rows re-drawn at random (a trained head's confusion level)
    y = head.scores_device(X)["labels"]
    flip = torch.rand(args.rows, device="cuda", generator=g) < 0.3
    y = torch.where(flip, torch.randint(0, K, (args.rows,), device="cuda", generator=g, dtype=torch.int32), y).contiguous()
    out = {}
    for exact in (True, False):
        head.evaluate_device(X[:10000].contiguous(), y[:10000].contiguous(), exact=exact)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hits, loss = head.evaluate_device(X, y, exact=exact)
        dt = time.perf_counter() - t0
        out[f"evaluate_{'exact' if exact else 'tc'}"] = {"rows_per_s": args.rows / dt, "ms": dt * 1e3, "acc": hits / args.rows,
                                                          "log_loss": loss / args.rows}
    Xr, yr = X[:args.ref].contiguous(), y[:args.ref].contiguous()
    proba = head.scores_device(Xr, want_proba=True)["proba"]
    platt_fit_device(proba[:5000].contiguous(), yr[:5000].contiguous())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, bb, l, passes = platt_fit_device(proba, yr)
    dt = time.perf_counter() - t0
    out["platt_fit"] = {"rows": args.ref, "classes": K, "passes": passes, "ms": dt * 1e3,
                        "finite": bool(np.isfinite(a).all() and np.isfinite(bb).all())}
    # CPU comparison on a handful of classes (scipy L-BFGS-B, what sklearn runs per class)
    from oracle import trainer as otr
    pk, yk = proba[:, :4].cpu().numpy(), yr.cpu().numpy()
    t0 = time.perf_counter()
    ref = [otr.sigmoid_calibration(pk[:, k], (yk == k).astype(int)) for k in range(4)]
    cpu_per_class = (time.perf_counter() - t0) / 4
    out["platt_fit"]["cpu_s_per_class"] = cpu_per_class
    out["platt_fit"]["cpu_s_all_classes_est"] = cpu_per_class * K
    q = np.linspace(0, 1, 101)
    out["platt_fit"]["max_curve_diff_vs_lbfgsb"] = float(max(
        np.abs(1 / (1 + np.exp(a[k] * q + bb[k])) - 1 / (1 + np.exp(ref[k][0] * q + ref[k][1]))).max() for k in range(4)))
    if args.full:
        out["trainer"] = full_run(args)
    print(json.dumps(out))


def full_run(args):
    """Whole MermaidTrainer call on HBM-resident splits: Gaussian clusters as in the reference's
    tests/pyspacer/test_mlp_benchmark.py:41-63 (centroids x 3.0, std 1.3), production head (500, 300, 100)."""
    from mermaid_classifier_b200.trainer import DeviceLabels, MermaidTrainer, TaskLabels

    K = args.classes
    g = torch.Generator(device="cuda").manual_seed(42)
    centers = torch.randn(K, 1280, device="cuda", generator=g) * 3.0
    classes = np.asarray([f"class_{i:03d}" for i in range(K)])

    def split(n):
        yi = torch.randint(0, K, (n,), device="cuda", generator=g)
        X = centers[yi] + torch.randn(n, 1280, device="cuda", generator=g) * 1.3
        return DeviceLabels(X, classes[yi.cpu().numpy()])

    labels = TaskLabels(train=split(args.train_rows), ref=split(args.ref), val=split(args.ref))
    seen = []
    trainer = MermaidTrainer(batch_size=100_000, on_epoch_end=seen.append, early_stopping_patience=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    clf_cal, val_results, msg = trainer(labels, args.epochs, [])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"train_rows": args.train_rows, "ref_rows": args.ref, "val_rows": args.ref, "epochs_run": len(seen),
            "seconds": dt, "train_samples_per_s_overall": args.train_rows * len(seen) / dt,
            "val_loss": [m["val_loss"] for m in seen], "ref_acc": msg.ref_accs, "calibrated_val_acc": msg.acc,
            "early_stop": trainer._early_stop_info}


if __name__ == "__main__":
    main()
