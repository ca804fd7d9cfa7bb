"""Data-parallel MLP training check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py

Parity mode: 2+ ranks with the NCCL all-reduce inside mc_mlp_partial_fit must reproduce the
single-process CPU oracle's loss curve and weights (the oracle is pinned to the reference).
Throughput mode: per-rank shards, global mini-batch = 200 * world; checked for rank-identical weights.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200.torch_classifier import DataParallel, TorchMLPClassifier  # noqa: E402
from oracle import head as ohead  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dp = DataParallel(device=local)
    rng = np.random.RandomState(42)
    K, D, n = 37, 1280, 1650
    centers = rng.randn(K, D) * 3.0
    y = rng.randint(0, K, size=n)
    X = (centers[y] + rng.randn(n, D) * 1.3).astype(np.float32)
    hidden = (200, 100)
    clf = TorchMLPClassifier(hidden_layer_sizes=hidden, learning_rate_init=1e-4, random_state=0).set_device(local)
    clf.enable_data_parallel(dp, "parity")
    for _ in range(2):
        clf.partial_fit(X, y, classes=list(range(K)))
    ws = [lin.weight.detach().numpy() for lin in clf._module.linears]
    flat = torch.from_numpy(np.concatenate([w.reshape(-1) for w in ws])).cuda()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = True
    if rank == 0:
        w, b = ohead.init_mlp(D, hidden, K, 0)
        adam = ohead.AdamState(w + b)
        want = [ohead.partial_fit(w, b, adam, X, y, lr=1e-4, random_state=0) for _ in range(2)]
        err_w = max(float(np.abs(a - r.numpy()).max()) for a, r in zip(ws, w))
        rel = max(abs(a - r) / abs(r) for a, r in zip(clf.loss_curve_, want))
        ok = same and rel < 1e-4 and err_w < 5e-5 and clf.n_steps_ == adam.t
        print(f"parity mode world={world}: loss {clf.loss_curve_} vs oracle {want} (rel {rel:.2e}); "
              f"max |dW| {err_w:.2e}; ranks identical {same}; steps {clf.n_steps_}", flush=True)
    # throughput mode: each rank its own shard
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    clf2 = TorchMLPClassifier(hidden_layer_sizes=hidden, learning_rate_init=1e-4, random_state=0).set_device(local)
    clf2.enable_data_parallel(dp, "throughput")
    clf2.partial_fit(X[lo:hi], y[lo:hi], classes=list(range(K)))
    flat = torch.from_numpy(np.concatenate([l.weight.detach().numpy().reshape(-1) for l in clf2._module.linears])).cuda()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same2 = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        print(f"throughput mode world={world}: loss {clf2.loss_curve_}, ranks identical {same2}, steps {clf2.n_steps_}", flush=True)
        ok = ok and same2
        print("DP_CHECK_OK" if ok else "DP_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
