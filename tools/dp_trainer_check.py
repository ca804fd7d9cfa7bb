"""MermaidTrainer under data parallelism, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_trainer_check.py

parity mode     : every rank holds the fixture's splits; the per-epoch metrics must be the reference run's
                  (tests/golden/trainer_eval.npz, same tolerances as the single-GPU test) and identical on all ranks.
throughput mode : every rank holds its own shard of each split; all ranks must report identical metrics, identical
                  calibrators and identical weights, and the reported val metrics must equal a single-process
                  evaluation of the final model over the whole val split.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mermaid_classifier_b200.torch_classifier import DataParallel  # noqa: E402
from mermaid_classifier_b200.trainer import DeviceLabels, MermaidTrainer, TaskLabels  # noqa: E402


def same_on_all_ranks(values, world):
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return all(torch.equal(parts[0], p) for p in parts)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dp = DataParallel(device=local)
    g = np.load(ROOT / "tests" / "golden" / "trainer_eval.npz")
    cls = g["classes"]
    ok = True

    # ---- parity -----------------------------------------------------------------------------------------------
    labels = TaskLabels(train=DeviceLabels(g["Xt"], cls[g["yt"]]), ref=DeviceLabels(g["Xr"], cls[g["yr"]]),
                        val=DeviceLabels(g["Xv"], cls[g["yv"]]))
    seen = []
    tr = MermaidTrainer(batch_size=int(g["chunk"]), on_epoch_end=seen.append, hidden_layer_sizes=(24, 16),
                        learning_rate_init=1e-3, device=local, data_parallel=dp, dp_mode="parity")
    cal, _, msg = tr(labels, int(g["epochs"]), [])
    vl = [m["val_loss"] for m in seen]
    a, b = cal.platt
    p_ok = (np.allclose(vl, g["val_loss"], rtol=2e-5) and np.allclose([m["training_loss"] for m in seen], g["train_loss"], rtol=2e-5)
            and np.allclose(msg.ref_accs, g["ref_acc"], atol=1.01 / len(g["yr"])) and np.allclose(a, g["platt_a"], rtol=5e-3)
            and same_on_all_ranks(vl + list(a) + list(b), world))
    if rank == 0:
        print(f"parity world={world}: val_loss {np.round(vl, 6).tolist()} vs reference run {np.round(g['val_loss'], 6).tolist()} -> {p_ok}", flush=True)
    ok = ok and p_ok

    # ---- throughput: rank-private shards ------------------------------------------------------------------------
    def shard(X, y):
        lo, hi = (len(y) * rank) // world, (len(y) * (rank + 1)) // world
        return DeviceLabels(X[lo:hi], cls[y[lo:hi]])

    nt = (len(g["yt"]) // (world * 100)) * world * 100  # equal shards, equal chunk counts
    labels_s = TaskLabels(train=shard(g["Xt"][:nt], g["yt"][:nt]), ref=shard(g["Xr"], g["yr"]), val=shard(g["Xv"], g["yv"]))
    seen2 = []
    tr2 = MermaidTrainer(batch_size=100, on_epoch_end=seen2.append, early_stopping_patience=3, hidden_layer_sizes=(24, 16),
                         learning_rate_init=1e-3, device=local, data_parallel=dp, dp_mode="throughput")
    cal2, _, msg2 = tr2(labels_s, 4, [])
    a2, b2 = cal2.platt
    w_flat = np.concatenate([l.weight.detach().numpy().reshape(-1) for l in cal2.estimator._module.linears])
    metrics = [m["val_loss"] for m in seen2] + [m["val_accuracy"] for m in seen2] + msg2.ref_accs
    t_same = same_on_all_ranks(metrics + list(a2) + list(b2) + w_flat.tolist(), world)
    # the reduced val metrics equal a single-process pass over the whole val split
    single = MermaidTrainer(batch_size=100, hidden_layer_sizes=(24, 16), device=local)
    acc_full, loss_full = single._calc_acc_and_log_loss_batched(cal2.estimator, DeviceLabels(g["Xv"], cls[g["yv"]]), list(cls))
    last = seen2[-1]
    restored_is_last = tr2._early_stop_info["best_val_epoch"] == tr2._early_stop_info["final_epoch"]
    t_ok = t_same and np.isfinite(a2).all() and (not restored_is_last or (
        abs(loss_full - last["val_loss"]) <= 1e-12 * abs(loss_full) and abs(acc_full - last["val_accuracy"]) < 1e-12))
    if rank == 0:
        print(f"throughput world={world}: val_loss {np.round([m['val_loss'] for m in seen2], 6).tolist()}, full-split check "
              f"{loss_full:.9f} vs {last['val_loss']:.9f}, ranks identical {t_same} -> {t_ok}", flush=True)
    ok = ok and t_ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_TRAINER_OK" if int(flag) else "DP_TRAINER_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
