"""C5: MLP-head training on synthetic 1280-d features, 500 classes, device-resident data; single GPU or
data-parallel under torchrun (NCCL gradient all-reduce inside the C library).

    python tools/bench_train.py [--rows 2000000] [--hidden 500,300,100] [--mode parity|throughput]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_train.py
"""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200.torch_classifier import DataParallel, TorchMLPClassifier

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000, help="rows per rank (throughput mode) / total rows (parity mode)")
ap.add_argument("--hidden", default="500,300,100")
ap.add_argument("--mode", default="throughput", choices=["parity", "throughput"])
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--cpu-rows", type=int, default=20000)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dp = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    dp = DataParallel(device=local)
hidden = tuple(int(x) for x in args.hidden.split(","))
K = 500
# Gaussian clusters as in the reference's tests/pyspacer/test_mlp_benchmark.py:41-63, generated on the device
g = torch.Generator(device=dev).manual_seed(42 + (rank if args.mode == "throughput" else 0))
centers = torch.randn((K, 1280), generator=torch.Generator(device=dev).manual_seed(7), device=dev) * 3.0
y = torch.randint(0, K, (args.rows,), generator=g, device=dev, dtype=torch.int64)
X = torch.empty((args.rows, 1280), dtype=torch.float32, device=dev)
for s in range(0, args.rows, 500_000):
    e = min(args.rows, s + 500_000)
    X[s:e] = centers[y[s:e]] + torch.randn((e - s, 1280), generator=g, device=dev) * 1.3
y32 = y.to(torch.int32)
clf = TorchMLPClassifier(hidden_layer_sizes=hidden, learning_rate_init=1e-4, random_state=0, alpha=1e-4).set_device(local)
clf.init_for(1280, list(range(K)))
if dp is not None:
    clf.enable_data_parallel(dp, args.mode)
clf.partial_fit_device(X[:20000], y32[:20000])  # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(args.epochs):
    clf.partial_fit_device(X, y32)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
rows_total = args.rows * (world if args.mode == "throughput" else 1) * args.epochs
if rank == 0:
    line = {"workload": f"C5: MLP{hidden} training, 500 classes, {args.rows} rows x 1280 {'per rank' if args.mode == 'throughput' else 'total'}, "
                        f"mini-batch 200{' per rank' if args.mode == 'throughput' and world > 1 else ''}, {args.epochs} passes, device-resident",
            "n_gpus": world, "mode": args.mode if world > 1 else "single", "samples_per_s": rows_total / dt,
            "adam_steps_per_s": (clf.n_steps_ - 100) / dt, "loss_curve": clf.loss_curve_[1:], "launches": clf.launches}
    if args.cpu_rows:
        from oracle import head as ohead
        Xc, yc = X[:args.cpu_rows].cpu().numpy(), y[:args.cpu_rows].cpu().numpy()
        w, b = ohead.init_mlp(1280, hidden, K, 0)
        adam = ohead.AdamState(w + b)
        t1 = time.perf_counter()
        ohead.partial_fit(w, b, adam, Xc, yc, lr=1e-4, random_state=0)
        line["cpu_oracle_samples_per_s"] = args.cpu_rows / (time.perf_counter() - t1)
        line["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
