"""C4: classify_features scoring -- N precomputed 1280-d features through the model.pt-style MLP(200,100)/Platt
head (500 classes) on one B200, labels on the device; a CPU-oracle subsample checks label agreement.

    python tools/bench_head.py [--rows 10000000] [--check 200000]
"""
import argparse, json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.inference import DeviceHead
from oracle import head as ohead

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--check", type=int, default=200_000)
ap.add_argument("--hidden", default="200,100")
args = ap.parse_args()
hidden = tuple(int(x) for x in args.hidden.split(","))
w, bb, a, b, _ = synth.synth_head(1280, hidden, 500, seed=0)
head = DeviceHead([x.numpy() for x in w], [x.numpy() for x in bb], a.numpy(), b.numpy())
dev = torch.device("cuda")
chunk = 1_000_000
# swish-like pooled features: generated on the device in 1M-row chunks (51 GB total at 10M rows)
g = torch.Generator(device=dev).manual_seed(0)
feats = torch.empty((args.rows, 1280), dtype=torch.float32, device=dev)
for s in range(0, args.rows, chunk):
    x = torch.randn((min(chunk, args.rows - s), 1280), generator=g, device=dev)
    feats[s:s + x.shape[0]] = torch.clamp(x, min=-0.28) * 0.5 + 0.1
torch.cuda.synchronize()
for _ in range(2):
    head.scores_device(feats[:chunk])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = head.scores_device(feats)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
labels = out["labels"]
n = min(args.check, args.rows)
t0 = time.perf_counter()
want = ohead.calibrated_proba(feats[:n].cpu().numpy(), w, bb, a, b).argmax(1)
cpu_s = time.perf_counter() - t0
agree = float((labels[:n].cpu().numpy() == want).mean())
print(json.dumps({"workload": f"C4: {args.rows} x 1280 fp32 features -> MLP{hidden}/Platt head, 500 classes, labels on device",
                  "features_per_s": args.rows / (ms / 1e3), "ms": ms, "hbm_GBps_algorithmic": args.rows * 5124 / (ms / 1e3) / 1e9,
                  "label_agreement_vs_cpu_oracle": agree, "checked_rows": n,
                  "cpu_oracle_features_per_s": n / cpu_s, "cpu_threads": torch.get_num_threads(), "launches": head.launches}))
