"""Bring-up check of the fused expand+depthwise kernel (MC_FUSE_MASK) against the two-kernel path.

    python tools/fused_check.py --mask 2 [--images 2] [--points 100] [--mode fp32]

Prints one JSON line: max-abs / bit-equality of the final features with the mask on vs off, and the per-layer
CUDA-event times of both runs for the blocks the mask selects.  Run each mask in its own process: a barrier
time-out traps and kills the CUDA context."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run(mask, args, sd):
    import torch

    from mermaid_classifier_b200 import _lib, synth
    from mermaid_classifier_b200.extractor import EfficientNetExtractor, synth_image_device

    if mask:
        os.environ["MC_FUSE_MASK"] = mask
    else:
        os.environ.pop("MC_FUSE_MASK", None)
    H, W = args.height, args.width
    ext = EfficientNetExtractor(state_dict=sd, mode=args.mode, max_batch=args.batch)
    images = [synth_image_device(13, i, H, W) for i in range(args.images)]
    pts = np.array([(i, r, c) for i in range(args.images)
                    for r, c in synth.synth_points(13, i, H, W, args.points, corners=True)], dtype=np.int32)
    lib = _lib.load()
    h = ext._ensure_handle()
    out = ext.extract_device(images, pts)
    torch.cuda.synchronize()
    _lib.check(lib.mc_extractor_profile(h, -2))
    for _ in range(args.reps):
        out = ext.extract_device(images, pts)
    torch.cuda.synchronize()
    ms = np.zeros(67)
    cnt = np.zeros(67, dtype=np.int64)
    _lib.check(lib.mc_extractor_profile_read(h, ms.ctypes.data, cnt.ctypes.data, 1))
    res = out.cpu().numpy()
    ext.close()
    return res, ms / args.reps, len(pts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mask", default="2")
    ap.add_argument("--images", type=int, default=2)
    ap.add_argument("--points", type=int, default=100)
    ap.add_argument("--height", type=int, default=600)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--batch", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mode", default="fp32")
    args = ap.parse_args()
    from mermaid_classifier_b200 import synth

    sd = synth.synth_backbone_state_dict()
    want, ms0, n = run("", args, sd)
    got, ms1, _ = run(args.mask, args, sd)
    m = int(args.mask, 16)
    layers = {}
    for b in range(16):
        if (m >> b) & 1:
            layers[f"b{b}"] = {"two_kernel_ms": round(float(ms0[1 + 4 * b] + ms0[2 + 4 * b]), 4),
                               "fused_ms": round(float(ms1[1 + 4 * b] + ms1[2 + 4 * b]), 4)}
    print(json.dumps({"mask": args.mask, "patches": n, "bit_equal": bool(np.array_equal(got, want)),
                      "max_abs": float(np.abs(got - want).max()), "nan": bool(np.isnan(got).any()),
                      "total_ms": [round(float(ms0.sum()), 3), round(float(ms1.sum()), 3)], "layers": layers}), flush=True)


if __name__ == "__main__":
    main()
