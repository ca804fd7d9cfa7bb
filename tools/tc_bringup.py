"""Bring-up aid: run the extractor with the tcgen05 GEMM enabled for ONE 1x1 conv at a time and
compare that layer's output tap with the all-SIMT run of the same mode (GPU box only)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import synth  # noqa: E402
from mermaid_classifier_b200.extractor import EfficientNetExtractor  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
layers = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3, 5, 32, 25]
sd = synth.synth_backbone_state_dict()
im = torch.from_numpy(synth.synth_image(synth.DEFAULT_SEED, 2, 500, 640)).cuda()
pts = synth.synth_points(synth.DEFAULT_SEED, 2, 500, 640, 5, corners=True)
p3 = np.array([(0, r, c) for r, c in pts], dtype=np.int32)
n = len(pts)


def tap_id(layer):
    if layer == 32:
        return 65
    b, which = divmod(layer, 2)
    return 1 + 4 * b + (0 if which == 0 else 3)


def run(mask, tap):
    os.environ["MC_TC_MASK"] = hex(mask)
    ext = EfficientNetExtractor(state_dict=sd, mode=mode, max_batch=16)
    buf = torch.zeros(n * 112 * 112 * 96, dtype=torch.float32, device="cuda")
    ext.set_tap(tap, buf)
    f = ext.extract_device([im], p3)
    torch.cuda.synchronize()
    out = buf.cpu().numpy(), f.cpu().numpy()
    ext.close()
    return out


for layer in layers:
    t = tap_id(layer)
    ref, fref = run(0, t)
    got, fgot = run(1 << layer, t)
    nz = np.flatnonzero(ref)
    ne = nz.max() + 1 if nz.size else 0
    err = np.abs(got[:ne] - ref[:ne]).max() if ne else -1
    print(f"mode {mode} layer {layer} tap {t}: elems {ne} ref absmax {np.abs(ref[:ne]).max():.4f} max abs diff {err:.3e} "
          f"feat diff {np.abs(fgot - fref).max():.3e} nan {np.isnan(got[:ne]).sum()}", flush=True)
