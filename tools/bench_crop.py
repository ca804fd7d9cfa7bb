"""Stand-alone crop (mc_crop_patches) bandwidth: bytes gathered + bytes written per second against the HBM peak."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mermaid_classifier_b200 import synth
from mermaid_classifier_b200.extractor import crop_patches_device, synth_image_device
n_img, n_pts = 200, 100
ims = [synth_image_device(synth.DEFAULT_SEED, i, 3000, 4000) for i in range(n_img)]
pts = np.array([(i, r, c) for i in range(n_img) for r, c in synth.synth_points(synth.DEFAULT_SEED, i, 3000, 4000, n_pts)], dtype=np.int32)
for _ in range(3):
    out = crop_patches_device(ims, pts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = crop_patches_device(ims, pts)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
n = pts.shape[0]
peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else 6650.0
gbs = n * 2 * 150528 / (ms / 1e3) / 1e9
print(json.dumps({"kernel": "crop_kernel (mc_crop_patches)", "patches": n, "ms": ms, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak,
                  "patches_per_s": n / (ms / 1e3)}))
